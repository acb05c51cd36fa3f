# round 2, call C: event-ordered side streams, background AdamW sweep, measured-config parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/train_sweep.py 2>&1 | tail -20
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2c.json 2>&1 | tail -2
