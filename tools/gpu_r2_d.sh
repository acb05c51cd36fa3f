# round 2, call D: grouped weight-gradient launch, sampler loop parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -3
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2d.json 2>&1 | tail -2
