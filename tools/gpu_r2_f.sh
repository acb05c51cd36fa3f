# round 2, call F: key side of every block on stream C (forward + backward), 4-deep masked-copy ring
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -3
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2f.json 2>&1 | tail -2
timeout 400 python tools/train_sweep.py 2>&1 | tail -20
