# round 2, call G: delta from the proj-dgrad epilogue; experiments: PDL fence after forks, weight-gradient CTA cap
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -2
MEBT_FORK_NOPDL=1 timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -2
MEBT_WGRAD_CTAS=64 timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -2
MEBT_WGRAD_CTAS=10000 timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -2
MEBT_FORK_NOPDL=1 MEBT_WGRAD_CTAS=64 timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -2
MEBT_FORK_NOPDL=1 timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2g_fence.json 2>&1 | tail -1
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2g.json 2>&1 | tail -1
