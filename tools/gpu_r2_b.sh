# round 2, call B: fused dropout / merged attention backward / two side streams
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 200 python tools/train_probe.py 6 0.1 2>&1 | tail -3
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r2b.json 2>&1 | tail -2
timeout 120 tools/gemm_bench train 2>&1 | tail -12
