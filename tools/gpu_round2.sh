set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_train16f_e.json
tail -3 gpurun_out/bench_err.log
