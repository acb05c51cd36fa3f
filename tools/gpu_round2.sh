set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8
timeout 60 tools/attn_bench
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_train16f_c.json
tail -3 gpurun_out/bench_err.log
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_sample128f_c.json
tail -3 gpurun_out/bench2_err.log
