set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_train16f_f.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['roofline']['ce_kernel_hbm'], j['roofline']['families_ms'])"
tail -3 gpurun_out/bench_err.log
