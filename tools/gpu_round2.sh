set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_train16f.json
python -c "import json; j=json.load(open('gpurun_out/r01_bench_train16f.json')); print(j['value'], j['ms_per_step'], j['roofline']['families_ms'])"
tail -3 gpurun_out/bench_err.log
