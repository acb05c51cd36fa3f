# round 2, call Z: full GPU suite + default bench line + reference arm on the current tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02z_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench_default.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02z_bench_default.json').read().strip().splitlines()[-1])
print('default', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'launches', j['gpu_launches'])
w=j['workloads']['sample128f']; print('sample128f', w['value'], w['ms_per_step'], w['roofline']['frac'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference.json 2>> gpurun_out/bench_err.log; tail -c 600 gpurun_out/r02z_bench_reference.json
