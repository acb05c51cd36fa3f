# round 2, last check of the committed tree: GPU suite, smoke, default bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'frac', round(j['roofline']['frac'],4))
for n,w in j['workloads'].items():
    print(' ', n, round(w.get('value',0)), round(w.get('ms_per_step',0),2), w.get('error'))
PY
