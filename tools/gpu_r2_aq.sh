# round 2, call AQ: default bench line with both secondary workloads
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'frac', round(j['roofline']['frac'],4), 'steps', j['steps'], j['warmup'])
for n,w in j['workloads'].items():
    print(' ', n, round(w['value']), round(w['ms_per_step'],2), 'e2e', round(w['e2e']['value']), round(w['roofline']['frac'],4), 'traffic', w['roofline'].get('traffic'))
PY
