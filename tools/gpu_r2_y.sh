# round 2, call Y: AdamW behind backward as short-lived CTAs, step on a high-priority stream
set -x
mkdir -p gpurun_out
for opts in "" "hp=1" "overlap=1,ctas=-1" "overlap=1,ctas=-1,hp=1" "overlap=1,ctas=-4,hp=1,buckets=8" "overlap=1,ctas=-1,hp=1,buckets=12" "overlap=1,ctas=-2,hp=1,buckets=6"; do
  MEBT_TRAIN_OPTS="$opts" timeout 300 python bench.py --workload train16f --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/tmp_train.json 2>> gpurun_out/bench_err.log
  python - "$opts" <<'PY'
import json,sys
j=json.loads(open('gpurun_out/tmp_train.json').read().strip().splitlines()[-1])
print('train16f [%s]' % sys.argv[1], round(j['ms_per_step'],3), 'ms', round(j['value']), 'e2e', round(j['e2e']['value']))
PY
done
