# round 2, call AD: SM partition between the backward chain and the side-stream weight-gradient launch
set -x
mkdir -p gpurun_out
run() {
  env $1 timeout 300 python bench.py --workload train16f --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/tmp_train.json 2>> gpurun_out/bench_err.log
  python - "$1" <<'PY'
import json,sys
j=json.loads(open('gpurun_out/tmp_train.json').read().strip().splitlines()[-1])
print('train16f [%s]' % sys.argv[1], round(j['ms_per_step'],3), 'ms', round(j['value']), 'e2e', round(j['e2e']['value']))
PY
}
run "MEBT_X=0"
run "MEBT_WGRAD_CTAS=52"
run "MEBT_WGRAD_CTAS=44"
run "MEBT_WGRAD_CTAS=52 MEBT_BWD_CHAIN_SMS=96"
run "MEBT_WGRAD_CTAS=36 MEBT_BWD_CHAIN_SMS=112"
run "MEBT_WGRAD_CTAS=72 MEBT_BWD_CHAIN_SMS=76"
run "MEBT_WGRAD_CTAS=28 MEBT_BWD_CHAIN_SMS=120"
