# round 2, call AI: two-issuer (DUAL) variant of the 128-wide GEMM: tests, then train16f with and without it
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -6
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_forward_gpu.py tests/test_training_gpu.py tests/test_stl_config_gpu.py -x -q 2>&1 | tail -6
run() {
  env $1 timeout 300 python bench.py --workload train16f --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/tmp_train.json 2>> gpurun_out/bench_err.log
  python - "$1" <<'PY'
import json,sys
j=json.loads(open('gpurun_out/tmp_train.json').read().strip().splitlines()[-1])
print('train16f [%s]' % sys.argv[1], round(j['ms_per_step'],3), 'ms', round(j['value']), 'e2e', round(j['e2e']['value']), 'gemm_ms', j['roofline']['families_ms']['gemm'], 'frac', round(j['roofline']['frac'],4))
PY
}
run "MEBT_GEMM_DUAL=0"
run "MEBT_GEMM_DUAL=1"
run "MEBT_GEMM_DUAL=0"
run "MEBT_GEMM_DUAL=1"
