# round 2, call AN: DUAL with ring-parity stage ownership: stress loops (VQGAN path, GEMM), tests, timing
set -x
mkdir -p gpurun_out
for i in 1 2 3; do
MEBT_CONV_DUAL=1 timeout 300 python tools/vqgan_stress.py 500 8 2>&1 | tail -2; echo "conv dual=1 rc=$?"
done
MEBT_CONV_DUAL=1 timeout 300 python tools/vqgan_stress.py 800 2 2>&1 | tail -2
timeout 600 python tools/gemm_dual_stress.py 2000 2>&1 | tail -9
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vqgan_gpu.py tests/test_backward_gpu.py tests/test_training_gpu.py -x -q 2>&1 | tail -3
for d in 0 1; do
MEBT_GEMM_DUAL=$d timeout 300 python bench.py --workload train16f --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/tmp_train.json 2>> gpurun_out/bench_err.log
python - $d <<'PY'
import json,sys
j=json.loads(open('gpurun_out/tmp_train.json').read().strip().splitlines()[-1])
print('train16f gemm dual', sys.argv[1], round(j['ms_per_step'],3), 'ms', round(j['value']))
PY
done
timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 > gpurun_out/r02_bench_vqgan16f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_vqgan16f.json').read().strip().splitlines()[-1])
print('vqgan16f', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'], j['cpu_baseline'])
PY
