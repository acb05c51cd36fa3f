# round 2, call AA: ROW-mode convolution (shifted descriptor starts) with / without the matrix-base-offset field, 128-wide tiles
set -x
mkdir -p gpurun_out
MEBT_CONV_BASEOFF=0 timeout 600 python -m pytest tests/test_vqgan_gpu.py -q -k "same_pad_conv3d" 2>&1 | tail -12
MEBT_CONV_BASEOFF=1 timeout 600 python -m pytest tests/test_vqgan_gpu.py -q -k "same_pad_conv3d" 2>&1 | tail -12
MEBT_CONV_ROW=0 timeout 600 python -m pytest tests/test_vqgan_gpu.py -q 2>&1 | tail -4
for bo in 0 1; do
MEBT_CONV_BASEOFF=$bo timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench_vqgan16f_$bo.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - $bo <<'PY'
import json,sys
j=json.loads(open('gpurun_out/r02aa_bench_vqgan16f_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('vqgan16f baseoff', sys.argv[1], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
done
