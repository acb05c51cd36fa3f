# round 2, call AR: row pairs (ROW = 2) in the convolution kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vqgan_gpu.py -x -q 2>&1 | tail -5
MEBT_CONV_ROW2=1 timeout 300 python tools/vqgan_stress.py 300 8 2>&1 | tail -2
for r in 0 1; do
MEBT_CONV_ROW2=$r timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ar_bench_vqgan16f_$r.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - $r <<'PY'
import json,sys
j=json.loads(open('gpurun_out/r02ar_bench_vqgan16f_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('vqgan16f row2', sys.argv[1], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
done
