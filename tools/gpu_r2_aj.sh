# round 2, call AJ: DUAL issuers in the per-tap convolution kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vqgan_gpu.py -x -q 2>&1 | tail -5
for d in 0 1 0 1; do
MEBT_CONV_DUAL=$d timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02aj_bench_vqgan16f_$d.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - $d <<'PY'
import json,sys
j=json.loads(open('gpurun_out/r02aj_bench_vqgan16f_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('vqgan16f dual', sys.argv[1], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_vqgan16f.csv python bench.py --workload vqgan16f --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_vqgan.log 2>&1
