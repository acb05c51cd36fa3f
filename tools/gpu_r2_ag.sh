# round 2, call AG: direct stores of the 8-channel last convolution; teacher-forced edit / fixed-context sampling loop
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vqgan_gpu.py tests/test_sampler_loops_gpu.py -q 2>&1 | tail -8
timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ag_bench_vqgan16f.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02ag_bench_vqgan16f.json').read().strip().splitlines()[-1])
print('vqgan16f', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_vqgan16f.csv python bench.py --workload vqgan16f --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_vqgan.log 2>&1
