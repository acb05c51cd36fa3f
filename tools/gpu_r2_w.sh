# round 2, call W: VQGAN encode + decode bench (vqgan16f), launch list, ncu pass over the convolution kernels
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 > gpurun_out/r02_bench_vqgan16f.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_vqgan16f.json').read().strip().splitlines()[-1])
print('vqgan16f', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'], j['roofline']['families_launches'], j['cpu_baseline'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_vqgan16f.csv python bench.py --workload vqgan16f --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_vqgan.log 2>&1
tail -2 gpurun_out/ncu_list_vqgan.log
timeout 300 ncu --set full --clock-control none -k 'regex:conv3d|pad_norm|groupnorm' -o /tmp/prof_vqgan -f python tools/ncu_targets.py --vqgan > gpurun_out/ncu_vqgan.log 2>&1
tail -3 gpurun_out/ncu_vqgan.log
ncu -i /tmp/prof_vqgan.ncu-rep --page raw --csv > gpurun_out/r02_prof_vqgan_raw.csv 2>/dev/null
