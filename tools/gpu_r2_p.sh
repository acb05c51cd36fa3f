# round 2, call P: headline numbers after the attention work
set -x
mkdir -p gpurun_out
timeout 300 python tools/train_sweep.py --base 2>&1 | tail -2
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_sample128f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02p_bench_sample128f.json').read().strip().splitlines()[-1])
print('sample128f', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
timeout 600 python bench.py --workload sample128f --batch 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_sample128f_b2.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02p_bench_sample128f_b2.json').read().strip().splitlines()[-1])
print('B=2', j['value'], j['ms_per_step'], j['roofline']['families_ms'])
PY
