# round 2, call K: split-KV attention for small-batch sampling
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_stl_config_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --workload sample128f --batch 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_sample128f_b2.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02k_bench_sample128f_b2.json').read().strip().splitlines()[-1])
print('B=2', j['value'], j['ms_per_step'], j['roofline']['families_ms'])
PY
timeout 600 python bench.py --workload sample128f --batch 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_sample128f_b1.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02k_bench_sample128f_b1.json').read().strip().splitlines()[-1])
print('B=1', j['value'], j['ms_per_step'], j['roofline']['families_ms'])
PY
