# round 2, call N: sampling fused into the head GEMM (Gumbel-max epilogue)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropin_gpu.py tests/test_sampler_loops_gpu.py tests/test_forward_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench_sample128f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02n_bench_sample128f.json').read().strip().splitlines()[-1])
print('sample128f', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
