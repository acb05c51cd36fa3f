# round 2, call M: pair-mode rule for 256-wide tiles from 96 tiles up; bf16 sampler logits
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_forward_gpu.py tests/test_backward_gpu.py tests/test_training_gpu.py tests/test_sampler_loops_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/train_sweep.py --base 2>&1 | tail -2
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_sample128f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02m_bench_sample128f.json').read().strip().splitlines()[-1])
print('sample128f', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['families_ms'], j['detail'])
PY
