# round 2, call O: attention forward issuer: four / eight MMAs per instruction block over split descriptor words
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py tests/test_dropout_gpu.py tests/test_stl_config_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 120 python tools/attn_probe.py 2>&1 | tail -8
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_sample128f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02o_bench_sample128f.json').read().strip().splitlines()[-1])
print('sample128f', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['families_ms'])
PY
