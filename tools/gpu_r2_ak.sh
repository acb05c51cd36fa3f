# round 2, call AK: pad_norm_act with four loads in flight per thread, up to 256 slabs for the GroupNorm statistics
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vqgan_gpu.py -x -q 2>&1 | tail -5
for b in 8 2; do
timeout 600 python bench.py --workload vqgan16f --batch $b --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ak_bench_vqgan16f_$b.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - $b <<'PY'
import json,sys
j=json.loads(open('gpurun_out/r02ak_bench_vqgan16f_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('vqgan16f B', sys.argv[1], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'], j['roofline'].get('layernorm_kernel_hbm'))
PY
done
