# round 2, final measurement batch on the final tree: GPU suite, default bench + reference arm, the other workloads, ncu
# launch lists, one --set full pass over the hot kernels, device timeline, SASS evidence is static (built here)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload sample128f --batch 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_sample128f_b2.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload maskgit16f --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_maskgit16f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload vq16f --steps 10 --warmup 3 > gpurun_out/r02_bench_vq16f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 > gpurun_out/r02_bench_vqgan16f.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
for n in ['default','sample128f_b2','maskgit16f','vq16f','vqgan16f']:
    try:
        j=json.loads(open('gpurun_out/r02_bench_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'frac', round(j['roofline']['frac'],4), 'traffic', j['roofline'].get('traffic'))
        if j.get('workloads'):
            w=j['workloads']['sample128f']; print('  sample128f', round(w['value']), round(w['ms_per_step'],2), round(w['roofline']['frac'],4))
    except Exception as e: print(n, 'ERR', e)
PY
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r02.json 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 900 --csv --log-file gpurun_out/r02_launches_train16f.csv python bench.py --workload train16f --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_list_train.log 2>&1
tail -1 gpurun_out/ncu_list_train.log | cut -c1-200
K='regex:gemm_bf16|gemm_grouped|latent_attention|attention_combine|attn_bwd|sample_stream|masked_ce|layernorm|embed_gather|vq_|colsum|row_gather'
timeout 600 ncu --set full --clock-control none -k "$K" -o /tmp/prof_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
tail -2 gpurun_out/ncu_targets.log
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/r02_prof_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r02_prof_kernels_raw.csv
