# round 2, final measurement batch, second pass (after the DUAL race fix; GEMM DUAL off by default)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_err.log; tail -2 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), 'frac', round(j['roofline']['frac'],4), 'launches', j['gpu_launches'], j['roofline']['families_ms'])
w=j['workloads']['sample128f']; print('  sample128f', round(w['value']), round(w['ms_per_step'],2), round(w['roofline']['frac'],4), w['roofline']['families_ms'])
PY
timeout 200 python tools/timeline.py --workload train16f --out gpurun_out/timeline_train16f_r02.json 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 900 --csv --log-file gpurun_out/r02_launches_train16f.csv python bench.py --workload train16f --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_list_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r02_launches_sample128f.csv python bench.py --workload sample128f --batch 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_sample.log 2>&1
K='regex:gemm_bf16|gemm_grouped|latent_attention|attention_combine|attn_bwd|sample_stream|masked_ce|layernorm|embed_gather|vq_|colsum|row_gather'
timeout 600 ncu --set full --clock-control none -k "$K" -o /tmp/prof_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
tail -2 gpurun_out/ncu_targets.log
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/r02_prof_kernels_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k 'regex:conv3d|pad_norm|groupnorm' -o /tmp/prof_vqgan -f python tools/ncu_targets.py --vqgan > gpurun_out/ncu_vqgan.log 2>&1
ncu -i /tmp/prof_vqgan.ncu-rep --page raw --csv > gpurun_out/r02_prof_vqgan_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_vqgan16f.csv python bench.py --workload vqgan16f --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_vqgan.log 2>&1
ls -la gpurun_out/r02_prof_kernels_raw.csv gpurun_out/r02_prof_vqgan_raw.csv
