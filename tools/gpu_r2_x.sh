# round 2, call X: vectorised GroupNorm / padding kernels (tests + vqgan16f bench), AdamW-overlap knobs on train16f,
# launch list + ncu pass over the convolution kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vqgan_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload vqgan16f --steps 5 --warmup 3 > gpurun_out/r02x_bench_vqgan16f.json 2> gpurun_out/bench_err.log; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02x_bench_vqgan16f.json').read().strip().splitlines()[-1])
print('vqgan16f', j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['families_ms'], j['roofline']['families_launches'], j['roofline'].get('layernorm_kernel_hbm'))
PY
for opts in "" "overlap=1" "overlap=1,ctas=24" "overlap=1,ctas=96,buckets=8" "overlap=1,ctas=16,buckets=12"; do
  MEBT_TRAIN_OPTS="$opts" timeout 300 python bench.py --workload train16f --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/tmp_train.json 2>> gpurun_out/bench_err.log
  python - "$opts" <<'PY'
import json,sys
j=json.loads(open('gpurun_out/tmp_train.json').read().strip().splitlines()[-1])
print('train16f [%s]' % sys.argv[1], round(j['ms_per_step'],3), 'ms', round(j['value']), 'e2e', round(j['e2e']['value']))
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_vqgan16f.csv python bench.py --workload vqgan16f --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_vqgan.log 2>&1
tail -2 gpurun_out/ncu_list_vqgan.log
timeout 400 ncu --set full --clock-control none -k 'regex:conv3d|pad_norm|groupnorm' -o /tmp/prof_vqgan -f python tools/ncu_targets.py --vqgan > gpurun_out/ncu_vqgan.log 2>&1
tail -3 gpurun_out/ncu_vqgan.log
ncu -i /tmp/prof_vqgan.ncu-rep --page raw --csv > gpurun_out/r02_prof_vqgan_raw.csv 2>/dev/null
ls -la gpurun_out/r02_prof_vqgan_raw.csv
