# round 2, call AB: VQGAN tests, launch list and ncu rows after ROW mode / 128-wide tiles
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vqgan_gpu.py -q 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_vqgan16f.csv python bench.py --workload vqgan16f --batch 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_vqgan.log 2>&1
tail -2 gpurun_out/ncu_list_vqgan.log
timeout 400 ncu --set full --clock-control none -k 'regex:conv3d|pad_norm|groupnorm' -o /tmp/prof_vqgan -f python tools/ncu_targets.py --vqgan > gpurun_out/ncu_vqgan.log 2>&1
tail -3 gpurun_out/ncu_vqgan.log
ncu -i /tmp/prof_vqgan.ncu-rep --page raw --csv > gpurun_out/r02_prof_vqgan_raw.csv 2>/dev/null
