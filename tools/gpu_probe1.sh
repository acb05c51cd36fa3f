set -x
python tools/gemm_probe.py train > gpurun_out/probe_train.log 2>&1
python tools/gemm_probe.py sample > gpurun_out/probe_sample.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1100 --csv --log-file gpurun_out/launches_train16f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_train.log 2>&1
tail -2 gpurun_out/ncu_list_train.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|latent_attention" -o gpurun_out/prof_gemm2 python tools/ncu_targets.py > gpurun_out/ncu_gemm2.log 2>&1
tail -2 gpurun_out/ncu_gemm2.log
ls -la gpurun_out/
