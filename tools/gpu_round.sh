set -x
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r1_b.json
tail -3 gpurun_out/bench_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 15200 -c 5100 --csv --log-file gpurun_out/launches_sample128f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6000 -c 4 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sample_logits|latent_attention|layernorm" -s 900 -c 6 -o gpurun_out/prof_misc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
ls -la gpurun_out/
