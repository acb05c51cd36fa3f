# Round-end measurement batch (1 GPU): benches, ncu launch list of the training step, one ncu --set full pass over the
# library's own hot kernels at BASELINE-sized shapes (report kept on the box, only its raw CSV page comes back).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_train16f.json; tail -2 gpurun_out/bench_err.log
timeout 900 python bench.py --workload sample128f --steps 2 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_sample128f.json; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload maskgit16f --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_maskgit16f.json; tail -2 gpurun_out/bench_err.log
timeout 600 python bench.py --workload vq16f --steps 10 --warmup 3 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_vq16f.json; tail -2 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_err.log > gpurun_out/r01_bench_train16f_reference.json; tail -2 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 1000 --csv --log-file gpurun_out/launches_train16f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_train.log 2>&1
K='regex:gemm_bf16|latent_attention|attn_bwd|attn_delta|sample_stream|masked_ce|layernorm|embed_gather|vq_argmin|colsum|dropout_rows|row_gather'
timeout 420 ncu --set full --clock-control none -k "$K" -o /tmp/prof_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
tail -3 gpurun_out/ncu_targets.log
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/prof_kernels.ncu-rep | tail -14
