# round 2, call I: tensor-core codebook search, streaming masked-CE kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropin_gpu.py tests/test_training_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --workload vq16f --no-cpu-baseline > gpurun_out/r02i_bench_vq16f.json 2> gpurun_out/r02i_bench_vq16f.err; tail -2 gpurun_out/r02i_bench_vq16f.err; head -c 900 gpurun_out/r02i_bench_vq16f.json
timeout 300 python tools/train_sweep.py --base 2>&1 | tail -2
timeout 300 python tools/train_probe.py 6 0.1 2>&1 | tail -2
