set -x
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -5
timeout 120 tools/gemm_bench train > gpurun_out/gb_train.log 2>&1
timeout 120 tools/gemm_bench sample > gpurun_out/gb_sample.log 2>&1
cat gpurun_out/gb_train.log gpurun_out/gb_sample.log
