# round 2, call AE: compute-sanitizer memcheck over the small kernel tests (convolutions incl. ROW mode, GroupNorm pass, GEMM
# modes, attention forward / backward, the micro training step)
set -x
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 9 --print-limit 20"
( timeout 900 $S python -m pytest tests/test_vqgan_gpu.py -x -q -k "pad_norm or same_pad_conv or transpose" 2>&1 | tail -25 ) > gpurun_out/r02_sanitizer_vqgan.log
tail -6 gpurun_out/r02_sanitizer_vqgan.log
( timeout 900 $S python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -25 ) > gpurun_out/r02_sanitizer_gemm.log
tail -6 gpurun_out/r02_sanitizer_gemm.log
( timeout 900 $S python -m pytest tests/test_kernels_gpu.py -x -q -k "latent_attention or masked_ce or layernorm or embed" 2>&1 | tail -25 ) > gpurun_out/r02_sanitizer_kernels.log
tail -6 gpurun_out/r02_sanitizer_kernels.log
( timeout 900 $S python -m pytest tests/test_backward_gpu.py -x -q 2>&1 | tail -25 ) > gpurun_out/r02_sanitizer_backward.log
tail -6 gpurun_out/r02_sanitizer_backward.log
