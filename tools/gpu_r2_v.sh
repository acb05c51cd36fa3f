# round 2, call V: VQGAN convolutions (5-D TMA implicit GEMM) - first contact with the hardware, under a short timeout
set -x
timeout 300 python -m pytest tests/test_vqgan_gpu.py -x -q -m gpu 2>&1 | tail -25
