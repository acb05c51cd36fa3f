set -x
timeout 300 python -m pytest tests/test_vqgan_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -3
