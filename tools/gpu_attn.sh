set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -m gpu -x -q --timeout 300 -k "attention" 2>&1 | tail -15
for b in p0 p1; do echo "== $b"; timeout 120 tools/attn_bench_$b; done
