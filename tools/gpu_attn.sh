set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -m gpu -x -q --timeout 300 -k "attention" 2>&1 | tail -5
for b in head t-1_p0 t-1_p1 t3_p1; do echo "== $b"; timeout 120 tools/attn_bench_$b; done
