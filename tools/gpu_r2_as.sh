set -x
timeout 300 python -m pytest tests/test_dropin_gpu.py -x -q -k "vtokens_false or pipelines" 2>&1 | tail -12
