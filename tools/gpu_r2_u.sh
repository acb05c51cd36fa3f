set -x
timeout 600 python -m pytest tests/test_training_gpu.py -x -q -m gpu -k fused 2>&1 | grep -E "^E|^>|passed|failed" | head -20
