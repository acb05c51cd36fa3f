# round 2, call R: attention backward with two column-split softmax warpgroups
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_training_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 120 python tools/attn_bwd_probe.py 2>&1 | tail -10
timeout 300 python tools/train_sweep.py --base 2>&1 | tail -2
