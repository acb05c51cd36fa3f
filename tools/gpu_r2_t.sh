# round 2, call T: attention backward, 2 vs 4 softmax warpgroups
set -x
timeout 120 python tools/attn_bwd_probe.py 2>&1 | tail -8
MEBT_ATTN_BWD_GROUPS=4 timeout 120 python tools/attn_bwd_probe.py 2>&1 | tail -8
MEBT_ATTN_BWD_GROUPS=4 timeout 300 python -m pytest tests/test_backward_gpu.py tests/test_dropout_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/train_sweep.py --base 2>&1 | tail -1
MEBT_ATTN_BWD_GROUPS=4 timeout 300 python tools/train_sweep.py --base 2>&1 | tail -1
