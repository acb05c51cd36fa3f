"""ncu target: the VQGAN's last convolution (64 -> 3 channels, k 3, 16 x 128 x 128, 2 videos) next to its 64 -> 64 twin."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from mebt_b200.vqgan import SamePadConv3d, to_channels_last
torch.manual_seed(0)
x = to_channels_last(torch.randn(2, 64, 16, 128, 128, device="cuda"))
for cout in (64, 3, 32):
    m = SamePadConv3d(64, cout, 3).cuda()
    for _ in range(2):
        m.forward_cl(x)
torch.cuda.synchronize()
