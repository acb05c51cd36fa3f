"""LayerNorm forward timing at the 128-frame token-row scale (CUDA events, 20 launches over a buffer larger than L2)."""
import sys
import torch
sys.path.insert(0, ".")
from mebt_b200 import ops  # noqa: E402
g, b = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
for rows in (4096, 65536, 131072):
    x = torch.randn(rows, 1024, device="cuda").bfloat16()
    y = torch.empty_like(x)
    for _ in range(3):
        ops.layernorm(x, g, b, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.layernorm(x, g, b, out=y)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"rows {rows:7d}: {us:8.1f} us  {rows * 1024 * 4 / us / 1e3:7.1f} GB/s")
