import sys, torch
sys.path.insert(0, ".")
from mebt_b200 import ops
for rows in (3072, 8192):
    lg = torch.randn(rows, 16384, device="cuda").bfloat16()
    tg = torch.randint(0, 16384, (rows,), device="cuda")
    for _ in range(3): ops.masked_ce(lg.clone(), tg, 0.0, dlogits=None)
    bufs = [lg.clone() for _ in range(8)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in bufs: ops.masked_ce(b, tg, 0.0, dlogits=b, grad_scale=1.0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 8 * 1e3
    print(f"rows {rows}: {us:.1f} us per call incl. the reduce kernel -> {rows*16384*4/us/1e6:.2f} TB/s")
