"""Timing of mebt_latent_attention_fwd at the 128-frame sampling shapes (CUDA events, 20 back-to-back launches)."""
import sys
import torch
sys.path.insert(0, ".")
from mebt_b200 import ops  # noqa: E402

bf = torch.bfloat16


def bench(name, B, H, NQ, NK, reps=20):
    D = H * 64
    q = torch.randn(B * NQ, D, device="cuda").to(bf)
    kv = torch.randn(B * NK, 2 * D, device="cuda").to(bf)
    for _ in range(3):
        ops.attention(q, 0, kv, 0, D, NK, None, 0, 0, 0, B, H, NQ)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.attention(q, 0, kv, 0, D, NK, None, 0, 0, 0, B, H, NQ)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    exps = B * H * NQ * NK
    print(f"{name:12s} B={B} NQ={NQ} NK={NK}: {us:8.1f} us  {4 * exps * 64 / us / 1e6:7.1f} TF  {exps / us / 1e3:6.2f} Gexp/ms "
          f"(MUFU bound {exps / (148 * 16 * 1.9e3):.1f} us at 1.9 GHz)", flush=True)


for B in (16, 6):
    bench("latent_enc", B, 16, 256, 8192 if B == 16 else 512)
    bench("latent_self", B, 16, 256, 256)
    bench("latent_dec", B, 16, 8192 if B == 16 else 512, 256)
