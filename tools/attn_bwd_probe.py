"""Timing of the attention backward (and the dropout forward) at the 16-frame training shapes, B = 6, H = 16 (CUDA events,
20 back-to-back launches; the backward call includes the 6 us delta pre-pass that the training engine gets from a GEMM
epilogue instead)."""
import sys
import torch
sys.path.insert(0, ".")
from mebt_b200 import ops  # noqa: E402

bf = torch.bfloat16


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def bench(name, B, H, NQ, NK1, NK2, p):
    D = H * 64
    r = lambda *s: torch.randn(*s, device="cuda").to(bf)  # noqa: E731
    q, kv1 = r(B * NQ, D), r(B * NK1, 2 * D)
    kv2 = r(B * NK2, 2 * D) if NK2 else None
    lse = torch.empty(B, H, NQ, device="cuda")
    o = ops.attention(q, 0, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ, lse=lse, drop_p=p, drop_seed=5)
    do = r(B * NQ, D)
    dq, dkv1 = torch.empty_like(q), torch.empty_like(kv1)
    dkv2 = torch.empty_like(kv2) if NK2 else None
    fwd = timed(lambda: ops.attention(q, 0, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ, out=o, lse=lse, drop_p=p, drop_seed=5))
    bwd = timed(lambda: ops.attention_bwd(q, 0, kv1, 0, D, NK1, kv2, 0, D, NK2, o, do, lse, dq, 0, dkv1, 0, D, dkv2, 0, D,
                                          B, H, NQ, drop_p=p, drop_seed=5))
    fl = 2.0 * B * H * NQ * (NK1 + NK2) * 64
    print(f"{name:12s} NQ={NQ} NK={NK1}+{NK2} p={p}: fwd {fwd:6.1f} us ({2 * fl / fwd / 1e6:6.1f} TF)  "
          f"bwd {bwd:6.1f} us ({5 * fl / bwd / 1e6:6.1f} TF algorithmic)", flush=True)


for p in (0.0, 0.1):
    bench("latent_enc", 6, 16, 256, 512, 0, p)
    bench("latent_self", 6, 16, 256, 256, 0, p)
    bench("latent_dec", 6, 16, 512, 256, 0, p)
    bench("lt2l", 6, 16, 256, 256, 512, p)
