"""The arithmetic behind `precision = "fp32"` (mebt_b200/csrc/precise.cu), replayed on the CPU: with hi = bf16(x) and
lo = bf16(x - hi), the three bf16 x bf16 products A_hi W_hi + A_lo W_hi + A_hi W_lo accumulated in fp32 reproduce the
fp32 GEMM to ~2^-16 per product - two orders of magnitude inside north_star's 1e-4 tolerance - while a plain bf16 GEMM is
two orders outside it.  (The kernel evaluates the sum as ONE bf16 GEMM over [hi | lo | hi] x [hi | hi | lo].)"""
import torch


def _split(x):
    hi = x.bfloat16().float()
    lo = (x - hi).bfloat16().float()
    return hi, lo


def test_bf16x3_split_reaches_fp32_accuracy():
    g = torch.Generator().manual_seed(0)
    for M, K, N, scale in ((64, 1024, 256, 0.02), (32, 4096, 128, 0.02), (48, 256, 512, 1.0)):
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) * scale
        ref = a.double() @ w.double().T
        a_hi, a_lo = _split(a)
        w_hi, w_lo = _split(w)
        assert float((a - a_hi - a_lo).abs().max()) <= 2.0 ** -16 * float(a.abs().max())     # what the split drops
        # the concatenated operands of the kernel: one GEMM with K' = 3K
        a3 = torch.cat([a_hi, a_lo, a_hi], 1)
        w3 = torch.cat([w_hi, w_hi, w_lo], 1)
        got = (a3 @ w3.T).double()
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        assert err < 2e-5, (M, K, N, err)
        plain = (a_hi @ w_hi.T).double()
        assert float((plain - ref).abs().max()) / float(ref.abs().max()) > 1e-3
