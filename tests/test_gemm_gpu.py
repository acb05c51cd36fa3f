"""tcgen05 GEMM (K2/K4) against a torch fp32 matmul of the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GELU, OUT_FP32, ACC, BN256, BN128, BN64, NO_SPLITK, NO_PAIR, PAIR = 1, 2, 4, 16, 32, 64, 128, 256, 512
DUAL = 1024    # two MMA-issuing threads, an accumulator half each (128-wide tiles; opt-in)

CASES = [
    # M, N, K, a_mn, b_mn, flags, bias, resid
    (128, 64, 64, 0, 0, BN64, False, False),
    (128, 128, 64, 0, 0, BN128, False, False),
    (128, 256, 256, 0, 0, BN256, False, False),
    (256, 512, 1024, 0, 0, BN256, True, False),
    (1536, 1024, 1024, 0, 0, 0, True, True),
    (1536, 4096, 1024, 0, 0, GELU, True, False),
    (1536, 1024, 4096, 0, 0, 0, True, True),
    (1000, 768, 1024, 0, 0, 0, True, True),            # ragged M
    (3, 256, 256, 0, 0, 0, True, False),               # tiny M
    (3072, 16384, 1024, 0, 0, OUT_FP32, False, False),  # head
    (333, 256, 256, 0, 0, OUT_FP32, True, False),
    (128, 64, 64, 0, 1, BN64, False, False),
    (128, 64, 64, 1, 0, BN64, False, False),
    (128, 128, 128, 1, 1, BN128, False, False),
    (1536, 1024, 4096, 0, 1, 0, False, False),          # dgrad shape
    (1024, 4096, 1536, 1, 1, OUT_FP32, False, False),   # wgrad shape (K = tokens)
    (1024, 1024, 1000, 1, 1, OUT_FP32 | ACC, False, False),  # ragged K + accumulate
    (1536, 1024, 1024, 0, 1, BN256, True, True),
    (1536, 1024, 1024, 1, 0, BN128, True, False),
    # split-K territory (few tiles, long K) and the same shapes with the split disabled
    (256, 512, 4096, 0, 0, 0, True, True),
    (256, 512, 4096, 0, 0, NO_SPLITK, True, True),
    (1536, 1024, 4096, 0, 0, GELU, True, False),
    (1536, 1024, 4096, 0, 1, NO_SPLITK, False, False),
    (104, 256, 4000, 1, 1, OUT_FP32 | ACC, False, False),
    (1024, 64, 2048, 0, 0, 0, True, False),
    # 2-CTA cluster variant (B tile shared by TMA multicast): forced on small shapes, odd tile rows, both B majors
    (256, 256, 256, 0, 0, PAIR, False, False),
    (384, 512, 1024, 0, 0, PAIR | GELU, True, False),
    (1000, 768, 1024, 0, 0, PAIR | BN128, True, True),
    (1536, 1024, 4096, 0, 1, PAIR, False, False),
    (640, 256, 520, 1, 1, PAIR | OUT_FP32, False, False),
    (128, 256, 128, 0, 0, PAIR, True, False),
    (8192, 4096, 1024, 0, 0, 0, True, False),            # large enough to take the pair path on its own
    (8192, 4096, 1024, 0, 0, NO_PAIR, True, False),
    # two-issuer variant: all operand majors, odd / even k-block counts, several tiles per CTA (2304 > 148 tiles), residual
    (1536, 1024, 4096, 0, 0, DUAL, True, True),
    (1536, 1024, 1024, 0, 1, DUAL | BN128, True, False),
    (1024, 1024, 1536, 1, 1, DUAL | BN128 | OUT_FP32, False, False),
    (512, 256, 448, 1, 0, DUAL | BN128, True, False),
    (4608, 8192, 320, 0, 0, DUAL | BN128 | NO_PAIR, True, False),
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,flags,bias,resid", CASES)
def test_gemm(M, N, K, a_mn, b_mn, flags, bias, resid):
    from mebt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    bias_t = torch.randn(N, device="cuda", generator=g) if bias else None
    res_t = torch.randn(M, N, device="cuda", generator=g).bfloat16() if resid else None
    out_fp32 = bool(flags & OUT_FP32)
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)
    C0 = None
    if flags & ACC:
        C0 = torch.randn(M, N, device="cuda", generator=g)
        C.copy_(C0)
    ops.gemm(A_st, B_st, bias_t, res_t, gelu=bool(flags & GELU), out=C, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn),
             accumulate=bool(flags & ACC), flags_extra=flags & (BN256 | BN128 | BN64 | NO_SPLITK | NO_PAIR | PAIR | DUAL))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + bias_t
    if flags & GELU:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + res_t.float()
    if C0 is not None:
        ref = ref + C0
    assert not torch.isnan(C.float()).any()
    rel = (C.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    # bf16 output: half an ulp of bf16 (2^-9) relative to the row scale; fp32 output: accumulation order only
    assert rel < (1e-4 if out_fp32 else 6e-3), rel
    # bitwise reproducible, split-K included (partials are summed in split order, not arrival order)
    C2 = torch.full_like(C, float("nan"))
    if C0 is not None:
        C2.copy_(C0)
    ops.gemm(A_st, B_st, bias_t, res_t, gelu=bool(flags & GELU), out=C2, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn),
             accumulate=bool(flags & ACC), flags_extra=flags & (BN256 | BN128 | BN64 | NO_SPLITK | NO_PAIR | PAIR | DUAL))
    assert torch.equal(C, C2)


@pytest.mark.parametrize("rows_q,rows_k,D", [(1536, 3072, 1024), (3072, 1536, 1024), (200, 72, 256)])
def test_grouped_wgrad(rows_q, rows_k, D):
    """One launch for the five weight gradients of a Block (csrc/gemm_grouped.cu) against torch fp32 matmuls of the same
    bf16 operands; strided operands (dqkv inside a [rows, 3D] buffer), accumulation, ragged reductions, and the fused
    dropout epilogue's companion: bit-identical results to the five separate mebt_gemm_bf16 launches."""
    from mebt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows_q + rows_k)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    d_mlp, u = rnd(rows_q, D), rnd(rows_q, 4 * D)
    da, h = rnd(rows_q, 4 * D), rnd(rows_q, D)
    d_proj, att = rnd(rows_q, D), rnd(rows_q, D)
    dqkv_buf, qn = rnd(rows_q, 3 * D), rnd(rows_q, D)
    dq = dqkv_buf[:, :D]                                        # a column slice: row stride 3D
    dkv, kn = rnd(rows_k, 2 * D), rnd(rows_k, D)
    w_qkv = torch.full((3 * D, D), float("nan"), device="cuda")
    outs = [torch.full((D, 4 * D), float("nan"), device="cuda"), torch.full((4 * D, D), float("nan"), device="cuda"),
            torch.randn(D, D, device="cuda", generator=g), w_qkv[:D], w_qkv[D:]]
    prev_proj = outs[2].clone()
    problems = [(d_mlp, u, outs[0], False), (da, h, outs[1], False), (d_proj, att, outs[2], True), (dq, qn, outs[3], False),
                (dkv, kn, outs[4], False)]
    ops.grouped_wgrad(problems)
    torch.cuda.synchronize()
    for i, (dy, x, dw, acc) in enumerate(problems):
        ref = dy.float().t() @ x.float()
        if acc:
            ref = ref + prev_proj
        err = (dw - ref).abs().max().item() / ref.abs().max().item()
        assert err < 2e-3, (i, err)
    # same arithmetic as the stand-alone GEMM (fp32 accumulation in TMEM over the same k-blocks in the same order)
    single = ops.gemm(d_mlp, u, out_dtype=torch.float32, a_mn_major=True, b_mn_major=True, flags_extra=16 | 128)
    assert torch.equal(single, outs[0])


def test_gemm_dual_repeated_launches_are_bit_identical():
    """The two-issuer variant sums in a fixed order (even ring positions into one accumulator half, odd ones into the other,
    one final add): 600 launches of the training step's 1536 x 1024 x 4096 shape next to a copy stream all equal the first
    bit for bit, and agree with the single-issuer kernel within the accumulation-order rounding."""
    from mebt_b200 import ops
    torch.manual_seed(3)
    a = torch.randn(1536, 4096, device="cuda").to(torch.bfloat16)
    b = torch.randn(1024, 4096, device="cuda").to(torch.bfloat16)
    ref = ops.gemm(a, b, flags_extra=DUAL | BN128)
    single = ops.gemm(a, b, flags_extra=BN128)
    assert (ref.float() - single.float()).abs().max() <= 2 ** -6 * single.float().abs().max()
    side = torch.cuda.Stream()
    big = torch.empty(2, 32 << 20, device="cuda", dtype=torch.uint8)
    for it in range(600):
        if it % 7 == 0:
            with torch.cuda.stream(side):
                big[1].copy_(big[0], non_blocking=True)
        out = ops.gemm(a, b, flags_extra=DUAL | BN128)
        if it % 5 == 0:
            assert torch.equal(out, ref), f"launch {it} differs"
    torch.cuda.synchronize()
