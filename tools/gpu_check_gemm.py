"""GPU bring-up check for mebt_gemm_bf16: every operand-major combination, tile width and epilogue
against a torch fp32 matmul of the same bf16-rounded operands.  Prints one line per case and never
stops at the first failure, so a single gpurun call yields a full picture.

    python tools/gpu_check_gemm.py [--quick]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from mebt_b200 import _lib  # noqa: E402

GELU, OUT_FP32, ACC, BN256, BN128, BN64 = 1, 2, 4, 16, 32, 64


def run_case(M, N, K, a_mn, b_mn, flags, bias, resid, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
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
    stream = torch.cuda.current_stream().cuda_stream
    rc = _lib.lib.mebt_gemm_bf16(A_st.data_ptr(), A_st.stride(0), int(a_mn), B_st.data_ptr(), B_st.stride(0), int(b_mn),
                                 C.data_ptr(), C.stride(0), M, N, K,
                                 bias_t.data_ptr() if bias else None, res_t.data_ptr() if resid else None,
                                 N if resid else 0, flags, stream)
    if rc != 0:
        return f"rc={rc} {_lib.last_error()}", float("inf")
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
    err = (C.float() - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    rel = err.max().item() / denom
    nan = torch.isnan(C.float()).sum().item()
    return f"max_abs={err.max().item():.4e} rel_to_max={rel:.3e} nan={nan}", rel if nan == 0 else float("inf")


def main():
    quick = "--quick" in sys.argv
    print(_lib.version(), "device_check:", _lib.lib.mebt_device_check(), torch.cuda.get_device_name(0), flush=True)
    cases = []
    # (M, N, K, a_mn, b_mn, flags, bias, resid)
    cases.append((128, 64, 64, 0, 0, BN64, False, False))
    cases.append((128, 128, 64, 0, 0, BN128, False, False))
    cases.append((128, 256, 64, 0, 0, BN256, False, False))
    cases.append((128, 256, 256, 0, 0, BN256, False, False))
    cases.append((256, 512, 1024, 0, 0, BN256, True, False))
    cases.append((1536, 1024, 1024, 0, 0, 0, True, True))
    cases.append((1536, 4096, 1024, 0, 0, 0, True, False))
    cases.append((1536, 4096, 1024, 0, 0, GELU, True, False))
    cases.append((1536, 1024, 4096, 0, 0, 0, True, True))
    cases.append((1000, 768, 1024, 0, 0, 0, True, True))          # ragged M
    cases.append((3072, 16384, 1024, 0, 0, OUT_FP32, False, False))  # head
    cases.append((333, 256, 256, 0, 0, OUT_FP32, True, False))
    # MN-major operands
    cases.append((128, 64, 64, 0, 1, BN64, False, False))
    cases.append((128, 64, 64, 1, 0, BN64, False, False))
    cases.append((128, 128, 128, 1, 1, BN128, False, False))
    cases.append((1536, 1024, 4096, 0, 1, 0, False, False))       # dgrad shape
    cases.append((1024, 4096, 1536, 1, 1, OUT_FP32, False, False))  # wgrad shape (K = tokens)
    cases.append((1024, 1024, 1000, 1, 1, OUT_FP32 | ACC, False, False))  # ragged K + accumulate
    cases.append((1536, 1024, 1024, 0, 1, BN256, True, True))
    cases.append((1536, 1024, 1024, 1, 0, BN128, True, False))
    if quick:
        cases = cases[:6]
    worst = 0.0
    nfail = 0
    for c in cases:
        t0 = time.time()
        try:
            msg, rel = run_case(*c)
        except Exception as e:  # noqa: BLE001
            msg, rel = f"EXC {type(e).__name__}: {e}", float("inf")
        ok = rel < 2e-2 if not (c[5] & OUT_FP32) else rel < 1e-3
        nfail += 0 if ok else 1
        worst = max(worst, rel if rel != float("inf") else 0)
        print(f"{'OK  ' if ok else 'FAIL'} M={c[0]} N={c[1]} K={c[2]} a_mn={c[3]} b_mn={c[4]} flags={c[5]} "
              f"bias={c[6]} res={c[7]} :: {msg} ({(time.time() - t0) * 1e3:.0f} ms)", flush=True)
        if "EXC" in msg or "rc=4" in msg:
            print("CUDA context likely poisoned; stopping", flush=True)
            break
    print(f"gemm check: {len(cases)} cases, {nfail} failed, worst rel {worst:.3e}", flush=True)

    if nfail == 0:
        # quick throughput probe (burst), CUDA events
        for (M, N, K) in [(1536, 4096, 1024), (1536, 1024, 4096), (8192, 4096, 1024), (8192, 16384, 1024)]:
            A = torch.randn(M, K, device="cuda").bfloat16()
            B = torch.randn(N, K, device="cuda").bfloat16()
            C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            st = torch.cuda.current_stream().cuda_stream
            args = (A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, M, N, K, None, None, 0, 0, st)
            for _ in range(5):
                _lib.lib.mebt_gemm_bf16(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                _lib.lib.mebt_gemm_bf16(*args)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"perf M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
            e0.record()
            for _ in range(20):
                torch.matmul(A, B.t())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"     cublas: {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
