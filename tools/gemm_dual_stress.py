"""Stress loop for the two-issuer (DUAL) 128-wide GEMM: the training step's N = 1024 shapes in all operand-major forms,
thousands of launches each with a background copy stream perturbing the memory system; every result must equal the first
launch's bit for bit (fixed summation order) and agree with the single-issuer kernel's within bf16 rounding.
usage: python tools/gemm_dual_stress.py [iterations]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from mebt_b200 import ops  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    dev = "cuda"
    torch.manual_seed(0)
    side = torch.cuda.Stream()
    big_a, big_b = torch.empty(64 << 20, device=dev, dtype=torch.uint8), torch.empty(64 << 20, device=dev, dtype=torch.uint8)
    t0 = time.time()
    for (m, n, k, a_mn, b_mn) in ((1536, 1024, 1024, False, False), (1536, 1024, 4096, False, False), (1536, 1024, 3072, False, True),
                                  (1536, 1024, 4096, False, True), (1024, 1024, 1536, True, True), (3072, 128, 1024, False, False)):
        a = torch.randn((k, m) if a_mn else (m, k), device=dev).to(torch.bfloat16)
        b = torch.randn((k, n) if b_mn else (n, k), device=dev).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        ref = ops.gemm(a, b, bias, a_mn_major=a_mn, b_mn_major=b_mn, flags_extra=1024 | 32)
        torch.cuda.synchronize()
        exact = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t()) + bias
        err = (ref.float() - exact).abs().max().item() / exact.abs().max().item()
        assert err < 1e-2, (m, n, k, err)
        bad = 0
        for it in range(iters):
            if it % 7 == 0:
                with torch.cuda.stream(side):
                    big_b.copy_(big_a, non_blocking=True)
            out = ops.gemm(a, b, bias, a_mn_major=a_mn, b_mn_major=b_mn, flags_extra=1024 | 32)
            if it % 4 == 0:
                bad += int(not torch.equal(out, ref))
            if it % 100 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        print(f"{m} x {n} x {k} a_mn={int(a_mn)} b_mn={int(b_mn)}: err {err:.2e}, {bad} of {iters // 4 + 1} compared launches differ")
        assert bad == 0
    print(f"ok ({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
