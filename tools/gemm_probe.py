"""Per-shape timing of mebt_gemm_bf16 (CUDA events, back-to-back launches, weights rotated through a pool larger
than L2 so every launch reads them cold, as inside the model).  Prints one line per (shape, tile width).
usage: python tools/gemm_probe.py [train|sample]"""
import sys

import torch

sys.path.insert(0, ".")
from mebt_b200 import _lib  # noqa: E402

BN = {256: 16, 128: 32, 64: 64, 0: 0, 'pair': 512, 'nopair': 256}


def bench(M, N, K, flag, reps=40):
    pool = max(2, int(200e6 // (N * K * 2)) + 1)
    pool = min(pool, 64)
    A = torch.randn(M, K, device="cuda").bfloat16()
    Ws = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(pool)]
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run(i):
        return _lib.lib.mebt_gemm_bf16(A.data_ptr(), K, 0, Ws[i % pool].data_ptr(), K, 0, C.data_ptr(), N, M, N, K,
                                       bias.data_ptr(), None, 0, flag, st)
    for i in range(5):
        if run(i) != 0:
            return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "train"
    if which == "train":      # STL-16f, B = 6: latent rows 1536, token rows 3072
        shapes = [(128, 256, 64), (128, 256, 1024), (1536, 1024, 1024), (1536, 3072, 1024), (1536, 4096, 1024), (1536, 1024, 4096), (3072, 2048, 1024),
                  (3072, 1024, 1024), (3072, 4096, 1024), (3072, 1024, 4096), (3072, 16384, 1024)]
    else:                     # 128f sampling, B = 4: latent rows 1024, token rows up to 32768
        shapes = [(1024, 1024, 1024), (1024, 3072, 1024), (1024, 4096, 1024), (1024, 1024, 4096), (16384, 2048, 1024),
                  (16384, 1024, 1024), (16384, 4096, 1024), (16384, 1024, 4096), (16384, 16384, 1024),
                  (32768, 4096, 1024)]
    for (M, N, K) in shapes:
        line = f"M={M:6d} N={N:6d} K={K:5d} |"
        for bn in (0, 'pair', 'nopair', 128):
            if isinstance(bn, int) and bn and N % bn:
                continue
            us = bench(M, N, K, BN[bn])
            if us is None:
                line += f" bn{bn}: n/a |"
            else:
                line += f" bn{bn or 'auto'}: {us:7.1f} us {2 * M * N * K / us / 1e6:7.1f} TF |"
        print(line, flush=True)


if __name__ == "__main__":
    main()
