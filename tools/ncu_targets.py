"""Launch each hot kernel once or twice at a BASELINE-sized shape so that `ncu --set full` can capture it cheaply:

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels python tools/ncu_targets.py

(about 25 kernel launches in total; the warm-up launches are also captured and can be ignored)."""
import sys

import torch

sys.path.insert(0, ".")
from mebt_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
bf = torch.bfloat16


def r(*s, dtype=bf):
    return torch.randn(*s, device=dev).to(dtype)


def main():
    # K2/K4 large GEMMs (128f sampling, 16 videos: token rows up to 131072; use 16384 rows)
    a = r(8192, 1024)
    for (n, k) in ((4096, 1024), (1024, 4096)):
        w = r(n, k)
        x = r(8192, k)
        for _ in range(1):
            ops.gemm(x, w, torch.zeros(n, device=dev))
    head = r(16384, 1024)
    for _ in range(1):
        logits = ops.gemm(a, head, out_dtype=torch.float32)                  # head GEMM, fp32 logits
    # K6 fused into the head GEMM (Gumbel-max epilogue, no logits written)
    ops.head_sample(a, head, 1.0, seed=1, offset=1)
    # K6 sampling (fast mode) and K5 masked CE over the materialised logits
    for _ in range(1):
        ops.sample_logits(logits, 1.0, None, None, noise=None, seed=1, offset=1)
    tg = torch.randint(0, 16384, (8192,), device=dev)
    lb = logits.to(bf)
    for _ in range(1):
        ops.masked_ce(lb, tg, 0.0, dlogits=lb, grad_scale=1.0)
    # LayerNorm and the stem gather at token-row scale
    g, b = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
    xs = r(32768, 1024)
    for _ in range(1):
        ops.layernorm(xs, g, b)
    B, N = 8, 8192
    xi = torch.randint(0, 16384, (B, N), device=dev)
    perm = torch.stack([torch.randperm(N, device=dev) for _ in range(B)])
    tok, pos = torch.randn(16384, 1024, device=dev), torch.randn(1, N, 1024, device=dev)
    for _ in range(1):
        ops.embed_gather(xi, perm[:, :4096], perm[:, 4096:], tok, pos, torch.randn(1, 1, 1024, device=dev),
                         torch.randn(1, 256, 1024, device=dev))
    # K3 attention: latent_enc shape (256 latents x 7168 contexts) and latent_dec shape (8192 targets x 256 latents)
    Bq, H, D = 4, 16, 1024
    q = r(Bq * 256, D)
    kv = r(Bq * 7168, 2 * D)
    for _ in range(1):
        ops.attention(q, 0, kv, 0, D, 7168, None, 0, 0, 0, Bq, H, 256)
    q2 = r(Bq * 8192, D)
    kv2 = r(Bq * 256, 2 * D)
    for _ in range(1):
        ops.attention(q2, 0, kv2, 0, D, 256, None, 0, 0, 0, Bq, H, 8192)
    # enc at the full 128-frame context (256 x 8192) and lt2l (256 latents x [256 latents + 8192 targets])
    kv3 = r(Bq * 8192, 2 * D)
    qkv3 = r(Bq * 256, 3 * D)
    for _ in range(1):
        ops.attention(q, 0, kv3, 0, D, 8192, None, 0, 0, 0, Bq, H, 256)
        ops.attention(qkv3, 0, qkv3, D, 2 * D, 256, kv3, 0, D, 8192, Bq, H, 256)
    # split-KV form (batch 2 at 8192 keys: 32 work items split 4 ways + the merge kernel)
    q2b, kv2b = r(2 * 256, D), r(2 * 8192, 2 * D)
    ops.attention(q2b, 0, kv2b, 0, D, 8192, None, 0, 0, 0, 2, H, 256)
    # training-step kernels at the 16-frame shapes (B = 6: 1536 latent rows, 3072 token rows)
    rows = 3072
    xs2, dy2 = r(rows, 1024), r(rows, 1024)
    _, mean, rstd = ops.layernorm(xs2, g, b, save_stats=True)
    for _ in range(1):
        ops.layernorm_bwd(dy2, xs2, mean, rstd, g)
        ops.colsum(r(rows, 4096))
        ops.dropout_rows_(dy2, 0.1, 1, 1)
    qb, kvb, dob = r(6 * 256, 3 * D), r(6 * 512, 2 * D), r(6 * 256, D)
    lse = torch.empty(6, H, 256, device=dev)
    for _ in range(1):
        ob = ops.attention(qb, D, kvb, 0, D, 512, None, 0, 0, 0, 6, H, 256, lse=lse, drop_p=0.1, drop_seed=3)
        ops.attention_bwd(qb, D, kvb, 0, D, 512, None, 0, 0, 0, ob, dob, lse, torch.zeros_like(qb), D, torch.zeros_like(kvb),
                          0, D, None, 0, 0, 6, H, 256, drop_p=0.1, drop_seed=3)
    # the training step's GEMM shapes (B = 6): 128-wide DUAL tiles (1536 x 1024 x 1024 / x 4096), pair-mode 256-wide tiles
    for (m, n, k) in ((1536, 1024, 1024), (1536, 1024, 4096), (1536, 4096, 1024), (3072, 1024, 4096), (1536, 3072, 1024)):
        ops.gemm(r(m, k), r(n, k), torch.zeros(n, device=dev))
    # masked CE at the training shape (3072 rows: the streaming kernel) and the grouped weight-gradient launch of a block
    lt = r(3072, 16384)
    ops.masked_ce(lt, torch.randint(0, 16384, (3072,), device=dev), 0.0, dlogits=lt, grad_scale=1.0)
    from mebt_b200 import _lib
    if hasattr(ops, "grouped_wgrad"):
        R = 1536
        probs = [(r(R, 1024), r(R, 4096), torch.empty(1024, 4096, device=dev), False), (r(R, 4096), r(R, 1024), torch.empty(4096, 1024, device=dev), False),
                 (r(R, 1024), r(R, 1024), torch.empty(1024, 1024, device=dev), False), (r(R, 3072), r(R, 1024), torch.empty(3072, 1024, device=dev), False)]
        ops.grouped_wgrad(probs)
    # K9/K10 codebook: tensor-core search (default) and the fp32 FFMA kernel
    E = torch.randn(16384, 256, device=dev)
    z = torch.randn(64, 256, 4, 16, 16, device=dev)
    for _ in range(1):
        enc = ops.vq_argmin(z, E)
        ops.vq_argmin(z[:8], E, tensor_cores=False)
    ops.row_gather(enc, E, channel_first=True)
    torch.cuda.synchronize()
    print("done")




def vqgan_targets():
    """VQGAN convolution kernels at the 16-frame shapes (n_hiddens 32, downsample 4 8 8): full-resolution 64-channel
    ResBlock convolution, the 256-channel bottleneck convolution, a strided down convolution, one parity class of an up
    convolution, and the GroupNorm + SiLU + padding pass."""
    import torch
    from mebt_b200.vqgan import SamePadConv3d, SamePadConvTranspose3d, to_channels_last, norm_args
    torch.manual_seed(0)
    dev = "cuda"
    for cin, cout, k, stride, dims in ((64, 64, 3, 1, (16, 128, 128)), (256, 256, 3, 1, (4, 16, 16)), (64, 128, 4, 2, (8, 64, 64))):
        m = SamePadConv3d(cin, cout, k, stride=stride).to(dev)
        x = to_channels_last(torch.randn(2, cin, *dims, device=dev))
        gn = torch.nn.GroupNorm(32, cin, eps=1e-6).to(dev)
        m.forward_cl(x, pre=norm_args(gn), act=1)
    up = SamePadConvTranspose3d(256, 128, 4, stride=(1, 2, 2)).to(dev)
    up.forward_cl(to_channels_last(torch.randn(2, 256, 4, 16, 16, device=dev)))
    torch.cuda.synchronize()


if __name__ == "__main__":
    if "--vqgan" in sys.argv:
        vqgan_targets()
    else:
        main()
