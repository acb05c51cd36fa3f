"""Backward kernels against torch autograd (fp32) on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


@pytest.mark.parametrize("rows,N", [(3072, 1024), (1000, 4096), (7, 256), (20000, 3072)])
def test_colsum(rows, N):
    from mebt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(rows, N, device="cuda", generator=g).bfloat16()
    out = ops.colsum(x)
    ref = x.float().sum(0)
    assert (out - ref).abs().max() < 1e-3 * max(1.0, ref.abs().max().item())
    base = torch.randn(N, device="cuda", generator=g)
    acc = base.clone()
    ops.colsum(x, out=acc, accumulate=True)
    assert (acc - (base + ref)).abs().max() < 1e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(ops.colsum(x), out)          # deterministic


@pytest.mark.parametrize("rows,D", [(3072, 1024), (1000, 256), (5, 128), (40000, 1024)])
def test_layernorm_bwd(rows, D):
    from mebt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = (torch.randn(rows, D, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
    dy = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    w = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    b = 0.1 * torch.randn(D, device="cuda", generator=g)
    y, mean, rstd = ops.layernorm(x, w, b, save_stats=True)
    xr = x.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), wr, br, 1e-5).backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, mean, rstd, w)
    assert _rel(dx, xr.grad) < 1e-2
    assert _rel(dg, wr.grad) < 2e-3 and _rel(db, br.grad) < 2e-3
    # accumulate variants
    dx2 = dx.clone()
    dg2, db2 = dg.clone(), db.clone()
    ops.layernorm_bwd(dy, x, mean, rstd, w, dx=dx2, accumulate_dx=True, dgamma=dg2, dbeta=db2, accumulate_params=True)
    assert _rel(dx2, 2 * xr.grad) < 1e-2 and _rel(dg2, 2 * wr.grad) < 2e-3


def test_gemm_gelu_aux_and_dgelu():
    from mebt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    M, K, N = 1000, 256, 1024
    h = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W1 = (0.05 * torch.randn(N, K, device="cuda", generator=g)).bfloat16()
    b1 = 0.1 * torch.randn(N, device="cuda", generator=g)
    a = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    u = ops.gemm_aux(h, W1, a, bias=b1, gelu=True)
    a_ref = h.float() @ W1.float().t() + b1
    assert _rel(a, a_ref) < 6e-3 and _rel(u, F.gelu(a_ref)) < 6e-3
    # da = (du @ W2) * gelu'(a)   with W2 [D_out, N] stored [K_red = D_out, N] -> b_mn_major
    Dout = 256
    du_up = torch.randn(M, Dout, device="cuda", generator=g).bfloat16()
    W2 = (0.05 * torch.randn(Dout, N, device="cuda", generator=g)).bfloat16()
    da = ops.gemm_aux(du_up, W2, a, dgelu=True, b_mn_major=True)
    ar = a.float().requires_grad_(True)
    (F.gelu(ar) * (du_up.float() @ W2.float())).sum().backward()
    assert _rel(da, ar.grad) < 8e-3


def test_embed_backward():
    from mebt_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, NC, NT, L, D, V, N = 3, 300, 212, 64, 128, 1000, 512
    x = torch.randint(0, V, (B, N), generator=g)
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    ctx_idx, tgt_idx = perm[:, :NC], perm[:, NC:]
    tok = torch.randn(V, D, generator=g, requires_grad=True)
    pos = torch.randn(1, N, D, generator=g, requires_grad=True)
    mask = torch.randn(1, 1, D, generator=g, requires_grad=True)
    sos = torch.randn(1, L, D, generator=g, requires_grad=True)
    d_ctx = torch.randn(B * NC, D, generator=g).bfloat16()
    d_tgt = torch.randn(B * NT, D, generator=g).bfloat16()
    d_lat = torch.randn(B * L, D, generator=g).bfloat16()
    ctx = tok[torch.gather(x, 1, ctx_idx)] + pos[0][ctx_idx]
    tgt = mask.expand(B, NT, -1) + pos[0][tgt_idx]
    lat = sos.expand(B, -1, -1)
    ((ctx.reshape(-1, D) * d_ctx.float()).sum() + (tgt.reshape(-1, D) * d_tgt.float()).sum()
     + (lat.reshape(-1, D) * d_lat.float()).sum()).backward()
    gt, gp = torch.zeros(V, D, device="cuda"), torch.zeros(N, D, device="cuda")
    gm, gs = torch.zeros(D, device="cuda"), torch.zeros(L, D, device="cuda")
    pc = perm.cuda()
    ops.embed_backward(x.cuda(), pc[:, :NC], pc[:, NC:], d_ctx.cuda(), d_tgt.cuda(), d_lat.cuda(), gt, gp, gm, gs)
    assert (gt.cpu() - tok.grad).abs().max() < 1e-4 and (gp.cpu() - pos.grad[0]).abs().max() < 1e-4
    assert (gm.cpu() - mask.grad.view(-1)).abs().max() < 2e-3 and (gs.cpu() - sos.grad[0]).abs().max() < 1e-4


@pytest.mark.parametrize("B,H,NQ,NK1,NK2", [(2, 4, 256, 300, 0), (1, 2, 256, 256, 0), (2, 2, 200, 256, 0),
                                             (2, 4, 256, 256, 724), (2, 2, 256, 0, 0), (1, 2, 128, 0, 130),
                                             (1, 16, 1000, 256, 0)])
def test_latent_attention_bwd(B, H, NQ, NK1, NK2):
    from mebt_b200 import ops
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(9)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    qbuf = rnd(B * NQ, 3 * D)
    kv1 = rnd(B * NK1, 2 * D) if NK1 else None
    kv2 = rnd(B * NK2, 2 * D) if NK2 else None
    do = rnd(B * NQ, D)
    lse = torch.empty(B, H, NQ, device="cuda")
    out = ops.attention(qbuf, D, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ, lse=lse)
    dqbuf = torch.zeros_like(qbuf)
    dkv1 = torch.zeros_like(kv1) if NK1 else None
    dkv2 = torch.zeros_like(kv2) if NK2 else None
    ops.attention_bwd(qbuf, D, kv1, 0, D, NK1, kv2, 0, D, NK2, out, do, lse, dqbuf, D, dkv1, 0, D, dkv2, 0, D, B, H, NQ)
    torch.cuda.synchronize()
    if NK1 + NK2 == 0:
        assert (dqbuf == 0).all()
        return
    q = qbuf[:, D:2 * D].float().view(B, NQ, H, 64).transpose(1, 2).requires_grad_(True)
    ks, vs = [], []
    for kv, nk in ((kv1, NK1), (kv2, NK2)):
        if nk:
            ks.append(kv[:, :D].float().view(B, nk, H, 64).transpose(1, 2))
            vs.append(kv[:, D:].float().view(B, nk, H, 64).transpose(1, 2))
    k = torch.cat(ks, 2).requires_grad_(True)
    v = torch.cat(vs, 2).requires_grad_(True)
    o = F.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v
    o.backward(do.float().view(B, NQ, H, 64).transpose(1, 2))
    dq_ref = q.grad.transpose(1, 2).reshape(B * NQ, D)
    assert _rel(dqbuf[:, D:2 * D], dq_ref) < 2e-2
    assert (dqbuf[:, :D] == 0).all() and (dqbuf[:, 2 * D:] == 0).all()      # only the Q slice is written
    off = 0
    for dkv, nk in ((dkv1, NK1), (dkv2, NK2)):
        if nk:
            dk_ref = k.grad[:, :, off:off + nk].transpose(1, 2).reshape(B * nk, D)
            dv_ref = v.grad[:, :, off:off + nk].transpose(1, 2).reshape(B * nk, D)
            assert _rel(dkv[:, :D], dk_ref) < 2e-2 and _rel(dkv[:, D:], dv_ref) < 2e-2
            off += nk
