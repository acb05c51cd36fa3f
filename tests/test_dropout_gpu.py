"""Training-mode dropout (nn.Dropout at mebt/modules/gpt.py:112-113,136,140,154,216,239-241) on the CUDA path.

The CUDA kernels draw their keep decisions from a counter-based hash of (seed, site, row, column) and regenerate them
in backward; torch's Philox stream cannot be reproduced, so parity is shown by REPLAY: the masks the kernels applied
are materialised (mebt_dropout_rows on a tensor of ones / mebt_attention_dropout_mask) and fed to the CPU oracle, whose
dropout placement is itself pinned against the reference (tests/test_oracle_golden.py::test_training_mode_dropout_*)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from helpers import model_configs

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12)).item()


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_rows_kernel(p):
    from mebt_b200 import ops
    rows, D = 777, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    res = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    keep = ops.dropout_rows_(torch.ones_like(x), p, 1234, 7).float()           # the factors: 0 or 1/(1-p_q)
    thr = round(p * 65536)
    inv = 65536.0 / (65536 - thr)
    vals = keep.unique()
    assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - inv) < 4e-3 * inv     # bf16(1/(1-p))
    frac = float((keep == 0).float().mean())
    sigma = (p * (1 - p) / keep.numel()) ** 0.5
    assert abs(frac - thr / 65536.0) < 5 * sigma, (frac, p)
    # columns and rows are decorrelated: per-row and per-column drop rates are all close to p
    assert float(((keep == 0).float().mean(0) - p).abs().max()) < 6 * (p * (1 - p) / rows) ** 0.5
    assert float(((keep == 0).float().mean(1) - p).abs().max()) < 6 * (p * (1 - p) / D) ** 0.5
    y = ops.dropout_rows_(x.clone(), p, 1234, 7)
    mask = keep != 0
    assert torch.equal(y[~mask], torch.zeros_like(y[~mask]))
    assert torch.equal(y[mask], (x.float() * inv).bfloat16()[mask])
    assert torch.equal(ops.dropout_rows_(x.clone(), p, 1234, 7), y)                         # deterministic
    assert not torch.equal(ops.dropout_rows_(x.clone(), p, 1234, 8), y)                     # other site, other mask
    assert not torch.equal(ops.dropout_rows_(x.clone(), p, 1235, 7), y)                     # other seed
    yr = ops.dropout_rows_(x.clone(), p, 1234, 7, resid=res)
    ref = torch.where(mask, torch.addcmul(res.float(), x.float(), torch.full_like(x.float(), inv)), res.float()).bfloat16()
    assert torch.equal(yr, ref)
    wide = torch.randn(rows, 3 * D, device="cuda", generator=g).bfloat16()                  # strided view
    got = ops.dropout_rows_(wide[:, D:2 * D], p, 1234, 7)
    assert torch.equal((got != 0) | (wide[:, D:2 * D] == 0), mask | (wide[:, D:2 * D] == 0))
    assert torch.equal(ops.dropout_rows_(x.clone(), 0.0, 1, 1), x)                          # p = 0 is the identity


@pytest.mark.parametrize("B,H,NQ,NK1,NK2,p", [(2, 4, 256, 300, 0, 0.1), (1, 2, 256, 256, 0, 0.5), (2, 2, 200, 256, 77, 0.1),
                                               (2, 4, 256, 256, 724, 0.1), (1, 2, 128, 0, 130, 0.25)])
def test_attention_dropout_fwd_bwd(B, H, NQ, NK1, NK2, p):
    """O = (softmax(S) .* keep/(1-p)) V and its gradients, with the mask the kernels regenerate in all three of them."""
    from mebt_b200 import ops
    D, seed = H * 64, 99
    g = torch.Generator(device="cuda").manual_seed(9)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    qbuf = rnd(B * NQ, 3 * D)
    kv1 = rnd(B * NK1, 2 * D) if NK1 else None
    kv2 = rnd(B * NK2, 2 * D) if NK2 else None
    do = rnd(B * NQ, D)
    lse = torch.empty(B, H, NQ, device="cuda")
    keep = ops.attention_dropout_mask(B, H, NQ, NK1, NK2, p, seed)
    frac = float((keep == 0).float().mean())
    assert abs(frac - p) < 6 * (p * (1 - p) / keep.numel()) ** 0.5 + 1e-5
    out = ops.attention(qbuf, D, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ, lse=lse, drop_p=p, drop_seed=seed)
    dqbuf = torch.zeros_like(qbuf)
    dkv1 = torch.zeros_like(kv1) if NK1 else None
    dkv2 = torch.zeros_like(kv2) if NK2 else None
    ops.attention_bwd(qbuf, D, kv1, 0, D, NK1, kv2, 0, D, NK2, out, do, lse, dqbuf, D, dkv1, 0, D, dkv2, 0, D, B, H, NQ,
                      drop_p=p, drop_seed=seed)
    torch.cuda.synchronize()
    q = qbuf[:, D:2 * D].float().view(B, NQ, H, 64).transpose(1, 2).requires_grad_(True)
    ks, vs = [], []
    for kv, nk in ((kv1, NK1), (kv2, NK2)):
        if nk:
            ks.append(kv[:, :D].float().view(B, nk, H, 64).transpose(1, 2))
            vs.append(kv[:, D:].float().view(B, nk, H, 64).transpose(1, 2))
    k = torch.cat(ks, 2).requires_grad_(True)
    v = torch.cat(vs, 2).requires_grad_(True)
    s = (q @ k.transpose(-1, -2)) * 0.125
    o = (F.softmax(s, -1) * keep) @ v
    o.backward(do.float().view(B, NQ, H, 64).transpose(1, 2))
    assert _rel(out, o.detach().transpose(1, 2).reshape(B * NQ, D)) < 1.5e-2
    assert (lse - torch.logsumexp(s.detach(), -1)).abs().max() < 1e-3               # the row sum ignores the mask
    assert _rel(dqbuf[:, D:2 * D], q.grad.transpose(1, 2).reshape(B * NQ, D)) < 2e-2
    off = 0
    for dkv, nk in ((dkv1, NK1), (dkv2, NK2)):
        if nk:
            dk_ref = k.grad[:, :, off:off + nk].transpose(1, 2).reshape(B * nk, D)
            dv_ref = v.grad[:, :, off:off + nk].transpose(1, 2).reshape(B * nk, D)
            assert _rel(dkv[:, :D], dk_ref) < 2e-2 and _rel(dkv[:, D:], dv_ref) < 2e-2
            off += nk
    # without the mask the same kernels give a clearly different answer (the mask is really applied)
    plain = ops.attention(qbuf, D, kv1, 0, D, NK1, kv2, 0, D, NK2, B, H, NQ)
    assert _rel(plain, out) > 0.1


def _dropout_state(p):
    from mebt_b200.training import TrainState
    from mebt_b200.transformer import Net2NetTransformer
    from oracle import mebt_oracle as O
    z, cfg = load_golden("grads_dropout_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    params, vq, mask = model_configs(cfg, "linear")
    params.embd_pdrop = params.resid_pdrop = params.attn_pdrop = p
    model = Net2NetTransformer(params, vq, mask)
    model.load_state_dict(P, strict=True)
    model = model.cuda().eval()
    model.transformer.train()                  # dropout follows the GPT module's training flag, as nn.Dropout does
    ts = TrainState(model, n_buckets=2)
    return z, cfg, P, model, ts


def _kernel_masks(ts, cfg, B, NC, NT, seed):
    """The keep factors the CUDA step applies for `seed`, keyed for oracle.gpt_forward(drop=...)."""
    from mebt_b200 import ops
    D, H, L = cfg["n_embd"], cfg["n_head"], cfg["sos_emb"]
    ones = lambda rows: torch.ones(rows, D, device="cuda", dtype=torch.bfloat16)
    drop = {}
    for name, n in (("lat", L), ("ctx", NC), ("tgt", NT)):
        if ts.embd_pdrop > 0:
            drop[("stem", name)] = ops.dropout_rows_(ones(B * n), ts.embd_pdrop, seed, ts.STEM_SITES[name]).float().view(B, n, D).cpu()
    for i, mode in enumerate(cfg["mode"]):
        nq = NT if mode == "latent_dec" else L
        nk1, nk2 = {"latent_enc": (NC, 0), "latent_self": (L, 0), "latent_dec": (L, 0), "lt2l": (L, NT)}[mode]
        if ts.attn_pdrop > 0:
            drop[(i, "attn")] = ops.attention_dropout_mask(B, H, nq, nk1, nk2, ts.attn_pdrop, seed + 4 * i).cpu()
        if ts.resid_pdrop > 0:
            drop[(i, "proj")] = ops.dropout_rows_(ones(B * nq), ts.resid_pdrop, seed, 4 * i + 1).float().view(B, nq, D).cpu()
            drop[(i, "mlp")] = ops.dropout_rows_(ones(B * nq), ts.resid_pdrop, seed, 4 * i + 2).float().view(B, nq, D).cpu()
    return drop


def test_training_step_with_dropout_matches_oracle_replay():
    """Full training step at p = 0.1 (the STL yaml's value): loss and all parameter gradients against torch autograd
    through the oracle with the kernels' own masks replayed."""
    from oracle import mebt_oracle as O
    p, seed, t = 0.1, 20261017, 0.4
    z, cfg, P, model, ts = _dropout_state(p)
    ts.dropout_seed = seed
    x, indices = torch.from_numpy(z["x"]), torch.from_numpy(z["indices"])
    out = ts.loss_and_backward(x.cuda(), indices.cuda(), t=t)
    torch.cuda.synchronize()
    B, NC, NT = ts._ctx[0], ts._ctx[1], ts._ctx[2]
    drop = _kernel_masks(ts, cfg, B, NC, NT, seed)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    r = O.shared_step(Pg, cfg, x, indices, t, "linear", drop=drop)
    r["loss"].backward()
    assert abs(float(out["loss"]) - float(r["loss"])) < 3e-3 * float(r["loss"])
    worst = 0.0
    for n, prm in model.named_parameters():
        ref = Pg[n].grad
        if ref is None or float(ref.norm()) < 1e-7:
            continue
        err = _rel(prm.grad, ref)
        worst = max(worst, err)
        assert err < 7e-2, (n, err)
    # the no-dropout gradients are far from these: the masks matter and were the right ones
    r0 = O.shared_step({k: v.clone().requires_grad_(True) for k, v in P.items()}, cfg, x, indices, t, "linear")
    assert float((r0["logits"] - r["logits"]).abs().max()) > 1e-2
    # same seed -> bitwise identical step; eval mode -> dropout off
    g1 = ts.flat_grad.clone()
    ts.loss_and_backward(x.cuda(), indices.cuda(), t=t)
    lo = ts.emb_slice[0]
    assert torch.equal(g1[:lo], ts.flat_grad[:lo])
    ts.dropout_seed = seed + 1
    ts.loss_and_backward(x.cuda(), indices.cuda(), t=t)
    assert not torch.equal(g1[:lo], ts.flat_grad[:lo])
    model.transformer.eval()
    assert ts._dropout()[0] is None


def test_training_with_dropout_learns():
    z, cfg, P, model, ts = _dropout_state(0.1)
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    torch.manual_seed(0)
    opt = ts.make_optimizer(lr=3e-3, weight_decay=0.0)
    losses = [float(ts.train_step(opt, x, indices, t=0.5)["loss"]) for _ in range(10)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0] - 0.5, losses
