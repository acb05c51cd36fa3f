"""Layer stack on the GPU (op-by-op composition of the C-ABI kernels) against the CPU oracle and against
logits recorded from the unmodified reference.  bf16 activations/operands, fp32 accumulation:
tolerance 1e-2 relative (BASELINE.json north_star), measured as max|err| / max|ref| and as relative L2."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-2


def _run_stack(cfg, P, x, ctx_idx, tgt_idx, engine=True, hoist=True):
    from mebt_b200 import ops
    from mebt_b200.stack import WeightPack, stack_forward, stack_forward_ops
    dev = {k: v.cuda() for k, v in P.items()}
    modes = list(cfg["mode"]) + ["maskgit"] * (cfg["n_layer"] - len(cfg["mode"]))
    pack = WeightPack(dev, modes, cfg["n_head"])
    B = x.shape[0]
    xi = x.reshape(B, -1).cuda()
    ctx, tgt, lat = ops.embed_gather(xi, ctx_idx.cuda(), tgt_idx.cuda(), dev["tok_emb.weight"], dev["pos_emb"],
                                     dev["mask_emb"], dev["sos_emb"])
    logits = stack_forward(pack, B, lat, ctx, tgt, hoist=hoist) if engine else stack_forward_ops(pack, B, lat, ctx, tgt)
    ops.check_index_errors()
    return logits.view(B, tgt_idx.shape[1], -1).cpu()


def _check(logits, ref):
    assert torch.isfinite(logits).all()
    err = (logits - ref).abs().max().item() / ref.abs().max().item()
    l2 = ((logits - ref).norm() / ref.norm()).item()
    assert err < TOL and l2 < TOL, (err, l2)
    return err, l2


@pytest.mark.parametrize("name", ["micro", "tiny", "tiny5"])
def test_stack_forward_vs_oracle_and_golden(name):
    from oracle import mebt_oracle as O
    z, cfg = load_golden(f"forward_{name}")
    P = O.make_weights(cfg, int(z["wseed"]))
    x = torch.from_numpy(z["x"])
    indices = torch.from_numpy(z["indices"])
    for nc in z["ncs"]:
        nc = int(nc)
        ctx_idx, tgt_idx = indices[:, :nc], indices[:, nc:]
        ref = O.reconstruct_mask(P, cfg, x, ctx_idx, tgt_idx)
        logits = _run_stack(cfg, P, x, ctx_idx, tgt_idx)
        _check(logits, ref)
        if nc == int(z["ncs"][-1]):
            # the one-call C++ engine (without the enc K|V hoist) and the op-by-op composition launch the same
            # kernels: identical bits; the hoisted form folds ln1's gamma into bf16 weights: same tolerance
            unhoisted = _run_stack(cfg, P, x, ctx_idx, tgt_idx, hoist=False)
            assert torch.equal(unhoisted, _run_stack(cfg, P, x, ctx_idx, tgt_idx, engine=False))
            _check(unhoisted, ref)
        # the fixture holds what the unmodified reference produced
        sub = torch.from_numpy(z[f"nc{nc}_sub"])
        scale = float(np.abs(z[f"nc{nc}_rowmax"]).max())
        assert (logits[:, ::7, ::113] - sub).abs().max().item() < TOL * scale
        assert np.abs(torch.logsumexp(logits, -1).numpy() - z[f"nc{nc}_lse"]).max() < TOL * scale


def test_maskgit_padding_mode():
    """A short `mode` list is padded with full-attention 'maskgit' blocks (gpt.py:208-209)."""
    from oracle import mebt_oracle as O
    cfg = dict(n_embd=128, n_head=2, sos_emb=64, block_size=256, shape=[1, 16, 16], n_layer=3, vocab_size=16384,
               mode=["latent_enc", "latent_dec"])
    P = O.make_weights(cfg, 7)
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 16384, (2, 256), generator=g)
    perm = torch.stack([torch.randperm(256, generator=g) for _ in range(2)])
    ref = O.reconstruct_mask(P, cfg, x, perm[:, :100], perm[:, 100:])
    logits = _run_stack(cfg, P, x, perm[:, :100], perm[:, 100:])
    _check(logits, ref)


@pytest.mark.parametrize("name", ["micro", "tiny", "tiny5"])
def test_fp32_precision_mode_matches_reference_to_1e_4(name):
    """`precision = "fp32"` (bf16x3 split GEMM on the tensor cores + fp32 attention): logits within 1e-4 (north_star's
    fp32 tolerance) of the fixtures recorded from the unmodified reference and of the CPU oracle; BASELINE configs[0]."""
    from oracle import mebt_oracle as O
    from helpers import build_model
    z, cfg = load_golden(f"forward_{name}")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P)
    model.precision = "fp32"
    x, indices = torch.from_numpy(z["x"]), torch.from_numpy(z["indices"])
    for nc in (int(v) for v in z["ncs"]):
        logits, _ = model.reconstruct_mask(x.cuda(), indices[:, :nc].cuda(), indices[:, nc:].cuda())
        logits = logits.float().cpu()
        ref = O.reconstruct_mask(P, cfg, x, indices[:, :nc], indices[:, nc:])
        scale = float(ref.abs().max())
        assert float((logits - ref).abs().max()) < 1e-4 * scale, (nc, float((logits - ref).abs().max()), scale)
        sub = torch.from_numpy(z[f"nc{nc}_sub"])
        assert float((logits[:, ::7, ::113] - sub).abs().max()) < 1e-4 * scale
        assert float(np.abs(torch.logsumexp(logits, -1).numpy() - z[f"nc{nc}_lse"]).max()) < 1e-4 * scale
    # loss through the same path (shared_step: masked CE on fp32 logits)
    model.precision = "bf16"
    lb, _ = model.reconstruct_mask(x.cuda(), indices[:, :nc].cuda(), indices[:, nc:].cuda())
    assert float((lb.float().cpu() - ref).abs().max()) > 1e-4 * scale          # the bf16 engine is NOT this accurate


def test_fp32_precision_kernels():
    from mebt_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(300, 256, device="cuda", generator=g)
    w = torch.randn(512, 256, device="cuda", generator=g) * 0.05
    b = torch.randn(512, device="cuda", generator=g)
    r = torch.randn(300, 512, device="cuda", generator=g)
    ref = (a.double() @ w.double().T + b.double())
    ws = ops.split3(w, True)
    out = ops.gemm_f32(a, ws, b)
    assert float((out.double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
    out = ops.gemm_f32(a, ws, b, residual=r)
    assert float((out.double() - (ref + r.double())).abs().max()) < 2e-5 * float(ref.abs().max())
    out = ops.gemm_f32(a, ws, b, gelu=True)
    assert float((out.double() - torch.nn.functional.gelu(ref)).abs().max()) < 2e-5 * float(ref.abs().max())
    plain = ops.gemm(a.bfloat16(), w.bfloat16(), b, out_dtype=torch.float32)
    assert float((plain.double() - ref).abs().max()) > 1e-3 * float(ref.abs().max())     # what the split buys
    B, H, NQ, NK1, NK2 = 2, 3, 100, 77, 130
    D = H * 64
    q = torch.randn(B * NQ, 2 * D, device="cuda", generator=g)
    kv1 = torch.randn(B * NK1, 2 * D, device="cuda", generator=g)
    kv2 = torch.randn(B * NK2, 3 * D, device="cuda", generator=g)
    o = ops.attention_f32(q, D, kv1, 0, D, NK1, kv2, D, 2 * D, NK2, B, H, NQ)
    hv = lambda t, n: t.view(B, n, H, 64).transpose(1, 2).double()
    k = torch.cat([hv(kv1[:, :D], NK1), hv(kv2[:, D:2 * D], NK2)], 2)
    v = torch.cat([hv(kv1[:, D:], NK1), hv(kv2[:, 2 * D:], NK2)], 2)
    ref = torch.softmax(hv(q[:, D:], NQ) @ k.transpose(-1, -2) * 0.125, -1) @ v
    assert float((o.double() - ref.transpose(1, 2).reshape(B * NQ, D)).abs().max()) < 1e-5
    assert float(ops.attention_f32(q, D, None, 0, 0, 0, None, 0, 0, 0, B, H, NQ).abs().max()) == 0.0
