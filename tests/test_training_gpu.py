"""Training step (forward with saved activations, fused CE, full backward) against gradients recorded from the
unmodified reference (tests/golden/grads_*.npz) and against torch autograd through the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import build_model

pytestmark = pytest.mark.gpu


def _state(name):
    from mebt_b200.training import TrainState
    from oracle import mebt_oracle as O
    z, cfg = load_golden(f"grads_{name}")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P)
    return z, cfg, P, model, TrainState(model, n_buckets=2)


@pytest.mark.parametrize("name", ["micro", "tiny5"])
def test_gradients_vs_reference_fixture(name):
    z, cfg, P, model, ts = _state(name)
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    out = ts.loss_and_backward(x, indices, t=float(z["t"]))
    torch.cuda.synchronize()
    assert abs(float(out["loss"]) - float(z["loss"])) < 3e-3 * float(z["loss"])
    names = [str(n) for n in z["grad_names"]]
    norms = z["grad_norms"]
    grads = {n: p.grad for n, p in model.named_parameters()}
    worst = 0.0
    for n, ref_norm in zip(names, norms):
        g = grads[n]
        assert g is not None and torch.isfinite(g).all(), n
        if ref_norm < 0:                                   # reference: grad is None (block cannot reach the logits)
            assert float(g.abs().max()) == 0.0, n
            continue
        got = float(g.norm())
        if ref_norm < 1e-7:            # mathematically zero (e.g. key bias: softmax is shift-invariant), fp noise only
            assert got < 1e-6, (n, got)
            continue
        rel = abs(got - ref_norm) / (ref_norm + 1e-12)
        worst = max(worst, rel)
        assert rel < 5e-2, (n, got, float(ref_norm))
    for key in z.files:
        if key.startswith("g:"):
            n = key[2:]
            ref = torch.from_numpy(z[key])
            got = grads[n].reshape(-1)[::17].cpu()
            if float(ref.norm()) < 1e-7:
                continue
            cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
            err = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
            assert cos > 0.998 and err < 6e-2, (n, cos, err)


def test_gradients_vs_oracle_autograd_nc0():
    """NC = 0 (t = 0 with the linear schedule): latent_enc key/value projections and tok_emb get exact-zero grads."""
    from oracle import mebt_oracle as O
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]), torch.from_numpy(z["indices"])
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    r = O.shared_step(Pg, cfg, x, indices, 0.0, "linear")
    r["loss"].backward()
    out = ts.loss_and_backward(x.cuda(), indices.cuda(), t=0.0)
    assert abs(float(out["loss"]) - float(r["loss"])) < 3e-3 * float(r["loss"])
    for n, p in model.named_parameters():
        ref = Pg[n].grad
        if ref is None:
            assert float(p.grad.abs().max()) == 0.0, n
            continue
        if float(ref.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, n          # exact zeros stay exact zeros
            continue
        if float(ref.norm()) < 1e-7:                             # key bias: zero up to rounding noise
            assert float(p.grad.norm()) < 1e-6, n
            continue
        err = ((p.grad.cpu() - ref).norm() / ref.norm()).item()
        assert err < 6e-2, (n, err)
    assert float(model.tok_emb.weight.grad.abs().max()) == 0.0


def test_train_step_decreases_loss_and_is_deterministic():
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    ts.loss_and_backward(x, indices, t=0.5)
    g1 = ts.flat_grad.clone()
    ts.loss_and_backward(x, indices, t=0.5)
    lo, hi = ts.emb_slice
    assert torch.equal(g1[:lo], ts.flat_grad[:lo])               # GEMM / LN / attention backward are bit-reproducible
    assert torch.allclose(g1[lo:hi], ts.flat_grad[lo:hi], atol=1e-5)   # embedding scatter-add uses fp32 atomics
    opt = ts.make_optimizer(lr=3e-3, weight_decay=0.0)
    losses = [float(ts.train_step(opt, x, indices, t=0.5)["loss"]) for _ in range(8)]
    assert losses[-1] < losses[0] - 0.5, losses
    # parameters stayed views of the flat buffer and the bf16 operands follow the masters
    w = model.transformer.blocks[0].mlp[0].weight
    assert w.data_ptr() == ts._view(ts.flat, "transformer.blocks.0.mlp.0.weight").data_ptr()
    assert torch.equal(ts._view(ts.flat_bf16, "transformer.blocks.0.mlp.0.weight"), w.detach().reshape(-1).bfloat16())
    # and the eval path sees the updated weights
    model.eval()
    logits, _ = model.reconstruct_mask(x, indices[:, :100], indices[:, 100:])
    assert torch.isfinite(logits).all()


def test_flat_adamw_matches_torch_fused_adamw():
    """mebt_adamw_flat against torch.optim.AdamW on the reference's parameter groups (configure_optimizers,
    transformer.py:749-798), same gradients for 5 steps; the bf16 operand copy follows the masters."""
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    flat_opt = ts.make_optimizer(lr=1e-3, weight_decay=0.05)
    decayed = {n for n, p in model.named_parameters() if any(p is q for g in flat_opt.param_groups if g["weight_decay"] > 0 for q in g["params"])}
    assert "transformer.blocks.0.mlp.0.weight" in decayed and "transformer.blocks.0.mlp.0.bias" not in decayed
    assert "pos_emb" not in decayed and "transformer.blocks.0.ln1.weight" not in decayed
    ref_params = {n: torch.nn.Parameter(p.detach().clone()) for n, p in model.named_parameters()}
    ref_opt = torch.optim.AdamW([{"params": [ref_params[n] for n in ref_params if n in decayed], "weight_decay": 0.05},
                                 {"params": [ref_params[n] for n in ref_params if n not in decayed], "weight_decay": 0.0}],
                                lr=1e-3, betas=(0.9, 0.95), fused=True)
    for step in range(5):
        ts.loss_and_backward(x, indices, t=0.5)
        for n, p in model.named_parameters():
            ref_params[n].grad = p.grad.detach().clone()
        flat_opt.step()
        ref_opt.step()
        for n, p in model.named_parameters():
            # while the masters agree the gradients are the same; tiny differences grow slowly once they do not
            assert torch.allclose(p.detach(), ref_params[n].detach(), rtol=2e-5 * (step + 1), atol=2e-7 * (step + 1)), (n, step)
    w = "transformer.blocks.1.attn.proj.weight"
    assert torch.equal(ts._view(ts.flat_bf16, w), ts._view(ts.flat, w).bfloat16())


def test_training_step_autograd_bridge_with_a_torch_optimizer():
    """The drop-in `training_step` (transformer.py:734-739) used the reference's way: loss.backward() + any optimizer over
    model.parameters(), zero_grad(set_to_none=True) included; gradients equal the direct TrainState path bit for bit."""
    import random
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    batch = dict(video=x, label=x, indices=indices)
    lo = ts.emb_slice[0]
    random.seed(1)
    loss = model.training_step(batch, 0)
    assert loss.requires_grad and loss.dim() == 0
    loss.backward()
    g = ts.flat_grad.clone()
    random.seed(1)
    out = ts.loss_and_backward(x, indices)
    assert float(out["loss"]) == float(loss) and torch.equal(g[:lo], ts.flat_grad[:lo])
    random.seed(1)
    model.zero_grad()                                           # consume the pending gradient: the next backward assigns
    (0.5 * model.training_step(batch, 0)).backward()            # a scaled loss scales the gradients
    assert torch.allclose(ts.flat_grad[:lo], 0.5 * g[:lo], rtol=1e-3, atol=1e-9)
    opt = torch.optim.AdamW(model.parameters(), lr=3e-3, betas=(0.9, 0.95), weight_decay=0.0)
    random.seed(0)
    losses = []
    for i in range(8):
        opt.zero_grad()                                         # set_to_none=True: the .grad views are re-linked by backward
        loss = model.training_step(batch, i)
        loss.backward()
        assert all(p.grad is not None for p in model.parameters())
        opt.step()                                              # in-place update of the fp32 masters; operands refresh lazily
        losses.append(float(loss))
    assert losses[-1] < losses[0] - 0.3, losses
    with torch.no_grad():
        val = model.validation_step(batch, 0)
    assert not val.requires_grad and torch.isfinite(val)


def test_overlapped_optimizer_update_matches_sequential():
    """train_step(overlap_update=True): per-chunk AdamW on the side stream gives the same parameters as backward followed
    by one whole-buffer step."""
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    opt = ts.make_optimizer(lr=1e-3, weight_decay=0.05)
    start = ts.flat.clone()
    for _ in range(3):
        ts.train_step(opt, x, indices, t=0.5, overlap_update=True)
    torch.cuda.synchronize()
    got = ts.flat.clone()
    ts.flat.copy_(start)
    ts.refresh_operands()
    opt2 = ts.make_optimizer(lr=1e-3, weight_decay=0.05)
    for _ in range(3):
        ts.train_step(opt2, x, indices, t=0.5)
    torch.cuda.synchronize()
    lo = ts.emb_slice[0]
    assert torch.equal(got[:lo], ts.flat[:lo])
    assert torch.allclose(got[lo:], ts.flat[lo:], rtol=1e-5, atol=1e-7)        # embedding grads: fp32 atomics


def test_eval_sees_flat_adamw_updates():
    """FlatAdamW writes the masters through raw pointers, which torch's version counters do not see;
    the inference operand caches (GPT.weight_pack, Block.layer_weights, CrossAttention._weights) must still follow."""
    from mebt_b200.stack import WeightPack
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    ctx, tgt = indices[:, :100], indices[:, 100:]
    model.eval()
    before, _ = model.reconstruct_mask(x, ctx, tgt)               # builds the packs
    lat = model.sos_emb.expand(x.shape[0], -1, -1)
    blk = model.transformer.blocks[1]
    blk_before = blk(lat, lat[:, :0], lat[:, :0])[0].clone()
    model.train()
    opt = ts.make_optimizer(lr=3e-3, weight_decay=0.0)
    for _ in range(4):
        ts.train_step(opt, x, indices, t=0.5)
    model.eval()
    after, _ = model.reconstruct_mask(x, ctx, tgt)
    assert float((after - before).abs().max()) > 1e-2            # the update is visible ...
    gpt = model.transformer
    params = {"transformer." + n: p for n, p in gpt.named_parameters()}
    fresh = WeightPack(params, [b.mode for b in gpt.blocks], gpt.config.n_head)
    gpt._pack = (gpt._pack[0], fresh)
    again, _ = model.reconstruct_mask(x, ctx, tgt)
    assert torch.equal(after, again)                              # ... and equals a pack built from scratch
    assert float((blk(lat, lat[:, :0], lat[:, :0])[0] - blk_before).abs().max()) > 1e-3   # module path too


def test_training_step_accumulates_like_autograd():
    """two loss.backward() calls without zero_grad add up (accumulate_grad_batches,
    train_transformer.py:47-50); zero_grad in either form makes the next backward assign."""
    import random
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    batch = dict(video=x, label=x, indices=indices)
    batch2 = dict(video=x.flip(0), label=x, indices=indices.flip(0))
    lo = ts.emb_slice[0]
    grads = []
    for b, seed in ((batch, 3), (batch2, 4)):
        model.zero_grad()
        random.seed(seed)
        model.training_step(b, 0).backward()
        grads.append(ts.flat_grad.clone())
    model.zero_grad()
    for b, seed in ((batch, 3), (batch2, 4)):
        random.seed(seed)
        model.training_step(b, 0).backward()
    want = grads[0] + grads[1]
    err = ((ts.flat_grad[:lo] - want[:lo]).norm() / want[:lo].norm()).item()
    assert err < 2e-3, err                                        # bf16 attention / LN paths re-round; fp32 accumulation
    for p in model.parameters():                                  # zero_grad(set_to_none=False): in-place zero
        p.grad.zero_()
    random.seed(3)
    model.training_step(batch, 0).backward()
    assert torch.allclose(ts.flat_grad[:lo], grads[0][:lo], rtol=1e-4, atol=1e-8)


def test_training_step_refuses_a_stale_backward():
    from mebt_b200._lib import MebtError
    z, cfg, P, model, ts = _state("micro")
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    batch = dict(video=x, label=x, indices=indices)
    first = model.training_step(batch, 0)
    second = model.training_step(batch, 1)
    with pytest.raises(MebtError, match="another forward"):
        first.backward()
    second.backward()


def test_maskgit_block_mode_training_is_refused():
    """The `maskgit` Block mode (gpt.py:176-178,191-192: one self-attending stream over cat[contexts, targets]) runs on the
    inference engine only; no shipped config of the reference uses it (configs/*/*.yaml list the four latent modes).
    Training such a stack must fail loudly and at construction, never fall back to another path."""
    from mebt_b200.training import TrainState
    cfg = dict(n_embd=128, n_head=2, sos_emb=16, block_size=64, shape=[1, 8, 8], n_layer=2, vocab_size=256, avg_loss=1.0,
               mode=["maskgit", "maskgit"])
    model = build_model(cfg).train()
    with pytest.raises(NotImplementedError, match="four latent block modes"):
        TrainState(model)


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_fused_wgrad_adamw_matches_backward_then_step(dropout):
    """train_step(fused_update=True): AdamW of the blocks' Linear weights inside the epilogue of their weight-gradient
    GEMMs (mebt_stack_backward_fused; lt2l's key|value rows as a two-segment reduction) gives the parameters, moments and
    bf16 operand copies of backward followed by one whole-buffer step (loss.backward(); optimizer.step() in the
    reference, mebt/transformer.py:665-681).  D = 256: the narrowest width the grouped kernel tiles."""
    from mebt_b200.training import TrainState
    from oracle import mebt_oracle as O
    cfg = dict(n_embd=256, n_head=4, sos_emb=64, block_size=256, shape=[1, 16, 16], n_layer=6, vocab_size=16384,
               mode=["latent_enc", "latent_self", "latent_dec", "lt2l", "latent_dec", "latent_self"], avg_loss=1.0,
               embd_pdrop=dropout, resid_pdrop=dropout, attn_pdrop=dropout)
    P = O.make_weights(cfg, 3)
    model = build_model(cfg, P)
    model.transformer.train()
    ts = TrainState(model, n_buckets=2)
    ts.dropout_seed = 11
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 16384, (3, 256), generator=g).cuda()
    indices = torch.stack([torch.randperm(256, generator=g) for _ in range(3)]).cuda()
    start = ts.flat.clone()

    def run(fused, steps):
        ts.flat.copy_(start)
        ts.refresh_operands()
        opt = ts.make_optimizer(lr=1e-3, weight_decay=0.05)
        losses = [float(ts.train_step(opt, x, indices, t=0.5, fused_update=fused)["loss"]) for _ in range(steps)]
        assert opt.steps == steps
        torch.cuda.synchronize()
        return losses, ts.flat.clone(), opt.m.clone(), opt.v.clone(), ts.flat_bf16.clone()

    # ONE step: weight gradients are leaves of the backward, so both paths see the same accumulators everywhere except
    # lt2l's key|value rows (block 3: one two-segment reduction instead of two launches added in fp32) and the
    # atomically accumulated embedding gradients: parameters, moments and operand copies are bit-identical elsewhere
    l1, p1, m1, v1, b1 = run(True, 1)
    l0, p0, m0, v0, b0 = run(False, 1)
    assert l1 == l0
    lo = ts.emb_slice[0]
    same = torch.ones(lo, dtype=torch.bool, device=p1.device)
    for suffix in ("attn.key.weight", "attn.value.weight"):
        o, k = ts.offsets[f"transformer.blocks.3.{suffix}"]
        same[o:o + k] = False
        assert (m1[o:o + k] - m0[o:o + k]).abs().max() <= 1e-4 * m0[o:o + k].abs().max()     # m = 0.1 g after one step
        assert (p1[o:o + k] - p0[o:o + k]).abs().max() <= 2.1e-3                              # Adam: at most a flipped +-lr
        assert ((p1[o:o + k] - p0[o:o + k]).abs() > 1e-6).float().mean() < 1e-2
    for a_, b_ in ((p1, p0), (m1, m0), (v1, v0), (b1, b0)):
        assert torch.equal(a_[:lo][same], b_[:lo][same]), float((a_[:lo][same].float() - b_[:lo][same].float()).abs().max())
    # three steps: the trajectories stay together
    l1, p1, _, _, _ = run(True, 3)
    l0, p0, _, _, _ = run(False, 3)
    assert np.allclose(l1, l0, rtol=1e-4), (l1, l0)
    # the block after the last latent_dec is dead: its weights still decay through the plain kernel
    o, k = ts.offsets["transformer.blocks.5.mlp.2.weight"]
    assert not torch.equal(p1[o:o + k], start[o:o + k])
