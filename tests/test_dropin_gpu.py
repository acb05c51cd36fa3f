"""The drop-in `Net2NetTransformer` / `MaskGen` / `Codebook` API on the GPU against the CPU oracle and the
reference-generated fixtures."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import build_model

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _rel(a, b):
    return (a - b).abs().max().item() / b.abs().max().item()


def test_reconstruct_mask_and_forward_vs_reference_fixture():
    from oracle import mebt_oracle as O
    z, cfg = load_golden("shared_step_tiny")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P)
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    for i, t in enumerate(z["ts"]):
        logits, target, NT_weight, seq_len = model(x, None, t=float(t), indices=indices)
        ce, loss, acc1, acc5, ratio, sl, ntw = z[f"t{i}_scalars"]
        assert (target.cpu().numpy() == z[f"t{i}_target"]).all()          # bit-exact mask split / target ids
        assert NT_weight == ntw and seq_len == sl
        assert logits.dtype == torch.float32 and logits.shape == (2, target.shape[1], 16384)
        lse = torch.logsumexp(logits, -1).cpu().numpy()
        assert np.abs(lse - z[f"t{i}_lse"]).max() < TOL * np.abs(z[f"t{i}_lse"]).max()
    # shared_step draws t from python's RNG: pin it
    import random
    random.seed(0)
    t0 = random.random()
    random.seed(0)
    acc1_g, acc5_g, loss_g, ratio_g = model.shared_step(dict(video=x, label=x, indices=indices), 0)
    r = O.shared_step(P, cfg, x.cpu(), indices.cpu(), float(torch.tensor(t0)), "linear", 0.0)
    assert abs(float(loss_g) - float(r["loss"])) < 2e-3 * float(r["loss"])
    assert abs(ratio_g - r["ratio"]) < 1e-12
    assert abs(float(acc1_g) - r["acc1"]) < 0.5 and abs(float(acc5_g) - r["acc5"]) < 0.5


def test_reconstruct_mask_vs_oracle_micro():
    from oracle import mebt_oracle as O
    z, cfg = load_golden("forward_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P)
    x, indices = torch.from_numpy(z["x"]), torch.from_numpy(z["indices"])
    for nc in (0, 100, 255):
        ref = O.reconstruct_mask(P, cfg, x, indices[:, :nc], indices[:, nc:])
        logits, idx = model.reconstruct_mask(x.cuda(), indices[:, :nc].cuda(), indices[:, nc:].cuda())
        assert idx is None and _rel(logits.cpu(), ref) < TOL


class CpuDraws:
    """Feeds the drop-in the draws the CPU oracle consumes: same torch CPU generator, same order."""

    def __call__(self, kind, shape, device):
        if kind == "exponential":
            return torch.empty(shape, dtype=torch.float32).exponential_().to(device)
        return torch.randn(shape).to(device)


def test_draft_and_revise_teacher_forced_parity():
    """Step-by-step parity with the oracle's draft-and-revise run: masks bit-exact; logits within tolerance at
    every step; sampled ids bit-exact given the oracle's logits and noise; write-back bit-exact."""
    from mebt_b200 import ops
    from oracle import mebt_oracle as O
    z, cfg = load_golden("sampling_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P, schedule="cosine")
    B, N = 2, 256
    rng = O.TorchRng(123)
    x = torch.zeros(B, N, dtype=torch.long)
    e, a = torch.empty(B, 0, dtype=torch.long), torch.arange(N).repeat(B, 1)
    for phase, n_steps, T, k in (("draft", 4, 1.0, None), ("revise", 4, 0.8, 32)):
        st = torch.get_rng_state()
        perms = torch.stack([rng.randperm(N) for _ in range(B)])
        ctxs, tgts = (O.gibbs_draft_mask if phase == "draft" else O.gibbs_revise_mask)(e, a, n_steps, perms)
        torch.set_rng_state(st)
        mk = model.mask_sampler.create_gibbs_draft_mask if phase == "draft" else model.mask_sampler.create_gibbs_revise_mask
        g_ctxs, g_tgts = mk(e.cuda(), a.cuda(), n_steps, "cuda")
        for c, t, gc, gt in zip(ctxs, tgts, g_ctxs, g_tgts):
            assert torch.equal(gc.cpu(), c) and torch.equal(gt.cpu(), t)            # masks: bit-exact
            ref_logits = O.reconstruct_mask(P, cfg, x, c, t)
            logits, _ = model.reconstruct_mask(x.cuda(), gc, gt)
            assert _rel(logits.cpu(), ref_logits) < TOL
            q = rng.exponential(ref_logits.shape)
            ref_ids, ref_probs = O.sample_from_logits(ref_logits, T, k, None, q)
            ids, scores, _ = ops.sample_logits(ref_logits.view(-1, 16384).cuda(), T, k, None, noise=q.view(-1, 16384).cuda())
            assert torch.equal(ids.cpu().view_as(ref_ids), ref_ids)                   # selections: bit-exact
            ref_scores = ref_probs.gather(-1, ref_ids.unsqueeze(-1)).squeeze(-1)
            assert torch.allclose(scores.cpu().view_as(ref_scores), ref_scores, rtol=5e-6, atol=1e-12)
            xg = x.cuda()
            ops.scatter_ids(xg, gt, ids.view(B, -1))
            x = x.scatter(1, t, ref_ids)
            assert torch.equal(xg.cpu(), x)                                            # write-back: bit-exact


@pytest.mark.parametrize("rng_mode", ["torch", "philox"])
def test_samplers_end_to_end(rng_mode):
    from oracle import mebt_oracle as O
    z, cfg = load_golden("sampling_micro")
    model = build_model(cfg, O.make_weights(cfg, int(z["wseed"])), schedule="cosine")
    model.rng_mode, model.rng_seed = rng_mode, 5
    B, N = 3, 256
    x0 = torch.zeros(B, 1, 16, 16, dtype=torch.long, device="cuda")

    def run():
        torch.manual_seed(0)
        model._rng_offset = 0
        model.mask_sampler.rng_offset = 0
        a = model.draft_and_revise(x0, None, n_draft=4, draft_t=1.0, n_revise=2, revise_t=0.7, M=2)
        b, ctx, tgt = model.sample(x0, None, 1.0, 32, None, n_steps=6, strategy="maskgit", context_temperature=4.5)
        return a, b, ctx, tgt

    a, b, ctx, tgt = run()
    a2, b2, _, _ = run()
    assert torch.equal(a, a2) and torch.equal(b, b2)                                  # seeded runs are reproducible
    assert a.shape == (B, N) and b.shape == (B, N) and a.dtype == torch.int64
    assert int(a.min()) >= 0 and int(a.max()) < 16384 and (a != 0).float().mean() > 0.99
    both = torch.cat([ctx, tgt], 1).sort(1).values.cpu()
    assert torch.equal(both, torch.arange(N).repeat(B, 1))                            # contexts + targets partition N
    assert (x0 == 0).all()                                                            # the input is not modified
    for strat in ("random", "bootstrap"):
        ids, c2, t2 = model.sample(x0, None, n_steps=3, strategy=strat)
        assert ids.shape == (B, N) and c2.shape[1] + t2.shape[1] == N
    out = model.sample(x0, None, n_steps=3, debug=True)
    assert len(out) == 6 and out[5].shape == (B, N, 16384) and len(out[3]) >= 2


def test_module_level_api():
    from mebt_b200.transformer import gumbel_sort, sample_from_logits, top_k_logits, top_p_probs
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = 2 * torch.randn(2, 5, 16384, device="cuda", generator=g)
    ids, probs = sample_from_logits(logits, 0.9, 50, None, return_probs=True)
    assert ids.shape == (2, 5) and probs.shape == logits.shape and ((probs > 0).sum(-1) == 50).all()
    assert torch.allclose(probs.sum(-1), torch.ones(2, 5, device="cuda"), atol=1e-5)
    ids2 = sample_from_logits(logits)
    assert ids2.shape == (2, 5)
    order = gumbel_sort(torch.softmax(logits[0], -1))
    assert order.shape == (5, 16384) and (order.sort(-1).values == torch.arange(16384, device="cuda")).all()
    assert ((top_k_logits(logits, 7) > -float("inf")).sum(-1) == 7).all()
    assert torch.allclose(top_p_probs(torch.softmax(logits, -1), 0.5).sum(-1), torch.ones(2, 5, device="cuda"), atol=1e-5)


def test_gpt_block_crossattention_modules():
    """The per-module API (GPT / Block / CrossAttention forward with [B,n,D] tensors) against the oracle."""
    from oracle import mebt_oracle as O
    z, cfg = load_golden("forward_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P)
    gpt = model.transformer
    g = torch.Generator().manual_seed(5)
    B, L, NC, NT, D = 2, 64, 70, 50, 128
    lat, ctx, tgt = (torch.randn(B, n, D, generator=g) for n in (L, NC, NT))
    for i, blk in enumerate(gpt.blocks):
        r_lat, r_ctx, r_tgt = O.block_forward(P, i, blk.mode, cfg["n_head"], lat, ctx, tgt)
        o_lat, o_ctx, o_tgt, bias, idx = blk(lat.cuda(), ctx.cuda(), tgt.cuda(), None, 0.)
        assert idx is None and o_lat.dtype == torch.float32
        assert _rel(o_lat.cpu(), r_lat) < 2e-2 and _rel(o_tgt.cpu(), r_tgt) < 2e-2 and _rel(o_ctx.cpu(), r_ctx) < 2e-2
    qn = torch.randn(B, L, D, generator=g)
    kn = torch.randn(B, NC, D, generator=g)
    ref = O.cross_attention(P, "transformer.blocks.0.attn.", qn, kn, cfg["n_head"])
    y, a, b_, c = gpt.blocks[0].attn(qn.cuda(), kn.cuda(), 0.)
    assert a is None and _rel(y.cpu(), ref) < 2e-2
    logits, none = gpt(lat.cuda(), ctx.cuda(), tgt.cuda(), None)
    assert none is None and _rel(logits.cpu(), O.gpt_forward(P, cfg, lat, ctx, tgt)) < TOL
    assert gpt.get_block_size() == cfg["block_size"]


def test_codebook_and_vqgan_boundary():
    from mebt_b200.modules.codebook import Codebook
    from mebt_b200.vqgan import VQGAN
    from oracle import mebt_oracle as O
    z, _ = load_golden("codebook")
    torch.manual_seed(int(z["cb_seed"]))
    cb = Codebook(16384, 256)
    E = cb.embeddings.clone()
    cb = cb.cuda().eval()
    g = torch.Generator().manual_seed(int(z["z_seed"]))
    zz = torch.randn(2, 256, 4, 16, 16, generator=g)
    out = cb(zz.cuda())
    ref = O.codebook_quantise(zz, E)
    mism = out["encodings"].cpu() != ref["encodings"]
    assert mism.sum() <= 2
    assert abs(float(out["commitment_loss"]) - float(z["commitment_loss"])) < 1e-4
    assert abs(float(out["perplexity"]) - float(z["perplexity"])) < 1.0
    if mism.sum() == 0:
        assert torch.equal(out["embeddings"].cpu(), ref["embeddings"])
    assert torch.equal(cb.dictionary_lookup(ref["encodings"].cuda()).cpu(), torch.nn.functional.embedding(ref["encodings"], E))
    vq = VQGAN(16384, 256).cuda().eval()
    vq.codebook.embeddings.copy_(E)
    enc = vq.encode(zz.cuda())
    assert (enc.cpu() != ref["encodings"]).sum() <= 2
    dec = vq.decode(ref["encodings"].cuda())
    assert torch.equal(dec.cpu(), O.codebook_decode_gather(ref["encodings"], E))


def test_fp32_precision_forward_and_loss_vs_reference_fixture():
    """BASELINE configs[0] (MeBT tiny, fwd + masked CE) with `precision = "fp32"`: logits' log-sum-exp and the
    loss within 1e-4 of what the unmodified reference produced; targets and mask split bit-exact."""
    from oracle import mebt_oracle as O
    z, cfg = load_golden("shared_step_tiny")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P)
    model.precision = "fp32"
    x, indices = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["indices"]).cuda()
    B = x.shape[0]
    for i, t in enumerate(z["ts"]):
        logits, target, NT_weight, seq_len = model(x, None, t=float(t), indices=indices)
        ce, loss, acc1, acc5, ratio, sl, ntw = z[f"t{i}_scalars"]
        assert (target.cpu().numpy() == z[f"t{i}_target"]).all()
        lse = torch.logsumexp(logits, -1).cpu().numpy()
        assert np.abs(lse - z[f"t{i}_lse"]).max() < 1e-4 * np.abs(z[f"t{i}_lse"]).max()
        ce_gpu = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), target.reshape(-1), reduction="sum")
        assert abs(float(ce_gpu) - ce) < 1e-4 * ce
        from mebt_b200 import ops
        stats, _ = ops.masked_ce(logits.reshape(-1, logits.shape[-1]), target.reshape(-1), 0.0)
        assert abs(float(stats[0]) - ce) < 1e-4 * ce                           # the K5 kernel on the fp32 logits
        loss_gpu = float(stats[0]) / (B * sl * ratio ** 1.0)
        assert abs(loss_gpu - loss) < 1e-4 * loss


def test_sampling_script_pipelines_on_the_cuda_model():
    """bidirect_sample / extrapolate (sample_vqgan_transformer_videos.py:22-157) over the CUDA sampler: window
    bookkeeping on real `model.sample` outputs (the call-by-call parity with the reference's functions is the CPU test
    tests/test_pipelines_cpu.py)."""
    from oracle import mebt_oracle as O
    from mebt_b200.pipelines import bidirect_sample, extrapolate
    z, cfg = load_golden("forward_tiny")
    model = build_model(cfg, O.make_weights(cfg, int(z["wseed"])), schedule="cosine")
    B, hw = 2, 256
    log = bidirect_sample(model, B, total_length=16, step_size=16, context_size=8, vid_n_steps=4, vid_c_temp=1.0, bootstrap=2)
    cm = log["code_maps"]
    assert cm.shape == (B, 4, 16, 16) and cm.dtype == torch.long and int(cm.min()) >= 0 and int(cm.max()) < 16384
    assert log["score"].shape == (B,) and torch.isfinite(log["score"]).all() and (log["score"] < 0).all()
    start = cm.clone()
    log = extrapolate(model, start, total_length=32, step_size=16, context_size=8, vid_n_steps=3, vid_c_temp=1.0)
    cm2 = log["code_maps"]
    assert cm2.shape == (B, 8, 16, 16)                       # 4 given frames + 2 jumps of 2 new frames
    assert torch.equal(cm2[:, :4], start)                    # the given tokens are kept
    assert int(cm2.min()) >= 0 and int(cm2.max()) < 16384
    assert not torch.equal(cm2[:, 4:6], cm2[:, 6:8])


def test_sample_debug_selected_probs_equal_the_gather_of_the_dense_map():
    """sample(debug=True, debug_probs="selected") returns the [B, N] probabilities of the finally chosen codes - the
    gather the sampling scripts apply to the dense [B, N, 16384] map (sample_vqgan_transformer_videos.py:85-89) - and is
    identical to that gather on the dense map of an identically seeded run."""
    from oracle import mebt_oracle as O
    z, cfg = load_golden("sampling_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    model = build_model(cfg, P, schedule="cosine")
    model.rng_mode, model.rng_seed = "philox", 77
    x0 = torch.zeros(2, 256, dtype=torch.long, device="cuda")
    outs = []
    for mode in ("dense", "selected"):
        model._rng_offset = 0
        model.mask_sampler.rng_offset = 0
        # top_k = V filters nothing but keeps both runs on the same sampling kernel (without it the run that needs no
        # probabilities takes the streaming inverse-CDF kernel, whose prefix sums round differently)
        outs.append(model.sample(x0, None, 1.0, 16384, None, n_steps=5, strategy="maskgit", context_temperature=4.5,
                                 debug=True, debug_probs=mode))
    (xd, cd, td, _, _, dense), (xs, cs, ts, _, _, sel) = outs
    assert torch.equal(xd, xs) and torch.equal(cd, cs) and torch.equal(td, ts)
    assert dense.shape == (2, 256, 16384) and sel.shape == (2, 256)
    gathered = torch.gather(dense, -1, xd.unsqueeze(-1)).squeeze(-1)
    assert torch.allclose(gathered, sel, rtol=1e-6, atol=0)


def test_vtokens_false_model_encodes_and_decodes_through_its_vqgan(tmp_path):
    """`vtokens: False` end to end on the GPU: the transformer's frozen first stage (loaded from a Lightning-format VQGAN
    checkpoint, transformer.py:180-192) turns videos into the token grid `encode_to_z` returns (:683-694), and the sampling
    script's `bidirect_sample` decodes its code map to pixels through the same first stage by default
    (sample_vqgan_transformer_videos.py:82)."""
    from helpers import STL_16F, model_configs, to_attr
    from mebt_b200 import pipelines, vqgan as V
    from mebt_b200.transformer import Net2NetTransformer
    from oracle import vqgan_oracle as VO
    vcfg = dict(embedding_dim=64, n_codes=128, n_hiddens=32, downsample=(2, 4, 4), image_channels=3, norm_type="group",
                padding_type="replicate", sequence_length=8, sample_every_n_frames=1, resolution=32)
    vq = V.VQGAN(V._Args(vcfg))
    shapes = {k: tuple(v.shape) for k, v in vq.state_dict().items() if not k.startswith("codebook.") or k == "codebook.embeddings"}
    vq.load_state_dict({**vq.state_dict(), **VO.make_weights(shapes, 3)})
    ckpt = tmp_path / "vqgan.ckpt"
    torch.save({"hyper_parameters": {"args": V._Args(vcfg)}, "state_dict": vq.state_dict()}, ckpt)
    cfg = dict(STL_16F, n_embd=64, n_head=1, sos_emb=16, n_layer=4, mode=["latent_enc", "latent_self", "lt2l", "latent_dec"],
               vocab_size=128, block_size=256, shape=[4, 8, 8])
    params, _, mask = model_configs(cfg, schedule="cosine")
    params.vtokens = False
    model = Net2NetTransformer(params, to_attr(dict(params=dict(ckpt_path=str(ckpt), ignore_keys=["loss"]))), mask).cuda().eval()
    video = torch.rand(2, 3, 8, 32, 32, generator=torch.Generator().manual_seed(1)) - 0.5
    emb, tokens = model.encode_to_z(video.cuda())
    assert tokens.shape == (2, 256) and emb.shape == (2, 4, 8, 8, 64)
    assert torch.equal(tokens.view(2, 4, 8, 8), model.first_stage_model.encode(video.cuda()))
    # the oracle's encoder agrees except across near-ties of the code distances
    P = {k: v for k, v in vq.state_dict().items()}
    ref = VO.pre_quant(P, video, (2, 4, 4))
    E = P["codebook.embeddings"]
    flat = ref.permute(0, 2, 3, 4, 1).reshape(-1, 64)
    ref_codes = ((flat ** 2).sum(1, keepdim=True) - 2 * flat @ E.t() + (E ** 2).sum(1)[None]).argmin(1).view(2, 4, 8, 8)
    assert (tokens.view(2, 4, 8, 8).cpu() != ref_codes).float().mean() < 0.1
    torch.manual_seed(0)
    # the script hard-codes 4 video frames per latent frame (:29); this small VQGAN has 2, so 16 "frames" = its 4 latent frames
    log = pipelines.bidirect_sample(model, 2, total_length=16, step_size=16, context_size=8, vid_n_steps=4)
    assert log["code_maps"].shape == (2, 4, 8, 8)
    assert log["samples"].shape == (2, 3, 8, 32, 32) and float(log["samples"].min()) >= 0.0 and float(log["samples"].max()) <= 1.0
    again = torch.clamp(model.first_stage_model.decode(log["code_maps"]), -0.5, 0.5) + 0.5
    assert torch.equal(again, log["samples"])
