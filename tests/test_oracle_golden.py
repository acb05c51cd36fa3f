"""Pins oracle/mebt_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import mebt_oracle as O
from conftest import load_golden

torch.set_num_threads(8)


def _digest_check(z, prefix, logits, atol=2e-5):
    lg = logits.float()
    np.testing.assert_allclose(lg[:, ::7, ::113].numpy(), z[prefix + "_sub"], atol=atol, rtol=0)
    np.testing.assert_allclose(torch.logsumexp(lg, -1).numpy(), z[prefix + "_lse"], atol=atol, rtol=0)
    np.testing.assert_allclose(lg.max(-1).values.numpy(), z[prefix + "_rowmax"], atol=atol, rtol=0)
    np.testing.assert_allclose(lg.mean(-1).numpy(), z[prefix + "_rowmean"], atol=atol, rtol=0)
    assert (lg.argmax(-1).numpy() == z[prefix + "_argmax"]).mean() > 0.999


@pytest.mark.parametrize("name", ["micro", "tiny", "tiny5"])
def test_forward_logits(name):
    z, cfg = load_golden(f"forward_{name}")
    P = O.make_weights(cfg, int(z["wseed"]))
    x = torch.from_numpy(z["x"])
    indices = torch.from_numpy(z["indices"])
    for nc in z["ncs"]:
        nc = int(nc)
        logits = O.reconstruct_mask(P, cfg, x, indices[:, :nc], indices[:, nc:])
        assert torch.isfinite(logits).all()
        _digest_check(z, f"nc{nc}", logits)


@pytest.mark.parametrize("name", ["tiny", "micro"])
def test_shared_step(name):
    z, cfg = load_golden(f"shared_step_{name}")
    P = O.make_weights(cfg, int(z["wseed"]))
    x = torch.from_numpy(z["x"])
    indices = torch.from_numpy(z["indices"])
    ls = float(z["label_smoothing"])
    for i, t in enumerate(z["ts"]):
        r = O.shared_step(P, cfg, x, indices, float(t), "linear", ls)
        ce, loss, acc1, acc5, ratio, seq_len, ntw = z[f"t{i}_scalars"]
        assert (r["z_targets"].numpy() == z[f"t{i}_target"]).all()
        assert abs(float(r["ce_sum"]) - ce) < 1e-4 * abs(ce)
        assert abs(float(r["loss"]) - loss) < 1e-5 * abs(loss) + 1e-6
        assert abs(r["acc1"] - acc1) < 1e-4 and abs(r["acc5"] - acc5) < 1e-4
        assert abs(r["ratio"] - ratio) < 1e-12
        np.testing.assert_allclose(torch.logsumexp(r["logits"], -1).numpy(), z[f"t{i}_lse"], atol=2e-5)


def test_training_mode_dropout_replayed_from_reference():
    """The reference run in training mode (p = 0.1) with its nn.Dropout masks recorded: replaying the masks through the
    oracle must give the reference's loss and gradients, i.e. the oracle applies dropout where the reference does."""
    from helpers import reference_dropout_masks
    z, cfg = load_golden("grads_dropout_micro")
    p = float(z["p"])
    drop = reference_dropout_masks(z, p)
    P = {k: v.clone().requires_grad_(True) for k, v in O.make_weights(cfg, int(z["wseed"])).items()}
    r = O.shared_step(P, cfg, torch.from_numpy(z["x"]), torch.from_numpy(z["indices"]), float(z["t"]), "linear", drop=drop)
    assert abs(float(r["loss"]) - float(z["loss"])) < 2e-6 * float(z["loss"])
    np.testing.assert_allclose(r["logits"].detach().reshape(-1)[::997].numpy(), z["logits_sample"], atol=2e-5, rtol=0)
    r["loss"].backward()
    for n, ref_norm in zip((str(n) for n in z["grad_names"]), z["grad_norms"]):
        g = P[n].grad
        if ref_norm < 0:
            assert g is None or float(g.abs().max()) == 0.0, n
            continue
        assert abs(float(g.norm()) - ref_norm) <= 1e-4 * ref_norm + 1e-9, (n, float(g.norm()), float(ref_norm))
    for key in z.files:
        if key.startswith("g:"):
            np.testing.assert_allclose(P[key[2:]].grad.reshape(-1)[::17].numpy(), z[key], atol=1e-7, rtol=1e-3)
    # and dropout really changed the result
    r0 = O.shared_step(P, cfg, torch.from_numpy(z["x"]), torch.from_numpy(z["indices"]), float(z["t"]), "linear")
    assert float((r0["logits"] - r["logits"]).abs().max()) > 1e-2


def test_sample_from_logits_bit_exact():
    z, _ = load_golden("sample_from_logits")
    g = torch.Generator().manual_seed(int(z["seed"]))
    logits = 3.0 * torch.randn(3, 40, 16384, generator=g)
    for tag, (T, k, p) in dict(plain=(1.0, None, None), temp=(0.7, None, None), topk=(1.0, 32, None),
                               topp=(0.9, None, 0.8), both=(0.8, 100, 0.9)).items():
        torch.manual_seed(123)
        q = torch.empty_like(logits).exponential_()
        ids, probs = O.sample_from_logits(logits, T, k, p, q)
        assert (ids.numpy() == z[f"{tag}_ids"]).all(), tag
        np.testing.assert_array_equal(probs.gather(-1, ids.unsqueeze(-1)).squeeze(-1).numpy(), z[f"{tag}_score"])
        assert ((probs > 0).sum(-1).numpy() == z[f"{tag}_nnz"]).all()
        np.testing.assert_array_equal(probs[:, ::5, ::211].numpy(), z[f"{tag}_psub"])


def test_schedules_and_divide_indices():
    z, _ = load_golden("maskgen")
    N = 1024
    g = torch.Generator().manual_seed(5)
    for sched in ("cosine", "linear", "quadratic", "sqrt", "square", "cube", "cosine_plus", "convex"):
        g = torch.Generator().manual_seed(5)
        indices = torch.stack([torch.randperm(N, generator=g) for _ in range(2)])
        sizes = []
        for t in z["ts"]:
            c, tg, sl = O.divide_indices_eval(indices, float(t), sched)
            sizes.append([c.shape[1], tg.shape[1], sl])
            assert (c == indices[:, : c.shape[1]]).all()
        assert (np.array(sizes) == z[f"{sched}_sizes"]).all(), sched
        n_masked = []
        for steps in (8, 32, 128):
            for t_next in np.linspace(0, 1, steps + 1)[1:]:
                t = torch.full((2,), fill_value=t_next)
                n_masked.append(float(torch.ceil(O.schedule(sched, t) * N)[0]))
        assert (np.array(n_masked) == z[f"{sched}_n_masked"]).all(), sched


def test_divide_indices_train_window():
    z, _ = load_golden("maskgen")
    indices = torch.from_numpy(z["train_indices"])
    seq_len, T, start = (int(v) for v in z["train_meta"])
    c, tg, sl = O.divide_indices_train(indices, 0.4, "linear", (4, 16, 16), 300, T, start)
    assert sl == seq_len
    assert (c.numpy() == z["train_ctx"]).all() and (tg.numpy() == z["train_tgt"]).all()


def test_generate_next_mask_and_gibbs():
    z, _ = load_golden("maskgen")
    ctx, tgt, score = (torch.from_numpy(z[k]) for k in ("gnm_ctx", "gnm_tgt", "gnm_score"))
    torch.manual_seed(77)
    q = torch.empty_like(score).exponential_()
    nc, nt = O.generate_next_mask(ctx, tgt, score, 700, 2.25, q)
    assert (nc.numpy() == z["gnm_next_ctx"]).all() and (nt.numpy() == z["gnm_next_tgt"]).all()
    N = 1024
    torch.manual_seed(78)
    perms = torch.stack([torch.randperm(N) for _ in range(2)])
    e, a = torch.empty(2, 0).long(), torch.arange(N).repeat(2, 1)
    cs, ts = O.gibbs_draft_mask(e, a, 4, perms)
    assert (cs[3].numpy() == z["draft_ctx3"]).all() and (ts[3].numpy() == z["draft_tgt3"]).all()
    assert (ts[0].numpy() == z["draft_tgt0"]).all()
    cs, ts = O.gibbs_revise_mask(e, a, 4, perms)
    assert (torch.stack(cs).numpy() == z["revise_ctx"]).all() and (torch.stack(ts).numpy() == z["revise_tgt"]).all()


@pytest.mark.parametrize("name", ["micro", "tiny"])
def test_samplers_bit_exact(name):
    z, cfg = load_golden(f"sampling_{name}")
    P = O.make_weights(cfg, int(z["wseed"]))
    B, seed = int(z["B"]), int(z["seed"])
    x0 = torch.zeros(B, *cfg["shape"], dtype=torch.long)
    out = O.draft_and_revise(P, cfg, x0, O.TorchRng(seed), n_draft=2, draft_t=1.0, n_revise=2, revise_t=0.7, M=2)
    assert (out.numpy() == z["dnr_ids"]).all()
    if name == "micro":
        out = O.draft_and_revise(P, cfg, x0, O.TorchRng(seed + 1), n_draft=4, draft_t=0.9, draft_k=32, n_revise=4,
                                 revise_t=1.0, M=1)
        assert (out.numpy() == z["dnr_topk_ids"]).all()
    strategies = (("maskgit", 6, 4.5), ("random", 4, 4.5), ("bootstrap", 3, 4.5)) if name == "micro" else (("maskgit", 6, 4.5),)
    for strat, steps, ctemp in strategies:
        ids, ctx, tgt = O.sample_maskgit(P, cfg, x0, O.TorchRng(seed + 2), n_steps=steps, strategy=strat,
                                         context_temperature=ctemp, schedule_name="cosine")
        assert (ids.numpy() == z[f"sample_{strat}_ids"]).all(), strat
        assert (ctx.numpy() == z[f"sample_{strat}_ctx"]).all(), strat
        assert (tgt.numpy() == z[f"sample_{strat}_tgt"]).all(), strat


def test_entp_sampler_bit_exact():
    """oracle.sample_entp against the unmodified reference's entp_sample (fixture entp_micro.npz), all three strategies."""
    z, cfg = load_golden("entp_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    B, seed = int(z["B"]), int(z["seed"])
    x0 = torch.zeros(B, *cfg["shape"], dtype=torch.long)
    for strat, steps in (("maskgit", 5), ("random", 4), ("bootstrap", 3)):
        ids, ctx, tgt = O.sample_entp(P, cfg, x0, O.TorchRng(seed), n_steps=steps, strategy=strat, schedule_name="cosine")
        assert (ids.numpy() == z[f"entp_{strat}_ids"]).all(), strat
        assert (ctx.numpy() == z[f"entp_{strat}_ctx"]).all(), strat
        assert (tgt.numpy() == z[f"entp_{strat}_tgt"]).all(), strat


def test_sample_with_fixed_context_and_edit_bit_exact():
    """oracle.sample_maskgit(context_indices, target_indices, edit) against the unmodified reference's
    `sample(..., edit=True / False)` from a given context (fixture edit_micro.npz; transformer.py:373-376,387-389): the
    sliding-window call of `extrapolate()`.  Ids, final context and target index tensors bit for bit."""
    z, cfg = load_golden("edit_micro")
    P = O.make_weights(cfg, int(z["wseed"]))
    seed = int(z["seed"])
    for tag in ("a", "b", "c"):
        x0, ctx0, tgt0 = (torch.from_numpy(z[f"{tag}_{k}"]) for k in ("x0", "ctx0", "tgt0"))
        ids, ctx, tgt = O.sample_maskgit(P, cfg, x0, O.TorchRng(seed), n_steps=int(z[f"{tag}_steps"]), strategy="maskgit",
                                         context_temperature=4.5, schedule_name="cosine", context_indices=ctx0,
                                         target_indices=tgt0, edit=bool(z[f"{tag}_edit"]))
        assert (ids.numpy() == z[f"{tag}_ids"]).all(), tag
        assert (ctx.numpy() == z[f"{tag}_ctx"]).all(), tag
        assert (tgt.numpy() == z[f"{tag}_tgt"]).all(), tag
        keep = int(z[f"{tag}_keep"])
        assert (ids.numpy()[:, :keep] == z[f"{tag}_x0"][:, :keep]).all()


def test_codebook_bit_exact():
    z, _ = load_golden("codebook")
    torch.manual_seed(int(z["cb_seed"]))
    E = torch.randn(16384, 256)
    g = torch.Generator().manual_seed(int(z["z_seed"]))
    zz = torch.randn(2, 256, 4, 16, 16, generator=g)
    out = O.codebook_quantise(zz, E)
    assert (out["encodings"].numpy() == z["encodings"]).all()
    assert abs(float(out["commitment_loss"]) - float(z["commitment_loss"])) < 1e-6
    assert abs(float(out["perplexity"]) - float(z["perplexity"])) < 1e-2
    np.testing.assert_array_equal(out["embeddings"][:, ::9, :, ::3, ::5].numpy(), z["emb_sub"])


def test_flop_model_matches_survey():
    cfg = dict(n_embd=1024, sos_emb=256, vocab_size=16384, n_layer=24,
               mode=(["latent_enc", "latent_self"] * 6 + ["latent_enc"] + ["latent_dec", "lt2l"] * 5 + ["latent_dec"]))
    assert abs(O.forward_flops(cfg, 512, 512) / 1e9 - 234.9) < 0.1
    assert abs(O.forward_flops(cfg, 4096, 4096) / 1e9 - 1054.1) < 0.1


@pytest.mark.parametrize("tag", ["small", "bn"])
def test_vqgan_oracle_vs_reference_golden(tag):
    """oracle.vqgan_oracle (the CPU restatement of the VQGAN encoder / decoder, mebt/vqgan.py:263-405) against the
    unmodified reference's outputs: pre-VQ latent of a random video and the decoded video of random code grids, GroupNorm
    and eval-BatchNorm variants, strides (2,4,4) and (4,2,2)."""
    import json

    from oracle import vqgan_oracle as VO
    z, cfg = load_golden(f"vqgan_{tag}")
    shapes = {k: tuple(v) for k, v in json.loads(str(z["shapes_json"])).items()}
    P = VO.make_weights(shapes, int(z["wseed"]))
    ds = tuple(cfg["downsample"])
    lat = VO.pre_quant(P, torch.from_numpy(z["x"]), ds)
    rec = VO.decode(P, torch.from_numpy(z["codes"]), ds)
    np.testing.assert_allclose(lat.numpy(), z["z"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rec.numpy(), z["rec"], rtol=1e-4, atol=1e-5)
