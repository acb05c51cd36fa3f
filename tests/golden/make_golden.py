"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference needs pytorch_lightning / h5py / imageio / skvideo, which are not installed; they are
stubbed in sys.modules — nothing on the hot path touches them.  Weights come from
`oracle.mebt_oracle.make_weights` (a per-tensor seeded recipe), loaded into the reference model with
`load_state_dict(strict=True)`, so fixtures hold only inputs and outputs (small).

Each fixture stores the config as JSON, the integer inputs, and outputs.  Logits are stored as a strided
subsample plus per-row log-sum-exp and argmax over the full 16384-way row, which pins every row.
"""
from __future__ import annotations

import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = os.environ.get("MEBT_REF", "/root/reference")


def install_stubs():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        global_step = 0
        current_epoch = 0

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

    class LightningDataModule:
        pass

    class Trainer:
        pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    pl.Trainer = Trainer
    cb = types.ModuleType("pytorch_lightning.callbacks")
    cb.ModelCheckpoint = type("ModelCheckpoint", (), {})
    cb.Callback = type("Callback", (), {})
    pl.callbacks = cb
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.callbacks"] = cb
    for name in ("h5py", "imageio", "skvideo", "skvideo.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skvideo"].io = sys.modules["skvideo.io"]


class AttrDict(dict):
    """Stands in for OmegaConf: attribute access + hasattr + `in` + .get."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


CONFIGS = {
    # 5 layers so that all four latent modes carry signal to the logits (SURVEY.md §8(d) cfg #1 note)
    "micro": dict(n_embd=128, n_head=2, sos_emb=64, block_size=256, shape=[1, 16, 16], n_layer=5, vocab_size=16384,
                  mode=["latent_enc", "latent_self", "latent_dec", "lt2l", "latent_dec"], avg_loss=1.0),
    # BASELINE.json configs[0]
    "tiny": dict(n_embd=256, n_head=4, sos_emb=256, block_size=1024, shape=[4, 16, 16], n_layer=4, vocab_size=16384,
                 mode=["latent_enc", "latent_self", "latent_dec", "lt2l"], avg_loss=1.0),
    "tiny5": dict(n_embd=256, n_head=4, sos_emb=256, block_size=1024, shape=[4, 16, 16], n_layer=5, vocab_size=16384,
                  mode=["latent_enc", "latent_self", "lt2l", "latent_dec", "latent_dec"], avg_loss=1.0),
}


def build_reference(cfg: dict, schedule: str, seed: int, pdrop: float = 0.0):
    from mebt.transformer import Net2NetTransformer  # the reference's class

    from oracle.mebt_oracle import make_weights

    params = to_attr(dict(
        unconditional=True, vocab_size=cfg["vocab_size"], first_stage_vocab_size=cfg["vocab_size"],
        block_size=cfg["block_size"], n_layer=cfg["n_layer"], n_head=cfg["n_head"], n_embd=cfg["n_embd"],
        n_unmasked=0, embd_pdrop=pdrop, resid_pdrop=pdrop, attn_pdrop=pdrop, sample_every_n_latent_frames=0,
        first_stage_key="video", cond_stage_key="label", vtokens=True, vtokens_pos=False, vis_epoch=100,
        sos_emb=cfg["sos_emb"], avg_loss=bool(cfg.get("avg_loss", 1.0)), mode=list(cfg["mode"]), class_cond_dim=None))
    mask = to_attr(dict(target="mebt.mask_sampler.MaskGen",
                        params=dict(iid=False, schedule=schedule, max_token=cfg["block_size"], method="mlm",
                                    shape=cfg["shape"], t_range=[0.0, 1.0], budget=cfg["block_size"])))
    vq = to_attr(dict(params=dict(ckpt_path="unused", ignore_keys=["loss"])))
    model = Net2NetTransformer(params, vq, mask)
    W = make_weights(cfg, seed)
    missing, unexpected = model.load_state_dict(W, strict=True)
    assert not missing and not unexpected
    return model.eval(), W


def logits_digest(logits: torch.Tensor) -> dict:
    lg = logits.detach().float()
    return dict(sub=lg[:, ::7, ::113].contiguous().numpy(), lse=torch.logsumexp(lg, -1).numpy(),
                argmax=lg.argmax(-1).numpy(), rowmax=lg.max(-1).values.numpy(), rowmean=lg.mean(-1).numpy())


def synth_tokens(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    N = int(np.prod(cfg["shape"]))
    x = torch.randint(0, cfg["vocab_size"], (B, *cfg["shape"]), generator=g)
    indices = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    return x, indices


def save(name, cfg, **arrays):
    out = {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    out["cfg_json"] = np.array(json.dumps(cfg))
    path = HERE / f"{name}.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path.name}: {path.stat().st_size / 1024:.0f} KiB")


def gen_forward(cfg_name, B, wseed, dseed, ncs):
    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "linear", wseed)
    x, indices = synth_tokens(cfg, B, dseed)
    N = indices.shape[1]
    arrays = dict(x=x, indices=indices, wseed=wseed, ncs=np.array(ncs))
    for nc in ncs:
        ctx, tgt = indices[:, :nc], indices[:, nc:]
        with torch.no_grad():
            logits, _ = model.reconstruct_mask(x, ctx, tgt)
        for k, v in logits_digest(logits).items():
            arrays[f"nc{nc}_{k}"] = v
    save(f"forward_{cfg_name}", cfg, **arrays)


def gen_shared_step(cfg_name, B, wseed, dseed, ts, label_smoothing=0.0):
    """forward() + shared_step() in eval mode with explicit t (deterministic)."""
    import torch.nn.functional as F

    from mebt.utils import accuracy

    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "linear", wseed)
    model.label_smoothing = label_smoothing
    x, indices = synth_tokens(cfg, B, dseed)
    arrays = dict(x=x, indices=indices, wseed=wseed, ts=np.array(ts), label_smoothing=label_smoothing)
    for i, t in enumerate(ts):
        with torch.no_grad():
            logits, target, NT_weight, seq_len = model(x, None, t=t, indices=indices)
            ratio = NT_weight / float(seq_len)
            ce = F.cross_entropy(logits.reshape(-1, logits.size(-1)), target.reshape(-1), reduction="sum",
                                 label_smoothing=label_smoothing)
            loss = ce / (B * seq_len * ratio ** model.config.avg_loss)
            acc1, acc5 = accuracy(logits.reshape(-1, logits.shape[-1]), target.reshape(-1), topk=(1, 5))
        arrays[f"t{i}_target"] = target
        arrays[f"t{i}_scalars"] = np.array([float(ce), float(loss), float(acc1), float(acc5), ratio, seq_len, NT_weight],
                                           dtype=np.float64)
        arrays[f"t{i}_lse"] = torch.logsumexp(logits.float(), -1)
    save(f"shared_step_{cfg_name}", cfg, **arrays)


def gen_grads(cfg_name, B, wseed, dseed, t):
    """Loss gradients w.r.t. a few parameters (digest: per-tensor L2 norms + strided samples)."""
    import torch.nn.functional as F

    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "linear", wseed)
    x, indices = synth_tokens(cfg, B, dseed)
    logits, target, NT_weight, seq_len = model(x, None, t=t, indices=indices)
    ratio = NT_weight / float(seq_len)
    ce = F.cross_entropy(logits.reshape(-1, logits.size(-1)), target.reshape(-1), reduction="sum")
    loss = ce / (B * seq_len * ratio ** model.config.avg_loss)
    loss.backward()
    arrays = dict(x=x, indices=indices, wseed=wseed, t=t, loss=float(loss))
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(float(p.grad.norm()) if p.grad is not None else -1.0)
        if p.grad is not None and ("blocks.0." in n or "blocks.3." in n or "_emb" in n or "ln_f" in n):
            arrays["g:" + n] = p.grad.reshape(-1)[::17].contiguous()
    arrays["grad_names"] = np.array(names)
    arrays["grad_norms"] = np.array(norms)
    save(f"grads_{cfg_name}", cfg, **arrays)


def gen_sampling(cfg_name, B, wseed, seed):
    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "cosine", wseed)
    shape = (B, *cfg["shape"])
    arrays = dict(wseed=wseed, seed=seed, B=B)
    x0 = torch.zeros(shape, dtype=torch.long)
    # draft_and_revise, defaults scaled down
    torch.manual_seed(seed)
    out = model.draft_and_revise(x0, None, n_draft=2, draft_t=1.0, n_revise=2, revise_t=0.7, M=2)
    arrays["dnr_ids"] = out
    torch.manual_seed(seed + 1)
    out = model.draft_and_revise(x0, None, n_draft=4, draft_t=0.9, draft_k=32, n_revise=4, revise_t=1.0, M=1)
    arrays["dnr_topk_ids"] = out
    # maskgit sample, three strategies
    for strat, steps, ctemp in (("maskgit", 6, 4.5), ("random", 4, 4.5), ("bootstrap", 3, 4.5)):
        torch.manual_seed(seed + 2)
        ids, ctx, tgt = model.sample(x0, None, temperature=1.0, top_k=None, top_p=None, n_steps=steps,
                                     strategy=strat, context_temperature=ctemp)
        arrays[f"sample_{strat}_ids"] = ids
        arrays[f"sample_{strat}_ctx"] = ctx
        arrays[f"sample_{strat}_tgt"] = tgt
    save(f"sampling_{cfg_name}", cfg, **arrays)


def gen_entp(cfg_name, B, wseed, seed):
    """Net2NetTransformer.entp_sample (transformer.py:449-542) in its three strategies, and `sample(debug=True)`'s
    probability bookkeeping (transformer.py:396,428-437): the per-position top-8 of the dense [B,N,16384] tensor."""
    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "cosine", wseed)
    x0 = torch.zeros((B, *cfg["shape"]), dtype=torch.long)
    arrays = dict(wseed=wseed, seed=seed, B=B)
    for strat, steps in (("maskgit", 5), ("random", 4), ("bootstrap", 3)):
        torch.manual_seed(seed)
        ids, ctx, tgt = model.entp_sample(x0, None, temperature=1.0, top_k=None, top_p=None, n_steps=steps, strategy=strat)
        arrays[f"entp_{strat}_ids"], arrays[f"entp_{strat}_ctx"], arrays[f"entp_{strat}_tgt"] = ids, ctx, tgt
    torch.manual_seed(seed + 1)
    out = model.sample(x0, None, temperature=0.9, top_k=64, top_p=None, n_steps=4, strategy="maskgit",
                       context_temperature=4.5, debug=True)
    probs = out[5]
    top = probs.topk(8, -1)
    arrays.update(debug_ids=out[0], debug_ctx=out[1], debug_tgt=out[2], debug_top_p=top.values, debug_top_i=top.indices,
                  debug_chosen_p=probs.gather(-1, out[0].unsqueeze(-1)).squeeze(-1))
    save(f"entp_{cfg_name}", cfg, **arrays)


def gen_edit(cfg_name, B, wseed, seed):
    """`sample(..., context_indices, target_indices, edit=...)` (transformer.py:373-376,387-389,399): the call the sliding
    window of `extrapolate()` makes (sample_vqgan_transformer_videos.py:95-157) - a given context that is never rewritten,
    and with edit=True a mask schedule sized by the number of targets."""
    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "cosine", wseed)
    N = int(np.prod(cfg["shape"]))
    arrays = dict(wseed=wseed, seed=seed, B=B)
    for tag, edit, steps, n_keep in (("a", True, 5, 192), ("b", True, 4, 64), ("c", False, 6, 128)):
        g = torch.Generator().manual_seed(7)
        x0 = torch.randint(0, cfg["vocab_size"], (B, N), generator=g)
        ctx0 = torch.stack([torch.randperm(n_keep, generator=g) for _ in range(B)])
        tgt0 = torch.stack([n_keep + torch.randperm(N - n_keep, generator=g) for _ in range(B)])
        torch.manual_seed(seed)
        ids, ctx, tgt = model.sample(x0.view(B, *cfg["shape"]), None, temperature=1.0, top_k=None, top_p=None, n_steps=steps,
                                     context_indices=ctx0, target_indices=tgt0, strategy="maskgit", context_temperature=4.5,
                                     edit=edit)
        arrays.update({f"{tag}_edit": int(edit), f"{tag}_steps": steps, f"{tag}_keep": n_keep, f"{tag}_x0": x0, f"{tag}_ctx0": ctx0,
                       f"{tag}_tgt0": tgt0, f"{tag}_ids": ids, f"{tag}_ctx": ctx, f"{tag}_tgt": tgt})
    save(f"edit_{cfg_name}", cfg, **arrays)


def gen_sample_from_logits():
    """sample_from_logits / top-k / top-p / gumbel on random logits: ids + probs digests."""
    from mebt.transformer import sample_from_logits

    g = torch.Generator().manual_seed(11)
    logits = 3.0 * torch.randn(3, 40, 16384, generator=g)
    arrays = dict(seed=11)
    for tag, (T, k, p) in dict(plain=(1.0, None, None), temp=(0.7, None, None), topk=(1.0, 32, None),
                               topp=(0.9, None, 0.8), both=(0.8, 100, 0.9)).items():
        torch.manual_seed(123)
        ids, probs = sample_from_logits(logits, T, k, p, return_probs=True)
        arrays[f"{tag}_ids"] = ids
        arrays[f"{tag}_score"] = probs.gather(-1, ids.unsqueeze(-1)).squeeze(-1)
        arrays[f"{tag}_nnz"] = (probs > 0).sum(-1)
        arrays[f"{tag}_psub"] = probs[:, ::5, ::211].contiguous()
    save("sample_from_logits", {}, **arrays)


def gen_maskgen():
    from mebt.mask_sampler import MaskGen

    arrays = {}
    N = 1024
    for sched in ("cosine", "linear", "quadratic", "sqrt", "square", "cube", "cosine_plus", "convex"):
        mg = MaskGen(schedule=sched, shape=(4, 16, 16), budget=1024).eval()
        g = torch.Generator().manual_seed(5)
        indices = torch.stack([torch.randperm(N, generator=g) for _ in range(2)])
        ts = [0.0, 0.1, 0.25, 1.0 / 3.0, 0.5, 0.625, 0.75, 0.9, 0.999, 1.0]
        sizes = []
        for t in ts:
            c, tg, sl = mg.divide_indices(indices, torch.tensor(t), None, None)
            sizes.append([c.shape[1], tg.shape[1], int(sl)])
        arrays[f"{sched}_sizes"] = np.array(sizes)
        # float32 schedule -> ceil, as sample() evaluates it (transformer.py:398-399)
        n_masked = []
        for steps in (8, 32, 128):
            for t_next in np.linspace(0, 1, steps + 1)[1:]:
                t = torch.full((2,), fill_value=t_next)
                n_masked.append(float(torch.ceil(mg.schedule_fn(t) * N)[0]))
        arrays[f"{sched}_n_masked"] = np.array(n_masked)
    arrays["ts"] = np.array(ts)
    # training-mode slicing (numpy RNG draws recorded)
    mg = MaskGen(schedule="linear", shape=(4, 16, 16), budget=300).train()
    g = torch.Generator().manual_seed(6)
    indices = torch.stack([torch.randperm(N, generator=g) for _ in range(3)])
    np.random.seed(3)
    vid_t = np.arange(4) + 1
    prior = np.array([0.1, 0.2, 0.3, 0.4])
    st = np.random.get_state()
    c, tg, sl = mg.divide_indices(indices, torch.tensor(0.4), vid_t, prior.copy())
    np.random.set_state(st)
    T = np.random.choice(vid_t, p=prior / prior.sum())
    start = 0 if T == 4 else np.random.randint(0, 4 - T + 1)
    arrays.update(train_indices=indices, train_ctx=c, train_tgt=tg, train_meta=np.array([int(sl), int(T), int(start)]))
    # generate_next_mask
    mg = MaskGen(schedule="cosine", shape=(4, 16, 16)).eval()
    g = torch.Generator().manual_seed(7)
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(2)])
    ctx, tgt = perm[:, :200], perm[:, 200:]
    score = torch.rand(2, 824, generator=g)
    torch.manual_seed(77)
    nc, nt = mg.generate_next_mask(ctx, tgt, score, 0.5, strategy="maskgit", context_temperature=2.25,
                                   n_masked_toks=torch.tensor([700.0, 700.0]))
    arrays.update(gnm_ctx=ctx, gnm_tgt=tgt, gnm_score=score, gnm_next_ctx=nc, gnm_next_tgt=nt)
    # gibbs masks
    torch.manual_seed(78)
    c_l, t_l = MaskGen.create_gibbs_draft_mask(torch.empty(2, 0).long(), torch.arange(N).repeat(2, 1), 4, "cpu")
    arrays.update(draft_ctx3=c_l[3], draft_tgt3=t_l[3], draft_tgt0=t_l[0])
    torch.manual_seed(78)
    c_s, t_s = MaskGen.create_gibbs_revise_mask(torch.empty(2, 0).long(), torch.arange(N).repeat(2, 1), 4, "cpu")
    arrays.update(revise_ctx=c_s, revise_tgt=t_s)
    save("maskgen", {}, **arrays)


def gen_codebook():
    from mebt.modules.codebook import Codebook

    torch.manual_seed(0)
    cb = Codebook(16384, 256)
    cb._need_init = False
    cb.eval()
    g = torch.Generator().manual_seed(4)
    z = torch.randn(2, 256, 4, 16, 16, generator=g)
    with torch.no_grad():
        out = cb(z)
    flat = z.permute(0, 2, 3, 4, 1).reshape(-1, 256)
    d = (flat ** 2).sum(1, keepdim=True) - 2 * flat @ cb.embeddings.t() + (cb.embeddings.t() ** 2).sum(0, keepdim=True)
    top2 = torch.topk(d, 2, dim=1, largest=False).values
    save("codebook", {}, z_seed=4, cb_seed=0, encodings=out["encodings"], commitment_loss=float(out["commitment_loss"]),
         perplexity=float(out["perplexity"]), emb_sub=out["embeddings"][:, ::9, :, ::3, ::5].contiguous(),
         gap=(top2[:, 1] - top2[:, 0]))


def gen_grads_dropout(cfg_name, B, wseed, dseed, t, p, mseed):
    """Training-mode (dropout p) loss and gradients of the reference with RECORDED dropout masks: nn.Dropout.forward is
    replaced by `x * keep / (1-p)` with keep drawn from a seeded generator, and every keep tensor is stored (bit-packed)
    under the name of the module that drew it, in call order.  Replaying them pins where the oracle applies dropout."""
    import torch.nn as nn
    import torch.nn.functional as F

    cfg = CONFIGS[cfg_name]
    model, _ = build_reference(cfg, "linear", wseed, pdrop=p)
    model.transformer.train()                     # the dropouts live in GPT; the mask sampler stays in eval mode
    names = {id(m): n for n, m in model.named_modules()}
    g = torch.Generator().manual_seed(mseed)
    record = []
    orig = nn.Dropout.forward

    def recorded(self, x):
        assert self.training and abs(self.p - p) < 1e-12
        keep = torch.rand(x.shape, generator=g) >= p
        record.append((names[id(self)], keep))
        return x * keep.to(x.dtype) / (1.0 - p)

    nn.Dropout.forward = recorded
    try:
        x, indices = synth_tokens(cfg, B, dseed)
        logits, target, NT_weight, seq_len = model(x, None, t=t, indices=indices)
        ratio = NT_weight / float(seq_len)
        ce = F.cross_entropy(logits.reshape(-1, logits.size(-1)), target.reshape(-1), reduction="sum")
        loss = ce / (B * seq_len * ratio ** model.config.avg_loss)
        loss.backward()
    finally:
        nn.Dropout.forward = orig
    arrays = dict(x=x, indices=indices, wseed=wseed, t=t, p=p, loss=float(loss), logits_sample=logits.detach().reshape(-1)[::997])
    arrays["mask_names"] = np.array([n for n, _ in record])
    for i, (n, keep) in enumerate(record):
        arrays[f"mask_shape:{i}"] = np.array(keep.shape)
        arrays[f"mask_bits:{i}"] = np.packbits(keep.numpy().reshape(-1))
    gnames, norms = [], []
    for n, prm in model.named_parameters():
        gnames.append(n)
        norms.append(float(prm.grad.norm()) if prm.grad is not None else -1.0)
        if prm.grad is not None and ("blocks.0." in n or "blocks.3." in n or "_emb" in n or "ln_f" in n):
            arrays["g:" + n] = prm.grad.reshape(-1)[::17].contiguous()
    arrays["grad_names"] = np.array(gnames)
    arrays["grad_norms"] = np.array(norms)
    save(f"grads_dropout_{cfg_name}", cfg, **arrays)


def gen_vqgan():
    """VQGAN encoder / decoder of the unmodified reference (mebt/vqgan.py) with oracle.vqgan_oracle.make_weights:
    pre-VQ latent of a random video and the decoded video of random code grids, group- and batch-norm variants."""
    import argparse
    try:                                   # mebt.modules.lpips imports torchvision / requests: stub what is missing
        import mebt.modules.lpips  # noqa: F401
    except Exception:
        for name in ("torchvision", "torchvision.models", "requests", "tqdm"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        if not hasattr(sys.modules["torchvision"], "models"):
            sys.modules["torchvision"].models = sys.modules["torchvision.models"]
    from mebt import vqgan as RV
    from oracle import vqgan_oracle as VO

    for tag, norm, ds in (("small", "group", (2, 4, 4)), ("bn", "batch", (4, 2, 2))):
        args = argparse.Namespace(embedding_dim=64, n_codes=128, n_hiddens=32, downsample=ds, image_channels=3, norm_type=norm,
                                  padding_type="replicate")
        enc = RV.Encoder(args.n_hiddens, args.downsample, args.image_channels, args.norm_type, args.padding_type)
        dec = RV.Decoder(args.n_hiddens, args.downsample, args.image_channels, args.norm_type)
        pre = RV.SamePadConv3d(enc.out_channels, args.embedding_dim, 1, padding_type=args.padding_type)
        post = RV.SamePadConv3d(args.embedding_dim, enc.out_channels, 1)
        mods = {"encoder": enc, "decoder": dec, "pre_vq_conv": pre, "post_vq_conv": post}
        shapes = {f"{n}.{k}": tuple(v.shape) for n, m in mods.items() for k, v in m.state_dict().items()}
        shapes["codebook.embeddings"] = (args.n_codes, args.embedding_dim)
        P = VO.make_weights(shapes, 7)
        for n, m in mods.items():
            m.load_state_dict({k[len(n) + 1:]: v for k, v in P.items() if k.startswith(n + ".")}, strict=True)
            m.eval()
        g = torch.Generator().manual_seed(11)
        x = torch.rand(1, 3, 8, 32, 32, generator=g) - 0.5
        lat = tuple(s // d for s, d in zip((8, 32, 32), ds))
        codes = torch.randint(0, args.n_codes, (1, *lat), generator=g)
        with torch.no_grad():
            z = pre(enc(x))
            h = torch.nn.functional.embedding(codes, P["codebook.embeddings"])
            rec = dec(post(RV.shift_dim(h, -1, 1)))
        cfg = dict(vars(args))
        np.savez_compressed(HERE / f"vqgan_{tag}.npz", cfg_json=json.dumps(cfg), wseed=7, x=x.numpy(), codes=codes.numpy(),
                            z=z.numpy(), rec=rec.numpy(), keys=np.array(sorted(shapes)),
                            shapes_json=json.dumps({k: list(v) for k, v in shapes.items()}))
        print("vqgan", tag, tuple(z.shape), tuple(rec.shape), float(z.abs().mean()), float(rec.abs().mean()))


def _reference_script_functions():
    """bidirect_sample / extrapolate exactly as written in the reference's sample_vqgan_transformer_videos.py: the two
    FunctionDefs are compiled out of the script's source (the module itself imports lightning, omegaconf, matplotlib...)."""
    import ast

    from einops import rearrange, repeat
    src = open(os.path.join(REF, "sample_vqgan_transformer_videos.py")).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("bidirect_sample", "extrapolate")]
    ns = dict(torch=torch, np=np, repeat=repeat, rearrange=rearrange)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "sample_vqgan_transformer_videos.py", "exec"), ns)
    return ns["bidirect_sample"], ns["extrapolate"]


def gen_pipelines():
    """Sliding-window pipelines of the sampling script driven by tests/helpers.py::FakeSampler (deterministic tokens,
    recorded calls): the reference's code maps, scores and the arguments of every model.sample call."""
    sys.path.insert(0, str(REPO / "tests"))
    from helpers import FakeSampler
    bidirect, extrapolate = _reference_script_functions()
    arrays = {}

    def record(tag, model, log):
        arrays[f"{tag}_code_maps"] = log["code_maps"]
        arrays[f"{tag}_samples"] = log["samples"]
        if "score" in log:
            arrays[f"{tag}_score"] = log["score"]
        arrays[f"{tag}_n_calls"] = len(model.calls)
        for i, c in enumerate(model.calls):
            arrays[f"{tag}_call{i}_x"], arrays[f"{tag}_call{i}_ctx"], arrays[f"{tag}_call{i}_tgt"] = c["x"], c["ctx"], c["tgt"]
            arrays[f"{tag}_call{i}_meta"] = np.array(json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()
                                                                 if k not in ("x", "ctx", "tgt")}))

    m = FakeSampler((4, 4, 4))
    record("bi_one", m, bidirect(m, 2, total_length=16, step_size=16, context_size=12, temperature=0.9, top_k=5,
                                 vid_n_steps=6, vid_c_temp=2.0, ctemp_schedule="cosine", strategy="maskgit", bootstrap=3))
    m = FakeSampler((4, 4, 4))
    record("bi_one_nb", m, bidirect(m, 3, total_length=16, step_size=16, context_size=8, vid_n_steps=4, strategy="random"))
    m = FakeSampler((4, 4, 4))
    torch.manual_seed(5)
    vq = torch.randint(0, FakeSampler.V, (2, 4, 4, 4))
    arrays["ex_input"] = vq
    record("ex", m, extrapolate(m, vq, total_length=40, step_size=16, context_size=8, temperature=0.8, top_p=0.9,
                                vid_n_steps=5, vid_c_temp=3.0))
    m = FakeSampler((4, 4, 4))
    record("ex_odd", m, extrapolate(m, vq, total_length=30, step_size=16, context_size=12, vid_n_steps=7))
    save("pipelines", dict(note="FakeSampler"), **arrays)


def main():
    install_stubs()
    sys.path.insert(0, REF)          # `import mebt` resolves to the reference; also its top-level utils.py
    sys.path.insert(1, str(REPO))    # oracle.*
    torch.set_num_threads(8)
    gen_maskgen()
    gen_sample_from_logits()
    gen_codebook()
    gen_forward("micro", 2, wseed=1, dseed=2, ncs=[0, 1, 100, 128, 255])
    gen_forward("tiny", 2, wseed=1, dseed=2, ncs=[0, 300, 512, 1023])
    gen_forward("tiny5", 2, wseed=3, dseed=4, ncs=[512])
    gen_shared_step("tiny", 2, wseed=1, dseed=2, ts=[0.5, 0.13, 0.9])
    gen_shared_step("micro", 3, wseed=5, dseed=6, ts=[0.5, 0.3], label_smoothing=0.1)
    gen_grads("tiny5", 2, wseed=3, dseed=4, t=0.5)
    gen_grads("micro", 2, wseed=1, dseed=2, t=0.4)
    gen_grads_dropout("micro", 2, wseed=1, dseed=2, t=0.4, p=0.1, mseed=11)
    gen_sampling("micro", 2, wseed=1, seed=9)
    gen_sampling("tiny", 2, wseed=1, seed=9)
    gen_entp("micro", 2, wseed=1, seed=13)
    gen_edit("micro", 2, wseed=1, seed=47)
    gen_pipelines()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vqgan":
        install_stubs()
        sys.path.insert(0, REF)
        sys.path.insert(1, str(REPO))
        gen_vqgan()
    elif len(sys.argv) > 1 and sys.argv[1] == "pipelines":
        sys.path.insert(1, str(REPO))
        gen_pipelines()
    elif len(sys.argv) > 1 and sys.argv[1] == "edit":        # only the fixture of the fixed-context / edit sampling loop
        install_stubs()
        sys.path.insert(0, REF)
        sys.path.insert(1, str(REPO))
        torch.set_num_threads(8)
        gen_edit("micro", 2, wseed=1, seed=47)
    elif len(sys.argv) > 1 and sys.argv[1] == "entp":        # only the fixture added in round 2
        install_stubs()
        sys.path.insert(0, REF)
        sys.path.insert(1, str(REPO))
        torch.set_num_threads(8)
        gen_entp("micro", 2, wseed=1, seed=13)
    elif len(sys.argv) > 1 and sys.argv[1] == "dropout":     # only the fixture added after the first generation
        install_stubs()
        sys.path.insert(0, REF)
        sys.path.insert(1, str(REPO))
        torch.set_num_threads(8)
        gen_grads_dropout("micro", 2, wseed=1, dseed=2, t=0.4, p=0.1, mseed=11)
    else:
        main()
