"""CPU oracle for the MeBT hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, module-free restatement (torch CPU fp32 + numpy) of what the reference computes on the
path SURVEY.md §8 scopes.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this file; nothing under `mebt_b200/` does, and the product path
fails loudly when its CUDA library is missing.

Parity status: PINNED.  `tests/golden/*.npz` were produced by importing the *unmodified* reference
from /root/reference (script: tests/golden/make_golden.py) and `tests/test_oracle_golden.py` checks
every function below against them (bit-exact for ids / masks / indices, ≤1e-5 abs for fp32 tensors).

Every function cites the reference lines (relative to /root/reference) it follows.  Parameters are
passed as a flat dict keyed by the reference's own state_dict names, e.g.
`transformer.blocks.3.attn.query.weight`.
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

LATENT_MODES = ("latent_enc", "latent_self", "latent_dec", "lt2l")


# ------------------------------------------------------------------------------------------------
# Synthetic weights (shared by oracle, golden generator and the CUDA-path tests)
# ------------------------------------------------------------------------------------------------
def param_shapes(cfg: dict) -> dict:
    """state_dict names/shapes of Net2NetTransformer (mebt/transformer.py:106-140, modules/gpt.py:98-221)."""
    D, V, L, N = cfg["n_embd"], cfg["vocab_size"], cfg["sos_emb"], cfg["block_size"]
    shapes = {"mask_emb": (1, 1, D), "sos_emb": (1, L, D), "pos_emb": (1, N, D), "tok_emb.weight": (V, D)}
    for i in range(cfg["n_layer"]):
        p = f"transformer.blocks.{i}."
        for ln in ("ln1", "ln2"):
            shapes[p + ln + ".weight"] = (D,)
            shapes[p + ln + ".bias"] = (D,)
        for lin in ("key", "query", "value", "proj"):
            shapes[p + f"attn.{lin}.weight"] = (D, D)
            shapes[p + f"attn.{lin}.bias"] = (D,)
        shapes[p + "mlp.0.weight"] = (4 * D, D)
        shapes[p + "mlp.0.bias"] = (4 * D,)
        shapes[p + "mlp.2.weight"] = (D, 4 * D)
        shapes[p + "mlp.2.bias"] = (D,)
    shapes["transformer.ln_f.weight"] = (D,)
    shapes["transformer.ln_f.bias"] = (D,)
    shapes["transformer.head.weight"] = (V, D)
    return shapes


def make_weights(cfg: dict, seed: int = 0, reference_init: bool = False) -> dict:
    """Deterministic per-tensor weights, independent of module construction order.

    reference_init=True reproduces the *distribution* of the reference init (gpt.py:225-232,
    transformer.py:126-140: N(0, 0.02) weights/embeddings, zero biases, unit LayerNorm).  The default
    perturbs biases and LayerNorm parameters too so that parity tests exercise them.
    """
    out = {}
    for name, shape in param_shapes(cfg).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2**63))
        is_ln = ".ln1." in name or ".ln2." in name or ".ln_f." in name
        if is_ln and name.endswith("weight"):
            t = torch.ones(shape) if reference_init else 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = torch.zeros(shape) if reference_init else (0.05 if is_ln else 0.02) * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.float()
    return out


# ------------------------------------------------------------------------------------------------
# Layer stack
# ------------------------------------------------------------------------------------------------
def layer_norm(x, w, b):
    """nn.LayerNorm(D), eps 1e-5 (gpt.py:147-148,216)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def cross_attention(P: dict, pre: str, q_in, k_in, n_head: int, attn_keep=None, resid_keep=None):
    """CrossAttention.forward, gpt.py:119-141 (attn_bias = 0.).  Dropout: eval mode / p = 0 when the keep tensors are
    None; otherwise `attn_keep` [B,h,NQ,NK] and `resid_keep` [B,NQ,C] hold nn.Dropout's factors (0 or 1/(1-p)) for
    attn_drop (gpt.py:136) and resid_drop (gpt.py:140), so that a recorded or replayed mask gives the training-mode result."""
    B, NQ, C = q_in.shape
    NK = k_in.shape[1]
    hs = C // n_head
    k = F.linear(k_in, P[pre + "key.weight"], P[pre + "key.bias"]).view(B, NK, n_head, hs).transpose(1, 2)
    q = F.linear(q_in, P[pre + "query.weight"], P[pre + "query.bias"]).view(B, NQ, n_head, hs).transpose(1, 2)
    v = F.linear(k_in, P[pre + "value.weight"], P[pre + "value.bias"]).view(B, NK, n_head, hs).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hs))
    att = F.softmax(att, dim=-1)          # NK == 0 -> empty softmax -> y == 0
    if attn_keep is not None:
        att = att * attn_keep
    y = (att @ v).transpose(1, 2).contiguous().view(B, NQ, C)
    y = F.linear(y, P[pre + "proj.weight"], P[pre + "proj.bias"])
    if resid_keep is not None:
        y = y * resid_keep
    return y


def block_forward(P: dict, i: int, mode: str, n_head: int, lat, ctx, tgt, drop=None):
    """Block.forward, gpt.py:159-195: shared ln1 on query and key; residual on the NORMALISED query.
    `drop`: optional dict of dropout keep factors {(i,"attn"), (i,"proj"), (i,"mlp")} (training mode, gpt.py:136,140,154)."""
    drop = drop or {}
    pre = f"transformer.blocks.{i}."
    if mode == "latent_self":
        q, k = lat, lat
    elif mode == "latent_enc":
        q, k = lat, ctx
    elif mode == "latent_dec":
        q, k = tgt, lat
    elif mode == "lt2l":
        q, k = lat, torch.cat([lat, tgt], 1)
    elif mode == "maskgit":
        q = torch.cat([ctx, tgt], 1)
        k = q
    else:
        raise ValueError(mode)
    qn = layer_norm(q, P[pre + "ln1.weight"], P[pre + "ln1.bias"])
    kn = layer_norm(k, P[pre + "ln1.weight"], P[pre + "ln1.bias"])
    x = qn + cross_attention(P, pre + "attn.", qn, kn, n_head, drop.get((i, "attn")), drop.get((i, "proj")))
    h = layer_norm(x, P[pre + "ln2.weight"], P[pre + "ln2.bias"])
    h = F.linear(h, P[pre + "mlp.0.weight"], P[pre + "mlp.0.bias"])
    h = F.gelu(h)                          # nn.GELU() = exact erf form (gpt.py:152)
    h = F.linear(h, P[pre + "mlp.2.weight"], P[pre + "mlp.2.bias"])
    if (i, "mlp") in drop:
        h = h * drop[(i, "mlp")]
    x = x + h
    if mode in ("latent_enc", "latent_self", "lt2l"):
        lat = x
    elif mode == "latent_dec":
        tgt = x
    else:
        NC = ctx.shape[1]
        ctx, tgt = x[:, :NC], x[:, NC:]
    return lat, ctx, tgt


def stack_modes(cfg: dict) -> list:
    """GPT.__init__ pads a short mode list with 'maskgit' (gpt.py:208-209)."""
    modes = list(cfg["mode"])
    return modes + ["maskgit"] * (cfg["n_layer"] - len(modes))


def stem(P: dict, cfg: dict, x_indices, ctx_idx, tgt_idx):
    """Embedding stem of forward/reconstruct_mask (transformer.py:255-277 / :298-317)."""
    B = x_indices.shape[0]
    z_ctx = torch.gather(x_indices, 1, ctx_idx)
    NT = tgt_idx.shape[1]
    pos = P["pos_emb"][0]
    ctx = P["tok_emb.weight"][z_ctx] + pos[ctx_idx]
    tgt = P["mask_emb"].expand(B, NT, -1) + pos[tgt_idx]
    lat = P["sos_emb"].expand(B, -1, -1)
    return lat, ctx, tgt


def gpt_forward(P: dict, cfg: dict, lat, ctx, tgt, return_hidden: bool = False, drop=None):
    """GPT.forward, gpt.py:234-253.  Eval mode / p=0 when `drop` is None (the stem dropouts are identity); otherwise
    `drop` maps ("stem","lat"|"ctx"|"tgt") and (i,"attn"|"proj"|"mlp") to keep factors (gpt.py:239-241 and the blocks)."""
    if drop:
        lat = lat * drop[("stem", "lat")] if ("stem", "lat") in drop else lat
        ctx = ctx * drop[("stem", "ctx")] if ("stem", "ctx") in drop else ctx
        tgt = tgt * drop[("stem", "tgt")] if ("stem", "tgt") in drop else tgt
    for i, mode in enumerate(stack_modes(cfg)):
        lat, ctx, tgt = block_forward(P, i, mode, cfg["n_head"], lat, ctx, tgt, drop)
    x = layer_norm(tgt, P["transformer.ln_f.weight"], P["transformer.ln_f.bias"])
    logits = F.linear(x, P["transformer.head.weight"])
    if return_hidden:
        return logits, lat, tgt
    return logits


def reconstruct_mask(P: dict, cfg: dict, x_indices, ctx_idx, tgt_idx, drop=None):
    """Net2NetTransformer.reconstruct_mask, transformer.py:288-324 -> logits [B,NT,V] fp32."""
    B = x_indices.shape[0]
    x_indices = x_indices.reshape(B, -1)
    lat, ctx, tgt = stem(P, cfg, x_indices, ctx_idx, tgt_idx)
    return gpt_forward(P, cfg, lat, ctx, tgt, drop=drop)


# ------------------------------------------------------------------------------------------------
# Mask bookkeeping (MaskGen)
# ------------------------------------------------------------------------------------------------
def schedule(name: str, t: torch.Tensor) -> torch.Tensor:
    """MaskGen schedules, mask_sampler.py:34-67.  `t` is a float32 torch tensor: the reference evaluates
    these in float32 torch ops (np.pi promotes to a python float, the tensor stays float32)."""
    if name == "cosine":
        return torch.cos(0.5 * np.pi * t)
    if name == "cosine_plus":
        return 0.5 * (1 + torch.cos(np.pi * t))
    if name == "linear":
        return 1.0 - t
    if name == "quadratic":
        return (1.0 - t) ** 2.0
    if name == "square":
        return 1.0 - t ** 2.0
    if name == "cube":
        return 1.0 - t ** 3.0
    if name == "sqrt":
        return 1.0 - t ** 0.5
    if name == "convex":
        return (1.0 - t) ** 3.0
    raise ValueError(name)


def divide_indices_eval(indices: torch.Tensor, t: float, schedule_name: str):
    """MaskGen.divide_indices in eval mode (mask_sampler.py:75-115 with `self.training` False)."""
    ratio = schedule(schedule_name, torch.tensor(t))
    seq_len = int(np.prod(indices.shape[1:]))
    n_masked = int(torch.ceil(ratio * seq_len).to(torch.long))
    n_ctx = seq_len - n_masked
    n_tgt = min(seq_len, seq_len - n_ctx)
    return indices[:, :n_ctx], indices[:, -n_tgt:], seq_len


def divide_indices_train(indices, t: float, schedule_name: str, shape, budget: int, T: int, start_t: int):
    """Training branch of divide_indices (mask_sampler.py:83-114) with the two numpy draws (frame count T,
    window start) supplied by the caller."""
    ratio = schedule(schedule_name, torch.tensor(t))
    max_T = shape[0]
    num_pos = int(np.prod(shape[1:]))
    if max_T != T:
        lo, hi = start_t * num_pos, (start_t + T) * num_pos
        indices = torch.stack([row[(row >= lo) & (row < hi)] for row in indices])
    seq_len = int(np.prod(indices.shape[1:]))
    n_masked = int(torch.ceil(ratio * seq_len).to(torch.long))
    n_ctx = seq_len - n_masked
    n_tgt = min(budget, seq_len - n_ctx)
    return indices[:, :n_ctx], indices[:, -n_tgt:], seq_len


def gibbs_draft_mask(ctx_idx, tgt_idx, n_steps: int, perms):
    """create_gibbs_draft_mask, mask_sampler.py:338-356; `perms` = the B randperm(N) draws, stacked."""
    N = tgt_idx.shape[1]
    assert N % n_steps == 0
    m = N // n_steps
    shuffled = torch.gather(tgt_idx, 1, perms)
    ctxs = [torch.cat([ctx_idx, shuffled[:, : i * m]], 1) for i in range(n_steps)]
    tgts = [shuffled[:, i * m:] for i in range(n_steps)]
    return ctxs, tgts


def gibbs_revise_mask(ctx_idx, tgt_idx, n_steps: int, perms):
    """create_gibbs_revise_mask, mask_sampler.py:317-336."""
    N = tgt_idx.shape[1]
    assert N % n_steps == 0
    m = N // n_steps
    shuffled = torch.gather(tgt_idx, 1, perms)
    ctxs = [torch.cat([ctx_idx, shuffled[:, (i + 1) * m:], shuffled[:, : i * m]], 1) for i in range(n_steps)]
    tgts = [shuffled[:, i * m:(i + 1) * m] for i in range(n_steps)]
    return ctxs, tgts


def remask_order(score: torch.Tensor, ctemp: float, q: torch.Tensor) -> torch.Tensor:
    """MaskGen.gumbel_top_k, mask_sampler.py:178-187, with the Exp(1) draw `q` supplied."""
    prob = score / score.sum(-1, keepdim=True)
    prob = prob / (q ** ctemp)
    return prob.sort(dim=-1, descending=True)[1]


def generate_next_mask(ctx_idx, tgt_idx, score, n_masked: int, ctemp: float, q, strategy="maskgit", randn=None):
    """MaskGen.generate_next_mask, mask_sampler.py:189-236 (maskgit / random / bootstrap strategies)."""
    B, NC = ctx_idx.shape
    NT = tgt_idx.shape[1]
    if strategy in ("random", "bootstrap"):
        score = randn
        ctemp = 0.0
    seq_len = NC + NT
    if strategy == "bootstrap":
        n_masked = NT - 1
    n_ctx = seq_len - n_masked
    if n_ctx <= NC:
        return ctx_idx, tgt_idx
    n_new = n_ctx - NC
    order = remask_order(score, ctemp, q)
    next_ctx = torch.cat([ctx_idx, torch.gather(tgt_idx, -1, order[:, :n_new])], 1)
    next_tgt = torch.gather(tgt_idx, -1, order[:, n_new:])
    return next_ctx, next_tgt


# ------------------------------------------------------------------------------------------------
# Logit head consumers
# ------------------------------------------------------------------------------------------------
def top_k_filter(logits, k: int):
    """top_k_logits, transformer.py:891-895."""
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[..., [-1]]] = -float("inf")
    return out


def top_p_filter(probs, p: float):
    """top_p_probs, transformer.py:898-910."""
    sp, si = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(sp, dim=-1)
    rm = cum >= p
    rm[..., 1:] = rm[..., :-1].clone()
    rm[..., 0] = 0
    remove = rm.scatter(-1, si, rm)
    probs = probs.masked_fill(remove, 0.0)
    return probs / probs.sum(-1, keepdim=True)


def sample_from_logits(logits, temperature, top_k, top_p, q):
    """sample_from_logits + gumbel_sort, transformer.py:843-889 / :826-841, with the Exp(1) noise `q`
    (same shape as logits) supplied.  Returns (ids int64, probs fp32) — probs are the softmax BEFORE the
    Gumbel renormalisation, as the reference returns them."""
    logits = logits.to(torch.float32) / (temperature + 1e-8)
    if top_k is not None:
        logits = top_k_filter(logits, top_k)
    logits = torch.where(torch.isnan(logits), torch.full_like(logits, -float("inf")), logits)
    probs = F.softmax(logits, dim=-1)
    if top_p is not None:
        probs = top_p_filter(probs, top_p)
    pr = probs / probs.sum(-1, keepdim=True)
    race = (pr / q) * (pr > 0).float()
    ids = race.sort(dim=-1, descending=True)[1][..., 0]
    return ids, probs


def masked_ce(logits, targets, label_smoothing: float = 0.0):
    """Loss/metrics core of shared_step (transformer.py:726,731; mebt/utils.py:80-94).
    Returns (ce_sum, n_top1, n_top5)."""
    V = logits.shape[-1]
    lg = logits.reshape(-1, V)
    tg = targets.reshape(-1)
    ce = F.cross_entropy(lg, tg, reduction="sum", label_smoothing=label_smoothing)
    _, pred = lg.topk(5, 1, True, True)
    hit = pred.eq(tg.view(-1, 1))
    return ce, int(hit[:, :1].sum()), int(hit.sum())


def shared_step(P, cfg, x_indices, indices, t: float, schedule_name: str, label_smoothing=0.0, drop=None):
    """forward + shared_step for a full-length clip (T == max_T, budget >= N), transformer.py:216-286,
    :717-732.  Returns dict(loss, acc1, acc5, ce_sum, logits, z_targets, ratio)."""
    B = x_indices.shape[0]
    x_indices = x_indices.reshape(B, -1)
    ctx_idx, tgt_idx, seq_len = divide_indices_eval(indices, t, schedule_name)
    budget = cfg.get("budget", seq_len)
    n_tgt = min(budget, tgt_idx.shape[1])
    tgt_idx = tgt_idx[:, -n_tgt:] if n_tgt > 0 else tgt_idx
    z_t = torch.gather(x_indices, 1, tgt_idx)
    logits = reconstruct_mask(P, cfg, x_indices, ctx_idx, tgt_idx, drop=drop)
    NT_weight = float(seq_len - ctx_idx.shape[1])
    ratio = NT_weight / float(seq_len)
    ce, n1, n5 = masked_ce(logits, z_t, label_smoothing)
    weight = ratio ** float(cfg.get("avg_loss", 0.0))
    loss = ce / (B * seq_len * weight)
    n = z_t.numel()
    return dict(loss=loss, acc1=100.0 * n1 / n, acc5=100.0 * n5 / n, ce_sum=ce, logits=logits, z_targets=z_t,
                ratio=ratio, context_indices=ctx_idx, target_indices=tgt_idx)


# ------------------------------------------------------------------------------------------------
# Samplers.  `rng` supplies the reference's draws in the reference's order:
#   rng.randperm(n) -> int64 [n]   (CPU generator; mask_sampler.py:331,351)
#   rng.exponential(shape) -> fp32 (one exponential_ per sample_from_logits / gumbel_top_k call)
#   rng.randn(shape)
# ------------------------------------------------------------------------------------------------
class TorchRng:
    """Draws from torch's global CPU generator exactly as the reference does on CPU."""

    def __init__(self, seed: int):
        torch.manual_seed(seed)

    def randperm(self, n):
        return torch.randperm(n)

    def exponential(self, shape):
        return torch.empty(shape, dtype=torch.float32).exponential_()

    def randn(self, shape):
        return torch.randn(shape)


def _write_back(x, tgt_idx, ids):
    """The sparse-COO write-back of transformer.py:413-439 / :571-585 is a scatter."""
    return x.scatter(1, tgt_idx, ids)


def draft(P, cfg, x, temperature, top_k, top_p, n_steps, rng, ctx_idx=None, tgt_idx=None):
    """Net2NetTransformer.draft, transformer.py:544-586."""
    B = x.shape[0]
    x = x.reshape(B, -1)
    N = x.shape[1]
    if ctx_idx is None:
        ctx_idx = torch.empty(B, 0, dtype=torch.long)
        tgt_idx = torch.arange(N).repeat(B, 1)
    perms = torch.stack([rng.randperm(tgt_idx.shape[1]) for _ in range(B)])
    ctxs, tgts = gibbs_draft_mask(ctx_idx, tgt_idx, n_steps, perms)
    for c, t in zip(ctxs, tgts):
        logits = reconstruct_mask(P, cfg, x, c, t)
        ids, _ = sample_from_logits(logits, temperature, top_k, top_p, rng.exponential(logits.shape))
        x = _write_back(x, t, ids)
    return x


def revise(P, cfg, x, temperature, top_k, top_p, n_steps, rng, ctx_idx=None, tgt_idx=None):
    """Net2NetTransformer.revise, transformer.py:588-630."""
    B = x.shape[0]
    x = x.reshape(B, -1)
    N = x.shape[1]
    if ctx_idx is None:
        ctx_idx = torch.empty(B, 0, dtype=torch.long)
        tgt_idx = torch.arange(N).repeat(B, 1)
    perms = torch.stack([rng.randperm(tgt_idx.shape[1]) for _ in range(B)])
    ctxs, tgts = gibbs_revise_mask(ctx_idx, tgt_idx, n_steps, perms)
    for c, t in zip(ctxs, tgts):
        logits = reconstruct_mask(P, cfg, x, c, t)
        ids, _ = sample_from_logits(logits, temperature, top_k, top_p, rng.exponential(logits.shape))
        x = _write_back(x, t, ids)
    return x


def draft_and_revise(P, cfg, x, rng, n_draft=8, draft_t=1.0, draft_k=None, draft_p=None, n_revise=8, revise_t=1.0,
                     revise_k=None, revise_p=None, M=2, skip_draft=False):
    """Net2NetTransformer.draft_and_revise, transformer.py:632-663."""
    B = x.shape[0]
    x = x.reshape(B, -1)
    if not skip_draft:
        x = draft(P, cfg, x, draft_t, draft_k, draft_p, n_draft, rng)
    for _ in range(M):
        x = revise(P, cfg, x, revise_t, revise_k, revise_p, n_revise, rng)
    return x


def sample_maskgit(P, cfg, x, rng, temperature=1.0, top_k=None, top_p=None, n_steps=8, strategy="maskgit",
                   context_temperature=4.5, schedule_name="cosine", context_indices=None, target_indices=None, edit=False):
    """Net2NetTransformer.sample, transformer.py:353-447 (ctemp_schedule='linear').  With `context_indices` /
    `target_indices` the loop starts from that split (:387-389); `edit=True` sizes the mask schedule by the number of
    TARGETS instead of the sequence length (:373-376,399) - the sliding-window extrapolation of
    sample_vqgan_transformer_videos.py:95-157 re-samples only the new frames."""
    B = x.shape[0]
    x = x.reshape(B, -1)
    N = x.shape[1]
    if context_indices is None:
        ctx_idx = torch.empty(B, 0, dtype=torch.long)
        tgt_idx = torch.arange(N).repeat(B, 1)
    else:
        ctx_idx, tgt_idx = context_indices.clone(), target_indices.clone()
    edit_N = tgt_idx.shape[1] if edit else N
    for t_next in np.linspace(0, 1, n_steps + 1)[1:]:
        t = torch.full((B,), fill_value=t_next)                      # float32 (transformer.py:398)
        n_masked_t = torch.ceil(schedule(schedule_name, t) * edit_N)
        if int((n_masked_t > tgt_idx.shape[-1]).sum()) == B:
            continue
        logits = reconstruct_mask(P, cfg, x, ctx_idx, tgt_idx)
        ids, probs = sample_from_logits(logits, temperature, top_k, top_p, rng.exponential(logits.shape))
        scores = probs.gather(-1, ids.unsqueeze(-1)).squeeze(-1)
        x = _write_back(x, tgt_idx, ids)
        ctemp = context_temperature * (1.0 - t_next)
        n_masked = int(n_masked_t[0].long())
        NC, NT = ctx_idx.shape[1], tgt_idx.shape[1]
        if strategy == "bootstrap":
            n_masked = NT - 1
        randn = rng.randn(scores.shape) if strategy in ("random", "bootstrap") else None  # mask_sampler.py:206-208
        if NC + NT - n_masked <= NC:
            continue                                                # no Exp draw (mask_sampler.py:222-225)
        q = rng.exponential(scores.shape)
        ctx_idx, tgt_idx = generate_next_mask(ctx_idx, tgt_idx, scores, n_masked, ctemp, q, strategy, randn)
    return x, ctx_idx, tgt_idx


def entropy_scores(probs: torch.Tensor) -> torch.Tensor:
    """The reveal score of entp_sample, transformer.py:499-500: sum_v (p - log(p + 1e-8)), subtracted from its row
    maximum (so the most peaked distribution scores highest... and the flattest one exactly 0)."""
    s = -(-probs + torch.log(probs + 1e-8)).sum(-1)
    return s.max(-1, keepdim=True)[0] - s


def sample_entp(P, cfg, x, rng, temperature=1.0, top_k=None, top_p=None, n_steps=8, strategy="maskgit",
                schedule_name="cosine"):
    """Net2NetTransformer.entp_sample, transformer.py:449-542, with MaskGen.generate_next_mask_entp
    (mask_sampler.py:248-303): like `sample`, but tokens are revealed in the order of `entropy_scores`, the re-masking
    temperature is 0 (the Exp(1) draw is still consumed), the mask size is re-derived from the schedule inside
    generate_next_mask_entp (no n_masked_toks argument), and only 'random' replaces the scores by randn ('bootstrap'
    keeps them and reveals one token per step)."""
    B = x.shape[0]
    x = x.reshape(B, -1)
    N = x.shape[1]
    ctx_idx = torch.empty(B, 0, dtype=torch.long)
    tgt_idx = torch.arange(N).repeat(B, 1)
    for t_next in np.linspace(0, 1, n_steps + 1)[1:]:
        t = torch.full((B,), fill_value=t_next)
        n_masked_t = torch.ceil(schedule(schedule_name, t) * N)
        if int((n_masked_t > tgt_idx.shape[-1]).sum()) == B:
            continue
        logits = reconstruct_mask(P, cfg, x, ctx_idx, tgt_idx)
        ids, probs = sample_from_logits(logits, temperature, top_k, top_p, rng.exponential(logits.shape))
        scores = entropy_scores(probs)
        x = _write_back(x, tgt_idx, ids)
        NC, NT = ctx_idx.shape[1], tgt_idx.shape[1]
        randn = rng.randn(scores.shape) if strategy == "random" else None        # mask_sampler.py:267-269
        n_masked = int(torch.ceil(schedule(schedule_name, t)[0] * (NC + NT)).to(torch.long))   # :273-275
        if strategy == "bootstrap":
            n_masked = NT - 1
        if NC + NT - n_masked <= NC:
            continue
        q = rng.exponential(scores.shape)
        ctx_idx, tgt_idx = generate_next_mask(ctx_idx, tgt_idx, scores, n_masked, 0.0, q,
                                              "random" if strategy == "random" else "maskgit", randn)
    return x, ctx_idx, tgt_idx


# ------------------------------------------------------------------------------------------------
# VQGAN codebook (eval path)
# ------------------------------------------------------------------------------------------------
def codebook_quantise(z: torch.Tensor, E: torch.Tensor):
    """Codebook.forward eval path, modules/codebook.py:48-62,91-97.  z [b,c,t,h,w] fp32, E [n_codes,c].
    Returns dict(encodings int64 [b,t,h,w], embeddings [b,c,t,h,w], commitment_loss, perplexity)."""
    b, c = z.shape[:2]
    flat = z.permute(0, 2, 3, 4, 1).contiguous().flatten(end_dim=-2)
    d = (flat ** 2).sum(1, keepdim=True) - 2 * flat @ E.t() + (E.t() ** 2).sum(0, keepdim=True)
    idx = torch.argmin(d, dim=1)
    enc = idx.view(b, *z.shape[2:])
    emb = F.embedding(enc, E).permute(0, 4, 1, 2, 3).contiguous()
    onehot_mean = torch.bincount(idx, minlength=E.shape[0]).float() / idx.numel()
    perplexity = torch.exp(-torch.sum(onehot_mean * torch.log(onehot_mean + 1e-10)))
    emb_st = (emb - z) + z                 # straight-through estimator value (codebook.py:91), rounding included
    return dict(encodings=enc, embeddings=emb_st, embeddings_raw=emb, commitment_loss=0.25 * F.mse_loss(z, emb),
                perplexity=perplexity, distances=d)


def codebook_decode_gather(enc: torch.Tensor, E: torch.Tensor):
    """VQGAN.decode's lookup, mebt/vqgan.py:91-92: F.embedding then channel-first."""
    return F.embedding(enc, E).permute(0, 4, 1, 2, 3).contiguous()


# ------------------------------------------------------------------------------------------------
# Work model
# ------------------------------------------------------------------------------------------------
def forward_flops(cfg: dict, NC: int, NT: int) -> float:
    """Forward FLOPs per sample: sum_blocks(20 NQ D^2 + 4 NK D^2 + 4 NQ NK D) + 2 NT D V."""
    D, L, V = cfg["n_embd"], cfg["sos_emb"], cfg["vocab_size"]
    f = 0.0
    for mode in stack_modes(cfg):
        if mode == "latent_enc":
            nq, nk = L, NC
        elif mode == "latent_self":
            nq, nk = L, L
        elif mode == "latent_dec":
            nq, nk = NT, L
        elif mode == "lt2l":
            nq, nk = L, L + NT
        else:
            nq = nk = NC + NT
        f += 20.0 * nq * D * D + 4.0 * nk * D * D + 4.0 * nq * nk * D
    return f + 2.0 * NT * D * V
