"""Drop-in for `mebt.modules.gpt` (reference: mebt/modules/gpt.py): GPT / Block / CrossAttention with the
same constructor signatures, parameter names and shapes (so reference checkpoints load), whose forward
passes run on the mebt_b200 CUDA kernels.

Modules keep their parameters in ordinary `nn.Linear` / `nn.LayerNorm` containers (fp32 masters); the
bf16 tensor-core operands are derived from them lazily and re-derived whenever a parameter changes.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops
from ..stack import (LayerWeights, PreciseWeightPack, WeightPack, attention_core, block_forward, stack_forward,
                     stack_forward_sample,
                     stack_forward_f32)


class GPTConfig:
    """Base GPT config (gpt.py:78-88)."""
    embd_pdrop = 0.1
    resid_pdrop = 0.1
    attn_pdrop = 0.1

    def __init__(self, vocab_size, block_size, **kwargs):
        self.vocab_size = vocab_size
        self.block_size = block_size
        for k, v in kwargs.items():
            setattr(self, k, v)


class GPT1Config(GPTConfig):
    n_layer = 12
    n_head = 12
    n_embd = 768


def _as_rows(t: torch.Tensor) -> torch.Tensor:
    """[B, n, D] (any float dtype) -> a fresh contiguous bf16 [B*n, D] buffer (the engine updates streams in place)."""
    B, n, D = t.shape
    return t.detach().reshape(B * n, D).to(torch.bfloat16, copy=True).contiguous()


def _versions(module: nn.Module):
    """Cache key of the operands derived from a module's parameters: storage + torch's version counters + the epoch of
    raw-pointer writers (FlatAdamW updates the masters in place without touching `_version`)."""
    return (_lib.write_epoch(),) + tuple((p.data_ptr(), p._version) for p in module.parameters())


def _dropout_seed(module: nn.Module, *ps):
    """None in eval mode / with p = 0; else a fresh 62-bit seed for the counter-based masks of this call, drawn from
    torch's CPU generator (so `torch.manual_seed` pins the masks the way it pins nn.Dropout's)."""
    if not module.training or not any(p > 0 for p in ps):
        return None
    return int(torch.randint(0, 2 ** 62, (1,)).item())


# stem sites of GPT.forward's embd dropout (sos_emb, contexts, targets; gpt.py:239-241) - the training engine's values
STEM_SITES = {"lat": 1 << 20, "ctx": (1 << 20) + 1, "tgt": (1 << 20) + 2}


class CrossAttention(nn.Module):
    """Multi-head attention with separate query/key/value/proj Linear(D, D) (gpt.py:91-141)."""

    def __init__(self, config):
        super().__init__()
        assert config.n_embd % config.n_head == 0
        self.key = nn.Linear(config.n_embd, config.n_embd)
        self.query = nn.Linear(config.n_embd, config.n_embd)
        self.value = nn.Linear(config.n_embd, config.n_embd)
        self.attn_drop = nn.Dropout(config.attn_pdrop)
        self.resid_drop = nn.Dropout(config.resid_pdrop)
        self.proj = nn.Linear(config.n_embd, config.n_embd)
        self.n_head = config.n_head
        self._cache = None

    def _weights(self):
        key = _versions(self)
        if self._cache is None or self._cache[0] != key:
            w_qkv = ops.cast_bf16(torch.cat([self.query.weight, self.key.weight, self.value.weight], 0).detach().float())
            b_qkv = torch.cat([self.query.bias, self.key.bias, self.value.bias]).detach().float().contiguous()
            w_proj = ops.cast_bf16(self.proj.weight.detach().float())
            self._cache = (key, w_qkv, b_qkv, w_proj, self.proj.bias.detach().float().contiguous())
        return self._cache[1:]

    def forward(self, query, key, attn_bias, context_size=0, mode="none"):
        """query [B,NQ,D], key [B,NK,D] -> (y [B,NQ,D], None, None, None); attn_bias must be 0 (it always is on the
        reference path, transformer.py:281,321)."""
        if torch.is_tensor(attn_bias) or attn_bias not in (0, 0.0, None):
            raise NotImplementedError("mebt_b200: a non-zero attn_bias is dead code in the reference and unsupported")
        seed = _dropout_seed(self, self.attn_drop.p, self.resid_drop.p)
        B, NQ, D = query.shape
        NK = key.shape[1]
        w_qkv, b_qkv, w_proj, b_proj = self._weights()
        lw = LayerWeights("none", None, None, None, None, w_qkv, b_qkv, w_proj, b_proj, None, None, None, None)
        att = attention_core(lw, self.n_head, B, _as_rows(query), _as_rows(key) if NK > 0 else None, NK,
                             attn_p=self.attn_drop.p if seed is not None else 0.0, attn_seed=seed or 0)
        y = ops.gemm(att, w_proj, b_proj)
        if seed is not None and self.resid_drop.p > 0:           # attn_drop on the probabilities, resid_drop on the output
            ops.dropout_rows_(y, self.resid_drop.p, seed, 1)
        return y.view(B, NQ, D).to(query.dtype), None, None, None


class Block(nn.Module):
    """Transformer block with a `mode`-selected (query, key) pair (gpt.py:143-195)."""

    def __init__(self, config, mode):
        super().__init__()
        self.ln1 = nn.LayerNorm(config.n_embd)
        self.ln2 = nn.LayerNorm(config.n_embd)
        self.attn = CrossAttention(config)
        self.mlp = nn.Sequential(
            nn.Linear(config.n_embd, 4 * config.n_embd),
            nn.GELU(),
            nn.Linear(4 * config.n_embd, config.n_embd),
            nn.Dropout(config.resid_pdrop),
        )
        self.mode = mode
        self._cache = None

    def layer_weights(self) -> LayerWeights:
        key = _versions(self)
        if self._cache is None or self._cache[0] != key:
            w_qkv, b_qkv, w_proj, b_proj = self.attn._weights()
            f32 = lambda t: t.detach().float().contiguous()
            self._cache = (key, LayerWeights(
                self.mode, f32(self.ln1.weight), f32(self.ln1.bias), f32(self.ln2.weight), f32(self.ln2.bias),
                w_qkv, b_qkv, w_proj, b_proj,
                ops.cast_bf16(f32(self.mlp[0].weight)), f32(self.mlp[0].bias),
                ops.cast_bf16(f32(self.mlp[2].weight)), f32(self.mlp[2].bias)))
        lw = self._cache[1]
        lw.mode = self.mode
        return lw

    def forward(self, sos_emb, contexts, targets, mask_emb=None, attn_bias=None):
        seed = _dropout_seed(self, self.attn.attn_drop.p, self.attn.resid_drop.p, self.mlp[3].p)
        drop = (self.attn.attn_drop.p, self.attn.resid_drop.p, seed, 0) if seed is not None else None
        B, NS, C = sos_emb.size()
        NC, NT = contexts.size(1), targets.size(1)
        lat, ctx, tgt = block_forward(self.layer_weights(), self.attn.n_head, B, _as_rows(sos_emb), _as_rows(contexts),
                                      _as_rows(targets), drop=drop)
        dt = sos_emb.dtype
        return (lat.view(B, NS, C).to(dt), ctx.view(B, NC, C).to(dt), tgt.view(B, NT, C).to(dt), attn_bias, None)


class GPT(nn.Module):
    """The MeBT layer stack: `mode`-driven Blocks, ln_f, and the vocabulary head on targets (gpt.py:198-253)."""

    def __init__(self, vocab_size, block_size, n_layer=12, n_head=8, n_embd=256, embd_pdrop=0., resid_pdrop=0.,
                 attn_pdrop=0., n_unmasked=0, vtokens_pos=False, mode=[]):
        super().__init__()
        config = GPTConfig(vocab_size=vocab_size, block_size=block_size, embd_pdrop=embd_pdrop, resid_pdrop=resid_pdrop,
                           attn_pdrop=attn_pdrop, n_layer=n_layer, n_head=n_head, n_embd=n_embd, n_unmasked=n_unmasked,
                           mode=list(mode))
        if len(config.mode) < n_layer:                     # short lists are padded with full-attention blocks
            config.mode += ["maskgit"] * (n_layer - len(config.mode))
        self.drop = nn.Dropout(config.embd_pdrop)
        assert config.n_layer == len(config.mode)
        self.blocks = nn.Sequential(*[Block(config, m) for m in config.mode])
        self.ln_f = nn.LayerNorm(config.n_embd)
        self.head = nn.Linear(config.n_embd, config.vocab_size, bias=False)
        self.block_size = config.block_size
        self.apply(self._init_weights)
        self.config = config
        self._pack = None

    def get_block_size(self):
        return self.block_size

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def weight_pack(self) -> WeightPack:
        key = _versions(self)
        if self._pack is None or self._pack[0] != key:
            params = {"transformer." + n: p for n, p in self.named_parameters()}
            self._pack = (key, WeightPack(params, [b.mode for b in self.blocks], self.config.n_head))
        return self._pack[1]

    def precise_pack(self) -> PreciseWeightPack:
        key = _versions(self)
        if getattr(self, "_ppack", None) is None or self._ppack[0] != key:
            params = {"transformer." + n: p for n, p in self.named_parameters()}
            self._ppack = (key, PreciseWeightPack(params, [b.mode for b in self.blocks], self.config.n_head))
        return self._ppack[1]

    def forward_rows(self, B, lat, ctx, tgt, logits_dtype=torch.float32):
        """Stack on 2-D streams -> logits [B*NT, V] (no reshapes / dtype round trips).  bf16 streams take the tcgen05
        engine (1e-2 tolerance); fp32 streams take the fp32-accurate path (1e-4 tolerance, `precision = "fp32"`)."""
        seed = _dropout_seed(self, self.config.embd_pdrop, self.config.resid_pdrop, self.config.attn_pdrop)
        if seed is not None:
            return self._forward_rows_dropout(B, lat, ctx, tgt, logits_dtype, seed)
        if lat.dtype == torch.float32:
            return stack_forward_f32(self.precise_pack(), B, lat, ctx, tgt).to(logits_dtype)
        return stack_forward(self.weight_pack(), B, lat, ctx, tgt, logits_dtype)

    def sample_rows(self, B, lat, ctx, tgt, temperature, seed, offset):
        """Stack + head + one categorical draw per target row in the head GEMM's epilogue -> ids int64 [B*NT] (eval mode,
        bf16 engine only): the sampler steps of draft / revise without materialised logits."""
        if self.training or lat.dtype != torch.bfloat16:
            raise _lib.MebtError("sample_rows: eval mode on the bf16 engine only")
        return stack_forward_sample(self.weight_pack(), B, lat, ctx, tgt, temperature, seed, offset)

    def _forward_rows_dropout(self, B, lat, ctx, tgt, logits_dtype, seed):
        """Training-mode forward with the configured dropout (`model.train(); model(x, c, indices=...)` with the STL
        yaml's p = 0.1, gpt.py:238-248): embd dropout on the three streams, attention / proj / mlp dropout inside
        every block, block by block on the kernels (same sites as the one-call training engine, so one seed gives one
        set of masks on both paths).  Inference-only output (no autograd graph): training runs through
        `training_step` / `TrainState`."""
        cfg = self.config
        if lat.dtype != torch.bfloat16:
            raise NotImplementedError('mebt_b200: training-mode dropout runs on the bf16 path (precision = "bf16")')
        if cfg.embd_pdrop > 0:
            for name, t in (("lat", lat), ("ctx", ctx), ("tgt", tgt)):
                ops.dropout_rows_(t, cfg.embd_pdrop, seed, STEM_SITES[name])
        for i, blk in enumerate(self.blocks):
            lat, ctx, tgt = block_forward(blk.layer_weights(), cfg.n_head, B, lat, ctx, tgt,
                                          drop=(cfg.attn_pdrop, cfg.resid_pdrop, seed, 4 * i))
        xf = ops.layernorm(tgt, self.ln_f.weight.detach().float().contiguous(), self.ln_f.bias.detach().float().contiguous())
        w_head = self.weight_pack().w_head
        return ops.gemm(xf, w_head, out_dtype=logits_dtype)

    def forward(self, sos_emb, contexts, targets, mask_emb, attn_bias=None, debug=False):
        B, NT = targets.shape[0], targets.shape[1]
        logits = self.forward_rows(B, _as_rows(sos_emb), _as_rows(contexts), _as_rows(targets))
        return logits.view(B, NT, -1), None


def complement_idx(idx, dim):
    """The indices of range(dim) that are NOT in the trailing dimension of `idx` ([N, *, K] -> [N, *, dim - K], ascending;
    mebt/modules/gpt.py:19-42).  The indices of a row must be distinct."""
    mask = torch.ones(*idx.shape[:-1], dim, dtype=torch.bool, device=idx.device)
    mask.scatter_(-1, idx, False)
    full = torch.arange(dim, device=idx.device).expand(*idx.shape[:-1], dim)
    return full[mask].view(*idx.shape[:-1], dim - idx.shape[-1])
