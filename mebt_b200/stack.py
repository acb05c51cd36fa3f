"""The MeBT layer stack composed from the C-ABI kernels, one kernel call per op.

This is the op-by-op composition used by the drop-in `GPT` / `Block` / `CrossAttention` modules.  Activations
are bf16 2-D buffers ([B*rows, D]); accumulation, LayerNorm statistics, softmax and logits are fp32.

Reference behaviour reproduced (mebt/modules/gpt.py:159-195, :234-253):
  * ln1 is applied to BOTH the query stream and the key stream of a block (same parameters);
  * the attention residual is added to the *normalised* query, `x = ln1(q) + attn`;
  * contexts are never updated by the four latent modes; the head reads `targets` only.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Mapping

import torch

from . import _lib, ops

MODE_IDS = {"latent_enc": 0, "latent_self": 1, "latent_dec": 2, "lt2l": 3, "maskgit": 4}
LATENT_MODES = ("latent_enc", "latent_self", "latent_dec", "lt2l")


@dataclass
class LayerWeights:
    mode: str
    ln1_w: torch.Tensor
    ln1_b: torch.Tensor
    ln2_w: torch.Tensor
    ln2_b: torch.Tensor
    w_qkv: torch.Tensor      # bf16 [3D, D] rows = (query, key, value)
    b_qkv: torch.Tensor      # fp32 [3D]
    w_proj: torch.Tensor     # bf16 [D, D]
    b_proj: torch.Tensor
    w_fc1: torch.Tensor      # bf16 [4D, D]
    b_fc1: torch.Tensor
    w_fc2: torch.Tensor      # bf16 [D, 4D]
    b_fc2: torch.Tensor


class WeightPack:
    """bf16 tensor-core operands built from fp32 parameters named as in the reference state_dict."""

    def __init__(self, params: Mapping[str, torch.Tensor], modes, n_head: int, prefix: str = "transformer."):
        self.modes = list(modes)
        self.n_head = n_head
        self.layers: list[LayerWeights] = []
        f32 = lambda t: t.detach().float().contiguous()
        for i, mode in enumerate(self.modes):
            p = f"{prefix}blocks.{i}."
            wq, wk, wv = (params[p + f"attn.{n}.weight"].detach().float() for n in ("query", "key", "value"))
            bq, bk, bv = (params[p + f"attn.{n}.bias"].detach().float() for n in ("query", "key", "value"))
            self.layers.append(LayerWeights(
                mode=mode,
                ln1_w=f32(params[p + "ln1.weight"]), ln1_b=f32(params[p + "ln1.bias"]),
                ln2_w=f32(params[p + "ln2.weight"]), ln2_b=f32(params[p + "ln2.bias"]),
                w_qkv=ops.cast_bf16(torch.cat([wq, wk, wv], 0)), b_qkv=torch.cat([bq, bk, bv]).contiguous(),
                w_proj=ops.cast_bf16(f32(params[p + "attn.proj.weight"])), b_proj=f32(params[p + "attn.proj.bias"]),
                w_fc1=ops.cast_bf16(f32(params[p + "mlp.0.weight"])), b_fc1=f32(params[p + "mlp.0.bias"]),
                w_fc2=ops.cast_bf16(f32(params[p + "mlp.2.weight"])), b_fc2=f32(params[p + "mlp.2.bias"])))
        self.lnf_w = f32(params[prefix + "ln_f.weight"])
        self.lnf_b = f32(params[prefix + "ln_f.bias"])
        self.w_head = ops.cast_bf16(f32(params[prefix + "head.weight"]))
        self.D = self.lnf_w.numel()
        self.V = self.w_head.shape[0]
        # host array of mebt_layer_t for the one-call engine
        arr = (_lib.LayerStruct * len(self.layers))()
        for i, w in enumerate(self.layers):
            if w.mode not in MODE_IDS:
                raise ValueError(f"unknown block mode {w.mode!r}")
            arr[i].mode = MODE_IDS[w.mode]
            for f in ("ln1_w", "ln1_b", "ln2_w", "ln2_b", "w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1",
                      "w_fc2", "b_fc2"):
                setattr(arr[i], f, getattr(w, f).data_ptr())
        self.c_layers = arr
        # latent_enc K|V hoist: per-block (key|value) weights with ln1's gamma folded in, bias absorbing W.beta
        enc = [i for i, m in enumerate(self.modes) if m == "latent_enc"]
        self.hoist = None
        if enc:
            D = self.D
            ws, bs = [], []
            for i in enc:
                p = f"{prefix}blocks.{i}."
                w_kv = torch.cat([params[p + "attn.key.weight"], params[p + "attn.value.weight"]], 0).detach().float()
                b_kv = torch.cat([params[p + "attn.key.bias"], params[p + "attn.value.bias"]]).detach().float()
                g, b = params[p + "ln1.weight"].detach().float(), params[p + "ln1.bias"].detach().float()
                ws.append(w_kv * g[None, :])
                bs.append(b_kv + w_kv @ b)
            self._w_enc_kv = ops.cast_bf16(torch.cat(ws, 0).contiguous())
            self._b_enc_kv = torch.cat(bs).contiguous()
            self._ones = torch.ones(D, device=self.lnf_w.device)
            self._zeros = torch.zeros(D, device=self.lnf_w.device)
            self.hoist = _lib.EncHoistStruct(len(enc), self._w_enc_kv.data_ptr(), self._b_enc_kv.data_ptr(),
                                             self._ones.data_ptr(), self._zeros.data_ptr())

    def last_live_layer(self) -> int:
        """Blocks after the last one writing `targets` cannot reach the logits (gpt.py:247 reads targets only)."""
        live = -1
        for i, m in enumerate(self.modes):
            if m in ("latent_dec", "maskgit"):
                live = i
        return live


def attention_core(w: LayerWeights, n_head: int, B: int, qn, kn1, nk1: int, kn2=None, nk2: int = 0, q_is_k1=False,
                   attn_p: float = 0.0, attn_seed: int = 0):
    """q/k/v projections + attention for one block.  qn: ln1(query) [B*NQ, D]; kn1/kn2: ln1(key sources).
    q_is_k1: the first key source is the query stream itself (latent_self, lt2l, maskgit) -> one fused QKV GEMM.
    attn_p > 0: training-mode dropout on the attention probabilities (attn_drop, gpt.py:136)."""
    D = qn.shape[1]
    NQ = qn.shape[0] // B
    if q_is_k1:
        qkv = ops.gemm(qn, w.w_qkv, w.b_qkv)                      # [B*NQ, 3D]
        q_buf, q_col, kv1, k1c, v1c = qkv, 0, qkv, D, 2 * D
    else:
        q_buf, q_col = ops.gemm(qn, w.w_qkv[:D], w.b_qkv[:D]), 0   # [B*NQ, D]
        kv1, k1c, v1c = None, 0, 0
        if nk1 > 0:
            kv1 = ops.gemm(kn1, w.w_qkv[D:], w.b_qkv[D:])          # [B*NK1, 2D]
            k1c, v1c = 0, D
    kv2 = None
    if nk2 > 0:
        kv2 = ops.gemm(kn2, w.w_qkv[D:], w.b_qkv[D:])
    return ops.attention(q_buf, q_col, kv1, k1c, v1c, nk1, kv2, 0, D, nk2, B, n_head, NQ, drop_p=attn_p,
                         drop_seed=attn_seed)


def block_forward(w: LayerWeights, n_head: int, B: int, lat, ctx, tgt, drop=None):
    """One Block (gpt.py:159-195) on 2-D bf16 streams; returns the updated (lat, ctx, tgt).
    drop = (attn_p, resid_p, seed, site): training-mode dropout with the training engine's sites (site + 0 attention
    probabilities, + 1 proj output, + 2 mlp output; csrc/engine_train.cu), so that for the same seed the op-by-op
    module path and `mebt_stack_forward_train_dropout` draw the same masks."""
    mode = w.mode
    attn_p, resid_p, seed, site = drop if drop is not None else (0.0, 0.0, 0, 0)
    att_kw = dict(attn_p=attn_p, attn_seed=seed + site)
    L = lat.shape[0] // B
    NC = ctx.shape[0] // B
    NT = tgt.shape[0] // B
    ln1 = lambda t: ops.layernorm(t, w.ln1_w, w.ln1_b)
    if mode == "latent_enc":
        qn = ln1(lat)
        att = attention_core(w, n_head, B, qn, ln1(ctx) if NC > 0 else None, NC, **att_kw)
    elif mode == "latent_self":
        qn = ln1(lat)
        att = attention_core(w, n_head, B, qn, None, L, q_is_k1=True, **att_kw)
    elif mode == "latent_dec":
        qn = ln1(tgt)
        att = attention_core(w, n_head, B, qn, ln1(lat), L, **att_kw)
    elif mode == "lt2l":
        qn = ln1(lat)
        att = attention_core(w, n_head, B, qn, None, L, ln1(tgt) if NT > 0 else None, NT, q_is_k1=True, **att_kw)
    elif mode == "maskgit":
        D = lat.shape[1]
        both = torch.cat([ctx.view(B, NC, D), tgt.view(B, NT, D)], 1).reshape(B * (NC + NT), D)
        qn = ln1(both)
        att = attention_core(w, n_head, B, qn, None, NC + NT, q_is_k1=True, **att_kw)
    else:
        raise ValueError(f"unknown block mode {mode!r}")
    if resid_p > 0.0:                                             # x = ln1(q) + drop(proj(attn)), x + drop(mlp(ln2(x)))
        x = ops.dropout_rows_(ops.gemm(att, w.w_proj, w.b_proj), resid_p, seed, site + 1, resid=qn)
        h = ops.layernorm(x, w.ln2_w, w.ln2_b)
        u = ops.gemm(h, w.w_fc1, w.b_fc1, gelu=True)
        x = ops.dropout_rows_(ops.gemm(u, w.w_fc2, w.b_fc2), resid_p, seed, site + 2, resid=x)
    else:
        x = ops.gemm(att, w.w_proj, w.b_proj, residual=qn)        # x = ln1(q) + proj(attn)
        h = ops.layernorm(x, w.ln2_w, w.ln2_b)
        u = ops.gemm(h, w.w_fc1, w.b_fc1, gelu=True)
        x = ops.gemm(u, w.w_fc2, w.b_fc2, residual=x)             # x + mlp(ln2(x))
    if mode in ("latent_enc", "latent_self", "lt2l"):
        lat = x
    elif mode == "latent_dec":
        tgt = x
    else:
        D = x.shape[1]
        x3 = x.view(B, NC + NT, D)
        ctx, tgt = x3[:, :NC].reshape(B * NC, D), x3[:, NC:].reshape(B * NT, D)
    return lat, ctx, tgt


_workspaces: dict = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    ws = _workspaces.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25), dtype=torch.uint8, device=device)
        _workspaces[device] = ws
    return ws


def stack_forward(pack: WeightPack, B: int, lat, ctx, tgt, logits_dtype=torch.float32, hoist=True):
    """GPT.forward (gpt.py:234-253) in eval mode through the one-call C++ engine (`mebt_stack_forward_hoisted`).
    lat/tgt are updated in place.  Returns logits [B*NT, V].  hoist=True computes the K|V projections of all
    latent_enc blocks in one GEMM over the once-normalised contexts."""
    D = pack.D
    L, NC, NT = lat.shape[0] // B, ctx.shape[0] // B, tgt.shape[0] // B
    for t in (lat, ctx, tgt):
        if t.dtype != torch.bfloat16 or not t.is_contiguous():
            raise _lib.MebtError("stack_forward streams must be contiguous bf16")
    logits = torch.empty(B * NT, pack.V, device=lat.device, dtype=logits_dtype)
    h = pack.hoist if (hoist and pack.hoist is not None and NC > 0) else None
    nbytes = _lib.lib.mebt_stack_forward_hoisted_workspace_bytes(B, L, NC, NT, D, h.n_enc if h is not None else 0)
    ws = _workspace(lat.device, nbytes)
    _lib.call("mebt_stack_forward_hoisted", pack.c_layers, len(pack.layers), pack.lnf_w.data_ptr(), pack.lnf_b.data_ptr(),
              pack.w_head.data_ptr(), ctypes.byref(h) if h is not None else None, B, L, NC, NT, D, pack.n_head, pack.V,
              lat.data_ptr(), ctx.data_ptr(),
              tgt.data_ptr(), logits.data_ptr(), ops._DT[logits_dtype], ws.data_ptr(), ws.numel(),
              torch.cuda.current_stream().cuda_stream)
    return logits


def stack_forward_sample(pack: WeightPack, B: int, lat, ctx, tgt, temperature: float, seed: int, offset: int, hoist=True):
    """The eval forward with the sampling step fused into the head GEMM (`mebt_stack_forward_sample`): returns one
    categorical draw per target row, ids int64 [B*NT], from softmax(logits / temperature) by the Gumbel-max rule with
    in-kernel counter-hash noise; the [B*NT, V] logits never reach HBM (gpt.py:248 + transformer.py:843-889)."""
    D = pack.D
    L, NC, NT = lat.shape[0] // B, ctx.shape[0] // B, tgt.shape[0] // B
    for t in (lat, ctx, tgt):
        if t.dtype != torch.bfloat16 or not t.is_contiguous():
            raise _lib.MebtError("stack_forward streams must be contiguous bf16")
    ids = torch.empty(B * NT, device=lat.device, dtype=torch.int64)
    h = pack.hoist if (hoist and pack.hoist is not None and NC > 0) else None
    nbytes = _lib.lib.mebt_stack_forward_hoisted_workspace_bytes(B, L, NC, NT, D, h.n_enc if h is not None else 0)
    ws = _workspace(lat.device, nbytes)
    _lib.call("mebt_stack_forward_sample", pack.c_layers, len(pack.layers), pack.lnf_w.data_ptr(), pack.lnf_b.data_ptr(),
              pack.w_head.data_ptr(), ctypes.byref(h) if h is not None else None, B, L, NC, NT, D, pack.n_head, pack.V,
              lat.data_ptr(), ctx.data_ptr(), tgt.data_ptr(), None, 0, ids.data_ptr(), float(temperature), int(seed),
              int(offset), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    return ids


def stack_forward_ops(pack: WeightPack, B: int, lat, ctx, tgt, logits_dtype=torch.float32, skip_dead=True):
    """Same computation composed op by op from Python (one C-ABI call per kernel); used by the per-module
    drop-in API and as a cross-check of the engine."""
    last = pack.last_live_layer() if skip_dead else len(pack.layers) - 1
    for i, w in enumerate(pack.layers):
        if i > last:
            break
        lat, ctx, tgt = block_forward(w, pack.n_head, B, lat, ctx, tgt)
    xf = ops.layernorm(tgt, pack.lnf_w, pack.lnf_b)
    return ops.gemm(xf, pack.w_head, out_dtype=logits_dtype)


# ---- fp32-accurate mode -----------------------------------------------------------------------------------------------
class PreciseWeightPack:
    """Weights of the stack as bf16 (hi | hi | lo) splits for the fp32-accurate GEMM (csrc/precise.cu); LayerNorm
    parameters and biases stay fp32."""

    def __init__(self, params: Mapping[str, torch.Tensor], modes, n_head: int, prefix: str = "transformer."):
        self.modes = list(modes)
        self.n_head = n_head
        f32 = lambda t: t.detach().float().contiguous()
        sp = lambda t: ops.split3(f32(t), True)
        self.layers = []
        for i, mode in enumerate(self.modes):
            p = f"{prefix}blocks.{i}."
            wq, wk, wv = (params[p + f"attn.{n}.weight"].detach().float() for n in ("query", "key", "value"))
            bq, bk, bv = (params[p + f"attn.{n}.bias"].detach().float() for n in ("query", "key", "value"))
            self.layers.append(LayerWeights(
                mode=mode,
                ln1_w=f32(params[p + "ln1.weight"]), ln1_b=f32(params[p + "ln1.bias"]),
                ln2_w=f32(params[p + "ln2.weight"]), ln2_b=f32(params[p + "ln2.bias"]),
                w_qkv=sp(torch.cat([wq, wk, wv], 0)), b_qkv=torch.cat([bq, bk, bv]).contiguous(),
                w_proj=sp(params[p + "attn.proj.weight"]), b_proj=f32(params[p + "attn.proj.bias"]),
                w_fc1=sp(params[p + "mlp.0.weight"]), b_fc1=f32(params[p + "mlp.0.bias"]),
                w_fc2=sp(params[p + "mlp.2.weight"]), b_fc2=f32(params[p + "mlp.2.bias"])))
        self.lnf_w = f32(params[prefix + "ln_f.weight"])
        self.lnf_b = f32(params[prefix + "ln_f.bias"])
        self.w_head = sp(params[prefix + "head.weight"])
        self.D = self.lnf_w.numel()
        self.V = self.w_head.shape[0]


def _attention_core_f32(w: LayerWeights, n_head: int, B: int, qn, kn1, nk1: int, kn2=None, nk2: int = 0, q_is_k1=False):
    D = qn.shape[1]
    NQ = qn.shape[0] // B
    if q_is_k1:
        qkv = ops.gemm_f32(qn, w.w_qkv, w.b_qkv)                  # [B*NQ, 3D]
        q_buf, kv1, k1c, v1c = qkv, qkv, D, 2 * D
    else:
        q_buf = ops.gemm_f32(qn, w.w_qkv[:D], w.b_qkv[:D])
        kv1, k1c, v1c = None, 0, 0
        if nk1 > 0:
            kv1 = ops.gemm_f32(kn1, w.w_qkv[D:], w.b_qkv[D:])
            k1c, v1c = 0, D
    kv2 = ops.gemm_f32(kn2, w.w_qkv[D:], w.b_qkv[D:]) if nk2 > 0 else None
    return ops.attention_f32(q_buf, 0, kv1, k1c, v1c, nk1, kv2, 0, D, nk2, B, n_head, NQ)


def stack_forward_f32(pack: PreciseWeightPack, B: int, lat, ctx, tgt):
    """GPT.forward (gpt.py:234-253, eval mode) on fp32 streams [B*rows, D] with fp32-accurate GEMMs and attention:
    the path behind the 1e-4 tolerance of BASELINE.json's configs[0].  Returns fp32 logits [B*NT, V]."""
    D = pack.D
    L, NC, NT = lat.shape[0] // B, ctx.shape[0] // B, tgt.shape[0] // B
    last = -1
    for i, m in enumerate(pack.modes):
        if m in ("latent_dec", "maskgit"):
            last = i
    for i, w in enumerate(pack.layers):
        if i > last:
            break
        ln1 = lambda t: ops.layernorm(t, w.ln1_w, w.ln1_b)
        mode = w.mode
        if mode == "latent_enc":
            qn = ln1(lat)
            att = _attention_core_f32(w, pack.n_head, B, qn, ln1(ctx) if NC > 0 else None, NC)
        elif mode == "latent_self":
            qn = ln1(lat)
            att = _attention_core_f32(w, pack.n_head, B, qn, None, L, q_is_k1=True)
        elif mode == "latent_dec":
            qn = ln1(tgt)
            att = _attention_core_f32(w, pack.n_head, B, qn, ln1(lat), L)
        elif mode == "lt2l":
            qn = ln1(lat)
            att = _attention_core_f32(w, pack.n_head, B, qn, None, L, ln1(tgt) if NT > 0 else None, NT, q_is_k1=True)
        elif mode == "maskgit":
            both = torch.cat([ctx.view(B, NC, D), tgt.view(B, NT, D)], 1).reshape(B * (NC + NT), D)
            qn = ln1(both)
            att = _attention_core_f32(w, pack.n_head, B, qn, None, NC + NT, q_is_k1=True)
        else:
            raise ValueError(f"unknown block mode {mode!r}")
        x = ops.gemm_f32(att, w.w_proj, w.b_proj, residual=qn)
        h = ops.layernorm(x, w.ln2_w, w.ln2_b)
        u = ops.gemm_f32(h, w.w_fc1, w.b_fc1, gelu=True)
        x = ops.gemm_f32(u, w.w_fc2, w.b_fc2, residual=x)
        if mode in ("latent_enc", "latent_self", "lt2l"):
            lat = x
        elif mode == "latent_dec":
            tgt = x
        else:
            x3 = x.view(B, NC + NT, D)
            ctx, tgt = x3[:, :NC].reshape(B * NC, D), x3[:, NC:].reshape(B * NT, D)
    xf = ops.layernorm(tgt, pack.lnf_w, pack.lnf_b)
    return ops.gemm_f32(xf, pack.w_head)
