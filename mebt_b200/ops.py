"""Tensor-level wrappers over the C ABI (include/mebt_b200.h).

torch is used for device memory and streams only: every function below hands raw device pointers to a
hand-written sm_100a kernel and raises `MebtError` on any failure.  There is no eager/PyTorch fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import MebtError, call

BF16, FP32 = 0, 1
GEMM_GELU, GEMM_OUT_FP32, GEMM_ACCUMULATE = 1, 2, 4

_DT = {torch.bfloat16: BF16, torch.float32: FP32}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise MebtError("mebt_b200 ops need CUDA tensors (there is no CPU fallback)")


def _rows2d(t: torch.Tensor):
    """(rows, cols, row_stride) of a tensor viewed as 2-D with a unit-stride last dim."""
    if t.dim() != 2 or t.stride(1) != 1:
        raise MebtError(f"expected a 2-D tensor with contiguous rows, got shape {tuple(t.shape)} stride {t.stride()}")
    return t.shape[0], t.shape[1], t.stride(0)


def gemm(a: torch.Tensor, b: torch.Tensor, bias=None, residual=None, gelu=False, out=None, out_dtype=torch.bfloat16,
         a_mn_major=False, b_mn_major=False, accumulate=False, flags_extra=0) -> torch.Tensor:
    """out[M,N] = act(A @ B^T + bias) + residual.  a: [M,K] (or [K,M] if a_mn_major); b: [N,K] (or [K,N])."""
    _need_cuda(a, b, bias, residual, out)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise MebtError("gemm operands must be bf16")
    ar, ac, lda = _rows2d(a)
    br, bc, ldb = _rows2d(b)
    M, K = (ac, ar) if a_mn_major else (ar, ac)
    N, Kb = (bc, br) if b_mn_major else (br, bc)
    if K != Kb:
        raise MebtError(f"gemm: reduction dims differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    orr, occ, ldc = _rows2d(out)
    if (orr, occ) != (M, N):
        raise MebtError("gemm: bad out shape")
    flags = flags_extra
    if gelu:
        flags |= GEMM_GELU
    if out.dtype == torch.float32:
        flags |= GEMM_OUT_FP32
    elif out.dtype != torch.bfloat16:
        raise MebtError("gemm: out must be bf16 or fp32")
    if accumulate:
        flags |= GEMM_ACCUMULATE
    ldres = 0
    if residual is not None:
        rr, rc, ldres = _rows2d(residual)
        if (rr, rc) != (M, N) or residual.dtype != torch.bfloat16:
            raise MebtError("gemm: residual must be bf16 [M,N]")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise MebtError("gemm: bias must be contiguous fp32 [N]")
    call("mebt_gemm_bf16", a.data_ptr(), lda, int(a_mn_major), b.data_ptr(), ldb, int(b_mn_major), out.data_ptr(), ldc,
         M, N, K, _ptr(bias), _ptr(residual), ldres, flags, _stream())
    return out


def embed_gather(x_indices, ctx_idx, tgt_idx, tok_emb, pos_emb, mask_emb, sos_emb, out_dtype=torch.bfloat16):
    """Stem of reconstruct_mask -> (contexts [B*NC,D], targets [B*NT,D], latents [B*L,D])."""
    _need_cuda(x_indices, ctx_idx, tgt_idx, tok_emb, pos_emb, mask_emb, sos_emb)
    B, N = x_indices.shape
    NC, NT = ctx_idx.shape[1], tgt_idx.shape[1]
    V, D = tok_emb.shape
    pos2 = pos_emb.reshape(-1, D)
    sos2 = sos_emb.reshape(-1, D)
    L = sos2.shape[0]
    for t in (x_indices, ctx_idx, tgt_idx):
        if t.dtype != torch.int64 or (t.shape[1] > 0 and t.stride(1) != 1):
            raise MebtError("index tensors must be int64 with unit inner stride")
    for t in (tok_emb, pos2, mask_emb, sos2):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise MebtError("embedding tables must be contiguous fp32")
    dev = x_indices.device
    ctx = torch.empty(B * NC, D, device=dev, dtype=out_dtype)
    tgt = torch.empty(B * NT, D, device=dev, dtype=out_dtype)
    lat = torch.empty(B * L, D, device=dev, dtype=out_dtype)
    call("mebt_embed_gather", x_indices.data_ptr(), x_indices.stride(0), ctx_idx.data_ptr(),
         ctx_idx.stride(0) if NC > 0 else 0, tgt_idx.data_ptr(), tgt_idx.stride(0) if NT > 0 else 0,
         tok_emb.data_ptr(), pos2.data_ptr(), mask_emb.data_ptr(), sos2.data_ptr(), ctx.data_ptr(), tgt.data_ptr(),
         lat.data_ptr(), B, NC, NT, L, D, V, pos2.shape[0], _DT[out_dtype], _stream())
    return ctx, tgt, lat


def layernorm(x, gamma, beta, out=None, out_dtype=None, eps=1e-5, save_stats=False):
    _need_cuda(x, gamma, beta)
    rows, D, ldx = _rows2d(x)
    if out is None:
        out = torch.empty(rows, D, device=x.device, dtype=out_dtype or x.dtype)
    _, _, ldy = _rows2d(out)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    call("mebt_layernorm", x.data_ptr(), ldx, _DT[x.dtype], gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), ldy,
         _DT[out.dtype], rows, D, float(eps), _ptr(mean), _ptr(rstd), _stream())
    return (out, mean, rstd) if save_stats else out


def attention(q, q_col0, kv1, k1_col0, v1_col0, nk1, kv2, k2_col0, v2_col0, nk2, B, H, NQ, out=None, lse=None,
              drop_p=0.0, drop_seed=0):
    """q: [B*NQ, ldq] bf16 buffer; kv1/kv2: [B*NK, ld] buffers (or None).  Returns O [B*NQ, H*64] bf16.
    drop_p > 0: attn_drop on the softmax output (training mode, gpt.py:136)."""
    _need_cuda(q, kv1, kv2)
    D = H * 64
    if out is None:
        out = torch.empty(B * NQ, D, device=q.device, dtype=torch.bfloat16)
    if drop_p == 0.0 and lse is None and 2 * B * H * (((NQ + 127) // 128 + 1) // 2) <= 148:          # B200: 148 SMs (the library targets sm_100a only)
        # few work items: hand the library scratch for the split-KV form (it splits when the key list is long enough;
        # the same rule as the one-call engine, so both paths produce the same bits)
        nbytes = _lib.lib.mebt_latent_attention_fwd_workspace_bytes(B, H, NQ)
        ws = _ws(q.device, nbytes, "attn_fwd")
        call("mebt_latent_attention_fwd_ws", q.data_ptr(), q.stride(0), q_col0,
             _ptr(kv1) if nk1 > 0 else None, kv1.stride(0) if nk1 > 0 else 0, k1_col0, v1_col0, nk1,
             _ptr(kv2) if nk2 > 0 else None, kv2.stride(0) if nk2 > 0 else 0, k2_col0, v2_col0, nk2,
             out.data_ptr(), out.stride(0), None, B, H, NQ, 64, ws.data_ptr(), ws.numel(), _stream())
        return out
    call("mebt_latent_attention_fwd_dropout", q.data_ptr(), q.stride(0), q_col0,
         _ptr(kv1) if nk1 > 0 else None, kv1.stride(0) if nk1 > 0 else 0, k1_col0, v1_col0, nk1,
         _ptr(kv2) if nk2 > 0 else None, kv2.stride(0) if nk2 > 0 else 0, k2_col0, v2_col0, nk2,
         out.data_ptr(), out.stride(0), _ptr(lse), B, H, NQ, 64, float(drop_p), int(drop_seed), _stream())
    return out


def head_sample(x, w_head, temperature=1.0, seed=0, offset=0):
    """x bf16 [rows, D], w_head bf16 [V, D] -> ids int64 [rows], one draw per row from softmax(x w^T / temperature): the
    head GEMM with the Gumbel-max epilogue (no logits written)."""
    _need_cuda(x, w_head)
    if x.dtype != torch.bfloat16 or w_head.dtype != torch.bfloat16:
        raise MebtError("head_sample: bf16 operands expected")
    rows, D, ldx = _rows2d(x)
    V, D2, ldw = _rows2d(w_head)
    if D != D2:
        raise MebtError("head_sample: inner dimensions differ")
    ids = torch.empty(rows, device=x.device, dtype=torch.int64)
    ws = _ws(x.device, max(8 * rows, 8), "head_sample")
    call("mebt_head_sample", x.data_ptr(), ldx, w_head.data_ptr(), ldw, rows, V, D, float(temperature), int(seed), int(offset),
         ids.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    return ids


def scatter_ids(x, tgt_idx, ids):
    """x[b, tgt_idx[b,i]] = ids[b,i] in place."""
    _need_cuda(x, tgt_idx, ids)
    B, N = x.shape
    NT = tgt_idx.shape[1]
    ids = ids.contiguous()
    call("mebt_scatter_ids", x.data_ptr(), x.stride(0), tgt_idx.data_ptr(), tgt_idx.stride(0) if NT else 0,
         ids.data_ptr(), B, NT, N, _stream())
    return x


def masked_ce(logits, targets, label_smoothing=0.0, dlogits=None, grad_scale=1.0):
    """-> (stats fp32[3] = {ce_sum, n_top1, n_top5}, row_loss fp32 [rows]).  logits: [rows, V] fp32/bf16."""
    _need_cuda(logits, targets)
    rows, V, ld = _rows2d(logits)
    targets = targets.reshape(-1).contiguous()
    row_loss = torch.empty(rows, device=logits.device, dtype=torch.float32)
    row_rank = torch.empty(rows, device=logits.device, dtype=torch.int32)
    ldd = 0
    if dlogits is not None:
        _, _, ldd = _rows2d(dlogits)
        if dlogits.dtype != logits.dtype:
            raise MebtError("dlogits dtype must match logits")
    call("mebt_masked_ce", logits.data_ptr(), ld, _DT[logits.dtype], targets.data_ptr(), rows, V,
         float(label_smoothing), row_loss.data_ptr(), row_rank.data_ptr(), _ptr(dlogits), ldd, float(grad_scale),
         _stream())
    stats = torch.empty(3, device=logits.device, dtype=torch.float32)
    call("mebt_ce_reduce", row_loss.data_ptr(), row_rank.data_ptr(), rows, stats.data_ptr(), _stream())
    return stats, row_loss


def sample_logits(logits, temperature=1.0, top_k=None, top_p=None, noise=None, seed=0, offset=0, return_probs=False):
    """logits [rows, V] -> (ids int64 [rows], scores fp32 [rows], probs fp32 [rows,V] | None)."""
    _need_cuda(logits, noise)
    rows, V, ld = _rows2d(logits)
    ids = torch.empty(rows, device=logits.device, dtype=torch.int64)
    scores = torch.empty(rows, device=logits.device, dtype=torch.float32)
    probs = torch.empty(rows, V, device=logits.device, dtype=torch.float32) if return_probs else None
    if noise is not None:
        if noise.dtype != torch.float32 or not noise.is_contiguous() or noise.numel() != rows * V:
            raise MebtError("noise must be contiguous fp32 [rows, V]")
    call("mebt_sample_logits", logits.data_ptr(), ld, _DT[logits.dtype], rows, V, float(temperature),
         int(top_k) if top_k else 0, float(top_p) if top_p is not None else 0.0, _ptr(noise), int(seed), int(offset),
         ids.data_ptr(), scores.data_ptr(), _ptr(probs), _stream())
    return ids, scores, probs


def remask_sort(score, ctx_idx, tgt_idx, n_new, ctemp, noise=None, seed=0, offset=0, want_order=False):
    """-> (next_ctx [B,NC+n_new], next_tgt [B,NT-n_new], order | None)."""
    _need_cuda(score, ctx_idx, tgt_idx, noise)
    B, NT = score.shape
    NC = ctx_idx.shape[1]
    score = score.contiguous().float()
    dev = score.device
    next_ctx = torch.empty(B, NC + n_new, device=dev, dtype=torch.int64)
    next_tgt = torch.empty(B, NT - n_new, device=dev, dtype=torch.int64)
    order = torch.empty(B, NT, device=dev, dtype=torch.int64) if want_order else None
    if noise is not None:
        noise = noise.contiguous().float()
    call("mebt_remask_sort", score.data_ptr(), _ptr(noise), float(ctemp), ctx_idx.data_ptr(),
         ctx_idx.stride(0) if NC else 0, tgt_idx.data_ptr(), tgt_idx.stride(0), B, NC, NT, int(n_new), int(seed),
         int(offset), next_ctx.data_ptr(), next_tgt.data_ptr(), _ptr(order), _stream())
    return next_ctx, next_tgt, order


def row_sqnorm(E):
    _need_cuda(E)
    K, C = E.shape
    out = torch.empty(K, device=E.device, dtype=torch.float32)
    call("mebt_row_sqnorm", E.data_ptr(), K, C, out.data_ptr(), _stream())
    return out


def vq_split_codebook(E):
    """E [K, C] fp32 -> the fp16 (hi | hi | lo) form [K, 3C] the tensor-core search multiplies by (once per codebook)."""
    _need_cuda(E)
    K, C = E.shape
    out = torch.empty(K, 3 * C, device=E.device, dtype=torch.float16)
    call("mebt_vq_split_codebook", E.contiguous().data_ptr(), K, C, out.data_ptr(), _stream())
    return out


def vq_argmin(z, E, e_sqnorm=None, e_split=None, tensor_cores=None):
    """z: [b, C, ...] fp32 channel-first; E: [K, C] fp32 -> encodings int64 [b, ...].
    tensor_cores (default: whenever C and K are multiples of 64): the fp16-split tcgen05 GEMM with the argmin epilogue;
    False: the fp32 FFMA kernel (kept as the cross-check of the split arithmetic)."""
    _need_cuda(z, E)
    z = z.contiguous()
    b, C = z.shape[:2]
    S = z[0, 0].numel()
    K = E.shape[0]
    if e_sqnorm is None:
        e_sqnorm = row_sqnorm(E)
    if tensor_cores is None:
        tensor_cores = C % 64 == 0 and K % 64 == 0 and C <= 1024
    if tensor_cores:
        if e_split is None:
            e_split = vq_split_codebook(E)
        out = torch.empty(b * S, device=z.device, dtype=torch.int64)
        ws_bytes = _lib.lib.mebt_vq_argmin_tc_workspace_bytes(b * S, C)
        ws = torch.empty(ws_bytes, device=z.device, dtype=torch.uint8)
        call("mebt_vq_argmin_tc", z.data_ptr(), b, C, S, e_split.data_ptr(), e_sqnorm.data_ptr(), K, out.data_ptr(),
             ws.data_ptr(), ws_bytes, _stream())
        return out.view(b, *z.shape[2:])
    out = torch.empty(b * S, device=z.device, dtype=torch.int64)
    ws_bytes = _lib.lib.mebt_vq_argmin_workspace_bytes(b * S)
    ws = torch.empty(ws_bytes, device=z.device, dtype=torch.uint8)
    call("mebt_vq_argmin", z.data_ptr(), b, C, S, E.data_ptr(), e_sqnorm.data_ptr(), K, out.data_ptr(), ws.data_ptr(),
         ws_bytes, _stream())
    return out.view(b, *z.shape[2:])


def row_gather(enc, E, channel_first=False):
    """F.embedding(enc, E); channel_first=True returns [b, C, ...] like shift_dim(h, -1, 1)."""
    _need_cuda(enc, E)
    enc = enc.contiguous()
    K, C = E.shape
    b = enc.shape[0]
    S = enc[0].numel()
    if channel_first:
        out = torch.empty(b, C, *enc.shape[1:], device=enc.device, dtype=torch.float32)
    else:
        out = torch.empty(*enc.shape, C, device=enc.device, dtype=torch.float32)
    call("mebt_row_gather", enc.data_ptr(), E.data_ptr(), out.data_ptr(), b, S, C, K, int(channel_first), _stream())
    return out


def cast_bf16(x):
    _need_cuda(x)
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    call("mebt_cast_f32_to_bf16", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def check_index_errors():
    call("mebt_check_index_errors", _stream())


# ---- backward ops (training step) --------------------------------------------------------------------------------
GEMM_DGELU = 8
_ws_cache: dict = {}


def _ws(device, nbytes: int, tag: str = "ws") -> torch.Tensor:
    key = (device, tag)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def gemm_aux(a, b, aux, bias=None, residual=None, gelu=False, dgelu=False, out=None, a_mn_major=False,
             b_mn_major=False):
    """gemm with the auxiliary pre-activation tensor: gelu=True stores it, dgelu=True multiplies by gelu'(aux)."""
    _need_cuda(a, b, aux)
    ar, ac, lda = _rows2d(a)
    br, bc, ldb = _rows2d(b)
    M, K = (ac, ar) if a_mn_major else (ar, ac)
    N = bc if b_mn_major else br
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    flags = (GEMM_GELU if gelu else 0) | (GEMM_DGELU if dgelu else 0) | (GEMM_OUT_FP32 if out.dtype == torch.float32 else 0)
    call("mebt_gemm_bf16_aux", a.data_ptr(), lda, int(a_mn_major), b.data_ptr(), ldb, int(b_mn_major), out.data_ptr(),
         out.stride(0), M, N, K, _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0,
         aux.data_ptr(), aux.stride(0), flags, _stream())
    return out


def grouped_wgrad(problems):
    """problems: list of (dY [rows, n_out] bf16, X [rows, k_in] bf16, dW [n_out, k_in] fp32, accumulate) -> one launch
    computing every dW (+)= dY^T X (disjoint outputs)."""
    n = len(problems)
    descs = (_lib.WgradDesc * n)()
    for d, (dy, x, dw, acc) in zip(descs, problems):
        _need_cuda(dy, x, dw)
        if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dw.dtype != torch.float32:
            raise MebtError("grouped_wgrad: dY / X must be bf16 and dW fp32")
        rows, n_out, ld_dy = _rows2d(dy)
        rows_x, k_in, ldx = _rows2d(x)
        if rows != rows_x or tuple(dw.shape) != (n_out, k_in):
            raise MebtError("grouped_wgrad: shape mismatch")
        d.dY, d.ld_dy, d.X, d.ldx, d.dW, d.ldw = dy.data_ptr(), ld_dy, x.data_ptr(), ldx, dw.data_ptr(), dw.stride(0)
        d.n_out, d.k_in, d.rows, d.accumulate = n_out, k_in, rows, int(bool(acc))
    call("mebt_gemm_grouped_wgrad", descs, n, _stream())


def colsum(x, out=None, accumulate=False):
    """out[n] (+)= sum_rows x[:, n]; x bf16 [rows, N]."""
    _need_cuda(x)
    rows, N, ld = _rows2d(x)
    if out is None:
        out = torch.zeros(N, device=x.device, dtype=torch.float32)
        accumulate = False
    nbytes = _lib.lib.mebt_colsum_workspace_bytes(N)
    ws = _ws(x.device, nbytes)
    call("mebt_colsum", x.data_ptr(), ld, rows, N, out.data_ptr(), int(accumulate), ws.data_ptr(), ws.numel(), _stream())
    return out


def layernorm_bwd(dy, x, mean, rstd, gamma, dx=None, accumulate_dx=False, dgamma=None, dbeta=None,
                  accumulate_params=False):
    _need_cuda(dy, x)
    rows, D = x.shape
    if dx is None:
        dx = torch.empty_like(x)
        accumulate_dx = False
    if dgamma is None:
        dgamma = torch.empty(D, device=x.device, dtype=torch.float32)
        dbeta = torch.empty(D, device=x.device, dtype=torch.float32)
        accumulate_params = False
    nbytes = _lib.lib.mebt_layernorm_bwd_workspace_bytes(D)
    ws = _ws(x.device, nbytes)
    call("mebt_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
         dx.data_ptr(), int(accumulate_dx), dgamma.data_ptr(), dbeta.data_ptr(), int(accumulate_params), rows, D,
         ws.data_ptr(), ws.numel(), _stream())
    return dx, dgamma, dbeta


def embed_backward(x_indices, ctx_idx, tgt_idx, d_ctx, d_tgt, d_lat, d_tok, d_pos, d_mask, d_sos):
    """Accumulates into the fp32 gradients of tok_emb / pos_emb / mask_emb / sos_emb."""
    B = x_indices.shape[0]
    NC, NT = ctx_idx.shape[1], tgt_idx.shape[1]
    D = d_tok.shape[-1]
    L = d_lat.shape[0] // B
    ws = _ws(x_indices.device, _lib.lib.mebt_colsum_workspace_bytes(D))
    call("mebt_embed_backward", x_indices.data_ptr(), x_indices.stride(0), ctx_idx.data_ptr(),
         ctx_idx.stride(0) if NC else 0, tgt_idx.data_ptr(), tgt_idx.stride(0) if NT else 0, d_ctx.data_ptr(),
         d_tgt.data_ptr(), d_lat.data_ptr(), d_tok.data_ptr(), d_pos.data_ptr(), d_mask.data_ptr(), d_sos.data_ptr(), B,
         NC, NT, L, D, ws.data_ptr(), ws.numel(), _stream())


def attention_bwd(q, q_col0, kv1, k1_col0, v1_col0, nk1, kv2, k2_col0, v2_col0, nk2, o, do, lse, dq, dq_col0, dkv1,
                  dk1_col0, dv1_col0, dkv2, dk2_col0, dv2_col0, B, H, NQ, drop_p=0.0, drop_seed=0):
    nbytes = _lib.lib.mebt_latent_attention_bwd_workspace_bytes(B, H, NQ)
    ws = _ws(q.device, nbytes, "attn")
    call("mebt_latent_attention_bwd_dropout", q.data_ptr(), q.stride(0), q_col0,
         _ptr(kv1) if nk1 else None, kv1.stride(0) if nk1 else 0, k1_col0, v1_col0, nk1,
         _ptr(kv2) if nk2 else None, kv2.stride(0) if nk2 else 0, k2_col0, v2_col0, nk2,
         o.data_ptr(), o.stride(0), do.data_ptr(), do.stride(0), lse.data_ptr(), dq.data_ptr(), dq.stride(0), dq_col0,
         _ptr(dkv1) if nk1 else None, dkv1.stride(0) if nk1 else 0, dk1_col0, dv1_col0,
         _ptr(dkv2) if nk2 else None, dkv2.stride(0) if nk2 else 0, dk2_col0, dv2_col0, B, H, NQ, 64, float(drop_p),
         int(drop_seed), ws.data_ptr(), ws.numel(), _stream())


def dropout_rows_(x, p: float, seed: int, site: int, resid=None):
    """In-place training-mode dropout of a bf16 [rows, D] activation: x <- (resid +) x * keep / (1-p), keep a pure
    function of (seed, site, row, column) (nn.Dropout, mebt/modules/gpt.py:216,239-241).  Backward = the same call on dy."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or x.dim() != 2 or x.stride(1) != 1:
        raise MebtError("dropout_rows_: bf16 [rows, D] with unit column stride expected")
    if x.shape[0] == 0:
        return x
    call("mebt_dropout_rows", x.data_ptr(), x.stride(0), resid.data_ptr() if resid is not None else None,
         resid.stride(0) if resid is not None else 0, x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], float(p),
         int(seed), int(site), _stream())
    return x


def attention_dropout_mask(B, H, NQ, NK1, NK2, p: float, seed: int, device="cuda"):
    """Keep factors (0 or 1/(1-p)) the attention kernels apply for (p, seed): fp32 [B, H, NQ, NK1+NK2] (test support)."""
    out = torch.empty(B, H, NQ, NK1 + NK2, device=device, dtype=torch.float32)
    call("mebt_attention_dropout_mask", out.data_ptr(), B, H, NQ, NK1, NK2, float(p), int(seed), _stream())
    return out


# ---- fp32-accurate mode (bf16x3 split GEMM + fp32 attention); see csrc/precise.cu -------------------------------------
def split3(x, weight_side: bool):
    """fp32 [rows, K] -> bf16 [rows, 3K] = hi|lo|hi (activations) or hi|hi|lo (weights)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise MebtError("split3: fp32 [rows, K] with unit column stride expected")
    rows, K = x.shape
    out = torch.empty(rows, 3 * K, device=x.device, dtype=torch.bfloat16)
    call("mebt_split_f32_bf16x3", x.data_ptr(), x.stride(0), rows, K, out.data_ptr(), int(weight_side), _stream())
    return out


def gemm_f32(a, w_split, bias=None, residual=None, gelu=False):
    """out[M,N] (fp32) = act(a @ W^T + bias) (+ residual) to fp32 accuracy: a fp32 [M,K]; w_split = split3(W, True)."""
    if residual is not None and gelu:
        raise MebtError("gemm_f32: residual and gelu are exclusive")
    out = residual.clone() if residual is not None else None          # C += A W^T + bias on top of the residual
    if out is None:
        out = torch.empty(a.shape[0], w_split.shape[0], device=a.device, dtype=torch.float32)
    return gemm(split3(a, False), w_split, bias, gelu=gelu, out=out, accumulate=residual is not None)


def attention_f32(q, q_col0, kv1, k1_col0, v1_col0, nk1, kv2, k2_col0, v2_col0, nk2, B, H, NQ):
    """mebt_latent_attention_fwd on fp32 buffers in fp32 arithmetic.  Returns O [B*NQ, H*64] fp32."""
    _need_cuda(q, kv1, kv2)
    out = torch.empty(B * NQ, H * 64, device=q.device, dtype=torch.float32)
    call("mebt_latent_attention_fwd_f32", q.data_ptr(), q.stride(0), q_col0,
         _ptr(kv1) if nk1 > 0 else None, kv1.stride(0) if nk1 > 0 else 0, k1_col0, v1_col0, nk1,
         _ptr(kv2) if nk2 > 0 else None, kv2.stride(0) if nk2 > 0 else 0, k2_col0, v2_col0, nk2,
         out.data_ptr(), out.stride(0), B, H, NQ, 64, _stream())
    return out


# ---- VQGAN encoder / decoder convolutions (csrc/conv3d.cu); activations channels-last bf16 [B, T, H, W, C] -------------------
def _ints(*v):
    return (ctypes.c_int * len(v))(*v)


def pad_norm_act(x, pad6, norm=0, act=0, groups=32, eps=1e-6, gamma=None, beta=None):
    """replicate_pad(act(norm(x))): x [B, T, H, W, C] bf16 (C % 8 == 0) -> [B, T + pads, H + pads, W + pads, C].
    pad6 = (t0, t1, h0, h1, w0, w1).  norm: 0 none, 1 GroupNorm(groups, eps; gamma, beta), 2 affine x * gamma + beta.
    act: 0 none, 1 SiLU.  (F.pad(mode='replicate') + Normalize + silu of mebt/vqgan.py:255-260,344-349,381.)"""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or x.dim() != 5 or not x.is_contiguous():
        raise MebtError("pad_norm_act: contiguous bf16 [B, T, H, W, C] expected")
    B, T, H, W, C = x.shape
    y = torch.empty(B, T + pad6[0] + pad6[1], H + pad6[2] + pad6[3], W + pad6[4] + pad6[5], C, device=x.device,
                    dtype=torch.bfloat16)
    ws = _ws(x.device, _lib.lib.mebt_groupnorm_workspace_bytes(B, groups) if norm == 1 else 16, "gn")
    call("mebt_pad_norm_act", x.data_ptr(), C, y.data_ptr(), C, B, T, H, W, C, _ints(*pad6), int(norm), int(act), int(groups),
         float(eps), _ptr(gamma), _ptr(beta), ws.data_ptr(), ws.numel(), _stream())
    return y


def conv3d_ndhwc(xp, w_packed, cin, cout, taps, step, odims, bias=None, resid=None, out=None, origin=(0, 0, 0),
                 ystep=(1, 1, 1), yorigin=(0, 0, 0)):
    """Implicit-GEMM 3-D convolution over an already padded channels-last input (mebt_conv3d_ndhwc).
    xp [B, Tp, Hp, Wp, C]; w_packed bf16 [ceil64(cout), prod(taps) * ceil64(cin)] (zero rows behind cout); odims = (To, Ho, Wo) positions computed; they
    are written to out[b, t * ystep + yorigin, ...] (out defaults to a dense [B, To, Ho, Wo, cout] tensor)."""
    _need_cuda(xp, w_packed, resid)
    B = xp.shape[0]
    if out is None:
        out = torch.empty(B, *odims, cout, device=xp.device, dtype=torch.bfloat16)
    call("mebt_conv3d_ndhwc", xp.data_ptr(), xp.shape[4], _ints(*xp.shape[:4]), w_packed.data_ptr(), int(cin), _ptr(bias),
         _ptr(resid), resid.shape[4] if resid is not None else 0, out.data_ptr(), out.shape[4], _ints(*out.shape[:4]), int(cout),
         _ints(*taps), _ints(*step), _ints(*origin), _ints(*ystep), _ints(*yorigin), _ints(*odims), _stream())
    return out
