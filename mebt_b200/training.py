"""Training step of the MeBT stack on the mebt_b200 kernels (BASELINE.json configs[1]).

Step = stem gather -> stack forward (activations saved) -> fused masked-CE (loss, top-1/5, dlogits in one pass) ->
stack backward (dgrad / wgrad through the MN-major tcgen05 GEMM modes, attention backward, LayerNorm backward) ->
stem scatter-add -> gradient all-reduce (NCCL, launched per finished chunk of blocks so it overlaps the rest of
backward) -> AdamW -> bf16 operand refresh.  This is what `training_step` + `loss.backward()` + DDP + `optimizer.step()`
do in the reference (mebt/transformer.py:717-739, train_transformer.py:39-41).

`TrainState` re-homes every parameter of a `Net2NetTransformer` into ONE flat fp32 buffer (the `nn.Parameter`
objects stay, their `.data` become views), so that
  * query|key|value weights of a block are adjacent and form the fused [3D, D] operand without a copy,
  * the bf16 tensor-core operands are one cast kernel over the whole buffer,
  * gradients live in one flat fp32 buffer whose contiguous per-block slices are the all-reduce buckets.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, ops, parallel
from ._lib import MebtError, call
from .stack import MODE_IDS

_BLOCK_ORDER = ("ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias", "attn.query.weight", "attn.key.weight",
                "attn.value.weight", "attn.query.bias", "attn.key.bias", "attn.value.bias", "attn.proj.weight",
                "attn.proj.bias", "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias")
_FIELD_OF = {"ln1.weight": "ln1_w", "ln1.bias": "ln1_b", "ln2.weight": "ln2_w", "ln2.bias": "ln2_b",
             "attn.query.weight": "w_qkv", "attn.query.bias": "b_qkv", "attn.proj.weight": "w_proj",
             "attn.proj.bias": "b_proj", "mlp.0.weight": "w_fc1", "mlp.0.bias": "b_fc1", "mlp.2.weight": "w_fc2",
             "mlp.2.bias": "b_fc2"}
_BF16_FIELDS = ("w_qkv", "w_proj", "w_fc1", "w_fc2")


class TrainState:
    def __init__(self, model, n_buckets: int = 4):
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise MebtError("TrainState needs the model on a CUDA device (no CPU fallback)")
        self.model = model
        self.device = p0.device
        gpt = model.transformer
        self.modes = [b.mode for b in gpt.blocks]
        for m in self.modes:
            if m not in ("latent_enc", "latent_self", "latent_dec", "lt2l"):
                raise NotImplementedError(f"training supports the four latent block modes, not {m!r}")
        cfg = gpt.config
        # nn.Dropout probabilities of the reference (gpt.py:112-113,154,216); applied only while model.training
        self.embd_pdrop = float(getattr(cfg, "embd_pdrop", 0.0))
        self.resid_pdrop = float(getattr(cfg, "resid_pdrop", 0.0))
        self.attn_pdrop = float(getattr(cfg, "attn_pdrop", 0.0))
        self.dropout_seed = None          # set to an int to pin the masks (tests); default: drawn from torch's RNG per step
        self.D, self.H, self.V = cfg.n_embd, cfg.n_head, gpt.head.weight.shape[0]
        self.L = model.sos_emb.shape[1]
        named = dict(model.named_parameters())
        # Flat layout: block 0 | pad | block 1 | pad | ... | ln_f + head | pad | embeddings | pad.  Every bucket (any run of
        # blocks, the head bucket, the embedding bucket) is padded to a multiple of 8 x the natural alignment of the
        # tensors inside it, so that it splits into 1, 2, 4 or 8 equal shards whose bounds are also bounds of the
        # optimizer's decay-flag blocks (reduce-scatter / sharded AdamW / all-gather, `_exchange_sharded`).  The padding
        # belongs to no parameter and stays zero.
        natural = 1 << 20
        for p_ in named.values():
            while p_.numel() % natural:
                natural >>= 1
        self.align = 8 * max(natural, 4)
        pad_to = lambda n: -(-n // self.align) * self.align
        order, placed = [], {}
        self.block_slices = []
        off = 0
        for i in range(len(self.modes)):
            start = off
            for s_ in _BLOCK_ORDER:
                n_ = f"transformer.blocks.{i}.{s_}"
                order.append(n_)
                placed[n_] = off
                off += named[n_].numel()
            off = pad_to(off)
            self.block_slices.append((start, off))
        head_start = off
        for n_ in ("transformer.ln_f.weight", "transformer.ln_f.bias", "transformer.head.weight"):
            order.append(n_)
            placed[n_] = off
            off += named[n_].numel()
        off = pad_to(off)
        emb_start = off
        for n_ in ("mask_emb", "sos_emb", "pos_emb", "tok_emb.weight"):
            order.append(n_)
            placed[n_] = off
            off += named[n_].numel()
        total = pad_to(off)
        assert set(order) == set(named), set(named) ^ set(order)
        self.flat = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.flat_bf16 = torch.zeros(total, device=self.device, dtype=torch.bfloat16)
        self.offsets = {}
        self._grad_views = []             # (parameter, its view of the flat gradient buffer)
        for n in order:
            p = named[n]
            k = p.numel()
            off = placed[n]
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.flat_grad[off:off + k].view(p.shape)
            self._grad_views.append((p, p.grad))
            self.offsets[n] = (off, k)
        self.order = order
        self.head_slice = (head_start, emb_start)                                       # ln_f + head (+ pad)
        self.emb_slice = (emb_start, total)
        # all-reduce buckets: contiguous groups of blocks, in the order backward finishes them (last block first)
        n_layers = len(self.modes)
        n_buckets = max(1, min(n_buckets, n_layers))
        edges = [round(i * n_layers / n_buckets) for i in range(n_buckets + 1)]
        self.chunks = [(edges[i], edges[i + 1]) for i in range(n_buckets)]
        self.masters_dirty = False        # sharded exchange: other ranks' shards of the fp32 masters are stale here
        self._build_structs()
        self.refresh_operands()
        self._saved = None
        self._bwd_ws = None
        self._ctx = None
        self._pending = False             # a forward whose backward has not run yet
        self._step_id = 0                 # stamps the saved activations: one forward owns them until its backward ran
        self._grads_dirty = False         # the flat gradient buffer holds a gradient no optimizer step / zero_grad consumed
        self._grad_torch_version = self.flat_grad._version
        self.comm_stream = torch.cuda.Stream(device=self.device)
        self.update_ctas = 48             # grid of the background AdamW kernels (train_step(overlap_update=True))
        self._comm_ev = None              # (before, after) events around the main stream's wait for the exchange stream
        self._sharded_last = False
        model.__dict__["_train_state"] = self          # found again by Net2NetTransformer.training_step

    # ---- pointer tables for the C engine ---------------------------------------------------------------------------
    def _view(self, buf, name):
        off, k = self.offsets[name]
        return buf[off:off + k]

    def _build_structs(self):
        n = len(self.modes)
        self.c_layers = (_lib.LayerStruct * n)()
        self.c_grads = (_lib.LayerGradsStruct * n)()
        for i, mode in enumerate(self.modes):
            self.c_layers[i].mode = MODE_IDS[mode]
            for suffix, field in _FIELD_OF.items():
                name = f"transformer.blocks.{i}.{suffix}"
                src = self.flat_bf16 if field in _BF16_FIELDS else self.flat
                setattr(self.c_layers[i], field, self._view(src, name).data_ptr())
                setattr(self.c_grads[i], field, self._view(self.flat_grad, name).data_ptr())

    def _param_versions(self):
        return tuple(p._version for p, _ in self._grad_views)

    def refresh_operands(self):
        """fp32 masters -> bf16 tensor-core operands, one kernel over the flat buffer (after every optimizer step)."""
        self.sync_masters()
        call("mebt_cast_f32_to_bf16", self.flat.data_ptr(), self.flat_bf16.data_ptr(), self.flat.numel() // 4 * 4,
             torch.cuda.current_stream().cuda_stream)
        self._op_version = self._param_versions()

    def relink_grads(self):
        """`optimizer.zero_grad(set_to_none=True)` (torch's default) drops the `.grad` views of the flat gradient buffer;
        put them back so that any torch optimizer sees what backward wrote."""
        for p, g in self._grad_views:
            if p.grad is not g:
                p.grad = g

    # ---- one step ---------------------------------------------------------------------------------------------------
    def _buffers(self, B, NC, NT):
        saved_bytes = _lib.lib.mebt_stack_train_saved_bytes(self.c_layers, len(self.modes), B, self.L, NC, NT, self.D, self.H)
        if self._saved is None or self._saved.numel() < saved_bytes:
            self._saved = torch.empty(int(saved_bytes * 1.1), dtype=torch.uint8, device=self.device)
        ws_bytes = _lib.lib.mebt_stack_backward_workspace_bytes(B, self.L, NC, NT, self.D, self.H)
        if self._bwd_ws is None or self._bwd_ws.numel() < ws_bytes:
            self._bwd_ws = torch.empty(int(ws_bytes * 1.1), dtype=torch.uint8, device=self.device)
        return self._saved, self._bwd_ws

    # stem dropout sites (GPT.forward drops sos_emb, contexts, targets before the blocks, gpt.py:239-241; the fourth
    # dropout, on mask_emb, feeds nothing in the latent modes)
    STEM_SITES = {"lat": 1 << 20, "ctx": (1 << 20) + 1, "tgt": (1 << 20) + 2}

    def _dropout(self):
        """-> (DropoutStruct or None, embd_p, seed) for this step: active only in training mode, like nn.Dropout."""
        if not self.model.transformer.training or max(self.embd_pdrop, self.resid_pdrop, self.attn_pdrop) <= 0.0:
            return None, 0.0, 0
        seed = self.dropout_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # CPU generator: follows torch.manual_seed
        d = _lib.DropoutStruct(self.attn_pdrop, self.resid_pdrop, seed)
        return d, self.embd_pdrop, seed

    def forward(self, x_indices, ctx_idx, tgt_idx, logits_dtype=torch.bfloat16):
        """Stem + stack forward keeping activations.  -> logits [B*NT, V]."""
        m = self.model
        B = x_indices.shape[0]
        NC, NT = ctx_idx.shape[1], tgt_idx.shape[1]
        saved, _ = self._buffers(B, NC, NT)
        if self._param_versions() != self._op_version:          # a torch optimizer (or the user) updated the masters in place
            self.refresh_operands()
        ctx, tgt, lat = ops.embed_gather(x_indices, ctx_idx, tgt_idx, m.tok_emb.weight, m.pos_emb, m.mask_emb, m.sos_emb)
        drop, embd_p, seed = self._dropout()
        if embd_p > 0.0:
            for name, stream_t in (("lat", lat), ("ctx", ctx), ("tgt", tgt)):
                ops.dropout_rows_(stream_t, embd_p, seed, self.STEM_SITES[name])
        logits = torch.empty(B * NT, self.V, device=self.device, dtype=logits_dtype)
        st = torch.cuda.current_stream().cuda_stream
        call("mebt_stack_forward_train_dropout", self.c_layers, len(self.modes),
             self._view(self.flat, "transformer.ln_f.weight").data_ptr(),
             self._view(self.flat, "transformer.ln_f.bias").data_ptr(),
             self._view(self.flat_bf16, "transformer.head.weight").data_ptr(), B, self.L, NC, NT, self.D, self.H, self.V,
             lat.data_ptr(), ctx.data_ptr(), tgt.data_ptr(), logits.data_ptr(), ops._DT[logits_dtype], saved.data_ptr(),
             saved.numel(), ctypes.byref(drop) if drop is not None else None, st)
        self._step_id += 1
        self._pending = True
        self._ctx = (B, NC, NT, x_indices, ctx_idx, tgt_idx, lat, ctx, tgt, drop, embd_p, seed)
        return logits

    def grads_pending(self) -> bool:
        """True when the flat gradient buffer still holds a gradient that no `zero_grad` / optimizer step consumed, i.e.
        the next backward must ADD to it (torch autograd semantics, the reference's `accumulate_grad_batches`,
        train_transformer.py:47-50).  zero_grad(set_to_none=True) shows as dropped `.grad` views, an in-place zero as
        a bump of the buffer's torch version counter; FlatAdamW clears the flag itself."""
        if not self._grads_dirty:
            return False
        if any(p.grad is None for p, _ in self._grad_views) or self.flat_grad._version != self._grad_torch_version:
            self._grads_dirty = False
        return self._grads_dirty

    def backward(self, dlogits, accumulate=False, world_size=1, optimizer=None, sharded=False, fuse=None):
        """Stack + stem backward into the flat gradient buffer; with world_size > 1 each finished chunk of blocks is
        all-reduced (averaged) on a side stream while the next chunk runs.  With a `FlatAdamW` passed as `optimizer` the
        parameter update of a finished chunk is issued on that side stream too, right behind its all-reduce: blocks
        that backward has left are never read again in this step, so their AdamW overlaps the rest of backward.
        fuse: a `FlatAdamW` whose step for the blocks' Linear weights runs INSIDE their weight-gradient GEMMs
        (`mebt_stack_backward_fused`; single GPU, no accumulation, NC > 0); the caller then finishes the step with
        `fuse.step_rest()`."""
        if not self._pending:
            raise MebtError("backward: no forward is pending (each forward's activations serve exactly one backward)")
        B, NC, NT, x_indices, ctx_idx, tgt_idx, lat, ctx, tgt, drop, embd_p, seed = self._ctx
        if dlogits.dtype != torch.bfloat16 or not dlogits.is_contiguous():
            raise MebtError("dlogits must be contiguous bf16 [B*NT, V]")
        m = self.model
        saved, ws = self._buffers(B, NC, NT)
        dev = self.device
        d_lat = torch.empty(B * self.L, self.D, device=dev, dtype=torch.bfloat16)
        d_ctx = torch.empty(B * NC, self.D, device=dev, dtype=torch.bfloat16)
        d_tgt = torch.empty(B * NT, self.D, device=dev, dtype=torch.bfloat16)
        cur = torch.cuda.current_stream()
        n = len(self.modes)
        if not accumulate:
            lo, hi = self.emb_slice
            self.flat_grad[lo:hi].zero_()                       # embedding gradients are scatter-added
        works = []
        fuse_struct = None
        if fuse is not None:
            if accumulate or world_size > 1 or sharded or optimizer is not None or NC == 0:
                raise MebtError("backward(fuse=...): single GPU, no gradient accumulation, NC > 0 only")
            fuse_struct = ctypes.byref(fuse.fused_struct())
        for ci in range(len(self.chunks) - 1, -1, -1):
            lb, le = self.chunks[ci]
            call("mebt_stack_backward_fused", self.c_layers, self.c_grads, n,
                 self._view(self.flat, "transformer.ln_f.weight").data_ptr(),
                 self._view(self.flat_grad, "transformer.ln_f.weight").data_ptr(),
                 self._view(self.flat_grad, "transformer.ln_f.bias").data_ptr(),
                 self._view(self.flat_bf16, "transformer.head.weight").data_ptr(),
                 self._view(self.flat_grad, "transformer.head.weight").data_ptr(), B, self.L, NC, NT, self.D, self.H, self.V,
                 lat.data_ptr(), ctx.data_ptr(), tgt.data_ptr(), dlogits.data_ptr(), saved.data_ptr(), saved.numel(),
                 d_lat.data_ptr(), d_ctx.data_ptr(), d_tgt.data_ptr(), lb, le, int(accumulate),
                 ctypes.byref(drop) if drop is not None else None, fuse_struct, ws.data_ptr(), ws.numel(), cur.cuda_stream)
            lo, hi = self.block_slices[lb][0], self.block_slices[le - 1][1]
            if sharded:
                self._exchange_sharded(optimizer, lo, hi, cur)
                if le == n:
                    self._exchange_sharded(optimizer, *self.head_slice, cur)
                continue
            if world_size > 1:
                works.append(self._all_reduce_async(lo, hi, cur))
                if le == n:
                    works.append(self._all_reduce_async(*self.head_slice, cur))
            if optimizer is not None:
                self._update_async(optimizer, lo, hi, cur, background=ci > 0)
                if le == n:
                    self._update_async(optimizer, *self.head_slice, cur)
        if embd_p > 0.0:                                        # backward of the stem dropout: the same masks
            for name, grad_t in (("lat", d_lat), ("ctx", d_ctx), ("tgt", d_tgt)):
                ops.dropout_rows_(grad_t, embd_p, seed, self.STEM_SITES[name])
        g = lambda name: self._view(self.flat_grad, name)
        ops.embed_backward(x_indices, ctx_idx, tgt_idx, d_ctx, d_tgt, d_lat, g("tok_emb.weight").view(-1, self.D),
                           g("pos_emb").view(-1, self.D), g("mask_emb"), g("sos_emb").view(-1, self.D))
        if sharded:
            self._exchange_sharded(optimizer, *self.emb_slice, cur)
        else:
            if world_size > 1:
                works.append(self._all_reduce_async(*self.emb_slice, cur))
            if optimizer is not None:
                self._update_async(optimizer, *self.emb_slice, cur, background=False)
        if world_size > 1 or optimizer is not None:
            if self._comm_ev is None:
                self._comm_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._comm_ev[0].record(cur)
            cur.wait_stream(self.comm_stream)
            self._comm_ev[1].record(cur)
        self._sharded_last = bool(sharded)
        self.relink_grads()
        self._pending = False
        self._grads_dirty = True
        self._grad_torch_version = self.flat_grad._version
        return works

    def _update_async(self, optimizer, lo, hi, producer_stream, background=True):
        """AdamW + operand refresh of parameters [lo, hi) on the side stream, after the kernels (and the all-reduce,
        which runs on the same stream) that produced their gradients.  background: as a few-CTA kernel that leaves the
        SMs and the L2 to the backward kernels it overlaps (`self.update_ctas`); the last ranges of a step, which
        nothing overlaps, take the full grid."""
        self.comm_stream.wait_stream(producer_stream)
        with torch.cuda.stream(self.comm_stream):
            optimizer.step_range(lo, hi, self.update_ctas if background else 0)

    def _exchange_sharded(self, optimizer, lo, hi, producer_stream):
        """One bucket of the sharded exchange, on the side stream behind the kernels that produced its gradients:
        reduce-scatter (mean over ranks, fp32) -> AdamW + bf16 operand refresh of THIS rank's 1/world shard ->
        all-gather of the bucket's bf16 operands.  What the reference's DDP all-reduce + replicated optimizer.step()
        compute (train_transformer.py:39-41, transformer.py:749-798), with the optimizer's HBM traffic divided by the
        world size and 3/4 of the bytes on the wire.  The fp32 masters of the other ranks' shards go stale here until
        `sync_masters()`; the bf16 operands, which are all the forward / backward read, are complete on every rank."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        a, b = parallel.shard_bounds(lo, hi, rank, world)
        self.comm_stream.wait_stream(producer_stream)
        with torch.cuda.stream(self.comm_stream):
            parallel.reduce_scatter_mean_(self.flat_grad[lo:hi])
            optimizer.step_range(a, b, 0)
            parallel.all_gather_shards_(self.flat_bf16[lo:hi])
        self.masters_dirty = True

    def sync_masters(self):
        """After sharded steps: all-gather the fp32 masters so that `model.parameters()` / `state_dict()` / a rebuilt
        inference pack see every rank's updates (a no-op otherwise).  Collective: every rank must call it."""
        if not self.masters_dirty:
            return
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.comm_stream)
        for lo, hi in parallel.bucket_slices(self.block_slices, self.chunks, self.head_slice, self.emb_slice):
            parallel.all_gather_shards_(self.flat[lo:hi])
        self.masters_dirty = False
        _lib.bump_write_epoch()

    def comm_report(self):
        """Exposed exchange time of the last step: how long the compute stream waited for the exchange stream at the end
        of backward (ms), and the exchange used."""
        if self._comm_ev is None:
            return None
        torch.cuda.synchronize()
        return dict(exposed_ms=self._comm_ev[0].elapsed_time(self._comm_ev[1]),
                    exchange="reduce-scatter / sharded AdamW / bf16 all-gather" if self._sharded_last else "all-reduce")

    def exchange_desc(self, world):
        nb = len(self.chunks) + 2
        return (f"{nb} buckets ({len(self.chunks)} block chunks + head + embeddings): fp32 NCCL reduce-scatter (AVG) -> fused AdamW "
                f"on the 1/{world} shard -> bf16 all-gather of the operands, per bucket on a side stream behind backward")

    def _all_reduce_async(self, lo, hi, producer_stream):
        """One gradient bucket: waits for the kernels that produced it, then averages it over ranks on the side
        stream so that it overlaps the rest of backward (the reference's DDP bucket all-reduce)."""
        self.comm_stream.wait_stream(producer_stream)
        with torch.cuda.stream(self.comm_stream):
            # async_op=False: the host does not block (NCCL collectives are enqueued), but the side stream is made to wait
            # for the collective, so that later work on it (the chunk's AdamW) and `wait_stream` see averaged gradients
            return parallel.allreduce_mean_(self.flat_grad[lo:hi], async_op=False)

    def loss_and_backward(self, x_indices, indices, t=None, world_size=1, defer_backward=False, optimizer=None, sharded=False,
                          fuse=None):
        """shared_step + backward fused: -> dict(loss, acc1, acc5, ratio) as device tensors / floats.
        loss = CE_sum / (B * seq_len * ratio**avg_loss) (mebt/transformer.py:723-730).
        defer_backward=True stops after the loss and returns (dict, dlogits) for a later `backward(dlogits)`
        (the autograd bridge behind Net2NetTransformer.training_step)."""
        m = self.model
        B = x_indices.shape[0]
        x_indices = x_indices.reshape(B, -1)
        t = m._draw_t(t, training=True)                        # uniform over t_range, or the annealed Beta(a, b) of beta configs
        prior_t = m.t_prior(m.t_lengths, m.global_step)
        ctx_idx, tgt_idx, seq_len = m.mask_sampler.divide_indices(indices, t, m.t_lengths, prior_t)
        z_targets = torch.gather(x_indices, 1, tgt_idx)
        NC, NT = ctx_idx.shape[1], tgt_idx.shape[1]
        ratio = float(seq_len - NC) / float(seq_len)
        scale = 1.0 / (B * seq_len * ratio ** m.config.avg_loss)
        logits = self.forward(x_indices, ctx_idx, tgt_idx)
        stats, _ = ops.masked_ce(logits, z_targets.reshape(-1), m.label_smoothing, dlogits=logits, grad_scale=scale)
        n = float(B * NT)
        out = dict(loss=stats[0] * scale, acc1=stats[1] * (100.0 / n), acc5=stats[2] * (100.0 / n), ratio=ratio)
        if defer_backward:
            return out, logits                                  # logits now hold d(loss)/d(logits)
        if fuse is not None and NC == 0:
            fuse = None                                          # the key|value weights of latent_enc get no GEMM: plain step
        self.backward(logits, world_size=world_size, optimizer=optimizer, sharded=sharded, fuse=fuse)
        out["fused_update"] = fuse is not None
        return out

    def make_optimizer(self, lr=1.08e-5, weight_decay=0.01, flat=True):
        """The reference's AdamW (configure_optimizers, transformer.py:749-798: betas (0.9, 0.95), decay only on the
        transformer's Linear weights).  flat=True: one fused kernel over the flat buffers that also refreshes the bf16
        operands (FlatAdamW); flat=False: torch.optim.AdamW(fused=True) on the same parameter groups."""
        self.model.learning_rate, self.model.weight_decay = lr, weight_decay
        ref = self.model.configure_optimizers()                  # the reference's four groups (transformer.py:790-797)
        groups = [{"params": g["params"], "weight_decay": g["weight_decay"]} for g in ref.param_groups]
        if not flat:
            return torch.optim.AdamW(groups, lr=lr, betas=(0.9, 0.95), fused=True)
        return FlatAdamW(self, groups, lr=lr, betas=(0.9, 0.95), weight_decay=weight_decay)

    def train_step(self, optimizer, x_indices, indices, t=None, world_size=1, overlap_update=False, sharded=True,
                   fused_update=False):
        """fwd + loss + bwd (+ all-reduce) + AdamW + operand refresh.  Returns the loss statistics.
        overlap_update=True (FlatAdamW only) issues the update of each finished chunk of blocks on the side stream while
        backward continues.  Measured on B200 at the 16-frame shapes (round 2, 11.3 ms per step without it): as a 48-CTA
        trickle 13.5 ms (the trickle becomes the critical path), as 96 CTAs with 8 chunks 11.85 ms, as short-lived CTAs
        over the whole range (`update_ctas = -1`) 11.25 ms, the same with the step on a high-priority stream 11.8 ms: the
        HBM-bound update slows the latency-bound backward kernels it overlaps by as much as it hides, so it is off by
        default.
        fused_update=True (FlatAdamW, one GPU, no pending accumulated gradient): AdamW of the blocks' Linear weights inside
        their weight-gradient GEMMs (`mebt_stack_backward_fused`); the gradients of those weights are then not
        materialised in `p.grad`.  Measured on B200 at the 16-frame shapes it LOSES 1.3 ms per step (11.37 -> 12.66 ms):
        the epilogue's row-strided 128-byte accesses to p / m / v (one weight row per thread, as the accumulator lies in
        tensor memory) reach a fraction of the 6.2 TB/s the flat kernel streams at, also with the state prefetched into
        L2 a tile ahead and with approximate sqrt / division; off by default, kept as a tested option."""
        if isinstance(optimizer, FlatAdamW) and world_size > 1 and sharded:
            # data parallel with the fused optimizer: reduce-scatter -> AdamW on this rank's shard -> bf16 all-gather
            optimizer.begin_step()
            out = self.loss_and_backward(x_indices, indices, t, world_size, optimizer=optimizer, sharded=True)
            self._grads_dirty = False
            return out
        if isinstance(optimizer, FlatAdamW) and overlap_update:
            optimizer.begin_step()
            return self.loss_and_backward(x_indices, indices, t, world_size, optimizer=optimizer)
        if (isinstance(optimizer, FlatAdamW) and fused_update and world_size == 1 and self.D % 256 == 0
                and not self.grads_pending()):
            # single GPU: the step of the blocks' Linear weights (302 M of the 337 M parameters of STL-16f) runs in the
            # epilogue of their weight-gradient GEMMs; one AdamW launch over everything else finishes the step
            optimizer.begin_step()
            out = self.loss_and_backward(x_indices, indices, t, world_size, fuse=optimizer)
            optimizer.step_rest(fused=out.pop("fused_update"))
            self._grads_dirty = False
            return out
        out = self.loss_and_backward(x_indices, indices, t, world_size)
        optimizer.step()
        self._grads_dirty = False                               # train_step owns the whole step: the gradient is consumed
        if not isinstance(optimizer, FlatAdamW):
            self.refresh_operands()
        return out


class FlatAdamW:
    """AdamW over TrainState's flat fp32 parameter / gradient buffers in ONE kernel (`mebt_adamw_flat`): torch's fused
    AdamW arithmetic, the reference's decay / no-decay split as a per-block flag table, and the bf16 operand copy written
    in the same pass (30 B per parameter instead of 28 + 6 for optimizer + cast).  Exposes `param_groups[i]["lr"]` so
    that the reference's warm-up code (optimizer_step, transformer.py:665-681) can drive it."""

    def __init__(self, ts: TrainState, groups, lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01):
        self.ts = ts
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.param_groups = [dict(g, lr=lr) for g in groups]
        self.m = torch.zeros_like(ts.flat)
        self.v = torch.zeros_like(ts.flat)
        self.steps = 0
        name_of = {id(p): n for n, p in ts.model.named_parameters()}
        decayed = {name_of[id(p)] for g in groups if g["weight_decay"] > 0 for p in g["params"]}
        edges = sorted({o for o, _ in ts.offsets.values()} | {o + k for o, k in ts.offsets.values()})
        shift = 2
        while all(e % (1 << (shift + 1)) == 0 for e in edges) and shift < 20 and (1 << (shift + 1)) <= ts.align // 8:
            shift += 1
        self.shift = shift
        n = ts.flat.numel()
        flags = np.zeros((n + (1 << shift) - 1) >> shift, dtype=np.uint8)
        for name in decayed:
            o, k = ts.offsets[name]
            flags[o >> shift:(o + k) >> shift] = 1
        self.flags = torch.from_numpy(flags).to(ts.device)
        # fused mode: bit 1 marks what the weight-gradient epilogues update (the Linear weights of every block that can
        # reach the logits: blocks behind the last latent_dec get zero gradients through the plain kernel)
        last = max((i for i, md in enumerate(ts.modes) if md == "latent_dec"), default=-1)
        fflags = flags.copy()
        for i in range(last + 1):
            for suffix in ("attn.query.weight", "attn.key.weight", "attn.value.weight", "attn.proj.weight", "mlp.0.weight",
                           "mlp.2.weight"):
                o, k = ts.offsets[f"transformer.blocks.{i}.{suffix}"]
                fflags[o >> shift:(o + k) >> shift] |= 2
        self.flags_fused = torch.from_numpy(fflags).to(ts.device)
        if n % 4:
            raise MebtError("FlatAdamW: the flat parameter buffer must hold a multiple of 4 elements")

    def zero_grad(self, set_to_none=False):
        self.ts.flat_grad.zero_()
        self.ts._grads_dirty = False

    def begin_step(self):
        self.steps += 1

    def step_range(self, lo, hi, max_ctas=0):
        """The update of parameters [lo, hi) of the flat buffer (tensor-aligned bounds) for the step opened by
        `begin_step`, on the current stream.  max_ctas > 0: as a background kernel of that many CTAs."""
        ts = self.ts
        if lo % (1 << self.shift) or (hi - lo) % 4:
            raise MebtError("FlatAdamW.step_range: bounds must be tensor boundaries of the flat buffer")
        lr = float(self.param_groups[0]["lr"])
        call("mebt_adamw_flat_bg", ts.flat.data_ptr() + 4 * lo, ts.flat_grad.data_ptr() + 4 * lo, self.m.data_ptr() + 4 * lo,
             self.v.data_ptr() + 4 * lo, ts.flat_bf16.data_ptr() + 2 * lo, self.flags.data_ptr() + (lo >> self.shift),
             self.shift, hi - lo, lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.steps,
             int(max_ctas), torch.cuda.current_stream().cuda_stream)
        _lib.bump_write_epoch()                # the masters changed behind torch's version counters: derived operands are stale
        ts._grads_dirty = False

    def step(self):
        self.begin_step()
        self.step_range(0, self.ts.flat.numel())

    def fused_struct(self):
        """mebt_fused_adamw_t of the step opened by `begin_step` (TrainState.backward(fuse=self))."""
        ts = self.ts
        self._fused = _lib.FusedAdamwStruct(ts.flat_grad.data_ptr(), ts.flat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                            ts.flat_bf16.data_ptr(), self.flags.data_ptr(), self.shift,
                                            float(self.param_groups[0]["lr"]), self.betas[0], self.betas[1], self.eps,
                                            self.weight_decay, self.steps)
        return self._fused

    def step_rest(self, fused=True):
        """Finishes a step whose block Linear weights were updated inside the backward: one launch over the flat buffers
        that skips them (bit 1 of the flag table).  fused=False: the backward could not fuse (NC = 0): the whole update."""
        ts = self.ts
        if not fused:
            return self.step_range(0, ts.flat.numel())
        lr = float(self.param_groups[0]["lr"])
        call("mebt_adamw_flat_bg", ts.flat.data_ptr(), ts.flat_grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
             ts.flat_bf16.data_ptr(), self.flags_fused.data_ptr(), self.shift, ts.flat.numel(), lr, self.betas[0], self.betas[1],
             self.eps, self.weight_decay, self.steps, 0, torch.cuda.current_stream().cuda_stream)
        _lib.bump_write_epoch()
        ts._grads_dirty = False

    def state_dict(self):
        return dict(m=self.m, v=self.v, steps=self.steps, lr=[g["lr"] for g in self.param_groups])

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.steps = int(sd["steps"])
        for g, lr in zip(self.param_groups, sd["lr"]):
            g["lr"] = lr


class TrainStepFunction(torch.autograd.Function):
    """Autograd bridge for the drop-in `training_step`: forward runs stem + stack + fused masked CE on the CUDA engine
    and returns the loss; `loss.backward()` runs the engine's backward, which writes every parameter gradient straight
    into `p.grad` (views of the flat gradient buffer) and, under torch.distributed, averages them over ranks the way
    the reference's DDP does (train_transformer.py:39-41).  Like autograd, a backward ADDS to gradients that no
    `zero_grad()` / optimizer step has consumed since the previous backward (gradient accumulation)."""

    @staticmethod
    def forward(ctx, anchor, ts, x_indices, indices, world_size):
        out, dlogits = ts.loss_and_backward(x_indices, indices, world_size=world_size, defer_backward=True)
        ctx.ts, ctx.dlogits, ctx.world_size, ctx.step_id = ts, dlogits, world_size, ts._step_id
        ctx.mark_non_differentiable(out["acc1"], out["acc5"])
        return out["loss"].reshape(()), out["acc1"], out["acc5"]

    @staticmethod
    def backward(ctx, g_loss, g_acc1, g_acc5):
        dlogits = ctx.dlogits
        if dlogits is None:
            raise MebtError("training_step: backward through the same step twice")
        ts = ctx.ts
        if ts._step_id != ctx.step_id or not ts._pending:
            raise MebtError("training_step: another forward ran before this loss's backward; the saved activations belong "
                            "to the later step (call loss.backward() before the next training_step)")
        dlogits.mul_(g_loss.to(dlogits.dtype))                   # 1.0 unless the caller scaled the loss
        ts.backward(dlogits, accumulate=ts.grads_pending(), world_size=ctx.world_size)
        ctx.dlogits = None
        return None, None, None, None, None
