"""Drop-in for `mebt.transformer` (reference: mebt/transformer.py): `Net2NetTransformer`, the maskgit /
draft-and-revise samplers and the module-level sampling helpers, running on the mebt_b200 CUDA kernels.

Same constructor, method signatures, return values and state_dict keys as the reference.  Host-side control
flow (time-step loops, RNG draw order, float32 mask-size arithmetic) follows the reference so that, given the
same seeds, masks and index tensors are identical; the per-step device work is:
    K1 stem gather -> layer stack (K2 GEMMs + K3 attention) -> K4 head -> K6 sample -> K8 scatter [-> K7 re-mask]
"""
from __future__ import annotations

import argparse
import copy
import random

import numpy as np
import torch
import torch.nn as nn

from . import ops, rng
from ._lib import MebtError
from .mask_sampler import MaskGen
from .modules.gpt import GPT
from .utils import get_obj_from_str

try:                                       # Lightning is optional: the shell (L4 in SURVEY.md) stays host code
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:                          # noqa: BLE001
    class _Base(nn.Module):
        """Minimal stand-in exposing what Net2NetTransformer uses from LightningModule."""
        global_step = 0
        current_epoch = 0

        def save_hyperparameters(self, *args, **kwargs):
            pass

        def log(self, *args, **kwargs):
            pass

        @property
        def device(self):
            return next(self.parameters()).device


def disabled_train(self, mode=True):
    """Freezes train/eval mode of the first-stage model (transformer.py:20-23)."""
    return self


# ---- video-length priors (resolved by name from config.t_prior, transformer.py:25-49,125) ---------------------
def uniform(vid_lengths, t):
    return np.ones_like(vid_lengths, dtype=float)


def _length_gaussian(vid_lengths, t, b, c):
    centre = (vid_lengths - 1) * b
    return np.exp(-((t - centre) ** 2) / (2 * (b * c) ** 2))


def gaussian(vid_lengths, t, b, c):
    return _length_gaussian(vid_lengths, t, b, c)


def gaussian100000_2(vid_lengths, t):
    return _length_gaussian(vid_lengths, t, 100000, 2)


def gaussian2(vid_lengths, t):
    return _length_gaussian(vid_lengths, t, 30000, 2)


def longest(vid_lengths, t):
    x = np.zeros_like(vid_lengths, dtype=float)
    x[-1] = 1.
    return x


# ---- context-temperature schedules (resolved by name from `ctemp_schedule`, transformer.py:51-58,440) ----------
def linear(t):
    return 1. - t


def constant(t):
    return 1.


def cosine(t):
    return np.cos(t * np.pi / 2.)


_T_PRIORS = dict(uniform=uniform, gaussian2=gaussian2, gaussian100000_2=gaussian100000_2, longest=longest)
_CTEMP_SCHEDULES = dict(linear=linear, constant=constant, cosine=cosine)


def _instantiate(config):
    target = config["target"].replace("tats.", "mebt.")
    if target.startswith("mebt."):
        target = "mebt_b200." + target[len("mebt."):]
    return get_obj_from_str(target)(**config.get("params", dict()))


def _refuse_ddp_wrapper(module):
    """The CUDA backward writes parameter gradients straight into `p.grad` and averages them over ranks itself, so no
    autograd hook of a DistributedDataParallel wrapper ever fires: a wrapped module would stop at DDP's 'Expected to
    have finished reduction in the prior iteration' on the second step (or reduce twice with find_unused_parameters).
    The reference's Lightning `DDPStrategy` (train_transformer.py:41) wraps the module - detect it and say what to
    use instead: one process per GPU with torch.distributed initialised and a non-wrapping strategy."""
    trainer = getattr(module, "_trainer", None) or module.__dict__.get("trainer")
    wrapped = getattr(getattr(trainer, "strategy", None), "model", None) if trainer is not None else None
    if isinstance(wrapped, torch.nn.parallel.DistributedDataParallel):
        raise MebtError(
            "mebt_b200.training_step: the module is wrapped in DistributedDataParallel (Lightning DDPStrategy). The CUDA "
            "backward all-reduces its own gradient buckets over torch.distributed; run one process per GPU with "
            "init_process_group('nccl') and a single-device strategy (see INTEGRATION.md, 'Data-parallel training')")


class Net2NetTransformer(_Base):
    def __init__(self, transformer_config, first_stage_config, mask_config, ckpt_path=None, ignore_keys=[],
                 first_stage_key="video", cond_stage_key="label", pkeep=1.0, sos_token=0):
        super().__init__()
        cfg = self.config = transformer_config
        self.class_cond_dim = cfg.class_cond_dim if hasattr(cfg, "class_cond_dim") else None
        self.be_unconditional = cfg.unconditional
        self.sos_token = sos_token
        self.first_stage_key = first_stage_key
        self.first_stage_vocab_size = cfg.vocab_size
        self.cond_stage_key = cond_stage_key
        self.vtokens = cfg.vtokens
        self.n_embd = cfg.n_embd
        self.vis_epoch = cfg.vis_epoch
        for name, default in (("avg_loss", 0.0), ("embd_pdrop", 0.0), ("resid_pdrop", 0.0), ("attn_pdrop", 0.0)):
            if not hasattr(cfg, name):
                setattr(cfg, name, default)
        cfg.avg_loss = float(cfg.avg_loss)
        self.sample_every_n_latent_frames = getattr(cfg, "sample_every_n_latent_frames", 0) \
            if hasattr(cfg, "sample_every_n_latent_frames") else 0
        self.label_smoothing = cfg.label_smoothing if hasattr(cfg, "label_smoothing") else 0.0

        self.init_first_stage_from_ckpt(first_stage_config)
        self.init_cond_stage_from_ckpt(cfg)

        gpt_vocab_size = self.first_stage_vocab_size + self.cond_stage_vocab_size
        self.transformer = GPT(gpt_vocab_size, cfg.block_size, n_layer=cfg.n_layer, n_head=cfg.n_head, n_embd=cfg.n_embd,
                               vtokens_pos=cfg.vtokens_pos, n_unmasked=cfg.n_unmasked, attn_pdrop=cfg.attn_pdrop,
                               embd_pdrop=cfg.embd_pdrop, resid_pdrop=cfg.resid_pdrop, mode=cfg.mode)
        self.mask_sampler = _instantiate(mask_config)

        if not hasattr(cfg, "beta_params"):
            mp = mask_config["params"] if "params" in mask_config else {}
            self.range = mp["t_range"] if "t_range" in mp else [0., 1.]
            self.beta = False
        else:
            self.beta_params = cfg.beta_params
            self.beta_iter = float(cfg.beta_iter)
            self.beta = True
        self.t_lengths = np.arange(self.mask_sampler.shape[0]) + 1

        if not hasattr(cfg, "t_prior"):
            cfg.t_prior = "longest"
        self.t_prior = _T_PRIORS[cfg.t_prior] if cfg.t_prior in _T_PRIORS else eval(cfg.t_prior)
        self.tok_emb = nn.Embedding(gpt_vocab_size, cfg.n_embd)
        self.tok_emb.weight.data.normal_(mean=0.0, std=0.02)
        self.mask_emb = nn.Parameter(torch.zeros(1, 1, cfg.n_embd))
        self.mask_emb.data.normal_(mean=0.0, std=0.02)
        if not hasattr(cfg, "sos_emb"):
            cfg.sos_emb = 1
        if cfg.sos_emb > 0:
            self.sos_emb = nn.Parameter(torch.zeros(1, cfg.sos_emb, cfg.n_embd))
            self.sos_emb.data.normal_(mean=0.0, std=0.02)
        else:
            raise NotImplementedError("mebt_b200: the latent-bottleneck stack needs sos_emb > 0 learned latents")
        self.num_pos = np.prod(self.mask_sampler.shape[1:])
        self.pos_emb = nn.Parameter(torch.zeros(1, cfg.block_size, cfg.n_embd))
        self.pos_emb.data.normal_(mean=0.0, std=0.02)
        self.n_head = cfg.n_head
        self.first_mode = cfg.mode[0]
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)
        self.pkeep = pkeep
        # Exp(1) noise of the Gumbel-max sampler: "torch" draws exponential_() from torch's device generator in the
        # reference's order (parity); "philox" generates it inside the kernel (no [B,NT,V] noise tensor in HBM).
        self.rng_mode = "torch"
        self.rng_seed = 0
        self._rng_offset = 0
        # dtype of the logits the head writes for the internal sampling / loss paths
        self.logits_dtype = torch.float32
        # ... and for the samplers' own forward -> sample steps (draft / revise / sample / entp_sample), whose logits never
        # leave the library: None = bf16 under precision "bf16" (half the head GEMM's store and the sampling kernel's read;
        # the rounding, 2^-9 relative, is below the bf16 forward's own 1e-2), fp32 under precision "fp32"
        self.sampler_logits_dtype = None
        # draft / revise steps with in-kernel noise: draw the token in the head GEMM's epilogue, no logits in HBM
        self.fused_head_sampling = True
        self.selected_probs_supported = True      # sample(debug=True, debug_probs="selected"), see mebt_b200.pipelines
        # "bf16": tcgen05 engine, logits within 1e-2 of the fp32 reference; "fp32": split-GEMM + fp32 attention path,
        # within 1e-4 (north_star tolerances).  Also settable from the config (`precision: fp32`).
        self.precision = str(getattr(transformer_config, "precision", "bf16"))
        self.save_hyperparameters()

    # ---- checkpoint / stage plumbing (transformer.py:170-214) ---------------------------------------------------
    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                print("Deleting key {} from state_dict.".format(k))
                del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    def init_first_stage_from_ckpt(self, config):
        """transformer.py:180-192: with `vtokens: False` the frozen VQGAN of `first_stage_config.params.ckpt_path` (a
        Lightning checkpoint of mebt.vqgan.VQGAN) encodes the videos to tokens; with `vtokens: True` the data already are
        token grids and no first stage is loaded."""
        if not self.vtokens:
            from .vqgan import load_vqgan
            params = config["params"] if isinstance(config, dict) else config.params
            ckpt = params["ckpt_path"] if isinstance(params, dict) else params.ckpt_path
            self.first_stage_model = load_vqgan(ckpt, device="cpu")           # moves with the module (.to / .cuda)
            for p in self.first_stage_model.parameters():
                p.requires_grad = False
            self.first_stage_model.codebook._need_init = False
            self.first_stage_model.eval()
            self.first_stage_model.train = disabled_train.__get__(self.first_stage_model)
            self.first_stage_vocab_size = self.first_stage_model.codebook.n_codes
            return
        self.first_stage_model = None
        self.first_stage_vocab_size = 16384

    def init_cond_stage_from_ckpt(self, args):
        if not self.be_unconditional:
            raise ValueError("conditional model %s is not implemented (nor is it in the reference)" % self.cond_stage_key)
        self.cond_stage_key = self.first_stage_key
        self.cond_stage_model = None
        self.cond_stage_vocab_size = 0

    # ---- device path ----------------------------------------------------------------------------------------------
    def _logits_rows(self, x_indices, context_indices, target_indices, logits_dtype=None):
        """K1 + stack + head: int64 ids/indices -> logits [B*NT, V]."""
        if not x_indices.is_cuda:
            raise MebtError("mebt_b200 runs on CUDA tensors only (no CPU fallback); move the model and inputs to cuda")
        B = x_indices.shape[0]
        if self.precision not in ("bf16", "fp32"):
            raise MebtError(f"precision must be 'bf16' or 'fp32', not {self.precision!r}")
        stream_dtype = torch.float32 if self.precision == "fp32" else torch.bfloat16
        ctx, tgt, lat = ops.embed_gather(x_indices, context_indices, target_indices, self.tok_emb.weight, self.pos_emb,
                                         self.mask_emb, self.sos_emb, out_dtype=stream_dtype)
        return self.transformer.forward_rows(B, lat, ctx, tgt, logits_dtype or self.logits_dtype)

    def _sample_rows(self, logits_rows, temperature, top_k, top_p, return_probs=False):
        noise = None
        seed = offset = 0
        if self.rng_mode == "philox":
            self._rng_offset += 1
            seed, offset = self.rng_seed, self._rng_offset
        else:
            noise = rng.exponential(logits_rows.shape, logits_rows.device)
        return ops.sample_logits(logits_rows, temperature, top_k, top_p, noise=noise, seed=seed, offset=offset,
                                 return_probs=return_probs)

    def _draw_t(self, t, training):
        """The masking time of one step (transformer.py:225-242): given, or uniform over `t_range`, or - for configs
        with `beta_params` - Beta(a, b) annealed towards Beta(1, 1) over `beta_iter` steps."""
        if t is not None:
            return torch.tensor(t)
        if training and self.beta:
            if self.global_step > self.beta_iter:
                a, b = 1., 1.
            else:
                a0, b0 = self.beta_params
                a = a0 - (a0 - 1.) * (self.global_step / self.beta_iter)
                b = b0 - (b0 - 1.) * (self.global_step / self.beta_iter)
            return torch.distributions.beta.Beta(a, b).sample()
        t = torch.tensor(random.random())                          # python RNG: identical on every DDP rank
        if training:
            t = self.range[0] + t * (self.range[1] - self.range[0])
        return t

    def forward(self, x, c, t=None, indices=None, vid_t=None, debug=False):
        """One masked-prediction step -> (logits [B,NT,V] fp32, z_targets, NT_weight, seq_len) (transformer.py:216-286)."""
        assert indices is not None
        _, x_indices = self.encode_to_z(x)
        B = x_indices.shape[0]
        t = self._draw_t(t, self.training or debug)
        if vid_t is None:
            prior_t = self.t_prior(self.t_lengths, self.global_step)
            vid_t = self.t_lengths
        else:
            assert len(vid_t) == 1
            prior_t = np.ones_like(vid_t, dtype=float)
        context_indices, target_indices, seq_len = self.mask_sampler.divide_indices(indices, t, vid_t, prior_t, debug)
        z_targets = torch.gather(x_indices, 1, target_indices)
        NC, NT = context_indices.shape[1], target_indices.shape[1]
        NT_weight = float(seq_len - NC)
        logits = self._logits_rows(x_indices, context_indices, target_indices, torch.float32)
        return logits.view(B, NT, -1), z_targets, NT_weight, seq_len

    def reconstruct_mask(self, x_indices, context_indices, target_indices, debug=False):
        """logits for given context/target index sets -> (logits [B,NT,V] fp32, None) (transformer.py:288-324)."""
        B = x_indices.shape[0]
        x_indices = x_indices.reshape(B, -1)
        logits = self._logits_rows(x_indices, context_indices, target_indices, torch.float32)
        return logits.view(B, target_indices.shape[1], -1), None

    def top_k_logits(self, logits, k):
        return top_k_logits(logits, k)

    def on_train_epoch_start(self):
        pass

    def on_validation_epoch_start(self):
        if (self.current_epoch + 1) % self.vis_epoch == 0 and self.first_stage_model is not None:
            saved = copy.deepcopy(self.mask_sampler.schedule)
            self.mask_sampler.schedule = "cosine"
            shape = (4, *self.mask_sampler.shape)
            x = torch.zeros(shape, dtype=torch.long, device=self.device)
            x = self.sample(x, None, 1.0, None, None, 32, None, None, context_temperature=6.0, skips=False)[0]
            frames = torch.cat([self.first_stage_model.decode(x.reshape(*shape)[i:i + 1]) for i in range(4)], 0)
            frames = (frames.clamp(-0.5, 0.5) + 0.5).permute(0, 2, 1, 3, 4)
            self.logger.experiment.add_video("sample", frames, self.current_epoch, fps=20)
            self.logger.experiment.flush()
            self.mask_sampler.schedule = saved

    # ---- samplers ---------------------------------------------------------------------------------------------------
    def _initial_masks(self, x, context_indices, target_indices, clone=False):
        B, N = x.shape
        if context_indices is None:
            context_indices = torch.empty(B, 0, dtype=torch.long, device=x.device)
            target_indices = torch.arange(N, device=x.device).repeat(B, 1)
        elif clone:
            context_indices, target_indices = context_indices.clone(), target_indices.clone()
        return context_indices, target_indices

    def _predict_and_write(self, partial, ctx_idx, tgt_idx, temperature, top_k, top_p, want_probs=False, want_scores=True):
        """One sampler step: forward, sample every target, write the ids back.  -> (ids, scores, probs)
        want_scores=False (the gibbs passes of draft / revise, which use only the ids) with in-kernel noise, no top-k /
        top-p and the bf16 engine takes the FUSED form: the head GEMM's epilogue draws the token (Gumbel-max over the
        fp32 accumulators, `mebt_stack_forward_sample`) and no logits are materialised (`fused_head_sampling`)."""
        B = partial.shape[0]
        NT = tgt_idx.shape[1]
        if (self.fused_head_sampling and not want_scores and not want_probs and self.rng_mode == "philox" and not top_k
                and top_p is None and self.precision == "bf16" and temperature > 0 and NT > 0
                and "_logits_rows" not in self.__dict__):
            if not partial.is_cuda:
                raise MebtError("mebt_b200 runs on CUDA tensors only (no CPU fallback); move the model and inputs to cuda")
            ctx, tgt, lat = ops.embed_gather(partial, ctx_idx, tgt_idx, self.tok_emb.weight, self.pos_emb, self.mask_emb,
                                             self.sos_emb, out_dtype=torch.bfloat16)
            self._rng_offset += 1
            ids = self.transformer.sample_rows(B, lat, ctx, tgt, float(temperature), self.rng_seed, self._rng_offset)
            ops.scatter_ids(partial, tgt_idx, ids.view(B, NT))
            return ids.view(B, NT), None, None
        dt = self.sampler_logits_dtype or (torch.bfloat16 if self.precision == "bf16" else torch.float32)
        logits = self._logits_rows(partial, ctx_idx, tgt_idx, dt)
        ids, scores, probs = self._sample_rows(logits, temperature, top_k, top_p, return_probs=want_probs)
        NT = tgt_idx.shape[1]
        ops.scatter_ids(partial, tgt_idx, ids.view(B, NT))
        return ids.view(B, NT), scores.view(B, NT), (probs.view(B, NT, -1) if probs is not None else None)

    @torch.no_grad()
    def sample(self, x, c, temperature=1.0, top_k=None, top_p=None, n_steps=8, context_indices=None, target_indices=None,
               strategy="maskgit", context_temperature=4.5, phase_history=None, refine_steps=1, forget_pivot=False,
               skips=[False, False, False], debug=False, ctemp_schedule="linear", edit=False, debug_probs="dense"):
        """Iterative maskgit-style decoding with confidence re-masking (transformer.py:353-447).
        debug=True returns, like the reference, the dense [B, N, 16384] fp32 map of the last predictive distribution of
        every position (2 GB at B = 32, 16 frames; 17 GB at 128 frames).  debug_probs="selected" (an extension; what the
        sampling scripts actually consume, sample_vqgan_transformer_videos.py:85-89) returns instead the [B, N] map of the
        probability of the token that was sampled at each position's last prediction (-1 where never predicted): the
        gather of the dense map at the final code, without the dense map."""
        B = x.shape[0]
        N = int(np.prod(x.shape[1:]))
        edit_N = target_indices.shape[1] if edit else N
        assert not self.transformer.training
        if strategy not in ("maskgit", "random", "mlm", "bootstrap"):
            return None
        partial = x.reshape(B, N).clone()
        context_indices, target_indices = self._initial_masks(partial, context_indices, target_indices, clone=True)
        ctemp_fn = _CTEMP_SCHEDULES[ctemp_schedule] if ctemp_schedule in _CTEMP_SCHEDULES else eval(ctemp_schedule)
        history, context_history, partial_probs = [], [], None
        sparse = debug and debug_probs == "selected"
        if debug:
            history.append(partial.clone())
            partial_probs = -torch.ones((B, N) if sparse else (B, N, 16384), device=x.device)
        self.mask_sampler.rng_mode, self.mask_sampler.rng_seed = self.rng_mode, self.rng_seed + 1
        for t_next in np.linspace(0, 1, n_steps + 1)[1:]:
            t = torch.full((B,), fill_value=t_next, device=x.device)              # float32
            # computed on the device like the reference, then read back ONCE per step (the skip test and the re-mask size
            # below both use the host copy: one sync per step instead of two)
            n_masked_toks = torch.ceil(self.mask_sampler.schedule_fn(t) * edit_N).cpu()
            if int((n_masked_toks > target_indices.shape[-1]).sum()) == B:
                continue                                                           # context already larger than asked
            target_indices = target_indices.view(B, -1)
            _, scores, probs = self._predict_and_write(partial, context_indices, target_indices, temperature, top_k,
                                                       top_p, want_probs=debug and not sparse)
            if sparse:
                partial_probs.scatter_(1, target_indices, scores)
            elif debug:
                partial_probs.scatter_(1, target_indices.unsqueeze(-1).expand(-1, -1, probs.shape[-1]), probs)
                history.append(partial.clone())
                context_history.append(context_indices)
            actual_temperature = context_temperature * ctemp_fn(t_next)
            context_indices, target_indices = self.mask_sampler.generate_next_mask(
                context_indices, target_indices, scores, t_next, strategy=strategy,
                context_temperature=actual_temperature, n_masked_toks=n_masked_toks)
        if debug:
            return partial.view(B, -1), context_indices, target_indices, history, context_history, partial_probs
        return partial.view(B, -1), context_indices, target_indices

    @torch.no_grad()
    def entp_sample(self, x, c, temperature=1.0, top_k=None, top_p=None, n_steps=8, context_indices=None,
                    target_indices=None, strategy="maskgit", context_temperature=4.5, phase_history=None, refine_steps=1,
                    forget_pivot=False, skips=[False, False, False], debug=False, ctemp_schedule="linear"):
        """Entropy-scored variant (transformer.py:449-542): tokens whose predictive distribution has the lowest
        entropy-like score are revealed first; re-masking uses context temperature 0."""
        B = x.shape[0]
        N = int(np.prod(x.shape[1:]))
        assert not self.transformer.training
        if strategy == "ar":
            raise NotImplementedError
        if strategy not in ("maskgit", "random", "mlm", "bootstrap"):
            return None
        partial = x.reshape(B, N).clone()
        context_indices, target_indices = self._initial_masks(partial, context_indices, target_indices, clone=True)
        history, context_history, partial_probs = [], [], None
        if debug:
            history.append(partial.clone())
            partial_probs = -torch.ones(B, N, 16384, device=x.device)
        for t_next in np.linspace(0, 1, n_steps + 1)[1:]:
            t = torch.full((B,), fill_value=t_next, device=x.device)
            n_masked_toks = torch.ceil(self.mask_sampler.schedule_fn(t) * N).cpu()
            if int((n_masked_toks > target_indices.shape[-1]).sum()) == B:
                continue
            target_indices = target_indices.view(B, -1)
            _, _, probs = self._predict_and_write(partial, context_indices, target_indices, temperature, top_k, top_p,
                                                  want_probs=True)
            scores = -(-probs + torch.log(probs + 1e-8)).sum(-1)
            scores = scores.max(-1, keepdim=True)[0] - scores
            if debug:
                partial_probs.scatter_(1, target_indices.unsqueeze(-1).expand(-1, -1, probs.shape[-1]), probs)
                history.append(partial.clone())
                context_history.append(context_indices)
            context_indices, target_indices = self.mask_sampler.generate_next_mask_entp(
                context_indices, target_indices, scores, t_next, strategy=strategy, context_temperature=0.0)
        if debug:
            return partial.view(B, -1), context_indices, target_indices, history, context_history, partial_probs
        return partial.view(B, -1), context_indices, target_indices

    def _gibbs_pass(self, x, temperature, top_k, top_p, n_steps, context_indices, target_indices, make_masks):
        B = x.shape[0]
        N = int(np.prod(x.shape[1:]))
        partial = x.reshape(B, N).clone()
        context_indices, target_indices = self._initial_masks(partial, context_indices, target_indices)
        ctxs, tgts = make_masks(context_indices, target_indices, n_steps, x.device)
        assert not self.transformer.training
        for ctx_idx, tgt_idx in zip(ctxs, tgts):
            self._predict_and_write(partial, ctx_idx, tgt_idx.view(B, -1), temperature, top_k, top_p, want_scores=False)
        return partial.view(B, -1)

    @torch.no_grad()
    def draft(self, x, c, temperature=1.0, top_k=None, top_p=None, n_steps=8, debug=False, context_indices=None,
              target_indices=None):
        """Draft phase: step i conditions on the first i/n of a random order and re-predicts all the rest
        (transformer.py:544-586)."""
        return self._gibbs_pass(x, temperature, top_k, top_p, n_steps, context_indices, target_indices,
                                self.mask_sampler.create_gibbs_draft_mask)

    @torch.no_grad()
    def revise(self, x, c, temperature=1.0, top_k=None, top_p=None, n_steps=8, debug=False, context_indices=None,
               target_indices=None):
        """Revise phase: each of n disjoint random groups is re-predicted given all other tokens
        (transformer.py:588-630)."""
        return self._gibbs_pass(x, temperature, top_k, top_p, n_steps, context_indices, target_indices,
                                self.mask_sampler.create_gibbs_revise_mask)

    @torch.no_grad()
    def draft_and_revise(self, x, c, n_draft=8, draft_t=1.0, draft_k=None, draft_p=None, n_revise=8, revise_t=1.0,
                         revise_k=None, revise_p=None, M=2, skip_draft=False, debug=False, context_indices=None,
                         target_indices=None, edit=False):
        """Draft once, then revise M times (transformer.py:632-663)."""
        B = x.shape[0]
        N = int(np.prod(x.shape[1:]))
        x = x.reshape(B, N)
        assert not self.transformer.training
        if not skip_draft:
            x = self.draft(x, c, draft_t, draft_k, draft_p, n_draft, debug, context_indices, target_indices)
        if edit:
            context_indices = target_indices = None
        for _ in range(M):
            x = self.revise(x, c, revise_t, revise_k, revise_p, n_revise, debug, context_indices, target_indices)
        return x.view(B, -1)

    # ---- training shell (transformer.py:665-798) ---------------------------------------------------------------
    def optimizer_step(self, epoch_nb, batch_nb, optimizer, optimizer_i, opt_closure, on_tpu=False,
                       using_native_amp=False, using_lbfgs=False):
        step = self.trainer.global_step
        lr_scale = 1.
        if step < self.warmup_steps:
            lr_scale = min(1., float(step + 1) / self.warmup_steps)
        elif self.cosine_lr:
            rad = float(step - self.warmup_steps) / float(self.trainer.max_steps - self.warmup_steps)
            assert rad >= 0
            lr_scale = 0.5 * (1 + np.cos(rad * np.pi))
        if step < self.warmup_steps or self.cosine_lr:
            for pg in optimizer.param_groups:
                pg["lr"] = self.learning_rate * lr_scale
        self.log("learning_rate", self.learning_rate * lr_scale, logger=True, on_step=True, sync_dist=True)
        optimizer.step(closure=opt_closure)

    @torch.no_grad()
    def encode_to_z(self, x):
        if self.vtokens:
            return x, x.reshape(x.shape[0], -1)
        emb, targets = self.first_stage_model.encode(x, include_embeddings=True)
        if self.sample_every_n_latent_frames > 0:
            emb = emb[:, :, ::self.sample_every_n_latent_frames]
            targets = targets[:, ::self.sample_every_n_latent_frames]
        return emb.movedim(1, -1).contiguous(), targets.reshape(targets.shape[0], -1)

    @torch.no_grad()
    def encode_to_c(self, c):
        quant_c, indices = self.cond_stage_model.encode(c, include_embeddings=True)
        if len(indices.shape) > 2:
            indices = indices.view(c.shape[0], -1)
        return quant_c, indices

    def get_input(self, key, batch):
        return batch[key]

    def get_xc(self, batch, N=None):
        x = self.get_input(self.first_stage_key, batch)
        c = self.get_input(self.cond_stage_key, batch)
        if N is not None:
            x, c = x[:N], c[:N]
        return x, c

    def shared_step(self, batch, batch_idx):
        """-> (acc1, acc5, loss, ratio); loss = CE_sum / (B * seq_len * ratio**avg_loss) (transformer.py:717-732).
        The logits feed the fused K5 kernel (loss + top-1/top-5 ranks in one pass) instead of three passes."""
        x, c = self.get_xc(batch)
        indices = self.get_input("indices", batch)
        logits, target, NT_weight, seq_len = self(x, c, indices=indices)
        ratio = NT_weight / float(seq_len)
        B, NT, V = logits.shape
        stats, _ = ops.masked_ce(logits.view(B * NT, V), target.reshape(-1), self.label_smoothing)
        weight = ratio ** self.config.avg_loss
        loss = stats[0] / (B * seq_len * weight)
        n = float(B * NT)
        return stats[1:2] * (100.0 / n), stats[2:3] * (100.0 / n), loss, ratio

    def training_step(self, batch, batch_idx):
        """-> loss with an autograd edge into the CUDA backward (transformer.py:734-739): `loss.backward()` fills every
        `p.grad`; any optimizer over `self.parameters()` then applies (the bf16 operand copies follow the masters)."""
        if next(self.parameters()).is_cuda and torch.is_grad_enabled():
            from .training import TrainState, TrainStepFunction
            ts = self.__dict__.get("_train_state") or TrainState(self)
            x, c = self.get_xc(batch)
            indices = self.get_input("indices", batch)
            import torch.distributed as dist
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            _refuse_ddp_wrapper(self)
            loss, acc1, acc5 = TrainStepFunction.apply(self.mask_emb, ts, x.reshape(x.shape[0], -1), indices, world)
        else:
            acc1, acc5, loss, ratio = self.shared_step(batch, batch_idx)
        for name, v in (("train/loss", loss), ("train/acc1", acc1), ("train/acc5", acc5)):
            self.log(name, v, prog_bar=True, logger=True, on_step=True, on_epoch=True, sync_dist=True)
        return loss

    # After data-parallel steps with the sharded exchange (TrainState.train_step, world > 1) each rank holds fresh fp32
    # masters only for its own shard; anything that reads the parameters themselves first gathers them (collective:
    # like the reference's DDP, every rank enters eval / state_dict together).
    def _sync_masters(self):
        ts = self.__dict__.get("_train_state")
        if ts is not None and ts.masters_dirty:
            ts.sync_masters()

    def train(self, mode: bool = True):
        if not mode:
            self._sync_masters()
        return super().train(mode)

    def state_dict(self, *args, **kwargs):
        self._sync_masters()
        return super().state_dict(*args, **kwargs)

    def validation_step(self, batch, batch_idx):
        acc1, acc5, loss, ratio = self.shared_step(batch, batch_idx)
        for name, v in (("val/loss", loss), ("val/acc1", acc1), ("val/acc5", acc5)):
            self.log(name, v, prog_bar=True, logger=True, on_step=True, on_epoch=True, sync_dist=True)
        return loss

    def configure_optimizers(self):
        """AdamW(betas=(0.9, 0.95)) with weight decay on the transformer's Linear weights only; four parameter
        groups in the reference's order: decayed weights, *_emb except pos_emb, biases + LayerNorm, pos_emb
        (transformer.py:749-798)."""
        decay, no_decay = set(), set()
        for mn, m in self.transformer.named_modules():
            for pn, _ in m.named_parameters(recurse=False):
                full = f"{mn}.{pn}" if mn else pn
                if pn.endswith("bias") or isinstance(m, (nn.LayerNorm, nn.Embedding)):
                    no_decay.add(full)
                elif pn.endswith("weight") and isinstance(m, nn.Linear):
                    decay.add(full)
        params = dict(self.transformer.named_parameters())
        assert not (decay & no_decay) and not (params.keys() - (decay | no_decay))
        emb = {n: p for n, p in self.named_parameters() if "_emb" in n and n != "pos_emb"}
        pos = {n: p for n, p in self.named_parameters() if "pos_emb" in n}
        groups = [
            {"params": [params[n] for n in sorted(decay)], "weight_decay": self.weight_decay},
            {"params": list(emb.values()), "weight_decay": 0.0},
            {"params": [params[n] for n in sorted(no_decay)], "weight_decay": 0.0},
            {"params": list(pos.values()), "weight_decay": 0.0},
        ]
        return torch.optim.AdamW(groups, lr=self.learning_rate, betas=(0.9, 0.95))

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
        add = parser.add_argument
        add("--vqvae", type=str, help="path to vqvae ckpt, or model name to download pretrained")
        add("--stft_vqvae", type=str, help="path to vqgan ckpt, or model name to download pretrained")
        add("--unconditional", action="store_true")
        add("--base_lr", type=float, default=4.5e-06)
        add("--vocab_size", type=int, default=16384)
        add("--first_stage_vocab_size", type=int, default=16384)
        add("--block_size", type=int, default=256)
        add("--n_layer", type=int, default=48)
        add("--n_head", type=int, default=24)
        add("--n_embd", type=int, default=1536)
        add("--n_unmasked", type=int, default=0)
        add("--sample_every_n_latent_frames", type=int, default=0)
        add("--first_stage_key", type=str, default="video", choices=["video"])
        add("--cond_stage_key", type=str, default="label", choices=["label", "text", "stft"])
        add("--iid", action="store_true")
        add("--schedule", type=str, default="cosine")
        add("--max_token", type=int, default=1024)
        add("--method", type=str, default=None)
        return parser


# ---- module-level sampling helpers (transformer.py:826-910) ---------------------------------------------------------
def gumbel_sort(prob):
    """Indices (*, C) of `prob` in the order of an exponential race: sort_desc((p / sum p) / q), q ~ Exp(1)."""
    shape = prob.shape
    flat = prob.reshape(-1, shape[-1]).float()
    return MaskGen.gumbel_top_k(flat, 1.0).view(shape)


def sample_from_logits(logits, temperature=1.0, top_k=None, top_p=None, return_probs=False):
    """Sample one id per row of `logits` (*, V) -> ids (*) [, probs (*, V) = softmax before renormalisation].
    One fused kernel (K6); the Exp(1) noise is torch's exponential_() over the logits' shape, like the reference."""
    shape = logits.shape
    rows = logits.reshape(-1, shape[-1])
    if rows.dtype not in (torch.float32, torch.bfloat16):
        rows = rows.float()
    noise = rng.exponential(rows.shape, rows.device)
    ids, _, probs = ops.sample_logits(rows.contiguous(), temperature, top_k, top_p, noise=noise, return_probs=return_probs)
    if return_probs:
        return ids.view(shape[:-1]), probs.view(shape)
    return ids.view(shape[:-1])


def top_k_logits(logits, k):
    """Everything below the k-th largest logit of a row becomes -inf (API helper; the samplers use the fused K6)."""
    v, _ = torch.topk(logits, k)
    out = logits.clone()
    out[out < v[..., [-1]]] = -float("Inf")
    return out


def top_p_probs(probs, p):
    """Nucleus filter + renormalisation (API helper, transformer.py:898-910)."""
    sp, si = torch.sort(probs, dim=-1, descending=True)
    drop = torch.cumsum(sp, dim=-1) >= p
    drop[..., 1:] = drop[..., :-1].clone()
    drop[..., 0] = 0
    probs = probs.masked_fill(drop.scatter(-1, si, drop), 0.0)
    return probs / torch.sum(probs, dim=-1, keepdim=True)
