"""Drop-in for `mebt.mask_sampler.MaskGen` (reference: mebt/mask_sampler.py).

Mask bookkeeping is host-side integer / float32 arithmetic and stays in torch ops, written to be
arithmetically identical to the reference (float32 schedule -> ceil; CPU-generator randperm).  The one
data-dependent device step — the confidence sort of `generate_next_mask` — runs the K7 kernel
(`mebt_remask_sort`).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import ops, rng
from ._lib import MebtError

_SCHEDULES = ("cosine", "linear", "quadratic", "sqrt", "square", "cube", "ar", "cosine_plus", "convex")
_METHODS = ("iid", "mlm", "partial_mlm", "block", "ar", "phase", "grid", "frame", "interpolate", "stochastic_phase",
            "stochastic_grid", "fdm", "softfdm", "clip")


class MaskGen(nn.Module):
    """Mask generator for a prescribed noise schedule (mask_sampler.py:9-32)."""

    _available_schedules = list(_SCHEDULES)
    _available_methods = list(_METHODS)

    def __init__(self, iid=False, schedule="cosine", max_token=256, method=None, shape=(4, 16, 16), t_range=(0., 1.),
                 budget=1024):
        super().__init__()
        if schedule not in _SCHEDULES:
            raise ValueError(f"Unsupported schedule: {schedule}")
        self.method = method if method is not None else ("iid" if iid else "mlm")
        if self.method not in _METHODS:
            raise ValueError(f"Unsupported method: {self.method}")
        self.schedule = schedule
        self.device = None
        self.shape = shape
        self.seq_len = np.prod(shape)
        self.max_token = max_token
        self.dense = True
        self.range = t_range if t_range is not None else (0., 1.)
        self.budget = budget
        # how the Exp(1) draw of gumbel_top_k is produced: "torch" = torch's device generator (the reference's
        # RNG stream), "philox" = in-kernel Philox4x32 keyed by (rng_seed, rng_offset)
        self.rng_mode = "torch"
        self.rng_seed = 0
        self.rng_offset = 0

    # ---- schedules (mask_sampler.py:34-67); t is a float32 tensor, arithmetic stays float32 -------------
    @staticmethod
    def cosine(t):
        return torch.cos(0.5 * np.pi * t)

    @staticmethod
    def cosine_plus(t):
        return 0.5 * (1 + torch.cos(np.pi * t))

    @staticmethod
    def linear(t):
        return 1.0 - t

    @staticmethod
    def quadratic(t):
        return (1.0 - t) ** 2.0

    @staticmethod
    def square(t):
        return 1.0 - t ** 2.0

    @staticmethod
    def cube(t):
        return 1.0 - t ** 3.0

    @staticmethod
    def sqrt(t):
        return 1.0 - t ** 0.5

    @staticmethod
    def convex(t):
        return (1.0 - t) ** 3.0

    @property
    def schedule_fn(self):
        return getattr(self, self.schedule)

    # ---- context / target split (mask_sampler.py:75-115) -------------------------------------------------
    def divide_indices(self, indices, t, vid_t, prior_t, debug=False):
        ratio = self.schedule_fn(t)
        slicing = self.training or debug
        if slicing:
            frames = self.shape[0]
            per_frame = int(np.prod(self.shape[1:]))
            prior_t = prior_t / prior_t.sum()
            T = np.random.choice(vid_t, p=prior_t)                       # numpy RNG, as the reference
            if frames != T:
                first = np.random.randint(0, frames - T + 1)
                lo, hi = first * per_frame, (first + T) * per_frame
                # keep permutation entries that fall inside the frame window, order and absolute positions kept
                rows = [row[(row >= lo) & (row < hi)] for row in indices]
                indices = torch.stack(rows).to(indices.device)
        seq_len = int(np.prod(indices.shape[1:]))
        n_masked = torch.ceil(ratio * seq_len).to(dtype=torch.long)
        n_ctx = seq_len - n_masked
        budget = self.budget if slicing else seq_len
        n_tgt = min(budget, seq_len - n_ctx)
        # NB `-0:` keeps everything, exactly like the reference (mask_sampler.py:114)
        return indices[:, :n_ctx], indices[:, -n_tgt:], seq_len

    def sample_mlm_mask(self, shape, ratios, max_token=None):
        """mask_sampler.py:117-144 (not used by Net2NetTransformer; kept for API parity)."""
        assert ratios.shape[0] == shape[0]
        seq_len = int(np.prod(shape[1:]))
        n_masked = torch.ceil(ratios[0] * seq_len).to(dtype=torch.long)
        n_ctx = int(seq_len - n_masked)
        n_tgt = int(min(self.budget if self.training else seq_len, seq_len - n_ctx))
        assert n_ctx + n_tgt <= seq_len
        ctx = torch.zeros(shape[0], n_ctx, dtype=torch.long, device=ratios.device)
        tgt = torch.zeros(shape[0], n_tgt, dtype=torch.long, device=ratios.device)
        for b in range(shape[0]):
            perm = torch.randperm(seq_len)
            ctx[b] = perm[:n_ctx]
            tgt[b] = perm[-n_tgt:]
        return ctx, tgt

    def sample_mask(self, shape, ratios, max_token, debug=False, context_ratios=None, method=None):
        if max_token is None:
            max_token = self.max_token
        if method in ("iid", "ar"):
            raise NotImplementedError
        if method in ("mlm", "maskgit", "random"):
            return self.sample_mlm_mask(shape, ratios, max_token)
        raise UnboundLocalError(f"method {method!r} produces no mask (as in the reference)")

    def forward(self, shape, t=None, max_token=None, device=None, debug=False, method=None):
        B = shape[0]
        method = self.method if method is None else method
        if t is None:
            t = torch.rand(B, device=device)
            if self.training:
                t = self.range[0] + t * (self.range[1] - self.range[0])
        if isinstance(t, float):
            t = torch.full((B,), fill_value=t, device=device)
        return self.sample_mask(shape, self.schedule_fn(t), max_token, debug=debug, method=method)

    # ---- confidence re-masking (mask_sampler.py:178-246) ------------------------------------------------
    def _exp_noise(self, like: torch.Tensor):
        if self.rng_mode == "philox":
            self.rng_offset += 1
            return None, self.rng_seed, self.rng_offset
        return rng.exponential(like.shape, like.device), 0, 0

    @staticmethod
    def _order(prob, context_temperature, noise, seed, offset):
        if not prob.is_cuda:
            raise MebtError("MaskGen.gumbel_top_k runs the K7 CUDA kernel; CPU tensors are not supported")
        B, NT = prob.shape
        empty = torch.empty(B, 0, dtype=torch.long, device=prob.device)
        ident = torch.arange(NT, device=prob.device).repeat(B, 1)
        return ops.remask_sort(prob, empty, ident, 0, float(context_temperature), noise=noise, seed=seed,
                               offset=offset, want_order=True)[2]

    @staticmethod
    def gumbel_top_k(prob, context_temperature=1.0):
        """Full descending order of (p / sum p) / q**ctemp with q ~ Exp(1) drawn from torch's device generator,
        like the reference's exponential_() (mask_sampler.py:178-187)."""
        q = rng.exponential(prob.shape, prob.device)
        return MaskGen._order(prob, context_temperature, q, 0, 0)

    def generate_next_mask(self, context_indices, target_indices, score, t, strategy="maskgit", context_temperature=4.5,
                           n_masked_toks=None, debug=False):
        if score is not None and strategy != "ar":
            assert target_indices.shape == score.shape
        B, NC = context_indices.shape
        NT = target_indices.shape[1]
        if strategy == "ar":
            next_ctx = torch.cat([context_indices, target_indices[:, :1]], 1)
            next_tgt = target_indices[:, 1:]
            locs = torch.zeros_like(target_indices[:, :1])
            return (next_ctx, next_tgt, locs) if debug else (next_ctx, next_tgt)
        if strategy not in ("maskgit", "random", "mlm", "bootstrap"):
            return None
        if isinstance(t, float):
            t = torch.full((B,), fill_value=t, device=score.device)
        if strategy in ("random", "bootstrap"):
            score = rng.randn(score.shape, score.device)            # drawn before the early return, as the reference
            context_temperature = 0.0
        seq_len = NC + NT
        if n_masked_toks is None:
            n_masked = int(torch.ceil(self.schedule_fn(t)[0] * seq_len).to(dtype=torch.long))
        else:
            n_masked = int(n_masked_toks[0].long())
        if strategy == "bootstrap":
            n_masked = NT - 1
        n_ctx = seq_len - n_masked
        if n_ctx <= NC:
            return (context_indices, target_indices, None) if debug else (context_indices, target_indices)
        n_new = n_ctx - NC
        if not score.is_cuda:
            raise MebtError("MaskGen.generate_next_mask runs the K7 CUDA kernel; CPU tensors are not supported")
        noise, seed, offset = self._exp_noise(score)
        next_ctx, next_tgt, order = ops.remask_sort(score, context_indices, target_indices, n_new,
                                                    float(context_temperature), noise=noise, seed=seed, offset=offset,
                                                    want_order=debug)
        if debug:
            return next_ctx, next_tgt, order[:, :n_new]
        return next_ctx, next_tgt

    def generate_next_mask_entp(self, context_indices, target_indices, score, t, strategy="maskgit",
                                context_temperature=4.5, n_masked_toks=None, debug=False):
        """Entropy variant (mask_sampler.py:248-303): same selection rule; only 'random' replaces the scores."""
        if strategy == "bootstrap":
            # the reference keeps the supplied scores for bootstrap here and forces n_masked = NT - 1
            B, NC = context_indices.shape
            NT = target_indices.shape[1]
            fake = torch.full((B,), float(NT - 1), device=score.device)
            return self.generate_next_mask(context_indices, target_indices, score, t, "maskgit", context_temperature,
                                           fake, debug)
        if n_masked_toks is not None and not torch.is_tensor(n_masked_toks):
            n_masked_toks = torch.tensor([float(n_masked_toks)])
        return self.generate_next_mask(context_indices, target_indices, score, t, strategy, context_temperature,
                                       n_masked_toks, debug)

    # ---- draft-and-revise partitions (mask_sampler.py:317-356); randperm from the CPU generator ----------
    @staticmethod
    def _shuffle_targets(target_indices, device, n_steps):
        B = target_indices.shape[0]
        N = int(np.prod(target_indices.shape[1:]))
        assert N % n_steps == 0
        perms = torch.stack([torch.randperm(N) for _ in range(B)]).to(device)
        return torch.gather(target_indices, 1, perms), N

    @staticmethod
    def create_gibbs_revise_mask(context_indices, target_indices, num_unit_gibbs_steps, device):
        n = num_unit_gibbs_steps
        shuffled, N = MaskGen._shuffle_targets(target_indices, device, n)
        m = N // n
        ctx = torch.stack([torch.cat([context_indices, shuffled[:, (i + 1) * m:], shuffled[:, :i * m]], 1)
                           for i in range(n)])
        tgt = torch.stack([shuffled[:, i * m:(i + 1) * m] for i in range(n)])
        return ctx, tgt

    @staticmethod
    def create_gibbs_draft_mask(context_indices, target_indices, num_unit_gibbs_steps, device):
        n = num_unit_gibbs_steps
        shuffled, N = MaskGen._shuffle_targets(target_indices, device, n)
        m = N // n
        ctx = [torch.cat([context_indices, shuffled[:, :i * m]], 1) for i in range(n)]
        tgt = [shuffled[:, i * m:] for i in range(n)]
        return ctx, tgt
