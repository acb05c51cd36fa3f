"""Noise draws of the samplers.

By default the Exp(1) / normal draws come from torch's generator on the tensor's device, in the order the
reference issues them (`exponential_()` over [B,NT,V] per sampling step, `randn_like` / `exponential_()` over
[B,NT] per re-masking step), so a seeded run consumes the same RNG stream as the reference on that device.
Tests install a hook to feed the exact draws the CPU oracle consumed.
"""
from __future__ import annotations

import torch

_hook = None


def set_hook(fn):
    """fn(kind: 'exponential' | 'randn', shape, device) -> fp32 tensor on `device`; None restores the default."""
    global _hook
    _hook = fn


def exponential(shape, device):
    if _hook is not None:
        return _hook("exponential", tuple(shape), device)
    return torch.empty(tuple(shape), dtype=torch.float32, device=device).exponential_()


def randn(shape, device):
    if _hook is not None:
        return _hook("randn", tuple(shape), device)
    return torch.randn(tuple(shape), dtype=torch.float32, device=device)
