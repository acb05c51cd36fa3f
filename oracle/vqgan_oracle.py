"""CPU restatement (torch fp32) of the reference's VQGAN encoder / decoder forward (mebt/vqgan.py:82-93, 263-405).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py and bench.py's CPU legs, never by mebt_b200/.
Pinned against the unmodified reference by tests/golden/vqgan_small.npz (tests/test_oracle_golden.py).

Everything is functional over a `state_dict`-shaped dict of tensors with the reference's key names.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def make_weights(keys_shapes, seed, scale=None):
    """Per-tensor seeded recipe, independent of module construction order: conv / convt weights ~ N(0, 1/sqrt(fan_in)),
    biases ~ N(0, 0.1), norm weights ~ 1 + N(0, 0.1), norm biases ~ N(0, 0.1), codebook rows ~ N(0, 1)."""
    out = {}
    for i, (k, shape) in enumerate(sorted(keys_shapes.items())):
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(shape, generator=g)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in k or "final_block" in k:
            out[k] = (1.0 if k.endswith("weight") else 0.0) + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("bias"):
            out[k] = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("weight"):
            fan = shape[1] * math.prod(shape[2:]) if "convt" not in k else shape[0] * math.prod(shape[2:]) / 8
            out[k] = torch.randn(shape, generator=g) / math.sqrt(fan)
        elif k == "codebook.N":
            out[k] = torch.zeros(shape)
        else:
            out[k] = torch.randn(shape, generator=g)
    return out


def _same_pad(kernel, stride):
    pad = []
    for k, s in list(zip(kernel, stride))[::-1]:          # F.pad starts from the last dimension (vqgan.py:369-373)
        p = k - s
        pad += [p // 2 + p % 2, p // 2]
    return pad


def same_pad_conv3d(P, name, x, stride=(1, 1, 1)):
    """SamePadConv3d.forward (vqgan.py:380-381)."""
    w = P[f"{name}.conv.weight"]
    return F.conv3d(F.pad(x, _same_pad(w.shape[2:], stride), mode="replicate"), w, P.get(f"{name}.conv.bias"), stride=stride)


def same_pad_convt3d(P, name, x, stride):
    """SamePadConvTranspose3d.forward (vqgan.py:403-404)."""
    w = P[f"{name}.convt.weight"]
    k = w.shape[2:]
    return F.conv_transpose3d(F.pad(x, _same_pad(k, stride), mode="replicate"), w, P.get(f"{name}.convt.bias"), stride=stride,
                              padding=tuple(kk - 1 for kk in k))


def normalize(P, name, x):
    """Normalize (vqgan.py:255-260): GroupNorm(32, eps 1e-6), or eval-mode (Sync)BatchNorm when running stats are present."""
    if f"{name}.running_mean" in P:
        return F.batch_norm(x, P[f"{name}.running_mean"], P[f"{name}.running_var"], P[f"{name}.weight"], P[f"{name}.bias"],
                            False, 0.0, 1e-5)
    return F.group_norm(x, 32, P[f"{name}.weight"], P[f"{name}.bias"], 1e-6)


def silu(x):
    return x * torch.sigmoid(x)


def res_block(P, name, x):
    """ResBlock.forward (vqgan.py:342-356), in == out channels (the only form Encoder / Decoder build)."""
    h = same_pad_conv3d(P, f"{name}.conv1", silu(normalize(P, f"{name}.norm1", x)))
    h = same_pad_conv3d(P, f"{name}.conv2", silu(normalize(P, f"{name}.norm2", h)))
    return x + h


def _strides(downsample):
    n = [int(math.log2(d)) for d in downsample]
    out = []
    for _ in range(max(n)):
        out.append(tuple(2 if d > 0 else 1 for d in n))
        n = [d - 1 for d in n]
    return out


def encoder(P, x, downsample, prefix="encoder"):
    """Encoder.forward (vqgan.py:293-300)."""
    h = same_pad_conv3d(P, f"{prefix}.conv_first", x)
    for i, st in enumerate(_strides(downsample)):
        h = same_pad_conv3d(P, f"{prefix}.conv_blocks.{i}.down", h, st)
        h = res_block(P, f"{prefix}.conv_blocks.{i}.res", h)
    return silu(normalize(P, f"{prefix}.final_block.0", h))


def decoder(P, x, upsample, prefix="decoder"):
    """Decoder.forward (vqgan.py:326-334)."""
    h = silu(normalize(P, f"{prefix}.final_block.0", x))
    for i, st in enumerate(_strides(upsample)):
        h = same_pad_convt3d(P, f"{prefix}.conv_blocks.{i}.up", h, st)
        h = res_block(P, f"{prefix}.conv_blocks.{i}.res1", h)
        h = res_block(P, f"{prefix}.conv_blocks.{i}.res2", h)
    return same_pad_conv3d(P, f"{prefix}.conv_last", h)


def pre_quant(P, x, downsample):
    """VQGAN.encode up to the codebook (vqgan.py:83): pre_vq_conv(encoder(x))."""
    return same_pad_conv3d(P, "pre_vq_conv", encoder(P, x, downsample))


def decode(P, encodings, downsample):
    """VQGAN.decode (vqgan.py:90-93)."""
    h = F.embedding(encodings, P["codebook.embeddings"])
    h = same_pad_conv3d(P, "post_vq_conv", h.permute(0, 4, 1, 2, 3))
    return decoder(P, h, downsample)
