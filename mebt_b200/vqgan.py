"""Drop-in for the inference side of `mebt.vqgan` (reference: mebt/vqgan.py): `VQGAN.encode` / `VQGAN.decode` with the
3-D convolutional `Encoder` / `Decoder` around the codebook, and the checkpoint / token-file formats of the sampling scripts.

The codebook lines (`Codebook.forward` = K9 + K10, `F.embedding` + `shift_dim` = K10 with a channel-first store) were the
hot-path rows; this module adds SURVEY.md 8(f) rank 4: the convolutions run as implicit GEMMs on the tensor cores
(`csrc/conv3d.cu`: 5-D TMA boxes of a replicate-padded channels-last activation, no im2col buffer), GroupNorm / eval
BatchNorm + SiLU + the replicate padding are one elementwise pass in front of each convolution, the ResBlock skip is an
epilogue operand, and a transposed convolution is one launch per output parity writing interleaved positions.

Module and parameter names follow the reference so that its checkpoints load (`encoder.conv_first.conv.weight`,
`decoder.conv_blocks.0.up.convt.weight`, ...): the `nn.Conv3d` / `nn.ConvTranspose3d` / `nn.GroupNorm` members only HOLD the
parameters; their torch forward is never called.  VQGAN training (discriminators, LPIPS, EMA codebook) is out of scope.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import MebtError, write_epoch
from .modules.codebook import Codebook


def _ceil(n, m):
    return -(-n // m) * m


def to_channels_last(x, ld=None):
    """[B, C, T, H, W] (any float dtype) -> bf16 [B, T, H, W, ld] with the channels zero-padded to a multiple of 8."""
    B, C = x.shape[:2]
    ld = ld or _ceil(C, 8)
    out = torch.zeros(B, *x.shape[2:], ld, device=x.device, dtype=torch.bfloat16)
    out[..., :C] = x.permute(0, 2, 3, 4, 1)
    return out


def to_channels_first(x, C):
    """bf16 [B, T, H, W, ld] -> fp32 [B, C, T, H, W]."""
    return x[..., :C].permute(0, 4, 1, 2, 3).float().contiguous()


class _Packed:
    """bf16 GEMM operand(s) derived from a parameter, rebuilt when the parameter changes."""

    def __init__(self):
        self.key, self.value = None, None

    def get(self, params, build):
        key = tuple((p.data_ptr(), p._version) for p in params) + (write_epoch(),)
        if self.key != key:
            self.key, self.value = key, build()
        return self.value


def _same_pad(kernel_size, stride):
    """pad_input of SamePadConv3d / SamePadConvTranspose3d (vqgan.py:368-374): (before, after) per dimension, t h w order."""
    out = []
    for k, s in zip(kernel_size, stride):
        p = k - s
        out.append((p // 2 + p % 2, p // 2))
    return out


class Normalize(nn.Module):
    """`Normalize(in_channels, norm_type)` (vqgan.py:255-260) as a parameter holder + the (norm, gamma, beta) triple the
    pad_norm_act kernel takes.  'group': GroupNorm(32, C, eps=1e-6); 'batch': SyncBatchNorm in eval mode, folded."""

    def __new__(cls, in_channels, norm_type="group"):
        assert norm_type in ("group", "batch")
        if norm_type == "group":
            m = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        else:
            m = nn.BatchNorm3d(in_channels)          # same parameters / buffers as SyncBatchNorm
        return m


def norm_args(m):
    if isinstance(m, nn.GroupNorm):
        return dict(norm=1, groups=m.num_groups, eps=m.eps, gamma=m.weight.detach().float().contiguous(),
                    beta=m.bias.detach().float().contiguous())
    scale = (m.weight / torch.sqrt(m.running_var + m.eps)).detach().float().contiguous()
    shift = (m.bias - m.running_mean * scale).detach().float().contiguous()
    return dict(norm=2, gamma=scale, beta=shift)


def silu(x):
    return x * torch.sigmoid(x)


class SiLU(nn.Module):
    def forward(self, x):
        return silu(x)


class SamePadConv3d(nn.Module):
    """vqgan.py:358-381.  `forward_cl` takes / returns channels-last bf16; `pre` = norm_args(...) + act of the
    Normalize / SiLU in front of the convolution (fused into the padding pass), `resid` = tensor added in the epilogue."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=True, padding_type="replicate"):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size,) * 3
        if isinstance(stride, int):
            stride = (stride,) * 3
        if padding_type != "replicate":
            raise NotImplementedError(f"mebt_b200.SamePadConv3d: padding_type {padding_type!r} (only 'replicate', the default)")
        self.kernel_size, self.stride = tuple(kernel_size), tuple(stride)
        self.pads = _same_pad(kernel_size, stride)
        self.pad_input = sum([self.pads[2], self.pads[1], self.pads[0]], tuple())       # F.pad order, as the reference stores it
        self.padding_type = padding_type
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=0, bias=bias)
        self._packed = _Packed()
        self._packed_w = _Packed()

    def _operands(self):
        conv = self.conv

        def build():
            w = conv.weight.detach().float()                                 # [Cout, Cin, kt, kh, kw]
            co, ci = w.shape[:2]
            cp, cop = _ceil(ci, 64), _ceil(co, 8)
            wp = torch.zeros(_ceil(cop, 64), *w.shape[2:], cp, device=w.device)     # rows in whole 64-wide tiles (no TMA fill)
            wp[:co, ..., :ci] = w.permute(0, 2, 3, 4, 1)
            b = torch.zeros(cop, device=w.device)
            if conv.bias is not None:
                b[:co] = conv.bias.detach().float()
            return wp.reshape(wp.shape[0], -1).to(torch.bfloat16).contiguous(), b.contiguous()
        return self._packed.get([conv.weight] + ([conv.bias] if conv.bias is not None else []), build)

    def _window(self, x):
        """Few-channel input (the RGB video), unit strides, output rows of 128 positions: the taps along w are packed
        into one 64-element K slice (8 positions x 8 channels; `mebt_conv3d_ndhwc` window mode)."""
        return (self.conv.in_channels <= 8 and x.shape[4] == 8 and self.stride == (1, 1, 1) and self.kernel_size[2] <= 8
                and x.shape[3] % 128 == 0 and self.conv.out_channels <= 64)

    def _operands_window(self):
        conv = self.conv

        def build():
            w = conv.weight.detach().float()                                 # [Cout, Cin, kt, kh, kw]
            co, ci, kt, kh, kw = w.shape
            cop = _ceil(co, 8)
            wp = torch.zeros(_ceil(cop, 64), kt, kh, 64, device=w.device)
            for dw in range(kw):
                wp[:co, :, :, dw * 8:dw * 8 + ci] = w[..., dw].permute(0, 2, 3, 1)
            b = torch.zeros(cop, device=w.device)
            if conv.bias is not None:
                b[:co] = conv.bias.detach().float()
            return wp.reshape(wp.shape[0], -1).to(torch.bfloat16).contiguous(), b.contiguous()
        return self._packed_w.get([conv.weight] + ([conv.bias] if conv.bias is not None else []), build)

    def forward_cl(self, x, pre=None, act=0, resid=None):
        pads = self.pads
        if self._window(x):
            w, b = self._operands_window()
            kt, kh, kw = self.kernel_size
            xp = ops.pad_norm_act(x, (pads[0][0], pads[0][1], pads[1][0], pads[1][1], pads[2][0], pads[2][1] + 8 - kw), act=act,
                                  **(pre or {}))
            return ops.conv3d_ndhwc(xp, w, 64, b.shape[0], (kt, kh, 1), (1, 1, 1), tuple(x.shape[1:4]), bias=b, resid=resid)
        w, b = self._operands()
        xp = ops.pad_norm_act(x, (pads[0][0], pads[0][1], pads[1][0], pads[1][1], pads[2][0], pads[2][1]), act=act,
                              **(pre or {}))
        odims = tuple(d // s for d, s in zip(x.shape[1:4], self.stride))
        return ops.conv3d_ndhwc(xp, w, self.conv.in_channels, b.shape[0], self.kernel_size, self.stride, odims, bias=b,
                                resid=resid)

    def forward(self, x):
        """[B, Cin, T, H, W] -> [B, Cout, T', H', W'] (fp32), like the reference module."""
        return to_channels_first(self.forward_cl(to_channels_last(x)), self.conv.out_channels)


class SamePadConvTranspose3d(nn.Module):
    """vqgan.py:384-405: ConvTranspose3d(kernel 4, stride s, padding 3) over the replicate-padded input = per output
    parity p of an up-sampled dimension a 2-tap stride-1 convolution out[2j + p] = sum_d xp[j + p + d] w[3 - p - 2d]
    (a dimension with stride 1 keeps all four taps, flipped)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=True, padding_type="replicate"):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size,) * 3
        if isinstance(stride, int):
            stride = (stride,) * 3
        if padding_type != "replicate":
            raise NotImplementedError(f"mebt_b200.SamePadConvTranspose3d: padding_type {padding_type!r}")
        if tuple(kernel_size) != (4, 4, 4) or any(s not in (1, 2) for s in stride):
            raise NotImplementedError("mebt_b200.SamePadConvTranspose3d: kernel 4, strides 1 or 2 (what the Decoder builds)")
        self.kernel_size, self.stride = tuple(kernel_size), tuple(stride)
        self.pads = _same_pad(kernel_size, stride)
        self.pad_input = sum([self.pads[2], self.pads[1], self.pads[0]], tuple())
        self.padding_type = padding_type
        self.convt = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=stride, bias=bias,
                                        padding=tuple(k - 1 for k in kernel_size))
        self._packed = _Packed()

    def _operands(self):
        convt = self.convt

        def build():
            w = convt.weight.detach().float()                                # [Cin, Cout, kt, kh, kw]
            ci, co = w.shape[:2]
            cp, cop = _ceil(ci, 64), _ceil(co, 8)
            b = torch.zeros(cop, device=w.device)
            if convt.bias is not None:
                b[:co] = convt.bias.detach().float()
            classes = []
            parities = [range(2) if s == 2 else range(1) for s in self.stride]
            for pt in parities[0]:
                for ph in parities[1]:
                    for pw in parities[2]:
                        par = (pt, ph, pw)
                        idx = [[3 - p - 2 * d for d in range(2)] if s == 2 else [3 - d for d in range(4)]
                               for p, s in zip(par, self.stride)]
                        sub = w[:, :, idx[0]][:, :, :, idx[1]][:, :, :, :, idx[2]]          # [Cin, Cout, nt, nh, nw]
                        wp = torch.zeros(_ceil(cop, 64), *sub.shape[2:], cp, device=w.device)
                        wp[:co, ..., :ci] = sub.permute(1, 2, 3, 4, 0)
                        classes.append((par, tuple(sub.shape[2:]), wp.reshape(wp.shape[0], -1).to(torch.bfloat16).contiguous()))
            return classes, b.contiguous()
        return self._packed.get([convt.weight] + ([convt.bias] if convt.bias is not None else []), build)

    def forward_cl(self, x, pre=None, act=0):
        classes, b = self._operands()
        pads = self.pads
        xp = ops.pad_norm_act(x, (pads[0][0], pads[0][1], pads[1][0], pads[1][1], pads[2][0], pads[2][1]), act=act,
                              **(pre or {}))
        B = x.shape[0]
        odims = tuple(x.shape[1:4])
        cop = b.shape[0]
        out = torch.empty(B, *(d * s for d, s in zip(odims, self.stride)), cop, device=x.device, dtype=torch.bfloat16)
        for par, taps, w in classes:
            ops.conv3d_ndhwc(xp, w, self.convt.in_channels, cop, taps, (1, 1, 1), odims, bias=b, out=out, origin=par,
                             ystep=self.stride, yorigin=par)
        return out

    def forward(self, x):
        return to_channels_first(self.forward_cl(to_channels_last(x)), self.convt.out_channels)


class ResBlock(nn.Module):
    """vqgan.py:325-356: x + conv2(silu(norm2(conv1(silu(norm1(x))))))."""

    def __init__(self, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, norm_type="group",
                 padding_type="replicate"):
        super().__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels, norm_type)
        self.conv1 = SamePadConv3d(in_channels, out_channels, kernel_size=3, padding_type=padding_type)
        self.dropout = torch.nn.Dropout(dropout)
        self.norm2 = Normalize(in_channels, norm_type)
        self.conv2 = SamePadConv3d(out_channels, out_channels, kernel_size=3, padding_type=padding_type)
        if self.in_channels != self.out_channels:
            self.conv_shortcut = SamePadConv3d(in_channels, out_channels, kernel_size=3, padding_type=padding_type)

    def forward_cl(self, x):
        h = self.conv1.forward_cl(x, pre=norm_args(self.norm1), act=1)
        skip = x if self.in_channels == self.out_channels else self.conv_shortcut.forward_cl(x)
        return self.conv2.forward_cl(h, pre=norm_args(self.norm2), act=1, resid=skip)

    def forward(self, x):
        return to_channels_first(self.forward_cl(to_channels_last(x)), self.out_channels)


class Encoder(nn.Module):
    """vqgan.py:263-300."""

    def __init__(self, n_hiddens, downsample, image_channel=3, norm_type="group", padding_type="replicate"):
        super().__init__()
        n_times_downsample = np.array([int(math.log2(d)) for d in downsample])
        self.conv_blocks = nn.ModuleList()
        max_ds = n_times_downsample.max()
        self.conv_first = SamePadConv3d(image_channel, n_hiddens, kernel_size=3, padding_type=padding_type)
        out_channels = n_hiddens
        for i in range(max_ds):
            block = nn.Module()
            in_channels = n_hiddens * 2 ** i
            out_channels = n_hiddens * 2 ** (i + 1)
            stride = tuple([2 if d > 0 else 1 for d in n_times_downsample])
            block.down = SamePadConv3d(in_channels, out_channels, 4, stride=stride, padding_type=padding_type)
            block.res = ResBlock(out_channels, out_channels, norm_type=norm_type)
            self.conv_blocks.append(block)
            n_times_downsample -= 1
        self.final_block = nn.Sequential(Normalize(out_channels, norm_type), SiLU())
        self.out_channels = out_channels

    def forward_cl(self, x):
        """channels-last in, channels-last out WITHOUT the final Normalize + SiLU: the caller's 1x1x1 convolution
        (`pre_vq_conv`) takes them as its fused `pre` (see VQGAN.encode)."""
        h = self.conv_first.forward_cl(x)
        for block in self.conv_blocks:
            h = block.down.forward_cl(h)
            h = block.res.forward_cl(h)
        return h

    def forward(self, x):
        h = self.forward_cl(to_channels_last(x))
        h = ops.pad_norm_act(h, (0,) * 6, act=1, **norm_args(self.final_block[0]))
        return to_channels_first(h, self.out_channels)


class Decoder(nn.Module):
    """vqgan.py:303-334."""

    def __init__(self, n_hiddens, upsample, image_channel, norm_type="group"):
        super().__init__()
        n_times_upsample = np.array([int(math.log2(d)) for d in upsample])
        max_us = n_times_upsample.max()
        in_channels = n_hiddens * 2 ** max_us
        self.final_block = nn.Sequential(Normalize(in_channels, norm_type), SiLU())
        self.conv_blocks = nn.ModuleList()
        out_channels = in_channels
        for i in range(max_us):
            block = nn.Module()
            in_channels = in_channels if i == 0 else n_hiddens * 2 ** (max_us - i + 1)
            out_channels = n_hiddens * 2 ** (max_us - i)
            us = tuple([2 if d > 0 else 1 for d in n_times_upsample])
            block.up = SamePadConvTranspose3d(in_channels, out_channels, 4, stride=us)
            block.res1 = ResBlock(out_channels, out_channels, norm_type=norm_type)
            block.res2 = ResBlock(out_channels, out_channels, norm_type=norm_type)
            self.conv_blocks.append(block)
            n_times_upsample -= 1
        self.conv_last = SamePadConv3d(out_channels, image_channel, kernel_size=3)
        self.image_channel = image_channel

    def forward_cl(self, x):
        h = x
        for i, block in enumerate(self.conv_blocks):
            # the first up-convolution takes final_block (Normalize + SiLU) as its fused prologue
            h = block.up.forward_cl(h, pre=norm_args(self.final_block[0]), act=1) if i == 0 else block.up.forward_cl(h)
            h = block.res1.forward_cl(h)
            h = block.res2.forward_cl(h)
        if len(self.conv_blocks) == 0:
            h = ops.pad_norm_act(h, (0,) * 6, act=1, **norm_args(self.final_block[0]))
        return self.conv_last.forward_cl(h)

    def forward(self, x):
        return to_channels_first(self.forward_cl(to_channels_last(x)), self.image_channel)


class VQGAN(nn.Module):
    """`mebt.vqgan.VQGAN` for inference.  Two ways to build it:
    * `VQGAN(args)` with the reference's hyper-parameter namespace (`embedding_dim, n_codes, n_hiddens, downsample,
      image_channels, norm_type, padding_type, no_random_restart, restart_thres`; vqgan.py:36-52): the full model;
    * `VQGAN(n_codes, embedding_dim, encoder=..., ...)`: the codebook with injected (or identity) callables around
      it, which is what the synthetic-latent benchmark (config #4) uses."""

    def __init__(self, args=16384, embedding_dim=256, encoder=None, pre_vq_conv=None, post_vq_conv=None, decoder=None,
                 n_codes=None):
        """`args`: the reference's hyper-parameter namespace (its only constructor argument, vqgan.py:36), or - second form -
        the number of codes (also accepted as `n_codes=`)."""
        super().__init__()
        self.args = None
        if n_codes is None:
            n_codes = args
        if hasattr(args, "n_hiddens"):
            self.args = args
            padding_type = getattr(args, "padding_type", "replicate")
            norm_type = getattr(args, "norm_type", "group")
            image_channels = getattr(args, "image_channels", 3)
            self.embedding_dim, self.n_codes = args.embedding_dim, args.n_codes
            self.encoder = Encoder(args.n_hiddens, args.downsample, image_channels, norm_type, padding_type)
            self.decoder = Decoder(args.n_hiddens, args.downsample, image_channels, norm_type)
            self.enc_out_ch = self.encoder.out_channels
            self.pre_vq_conv = SamePadConv3d(self.enc_out_ch, args.embedding_dim, 1, padding_type=padding_type)
            self.post_vq_conv = SamePadConv3d(args.embedding_dim, self.enc_out_ch, 1)
            self.codebook = Codebook(args.n_codes, args.embedding_dim,
                                     no_random_restart=getattr(args, "no_random_restart", False),
                                     restart_thres=getattr(args, "restart_thres", 1.0))
            self.codebook._need_init = False
            return
        self.embedding_dim, self.n_codes = embedding_dim, n_codes
        self.codebook = Codebook(n_codes, embedding_dim)
        self.codebook._need_init = False
        self.encoder = encoder if encoder is not None else nn.Identity()
        self.pre_vq_conv = pre_vq_conv if pre_vq_conv is not None else nn.Identity()
        self.post_vq_conv = post_vq_conv if post_vq_conv is not None else nn.Identity()
        self.decoder = decoder if decoder is not None else nn.Identity()

    # hyper-parameters of the reference's command line (vqgan.py:229-252): (flag, type, default[, choices]).  The GAN-training
    # ones are accepted (they are part of the namespace a checkpoint stores) and unused here.
    _HPARAMS = (("embedding_dim", int, 256), ("n_codes", int, 2048), ("n_hiddens", int, 240), ("lr", float, 3e-4),
                ("downsample", "ints", (4, 4, 4)), ("disc_channels", int, 64), ("disc_layers", int, 3),
                ("discriminator_iter_start", int, 50000), ("disc_loss_type", str, "hinge", ("hinge", "vanilla")),
                ("image_gan_weight", float, 1.0), ("video_gan_weight", float, 1.0), ("l1_weight", float, 4.0),
                ("gan_feat_weight", float, 0.0), ("perceptual_weight", float, 0.0), ("i3d_feat", "flag", False),
                ("restart_thres", float, 1.0), ("no_random_restart", "flag", False),
                ("norm_type", str, "group", ("batch", "group")),
                ("padding_type", str, "replicate", ("replicate", "constant", "reflect", "circular")))

    @staticmethod
    def add_model_specific_args(parent_parser):
        import argparse
        parser = argparse.ArgumentParser(parents=[parent_parser], add_help=False)
        for name, kind, default, *choices in VQGAN._HPARAMS:
            if kind == "flag":
                parser.add_argument(f"--{name}", action="store_true")
            elif kind == "ints":
                parser.add_argument(f"--{name}", nargs="+", type=int, default=default)
            else:
                parser.add_argument(f"--{name}", type=kind, default=default, **({"choices": list(choices[0])} if choices else {}))
        return parser

    @property
    def latent_shape(self):
        a = self.args
        input_shape = (a.sequence_length // a.sample_every_n_frames, a.resolution, a.resolution)
        return tuple(s // d for s, d in zip(input_shape, a.downsample))

    def _native(self):
        return isinstance(self.encoder, Encoder) and isinstance(self.decoder, Decoder)

    def pre_quant(self, x):
        """x [B, C, T, H, W] -> the pre-VQ latent [B, embedding_dim, t, h, w] fp32 (vqgan.py:83)."""
        if not self._native():
            return self.pre_vq_conv(self.encoder(x))
        h = self.encoder.forward_cl(to_channels_last(x))
        h = self.pre_vq_conv.forward_cl(h, pre=norm_args(self.encoder.final_block[0]), act=1)
        return to_channels_first(h, self.embedding_dim)

    def encode(self, x, include_embeddings=False):
        vq_output = self.codebook(self.pre_quant(x))
        if include_embeddings:
            return vq_output["embeddings"], vq_output["encodings"]
        return vq_output["encodings"]

    def decode(self, encodings):
        h = ops.row_gather(encodings, self.codebook.embeddings, channel_first=True)   # embedding + shift_dim fused
        if not self._native():
            return self.decoder(self.post_vq_conv(h))
        h = self.post_vq_conv.forward_cl(to_channels_last(h))
        return to_channels_first(self.decoder.forward_cl(h), self.decoder.image_channel)

    def forward(self, x, optimizer_idx=None, log_image=False):
        raise NotImplementedError("mebt_b200.VQGAN: GAN training (discriminators, LPIPS, EMA codebook) is out of scope")


# ---- file formats of the sampling scripts ------------------------------------------------------------------------------------
class _Args(dict):
    """argparse.Namespace-like view of a dict (hasattr / getattr defaults work)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = dict.__setitem__


def load_vqgan(ckpt, device="cuda"):
    """The reference's `load_vqgan` (mebt/download.py / utils: VQGAN.load_from_checkpoint): a Lightning checkpoint holding
    `hyper_parameters` (the argparse namespace, as `args` or flat) and `state_dict`.  Discriminator / LPIPS weights in the
    file are ignored (inference)."""
    sd = torch.load(ckpt, map_location="cpu", weights_only=False) if not isinstance(ckpt, dict) else ckpt
    hp = sd.get("hyper_parameters", {})
    args = hp.get("args", hp)
    args = _Args(vars(args) if hasattr(args, "__dict__") and not isinstance(args, dict) else dict(args))
    model = VQGAN(args)
    own = model.state_dict()
    state = {k: v for k, v in sd["state_dict"].items() if k in own}
    missing = [k for k in own if k not in state and not k.startswith("codebook.")]
    if missing:
        raise MebtError(f"load_vqgan: the checkpoint lacks {missing[:4]}{'...' if len(missing) > 4 else ''}")
    model.load_state_dict(state, strict=False)
    return model.to(device).eval()


def save_codemaps(save_np, code_maps):
    """`np.save(save_np + '_codemap', np.concatenate(all_code, 0))` of draft_and_revise_videos.py:181-185: the sampled code
    grids of a run (a list of [b, T, H, W] int64 batches) as one array.  Returns the file name written."""
    arr = np.concatenate([c.detach().cpu().numpy() if torch.is_tensor(c) else np.asarray(c) for c in code_maps], 0)
    np.save(save_np + "_codemap", arr)
    return save_np + "_codemap.npy"


def load_codemaps(path, device="cuda"):
    """The `--np_draft` input of the sampling scripts: a `_codemap.npy` written by a previous run -> int64 tensor."""
    return torch.from_numpy(np.load(path)).long().to(device)


def save_samples(save_np, samples, total_length, resolution, n_sample=None, rng=None):
    """draft_and_revise_videos.py:191-198: the decoded videos (a list of arrays / tensors with values in [0, 1], any leading
    shape over [3, total_length, resolution, resolution]) as ONE uint8 array [n, T, H, W, 3], randomly permuted and cut
    to n_sample like the reference (np.random, or `rng`)."""
    data = np.array([s.detach().cpu().numpy() if torch.is_tensor(s) else np.asarray(s) for s in samples])
    data = np.transpose(data.reshape(-1, 3, total_length, resolution, resolution), (0, 2, 3, 4, 1))
    n_total = data.shape[0]
    perm = (rng or np.random).permutation(n_total)[:n_sample if n_sample is not None else n_total]
    data = (data * 255).astype(np.uint8)[perm]
    np.save(save_np, data)
    return data
