"""The VQGAN boundary of the hot path (reference: mebt/vqgan.py:82-93).

Only the two lines that touch the codebook are on the path: `encode` hands the pre-VQ latent to
`Codebook.forward` (K9 + K10) and `decode` starts with `F.embedding(encodings, codebook.embeddings)` followed
by `shift_dim(h, -1, 1)` (K10 with a channel-first store).  The 3-D conv encoder/decoder, discriminators and
losses are cuDNN-class work that needs a checkpoint and is out of scope: they are taken
as injected callables.
"""
from __future__ import annotations

import torch.nn as nn

from . import ops
from .modules.codebook import Codebook


class VQGAN(nn.Module):
    """Codebook-facing part of `mebt.vqgan.VQGAN`.  `encoder`, `pre_vq_conv`, `post_vq_conv`, `decoder` default to
    identity so that latents / embeddings pass straight through (synthetic-latent benchmarks, config #4)."""

    def __init__(self, n_codes=16384, embedding_dim=256, encoder=None, pre_vq_conv=None, post_vq_conv=None,
                 decoder=None):
        super().__init__()
        self.codebook = Codebook(n_codes, embedding_dim)
        self.codebook._need_init = False
        self.encoder = encoder if encoder is not None else nn.Identity()
        self.pre_vq_conv = pre_vq_conv if pre_vq_conv is not None else nn.Identity()
        self.post_vq_conv = post_vq_conv if post_vq_conv is not None else nn.Identity()
        self.decoder = decoder if decoder is not None else nn.Identity()

    def encode(self, x, include_embeddings=False):
        h = self.pre_vq_conv(self.encoder(x))
        vq_output = self.codebook(h)
        if include_embeddings:
            return vq_output["embeddings"], vq_output["encodings"]
        return vq_output["encodings"]

    def decode(self, encodings):
        h = ops.row_gather(encodings, self.codebook.embeddings, channel_first=True)   # embedding + shift_dim fused
        return self.decoder(self.post_vq_conv(h))
