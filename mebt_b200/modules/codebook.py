"""Drop-in for `mebt.modules.codebook.Codebook` (reference: mebt/modules/codebook.py), eval path.

MeBT freezes the VQGAN (mebt/transformer.py:184-188), so the hot path is quantise (fused fp32 distance +
argmin, K9) and lookup (row gather, K10).  The EMA update / random restart of training mode belongs to VQGAN
training, which is out of scope; calling forward in training mode raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class Codebook(nn.Module):
    def __init__(self, n_codes, embedding_dim, no_random_restart=False, restart_thres=1.0):
        super().__init__()
        self.register_buffer("embeddings", torch.randn(n_codes, embedding_dim))
        self.register_buffer("N", torch.zeros(n_codes))
        self.register_buffer("z_avg", self.embeddings.data.clone())
        self.n_codes = n_codes
        self.embedding_dim = embedding_dim
        self._need_init = True
        self.no_random_restart = no_random_restart
        self.restart_thres = restart_thres
        self._sq = None

    def _sqnorm(self):
        """(|E|^2 per code, the codebook's fp16 (hi | hi | lo) form for the tensor-core search), per codebook version."""
        key = (self.embeddings.data_ptr(), self.embeddings._version)
        if self._sq is None or self._sq[0] != key:
            tc = self.embedding_dim % 64 == 0 and self.n_codes % 64 == 0 and self.embedding_dim <= 1024
            self._sq = (key, ops.row_sqnorm(self.embeddings), ops.vq_split_codebook(self.embeddings) if tc else None)
        return self._sq[1], self._sq[2]

    def forward(self, z):
        """z [b, c, t, h, w] fp32 -> dict(embeddings, encodings, commitment_loss, perplexity) (codebook.py:48-97)."""
        if self.training:
            raise NotImplementedError("mebt_b200.Codebook: EMA codebook training is out of scope; call .eval()")
        z = z.float().contiguous()
        sq, split = self._sqnorm()
        enc = ops.vq_argmin(z, self.embeddings, sq, split)                      # [b, t, h, w] int64
        emb = ops.row_gather(enc, self.embeddings, channel_first=True)           # [b, c, t, h, w]
        commitment_loss = 0.25 * torch.mean((z - emb) ** 2)
        emb_st = (emb - z).detach() + z                                           # straight-through value
        counts = torch.bincount(enc.reshape(-1), minlength=self.n_codes).float()
        avg = counts / enc.numel()
        perplexity = torch.exp(-torch.sum(avg * torch.log(avg + 1e-10)))          # histogram instead of one-hot
        return dict(embeddings=emb_st, encodings=enc, commitment_loss=commitment_loss, perplexity=perplexity)

    def dictionary_lookup(self, encodings):
        return ops.row_gather(encodings, self.embeddings, channel_first=False)
