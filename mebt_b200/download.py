"""Checkpoint loaders of `mebt.download` (reference: mebt/download.py:50-61).  The Google-Drive download helpers of that file
need a network and are not provided.

A Lightning checkpoint is a pickled dict with `state_dict` and `hyper_parameters`; `LightningModule.load_from_checkpoint`
re-creates the module from the latter (the constructor arguments recorded by `save_hyperparameters()`,
mebt/transformer.py:72) and loads the former.  That is what `load_transformer` does here, without Lightning.
"""
from __future__ import annotations

import inspect

import torch

from ._lib import MebtError
from .transformer import Net2NetTransformer
from .vqgan import load_vqgan as _load_vqgan


def load_vqgan(vqgan_ckpt, device=torch.device("cpu")):
    """`VQGAN.load_from_checkpoint(vqgan_ckpt).to(device).eval()` (download.py:50-54)."""
    return _load_vqgan(vqgan_ckpt, device=device)


def load_transformer(gpt_ckpt, vqgan_ckpt=None, device=torch.device("cpu")):
    """`Net2NetTransformer.load_from_checkpoint(gpt_ckpt).eval()` (download.py:56-61).  As in the reference `vqgan_ckpt` is
    accepted and NOT used: the first stage comes from `first_stage_config.params.ckpt_path` stored in the checkpoint's
    hyper-parameters (only read when `vtokens` is false)."""
    ckpt = torch.load(gpt_ckpt, map_location="cpu", weights_only=False) if not isinstance(gpt_ckpt, dict) else gpt_ckpt
    hp = ckpt.get("hyper_parameters")
    if hp is None or "state_dict" not in ckpt:
        raise MebtError("load_transformer: not a Lightning checkpoint (needs `hyper_parameters` and `state_dict`)")
    names = [p for p in inspect.signature(Net2NetTransformer.__init__).parameters if p != "self"]
    missing = [n for n in ("transformer_config", "first_stage_config", "mask_config") if n not in hp]
    if missing:
        raise MebtError(f"load_transformer: the checkpoint's hyper_parameters lack {missing}")
    kwargs = {k: hp[k] for k in names if k in hp}
    kwargs["ckpt_path"] = None                                  # the weights come from THIS checkpoint's state_dict
    model = Net2NetTransformer(**kwargs)
    model.load_state_dict(ckpt["state_dict"], strict=True)
    return model.to(device).eval()
