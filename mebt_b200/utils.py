"""Small host helpers on the path (reference: mebt/utils.py:28-94, utils.py:3-14)."""
from __future__ import annotations

import importlib

import torch


def shift_dim(x, src_dim=-1, dest_dim=-1, make_contiguous=True):
    """Move dim `src_dim` to position `dest_dim`, e.g. (b,c,t,h,w) -> (b,t,h,w,c) for (1,-1)."""
    n = x.dim()
    src = src_dim % n
    dest = dest_dim % n
    order = [d for d in range(n) if d != src]
    order.insert(dest, src)
    x = x.permute(order)
    return x.contiguous() if make_contiguous else x


def accuracy(output, target, topk=(1,)):
    """Top-k accuracies in percent (mebt/utils.py:80-94)."""
    with torch.no_grad():
        k_max = max(topk)
        pred = output.topk(k_max, 1, True, True)[1]
        hit = pred.eq(target.reshape(-1, 1))
        return [hit[:, :k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]


def get_obj_from_str(string, reload=False):
    module, cls = string.rsplit(".", 1)
    mod = importlib.import_module(module)
    if reload:
        importlib.reload(mod)
    return getattr(mod, cls)


def instantiate_from_config(config):
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    config["target"] = config["target"].replace("tats.", "mebt.")   # legacy prefix, as the reference
    return get_obj_from_str(config["target"])(**config.get("params", dict()))
