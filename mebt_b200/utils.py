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


# ---- small generic helpers of mebt/utils.py that scripts import (not on the hot path) ------------------------------------
def view_range(x, i, j, shape):
    """Reshape the dims [i, j) of x to `shape`: (b, thw, c) with (1, 2, (t, h, w)) -> (b, t, h, w, c) (mebt/utils.py:55-76)."""
    n = x.dim()
    i = i + n if i < 0 else i
    j = n if j is None else (j + n if j < 0 else j)
    if not 0 <= i < j <= n:
        raise AssertionError(f"view_range: bad dims ({i}, {j}) for {n} dimensions")
    return x.view(*x.shape[:i], *tuple(shape), *x.shape[j:])


def tensor_slice(x, begin, size):
    """x[b0:b0+s0, b1:b1+s1, ...]; a size of -1 runs to the end of that dimension (mebt/utils.py:110-117)."""
    if any(b < 0 for b in begin):
        raise AssertionError("tensor_slice: negative begin")
    ends = [x.shape[d] if s == -1 else b + s for d, (b, s) in enumerate(zip(begin, size))]
    if any(e < b for b, e in zip(begin, ends)):
        raise AssertionError("tensor_slice: negative size")
    return x[tuple(slice(b, e) for b, e in zip(begin, ends))]


def correct(output, target, topk=(1,)):
    """Top-k hit COUNTS (the un-normalised form of `accuracy`, mebt/utils.py:96-108)."""
    with torch.no_grad():
        pred = output.topk(max(topk), 1, True, True)[1]
        hit = pred.eq(target.reshape(-1, 1))
        return [hit[:, :k].reshape(-1).float().sum(0, keepdim=True) for k in topk]


def adopt_weight(global_step, threshold=0, value=0.0):
    """`value` before `threshold` steps, 1 afterwards (the GAN-loss warm-up switch, mebt/utils.py:120-124)."""
    return value if global_step < threshold else 1


def comp_getattr(args, attr_name, default=None):
    return getattr(args, attr_name, default)
