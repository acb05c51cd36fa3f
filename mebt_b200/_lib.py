"""ctypes binding of libmebt_b200.so (the C ABI declared in include/mebt_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a Python exception
is raised.  Loading the library does not need a GPU (it links cudart statically); calling a compute
entry point without an sm_100 device fails with the library's own error message.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_uint64, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libmebt_b200.so"


class MebtError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not _LIB_PATH.exists():
        raise MebtError(
            f"{_LIB_PATH} is missing: build it with `python -m mebt_b200.build` "
            "(mebt_b200 has no CPU / PyTorch fallback path)")
    return ctypes.CDLL(str(_LIB_PATH))


_lib = _load()
_lib.mebt_last_error.restype = c_char_p
_lib.mebt_version.restype = c_char_p

# name -> argtypes; every function returns int
_SIGNATURES: dict[str, list] = {
    "mebt_device_check": [],
    "mebt_gemm_bf16": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                       c_void_p, c_void_p, c_int, c_int, c_void_p],
    "mebt_gemm_bf16_aux": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                           c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "mebt_colsum": [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p],
    "mebt_layernorm_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                           c_int, c_int, c_void_p, c_size_t, c_void_p],
    "mebt_embed_backward": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                            c_void_p],
    "mebt_latent_attention_bwd": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                  c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                  c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p, c_size_t, c_void_p],
    "mebt_embed_gather": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_void_p],
    "mebt_layernorm": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                       c_void_p, c_void_p, c_void_p],
    "mebt_scatter_ids": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    "mebt_row_gather": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "mebt_cast_f32_to_bf16": [c_void_p, c_void_p, c_longlong, c_void_p],
    "mebt_check_index_errors": [c_void_p],
    "mebt_masked_ce": [c_void_p, c_longlong, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                       c_longlong, c_float, c_void_p],
    "mebt_ce_reduce": [c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "mebt_sample_logits": [c_void_p, c_longlong, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p, c_uint64,
                           c_uint64, c_void_p, c_void_p, c_void_p, c_void_p],
    "mebt_remask_sort": [c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                         c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p],
    "mebt_row_sqnorm": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "mebt_vq_argmin": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                       c_void_p],
    "mebt_vq_split_codebook": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "mebt_vq_argmin_tc": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                          c_void_p],
    "mebt_latent_attention_fwd": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                  c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "mebt_latent_attention_fwd_ws": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                     c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                     c_void_p],
}


def _bind():
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(_lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int




class LayerStruct(ctypes.Structure):
    """mebt_layer_t"""
    _fields_ = [("mode", c_int)] + [(n, c_void_p) for n in (
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


class LayerGradsStruct(ctypes.Structure):
    """mebt_layer_grads_t"""
    _fields_ = [(n, c_void_p) for n in (
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


_SIGNATURES["mebt_stack_forward_train"] = [ctypes.POINTER(LayerStruct), c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                           c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_void_p, c_size_t, c_void_p]
_SIGNATURES["mebt_stack_backward"] = [ctypes.POINTER(LayerStruct), ctypes.POINTER(LayerGradsStruct), c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                                      c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]
class DropoutStruct(ctypes.Structure):
    """mebt_dropout_t"""
    _fields_ = [("attn_p", c_float), ("resid_p", c_float), ("seed", c_uint64)]


_SIGNATURES["mebt_stack_forward_train_dropout"] = _SIGNATURES["mebt_stack_forward_train"][:-1] + [
    ctypes.POINTER(DropoutStruct), c_void_p]
_SIGNATURES["mebt_stack_backward_dropout"] = _SIGNATURES["mebt_stack_backward"][:-3] + [
    ctypes.POINTER(DropoutStruct), c_void_p, c_size_t, c_void_p]
_I3, _I4, _I6 = ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int)
_SIGNATURES["mebt_pad_norm_act"] = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, _I6, c_int, c_int,
                                    c_int, c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
_SIGNATURES["mebt_conv3d_ndhwc"] = [c_void_p, c_int, _I4, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, _I4,
                                    c_int, _I3, _I3, _I3, _I3, _I3, _I3, c_void_p]


class FusedAdamwStruct(ctypes.Structure):
    """mebt_fused_adamw_t"""
    _fields_ = [("grad_base", c_void_p), ("p", c_void_p), ("m", c_void_p), ("v", c_void_p), ("p_bf16", c_void_p),
                ("decay_blocks", c_void_p), ("block_shift", c_int), ("lr", c_float), ("beta1", c_float), ("beta2", c_float),
                ("eps", c_float), ("weight_decay", c_float), ("step", c_int)]


_SIGNATURES["mebt_stack_backward_fused"] = _SIGNATURES["mebt_stack_backward"][:-3] + [
    ctypes.POINTER(DropoutStruct), ctypes.POINTER(FusedAdamwStruct), c_void_p, c_size_t, c_void_p]
_SIGNATURES["mebt_split_f32_bf16x3"] = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
_SIGNATURES["mebt_latent_attention_fwd_f32"] = [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                                c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_void_p]
_SIGNATURES["mebt_adamw_flat"] = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_float,
                                  c_float, c_float, c_float, c_float, c_int, c_void_p]
_SIGNATURES["mebt_adamw_flat_bg"] = _SIGNATURES["mebt_adamw_flat"][:-1] + [c_int, c_void_p]
_SIGNATURES["mebt_dropout_rows"] = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_uint64,
                                    c_uint64, c_void_p]
_SIGNATURES["mebt_latent_attention_fwd_dropout"] = _SIGNATURES["mebt_latent_attention_fwd"][:-1] + [c_float, c_uint64,
                                                                                                    c_void_p]
_SIGNATURES["mebt_latent_attention_bwd_dropout"] = _SIGNATURES["mebt_latent_attention_bwd"][:-3] + [
    c_float, c_uint64, c_void_p, c_size_t, c_void_p]
_SIGNATURES["mebt_attention_dropout_mask"] = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_uint64, c_void_p]


class WgradDesc(ctypes.Structure):
    """mebt_wgrad_desc_t"""
    _fields_ = [("dY", c_void_p), ("ld_dy", c_int), ("X", c_void_p), ("ldx", c_int), ("dW", c_void_p), ("ldw", c_int),
                ("n_out", c_int), ("k_in", c_int), ("rows", c_int), ("accumulate", c_int)]


_SIGNATURES["mebt_gemm_grouped_wgrad"] = [ctypes.POINTER(WgradDesc), c_int, c_void_p]


class EncHoistStruct(ctypes.Structure):
    """mebt_enc_hoist_t"""
    _fields_ = [("n_enc", c_int), ("w_enc_kv", c_void_p), ("b_enc_kv", c_void_p), ("ones", c_void_p), ("zeros", c_void_p)]


_SIGNATURES["mebt_stack_forward_hoisted"] = [ctypes.POINTER(LayerStruct), c_int, c_void_p, c_void_p, c_void_p,
                                             ctypes.POINTER(EncHoistStruct), c_int, c_int, c_int, c_int, c_int, c_int,
                                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                             c_void_p]
_SIGNATURES["mebt_stack_forward_sample"] = [ctypes.POINTER(LayerStruct), c_int, c_void_p, c_void_p, c_void_p,
                                            ctypes.POINTER(EncHoistStruct), c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_float,
                                            c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]
_SIGNATURES["mebt_head_sample"] = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_uint64, c_uint64, c_void_p,
                                   c_void_p, c_size_t, c_void_p]
_SIGNATURES["mebt_stack_forward"] = [ctypes.POINTER(LayerStruct), c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p, c_size_t, c_void_p]
_bind()
for _n, _a in (("mebt_colsum_workspace_bytes", [c_int]), ("mebt_layernorm_bwd_workspace_bytes", [c_int]),
               ("mebt_latent_attention_bwd_workspace_bytes", [c_int, c_int, c_int])):
    getattr(_lib, _n).argtypes = _a
    getattr(_lib, _n).restype = c_size_t
_lib.mebt_groupnorm_workspace_bytes.argtypes = [c_int, c_int]
_lib.mebt_groupnorm_workspace_bytes.restype = c_size_t
_lib.mebt_stack_train_saved_bytes.argtypes = [ctypes.POINTER(LayerStruct), c_int, c_int, c_int, c_int, c_int, c_int, c_int]
_lib.mebt_stack_train_saved_bytes.restype = c_size_t
_lib.mebt_stack_backward_workspace_bytes.argtypes = [c_int] * 6
_lib.mebt_stack_backward_workspace_bytes.restype = c_size_t
_lib.mebt_launch_count.argtypes = []
_lib.mebt_launch_count.restype = ctypes.c_ulonglong
_lib.mebt_profile_enable.argtypes = [c_int]
_lib.mebt_profile_enable.restype = None
_SIGNATURES["mebt_profile_report"] = [c_void_p, c_void_p, c_void_p]
_lib.mebt_profile_report.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_longlong)]
_lib.mebt_profile_report.restype = c_int
_lib.mebt_stack_forward_hoisted_workspace_bytes.argtypes = [c_int] * 6
_lib.mebt_stack_forward_hoisted_workspace_bytes.restype = c_size_t
_lib.mebt_stack_forward_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
_lib.mebt_stack_forward_workspace_bytes.restype = c_size_t
_lib.mebt_latent_attention_fwd_workspace_bytes.argtypes = [c_int, c_int, c_int]
_lib.mebt_latent_attention_fwd_workspace_bytes.restype = c_size_t
_lib.mebt_head_sample_workspace_bytes.argtypes = [c_longlong]
_lib.mebt_head_sample_workspace_bytes.restype = c_size_t
_lib.mebt_vq_argmin_workspace_bytes.argtypes = [c_longlong]
_lib.mebt_vq_argmin_workspace_bytes.restype = c_size_t
_lib.mebt_vq_argmin_tc_workspace_bytes.argtypes = [c_longlong, c_int]
_lib.mebt_vq_argmin_tc_workspace_bytes.restype = c_size_t
_lib.mebt_vq_codebook_split_bytes.argtypes = [c_int, c_int]
_lib.mebt_vq_codebook_split_bytes.restype = c_size_t


def version() -> str:
    return _lib.mebt_version().decode()


def last_error() -> str:
    return _lib.mebt_last_error().decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise MebtError(f"{what or 'mebt_b200'} failed (code {rc}): {last_error()}")


def call(name: str, *args) -> None:
    check(getattr(_lib, name)(*args), name)


# Parameters are also written through raw device pointers (FlatAdamW's kernel updates the fp32 masters in place), which
# torch's per-tensor version counters do not see.  Every such writer bumps this epoch; the caches of derived operands
# (bf16 casts, folded weights) key on it next to (data_ptr, _version).
_WRITE_EPOCH = [0]


def bump_write_epoch() -> None:
    _WRITE_EPOCH[0] += 1


def write_epoch() -> int:
    return _WRITE_EPOCH[0]


FAMILIES = ("gemm", "attention", "layernorm", "embed", "sample", "ce", "remask", "scatter", "vq", "other")


def launch_count() -> int:
    return int(_lib.mebt_launch_count())


def profile_enable(on: bool) -> None:
    _lib.mebt_profile_enable(1 if on else 0)


def profile_report() -> dict:
    """-> {family: dict(ms=..., work=..., launches=...)} since the last report (synchronises the device)."""
    n = len(FAMILIES)
    t = (ctypes.c_double * n)()
    w = (ctypes.c_double * n)()
    c = (ctypes.c_longlong * n)()
    check(_lib.mebt_profile_report(t, w, c), "mebt_profile_report")
    return {f: dict(ms=t[i], work=w[i], launches=int(c[i])) for i, f in enumerate(FAMILIES)}


lib = _lib
