"""Import alias: `mebt.*` resolves to the mebt_b200 drop-in so that reference-facing scripts and config
`target:` strings (`mebt.transformer.Net2NetTransformer`, `mebt.mask_sampler.MaskGen`) work unchanged."""
from mebt_b200.download import load_transformer, load_vqgan  # noqa: F401
from mebt_b200.mask_sampler import MaskGen  # noqa: F401
from mebt_b200.transformer import Net2NetTransformer  # noqa: F401
from mebt_b200.vqgan import VQGAN  # noqa: F401


def __getattr__(name):
    """`from mebt import VideoData`: the data pipeline is the reference's own (out of scope here, mebt/_reference.py)."""
    if name == "VideoData":
        from mebt import data
        return data.VideoData
    raise AttributeError(f"module 'mebt' has no attribute {name!r}")
