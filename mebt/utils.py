from mebt_b200.utils import *  # noqa: F401,F403


def __getattr__(name):
    """File / video writers and debugging helpers (`save_video_grid`, `save_image_grid`, `visualize_tensors`, `ForkedPdb`)
    are not re-implemented: they resolve to the reference's own mebt/utils.py when `MEBT_REF` points at a checkout."""
    from mebt._reference import load
    try:
        return getattr(load("utils"), name)
    except ImportError as exc:
        raise AttributeError(f"mebt.utils.{name}: {exc}") from None
