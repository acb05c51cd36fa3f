"""`mebt.data` (datasets, `VideoData`, `preprocess`): the reference's own module, loaded from `$MEBT_REF` (mebt/_reference.py)."""
from mebt._reference import load as _load

_mod = _load("data")
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
