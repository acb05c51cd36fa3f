"""Components the drop-in does not re-implement (data loaders, video / image file writers) stay the
reference's own: with `MEBT_REF=<path of a MeBT checkout>` they are loaded from that checkout by file path, so that its scripts
(`from mebt import VideoData`, `from mebt.data import preprocess`, `from mebt.utils import save_video_grid`) run unchanged."""
import importlib.util
import os
import sys


def load(module: str):
    """The reference's `mebt/<module>.py` as a module object (cached), or ImportError naming what to set."""
    ref = os.environ.get("MEBT_REF")
    path = os.path.join(ref, "mebt", module + ".py") if ref else None
    if not path or not os.path.exists(path):
        raise ImportError(f"mebt.{module}: not part of mebt_b200 (out of scope: data pipeline / file writers); set "
                          f"MEBT_REF=<MeBT checkout> to use the reference's own mebt/{module}.py")
    name = f"mebt._reference_{module}"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[name]
        raise
    return mod
