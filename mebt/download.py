from mebt_b200.download import *  # noqa: F401,F403
from mebt_b200.download import load_transformer, load_vqgan  # noqa: F401
