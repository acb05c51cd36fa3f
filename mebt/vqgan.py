from mebt_b200.vqgan import *  # noqa: F401,F403
