from mebt_b200.mask_sampler import *  # noqa: F401,F403
