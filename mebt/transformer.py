from mebt_b200.transformer import *  # noqa: F401,F403
from mebt_b200.transformer import (Net2NetTransformer, sample_from_logits, gumbel_sort, top_k_logits, top_p_probs,  # noqa: F401
                                   uniform, gaussian, gaussian2, gaussian100000_2, longest, linear, constant, cosine, disabled_train)
