from mebt_b200.modules.gpt import *  # noqa: F401,F403
from mebt_b200.modules.gpt import GPT, Block, CrossAttention, GPTConfig, GPT1Config  # noqa: F401
