from mebt_b200.modules.codebook import Codebook  # noqa: F401
