from .codebook import Codebook  # noqa: F401
