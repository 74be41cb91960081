"""Drop-in for `core/extractor.py`: BasicEncoder(output_dim, norm_fn, dropout)."""
from rnc.modules import BasicEncoder, ResidualBlock  # noqa: F401
