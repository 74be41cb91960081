"""Drop-in for `core/corr.py`: CorrBlock(fmap1, fmap2, num_levels=4, radius=4)(coords)."""
from rnc.modules import CorrBlock  # noqa: F401
