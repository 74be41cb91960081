"""Drop-in for the reference's `core/raft_nc_dbl.py` module: `from raft_nc_dbl import RAFT` (evaluate.py:19, train.py:17)."""
from rnc.model import RAFTNcup as RAFT  # noqa: F401
