"""Drop-in for the reference's `core/raft.py` module: `from raft import RAFT` (evaluate.py:13, train.py:13)."""
from rnc.model import RAFTConvex as RAFT  # noqa: F401
