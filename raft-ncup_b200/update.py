"""Drop-in for `core/update.py`: BasicUpdateBlock(args).forward(net, inp, corr, flow)."""
from rnc.modules import BasicMotionEncoder, BasicUpdateBlock, FlowHead, SepConvGRU  # noqa: F401
