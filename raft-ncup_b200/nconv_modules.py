"""Drop-in for `core/nconv_modules.py`: NConvUNet / NConv2d parameter containers."""
from rnc.modules import NConv2d, NConvUNet  # noqa: F401
