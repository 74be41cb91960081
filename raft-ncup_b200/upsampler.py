"""Drop-in for `core/upsampler.py`: get_upsampler(in_ch, guidance_ch, args), NConvUpsampler."""
from rnc.modules import NConvUpsampler, get_upsampler  # noqa: F401
