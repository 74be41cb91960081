"""Drop-in for `core/interp_weights_est.py`: the `Simple` weights-estimation net."""
from rnc.modules import Simple  # noqa: F401
