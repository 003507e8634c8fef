"""`jax.numpy` -> numpy."""
import numpy as _np
from numpy import *  # noqa: F401,F403

float_ = _np.float64
ndarray = _np.ndarray


def array(x, dtype=None):
    return _np.asarray(x, dtype=dtype)
