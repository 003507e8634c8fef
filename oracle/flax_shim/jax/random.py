import numpy as _np

KeyArray = object


def normal(key, shape, dtype=None):
    return _np.zeros(shape)  # initialisers are never used: parameters are always supplied


def PRNGKey(seed):
    return _np.array([0, seed], dtype=_np.uint32)
