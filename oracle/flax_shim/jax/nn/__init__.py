import numpy as _np

from . import initializers  # noqa: F401


def relu(x):
    return _np.maximum(x, 0)


def gelu(x, approximate=True):
    if not approximate:
        raise NotImplementedError
    return 0.5 * x * (1.0 + _np.tanh(_np.sqrt(2.0 / _np.pi) * (x + 0.044715 * x ** 3)))


def softplus(x):
    return _np.logaddexp(x, 0.0)


def softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    e = _np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)
