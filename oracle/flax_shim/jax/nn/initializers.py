def _init(*a, **k):
    def init(key, shape, dtype=None):
        import numpy as np

        return np.zeros(shape)

    return init


normal = _init
zeros = _init()
ones = _init()


def variance_scaling(*a, **k):
    return _init()


def lecun_normal(*a, **k):
    return _init()
