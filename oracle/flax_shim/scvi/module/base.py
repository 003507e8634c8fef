import flax.linen as nn


class JaxBaseModuleClass(nn.Module):
    """Only what MrVAE needs to run `inference` on supplied parameters."""


class LossOutput:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def flax_configure(cls):
    return cls
