# Standard-normal draws for Normal.rsample are supplied by the golden generator (oracle/gen_golden.py) through
# NOISE_PROVIDER(key, shape): the JAX PRNG streams of the reference cannot be reproduced without jax, so the
# draws are explicit inputs of the sampled-path fixtures.
NOISE_PROVIDER = None


class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    @property
    def mean(self):
        return self.loc

    def rsample(self, key, sample_shape=()):
        if NOISE_PROVIDER is None:
            raise RuntimeError("the shimmed path is deterministic (use_mean=True) unless a NOISE_PROVIDER is installed")
        import numpy as np

        shape = tuple(sample_shape) + tuple(np.shape(self.loc))
        noise = np.asarray(NOISE_PROVIDER(key, shape))
        assert noise.shape == shape, (noise.shape, shape)
        return self.loc + self.scale * noise
