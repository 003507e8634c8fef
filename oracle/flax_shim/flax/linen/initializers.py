from jax.nn.initializers import lecun_normal, normal, ones, variance_scaling, zeros  # noqa: F401
