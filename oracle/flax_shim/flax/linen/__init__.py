"""numpy stand-in for the subset of flax.linen used by the reference (see ../README.md).

Module mechanics reproduced: dataclass-style fields, `@compact`, construction-order auto-naming
(`ClassName_i` per parent), `setup()` with attribute-named submodules, `apply({"params": tree}, ...,
method=...)`, `self.param(...)`.  Primitives (restated from the flax documentation / source):
DenseGeneral, Dense, LayerNorm (eps 1e-6, fast variance), Embed, MultiHeadDotProductAttention
(DenseGeneral q/k/v projections to (heads, head_dim), query / sqrt(head_dim), softmax, out projection).
"""
from __future__ import annotations

import dataclasses
import functools
from typing import Any, Optional

import numpy as np
from jax.nn import gelu, relu, softmax, softplus  # noqa: F401

from . import initializers  # noqa: F401

_STACK: list = []  # modules whose compact method / setup is running


def compact(fn):
    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        self._counters = {}
        _STACK.append(self)
        try:
            return fn(self, *args, **kwargs)
        finally:
            _STACK.pop()

    return wrapped


def merge_param(name, a, b):
    if a is None and b is None:
        raise ValueError(f"Parameter {name!r} must be passed to the constructor or at call time.")
    if a is not None and b is not None:
        raise ValueError(f"Parameter {name!r} was passed to the constructor and at call time.")
    return a if a is not None else b


@dataclasses.dataclass
class Module:
    name: Optional[str] = dataclasses.field(default=None, kw_only=True)
    parent: Any = dataclasses.field(default=None, kw_only=True, repr=False)

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        # drop annotations that are not dataclass fields in flax either (none needed here) and transform
        dataclasses.dataclass(cls, eq=False, repr=False)

    def __post_init__(self):
        self._counters = {}
        self._scope = None
        self._setup_done = False
        if _STACK and self.parent is None:
            parent = _STACK[-1]
            object.__setattr__(self, "parent", parent)
            if self.name is None:
                base = type(self).__name__
                i = parent._counters.get(base, 0)
                parent._counters[base] = i + 1
                object.__setattr__(self, "name", f"{base}_{i}")

    # submodules assigned in setup() are named after the attribute
    def __setattr__(self, key, value):
        if isinstance(value, Module) and _STACK and _STACK[-1] is self and not key.startswith("_") and key != "parent":
            if value.name is None or value.parent is self:
                object.__setattr__(value, "name", key)
                object.__setattr__(value, "parent", self)
        object.__setattr__(self, key, value)

    # ---- parameter scope
    def _params(self):
        if self._scope is not None:
            return self._scope
        return self.parent._params()[self.name]

    def param(self, name, init_fn, *shape_args):
        p = self._params()
        if name in p:
            return np.asarray(p[name])
        shape = shape_args[0] if shape_args else ()
        return np.zeros(shape)

    def make_rng(self, name):
        return name  # the stream name; numpyro.distributions.NOISE_PROVIDER turns it into explicit draws

    def _ensure_setup(self):
        if not self._setup_done and hasattr(self, "setup"):
            self._setup_done = True
            self._counters = {}
            _STACK.append(self)
            try:
                self.setup()
            finally:
                _STACK.pop()

    def apply(self, variables, *args, method=None, rngs=None, **kwargs):
        self._scope = variables["params"]
        self._ensure_setup()
        fn = method if method is not None else self.__call__
        if hasattr(fn, "__self__"):
            return fn(*args, **kwargs)
        return fn(self, *args, **kwargs)


def _prod(t):
    out = 1
    for v in t:
        out *= int(v)
    return out


class DenseGeneral(Module):
    features: Any = None
    axis: Any = -1
    use_bias: bool = True
    kernel_init: Any = None
    bias_init: Any = None

    def __call__(self, x):
        p = self._params()
        k = np.asarray(p["kernel"])
        axis = (self.axis,) if isinstance(self.axis, int) else tuple(self.axis)
        n_in = len(axis)
        assert all(a < 0 for a in axis) and sorted(axis) == list(range(-n_in, 0)), "only trailing contraction axes"
        y = np.tensordot(x, k, axes=(list(range(x.ndim - n_in, x.ndim)), list(range(n_in))))
        if self.use_bias:
            y = y + np.asarray(p["bias"])
        return y


class Dense(Module):
    features: int = None
    use_bias: bool = True
    kernel_init: Any = None

    def __call__(self, x):
        p = self._params()
        y = x @ np.asarray(p["kernel"])
        return y + np.asarray(p["bias"]) if self.use_bias else y


class LayerNorm(Module):
    epsilon: float = 1e-6
    use_bias: bool = True
    use_scale: bool = True
    use_fast_variance: bool = True

    def __call__(self, x):
        mu = x.mean(-1, keepdims=True)
        if self.use_fast_variance:
            var = np.maximum((x * x).mean(-1, keepdims=True) - mu * mu, 0.0)
        else:
            var = ((x - mu) ** 2).mean(-1, keepdims=True)
        mul = 1.0 / np.sqrt(var + self.epsilon)
        p = self._params() if (self.use_scale or self.use_bias) else {}
        if self.use_scale:
            mul = mul * np.asarray(p["scale"])
        y = (x - mu) * mul
        if self.use_bias:
            y = y + np.asarray(p["bias"])
        return y


class BatchNorm(Module):
    use_bias: bool = True
    use_scale: bool = True

    def __call__(self, x, use_running_average=True):
        raise NotImplementedError("normalization_type='batch' is not on the default path")


class Embed(Module):
    num_embeddings: int = None
    features: int = None
    embedding_init: Any = None

    def __call__(self, idx):
        return np.asarray(self._params()["embedding"])[np.asarray(idx).astype(int)]


class Sequential(Module):
    layers: Any = None

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class MultiHeadDotProductAttention(Module):
    num_heads: int = None
    qkv_features: Optional[int] = None
    out_features: Optional[int] = None
    dropout_rate: float = 0.0
    use_bias: bool = True

    @compact
    def __call__(self, inputs_q, inputs_kv=None, deterministic=None, **kw):
        if inputs_kv is None:
            inputs_kv = inputs_q
        if not deterministic and self.dropout_rate > 0:
            raise RuntimeError("attention dropout is active: the module is not in eval mode")
        qkv = self.qkv_features or inputs_q.shape[-1]
        hd = qkv // self.num_heads
        q = DenseGeneral((self.num_heads, hd), use_bias=self.use_bias, name="query")(inputs_q)
        k = DenseGeneral((self.num_heads, hd), use_bias=self.use_bias, name="key")(inputs_kv)
        v = DenseGeneral((self.num_heads, hd), use_bias=self.use_bias, name="value")(inputs_kv)
        q = q / np.sqrt(hd)
        w = softmax(np.einsum("...qhd,...khd->...hqk", q, k), axis=-1)
        a = np.einsum("...hqk,...khd->...qhd", w, v)
        out_f = self.out_features or inputs_q.shape[-1]
        return DenseGeneral(out_f, axis=(-2, -1), use_bias=self.use_bias, name="out")(a)
