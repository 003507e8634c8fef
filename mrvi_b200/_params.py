"""Flat parameter schema of the local-sample-distance path.

The reference keeps its trained state as a flax pytree, ``{"params": module.params,
**module.state}`` (`_model.py:324`).  This module flattens that tree into
``{"qu/Dense_0/kernel": ndarray, ...}`` (float32, C-contiguous), which is what crosses the
C-ABI (`include/mrvi_b200.h`).  Only the sub-trees the path reads are kept: ``qu``
(`_module.py:173-202`) and ``qz`` (`_module.py:114-170`).  Shapes follow flax's auto-naming of
the modules in `_components.py`.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Mapping, Optional

import numpy as np

from ._constants import RESNET_HIDDEN


@dataclass(frozen=True)
class MrviDims:
    """Hyper-parameters that shape the hot path (reference `_module.py:209-230`, `:114-128`)."""

    n_input: int  # G
    n_sample: int  # S
    n_latent: int = 30  # L
    n_latent_u: Optional[int] = 10  # Lu; None <=> isomorphic u/z (`_module.py:245-246`)
    encoder_n_hidden: int = 128  # H
    encoder_n_layers: int = 2
    n_latent_sample: int = 16  # E  (outerprod_dim)
    n_channels: int = 4  # C
    n_heads: int = 2  # nh
    qz_n_hidden: int = 32  # M  (ResnetBlock *output* width inside MLP, `_components.py:68-75`)
    qz_n_layers: int = 1  # n_layers of the second MLP (`_components.py:223-229`)
    use_map: bool = True

    def __post_init__(self):
        if self.n_latent_u is not None and self.n_latent_u == self.n_latent:
            object.__setattr__(self, "n_latent_u", None)

    @property
    def n_latent_q(self) -> int:
        """Width of u (Lq): n_latent_u, or n_latent when isomorphic."""
        return self.n_latent if self.n_latent_u is None else self.n_latent_u

    @property
    def head_dim(self) -> int:
        # flax MHA: qkv_features // num_heads, qkv_features = n_channels * n_heads
        return self.n_channels

    @property
    def qz_out_dim(self) -> int:
        return self.n_latent * (1 if self.use_map else 2)

    def as_dict(self):
        return asdict(self)


def _resnet_block_shapes(prefix: str, n_in: int, n_out: int, out: "OrderedDict[str, tuple]"):
    """ResnetBlock (`_components.py:27-51`): Dense(R) LN [skip Dense(R) iff n_in != R] Dense(n_out) LN."""
    R = RESNET_HIDDEN
    out[f"{prefix}/Dense_0/kernel"] = (n_in, R)
    out[f"{prefix}/Dense_0/bias"] = (R,)
    out[f"{prefix}/LayerNorm_0/scale"] = (R,)
    out[f"{prefix}/LayerNorm_0/bias"] = (R,)
    k = 1
    if n_in != R:
        out[f"{prefix}/Dense_1/kernel"] = (n_in, R)
        out[f"{prefix}/Dense_1/bias"] = (R,)
        k = 2
    out[f"{prefix}/Dense_{k}/kernel"] = (R, n_out)
    out[f"{prefix}/Dense_{k}/bias"] = (n_out,)
    out[f"{prefix}/LayerNorm_1/scale"] = (n_out,)
    out[f"{prefix}/LayerNorm_1/bias"] = (n_out,)


def _mlp_shapes(prefix: str, n_in: int, n_hidden: int, n_layers: int, n_out: int, out):
    """MLP (`_components.py:54-75`): n_layers x ResnetBlock(n_out=n_hidden) then Dense(n_out)."""
    d = n_in
    for i in range(n_layers):
        _resnet_block_shapes(f"{prefix}/ResnetBlock_{i}", d, n_hidden, out)
        d = n_hidden
    out[f"{prefix}/Dense_0/kernel"] = (d, n_out)
    out[f"{prefix}/Dense_0/bias"] = (n_out,)


#: parameters read only by the sampled paths (use_mean=False / normalize_distances=True); optional at load time
OPTIONAL_PARAMS = ("qu/NormalDistOutputNN_0/Dense_1/kernel", "qu/NormalDistOutputNN_0/Dense_1/bias")


def expected_param_shapes(dims: MrviDims, with_scale_head: bool = True) -> "OrderedDict[str, tuple]":
    """Name -> shape of every parameter the path reads, in a fixed order."""
    G, S, H, Lq, L = dims.n_input, dims.n_sample, dims.encoder_n_hidden, dims.n_latent_q, dims.n_latent
    E, C, nh, hd = dims.n_latent_sample, dims.n_channels, dims.n_heads, dims.head_dim
    o: "OrderedDict[str, tuple]" = OrderedDict()
    # ---- qu = _EncoderXU (`_module.py:173-202`)
    o["qu/Dense_0/kernel"] = (G, H)
    o["qu/Dense_0/bias"] = (H,)
    o["qu/Dense_1/kernel"] = (H, H)
    o["qu/Dense_1/bias"] = (H,)
    for i in range(2):
        o[f"qu/ConditionalNormalization_{i}/gamma_conditional/embedding"] = (S, H)
        o[f"qu/ConditionalNormalization_{i}/beta_conditional/embedding"] = (S, H)
    o["qu/Embed_0/embedding"] = (S, H)
    d = H
    for i in range(dims.encoder_n_layers):
        _resnet_block_shapes(f"qu/NormalDistOutputNN_0/ResnetBlock_{i}", d, H, o)
        d = H
    o["qu/NormalDistOutputNN_0/Dense_0/kernel"] = (H, Lq)
    o["qu/NormalDistOutputNN_0/Dense_0/bias"] = (Lq,)
    if with_scale_head:  # softplus scale head of q(u|x) (`_components.py:96-97`): only the sampled paths read it
        o["qu/NormalDistOutputNN_0/Dense_1/kernel"] = (H, Lq)
        o["qu/NormalDistOutputNN_0/Dense_1/bias"] = (Lq,)
    # ---- qz = EncoderUZ (`_module.py:114-170`) + AttentionBlock (`_components.py:165-230`)
    o["qz/u_ln/scale"] = (Lq,)
    o["qz/u_ln/bias"] = (Lq,)
    o["qz/Embed_0/embedding"] = (S, E)
    o["qz/sample_embed_ln/scale"] = (E,)
    o["qz/sample_embed_ln/bias"] = (E,)
    ab = "qz/AttentionBlock_0"
    o[f"{ab}/DenseGeneral_0/kernel"] = (Lq, E, 1)
    o[f"{ab}/DenseGeneral_1/kernel"] = (E, E, 1)
    mha = f"{ab}/MultiHeadDotProductAttention_0"
    for nm in ("query", "key", "value"):
        o[f"{mha}/{nm}/kernel"] = (1, nh, hd)
        o[f"{mha}/{nm}/bias"] = (nh, hd)
    o[f"{mha}/out/kernel"] = (nh, hd, C)
    o[f"{mha}/out/bias"] = (C,)
    _mlp_shapes(f"{ab}/MLP_0", E * C, dims.qz_n_hidden, 1, E, o)
    _mlp_shapes(f"{ab}/MLP_1", Lq + E, dims.qz_n_hidden, dims.qz_n_layers, dims.qz_out_dim, o)
    if dims.n_latent_u is not None:
        o["qz/Dense_0/kernel"] = (dims.n_latent_u, L)
        o["qz/Dense_0/bias"] = (L,)
    return o


def flatten_params(tree: Mapping, prefix: str = "") -> "OrderedDict[str, np.ndarray]":
    """Flatten a (possibly nested / FrozenDict / already flat) pytree to ``path -> float32 array``.

    Accepts what the reference holds in ``module.params`` (optionally wrapped in
    ``{"params": ...}``); leaves are converted with ``np.asarray`` so jax arrays work
    without importing jax.
    """
    flat: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def rec(node, path):
        if isinstance(node, Mapping) or hasattr(node, "items"):
            for k, v in node.items():
                rec(v, f"{path}/{k}" if path else str(k))
        else:
            flat[path] = np.ascontiguousarray(np.asarray(node), dtype=np.float32)

    rec(tree, prefix)
    if flat and all(k.startswith("params/") for k in flat):
        flat = OrderedDict((k[len("params/"):], v) for k, v in flat.items())
    return flat


def infer_dims(flat: Mapping[str, np.ndarray], **overrides) -> MrviDims:
    """Recover the hyper-parameters from parameter shapes (what a trained model pins down)."""
    G, H = flat["qu/Dense_0/kernel"].shape
    S, E = flat["qz/Embed_0/embedding"].shape
    Lq = flat["qz/u_ln/scale"].shape[0]
    mha = "qz/AttentionBlock_0/MultiHeadDotProductAttention_0"
    _, nh, hd = flat[f"{mha}/query/kernel"].shape
    C = flat[f"{mha}/out/bias"].shape[0]
    n_enc = 0
    while f"qu/NormalDistOutputNN_0/ResnetBlock_{n_enc}/Dense_0/kernel" in flat:
        n_enc += 1
    n_qz = 0
    while f"qz/AttentionBlock_0/MLP_1/ResnetBlock_{n_qz}/Dense_0/kernel" in flat:
        n_qz += 1
    M, out_dim = flat["qz/AttentionBlock_0/MLP_1/Dense_0/kernel"].shape
    if "qz/Dense_0/kernel" in flat:
        Lu, L = flat["qz/Dense_0/kernel"].shape
    else:
        Lu, L = None, Lq
    if out_dim == L:
        use_map = True
    elif out_dim == 2 * L:
        use_map = False
    else:
        raise ValueError(f"qz output width {out_dim} is neither n_latent={L} nor 2*n_latent")
    if hd != C:
        raise ValueError("MultiHeadDotProductAttention head_dim must equal n_channels")
    kw = dict(
        n_input=int(G), n_sample=int(S), n_latent=int(L), n_latent_u=None if Lu is None else int(Lu),
        encoder_n_hidden=int(H), encoder_n_layers=n_enc, n_latent_sample=int(E), n_channels=int(C),
        n_heads=int(nh), qz_n_hidden=int(M), qz_n_layers=n_qz, use_map=use_map,
    )
    kw.update(overrides)
    return MrviDims(**kw)


def validate_params(flat: Mapping[str, np.ndarray], dims: MrviDims) -> "OrderedDict[str, np.ndarray]":
    """Check every expected parameter is present with the right shape; return them in schema order."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in expected_param_shapes(dims).items():
        if name not in flat and name in OPTIONAL_PARAMS:
            continue
        if name not in flat:
            raise KeyError(f"missing parameter {name!r} (expected shape {shape})")
        a = np.ascontiguousarray(np.asarray(flat[name]), dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"parameter {name!r} has shape {tuple(a.shape)}, expected {tuple(shape)}")
        if not np.all(np.isfinite(a)):
            raise ValueError(f"parameter {name!r} contains non-finite values")
        out[name] = a
    return out


def init_params(dims: MrviDims, seed: int = 0, jitter: float = 0.1, embed_std: float = 0.1
                ) -> "OrderedDict[str, np.ndarray]":
    """Random parameters following the reference initialisers.

    ``Dense`` kernels: U(+-1/sqrt(fan_in)) (`_components.py:18-22`); flax ``Dense``/``DenseGeneral``
    /MHA kernels: lecun_normal; embeddings N(0, 0.1^2) (`_module.py:37`); gamma 1+0.02 N(0,1),
    beta 0 (`_components.py:109-127`).  ``jitter`` > 0 additionally perturbs every bias and every
    LayerNorm scale/bias with N(0, jitter^2) so that no parameter is at a value (0 or 1) that would
    hide a missing term in a kernel.  There is no trained checkpoint in this environment.
    """
    main_rng = np.random.default_rng(seed)
    opt_rng = np.random.default_rng([seed, 1])  # own stream: adding the optional heads leaves every other draw unchanged
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in expected_param_shapes(dims).items():
        rng = opt_rng if name in OPTIONAL_PARAMS else main_rng
        leaf = name.rsplit("/", 1)[1]
        if leaf == "embedding":
            if "gamma_conditional" in name:
                a = 1.0 + 0.02 * rng.standard_normal(shape)
            elif "beta_conditional" in name:
                a = jitter * rng.standard_normal(shape)
            else:
                a = embed_std * rng.standard_normal(shape)
        elif leaf == "kernel":
            is_ref_dense = ("DenseGeneral" not in name and "MultiHead" not in name
                            and name != "qz/Dense_0/kernel")
            if "MultiHeadDotProductAttention_0/out" in name:
                fan_in = shape[0] * shape[1]
            else:
                fan_in = shape[0]
            if is_ref_dense:
                b = 1.0 / np.sqrt(fan_in)
                a = rng.uniform(-b, b, size=shape)
            else:  # lecun_normal (truncated normal in flax; plain normal is close enough for synthetic weights)
                a = rng.standard_normal(shape) / np.sqrt(fan_in)
        elif leaf == "scale":
            a = 1.0 + jitter * rng.standard_normal(shape)
        elif leaf == "bias":
            a = jitter * rng.standard_normal(shape)
        else:  # pragma: no cover
            raise AssertionError(name)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


def save_params(path: str, flat: Mapping[str, np.ndarray], dims: MrviDims) -> None:
    """Write the flat parameters + dims to an ``.npz`` (the exchange format of INTEGRATION.md)."""
    meta = {f"__dims__/{k}": np.asarray(-1 if v is None else v) for k, v in dims.as_dict().items()}
    np.savez(path, **{k.replace("/", "|"): v for k, v in flat.items()},
             **{k.replace("/", "|"): v for k, v in meta.items()})


def load_params(path: str):
    """Inverse of :func:`save_params`: returns ``(flat, dims)``."""
    z = np.load(path)
    flat, meta = OrderedDict(), {}
    for k in z.files:
        name = k.replace("|", "/")
        if name.startswith("__dims__/"):
            v = z[k].item()
            meta[name[len("__dims__/"):]] = v
        else:
            flat[name] = z[k].astype(np.float32)
    if meta:
        if meta.get("n_latent_u", -1) == -1:
            meta["n_latent_u"] = None
        meta["use_map"] = bool(meta["use_map"])
        dims = MrviDims(**meta)
    else:
        dims = infer_dims(flat)
    return flat, dims
