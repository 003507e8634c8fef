"""CPU oracle for MrVI's counterfactual local-sample-distance path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The shipped path
(``mrvi_b200``) never imports ``oracle`` and raises if its CUDA library is missing.

It restates, operation by operation and in numpy, the reference's
``MrVI.get_local_sample_distances`` (``use_mean=True``, eval mode).  Every function cites the
reference ``file:line`` it follows (paths relative to the reference repo root,
``src/mrvi/...``).

PARITY PINNED AGAINST THE REFERENCE SOURCE, UNPINNED ONLY at the third-party boundary.  The reference executes its arithmetic through
flax.linen / jax / numpyro / scvi-tools, none of which is installed here (no wheel, no network),
and its own tests hold no numeric golden vectors for this path (``tests/test_model.py:233-241``
check only shapes and symmetry).  So the semantics of the flax primitives below
(``LayerNorm`` eps=1e-6 with fast variance, tanh-``gelu``, ``DenseGeneral``, ``Embed``,
``MultiHeadDotProductAttention`` with 1/sqrt(head_dim) query scaling, no mask, dropout off) are
restated from their published definitions, not checked against a run of the reference.  What IS
pinned: (i) the composition of those primitives is cross-checked by executing the reference's own
``_components.py`` / ``_module.py`` source on a numpy shim of the flax API
(``oracle/flax_shim`` + ``oracle/gen_golden.py``; fixtures in ``tests/golden``), (ii) two
independent restatements (this sequential-over-samples numpy one and the vectorised torch one in
``oracle/mrvi_oracle_torch.py``) agree, (iii) the invariants of SURVEY.md section 8(c).

Arithmetic dtype is a parameter: float64 is the numeric oracle, float32 mirrors the reference's
JAX default (x64 disabled).
"""
from __future__ import annotations

import numpy as np

LN_EPS = 1e-6  # flax.linen.LayerNorm default epsilon
R_HIDDEN = 128  # ResnetBlock.n_hidden default, `_components.py:31`


# --------------------------------------------------------------------------- primitives (flax)
def layer_norm(v, scale=None, bias=None, eps=LN_EPS):
    """flax.linen.LayerNorm over the last axis, ``use_fast_variance=True``:
    mean = E[v], var = max(0, E[v^2] - mean^2), y = (v - mean) * rsqrt(var + eps) [* scale] [+ bias]."""
    mu = v.mean(axis=-1, keepdims=True)
    var = np.maximum((v * v).mean(axis=-1, keepdims=True) - mu * mu, 0)
    mul = 1.0 / np.sqrt(var + v.dtype.type(eps))
    if scale is not None:
        mul = mul * scale
    y = (v - mu) * mul
    if bias is not None:
        y = y + bias
    return y


def gelu(v):
    """jax.nn.gelu(approximate=True) == flax ``nn.gelu`` default (`_module.py:56,128,178`; `_components.py:178`)."""
    t = v.dtype.type
    c = t(np.sqrt(2.0 / np.pi))
    return t(0.5) * v * (t(1.0) + np.tanh(c * (v + t(0.044715) * v * v * v)))


def relu(v):
    return np.maximum(v, 0)


def dense(v, p, name):
    """``Dense`` / ``nn.Dense``: v @ kernel + bias (`_components.py:15-24`)."""
    return v @ p[f"{name}/kernel"] + p[f"{name}/bias"]


def softmax(v, axis=-1):
    m = v.max(axis=axis, keepdims=True)
    e = np.exp(v - m)
    return e / e.sum(axis=axis, keepdims=True)


# --------------------------------------------------------------------------- building blocks
def resnet_block(h, p, prefix, act_internal, act_output):
    """ResnetBlock `_components.py:36-51`."""
    n_in = h.shape[-1]
    t = dense(h, p, f"{prefix}/Dense_0")
    t = layer_norm(t, p[f"{prefix}/LayerNorm_0/scale"], p[f"{prefix}/LayerNorm_0/bias"])
    t = act_internal(t)
    if n_in != R_HIDDEN:
        t = t + dense(h, p, f"{prefix}/Dense_1")
        k = 2
    else:
        t = t + h
        k = 1
    t = dense(t, p, f"{prefix}/Dense_{k}")
    t = layer_norm(t, p[f"{prefix}/LayerNorm_1/scale"], p[f"{prefix}/LayerNorm_1/bias"])
    return act_output(t)


def mlp(h, p, prefix, act):
    """MLP `_components.py:63-75`: ResnetBlock_* then a final Dense."""
    i = 0
    while f"{prefix}/ResnetBlock_{i}/Dense_0/kernel" in p:
        h = resnet_block(h, p, f"{prefix}/ResnetBlock_{i}", act, act)
        i += 1
    return dense(h, p, f"{prefix}/Dense_0")


def conditional_normalization(h, sid, p, prefix):
    """ConditionalNormalization `_components.py:130-162` (normalization_type="layer")."""
    h = layer_norm(h)  # use_bias=False, use_scale=False
    gamma = p[f"{prefix}/gamma_conditional/embedding"][sid]
    beta = p[f"{prefix}/beta_conditional/embedding"][sid]
    return gamma * h + beta


def _encoder_xu_trunk(p, x, sid):
    """_EncoderXU `_module.py:182-200` up to the last ResnetBlock of NormalDistOutputNN (`_components.py:92-94`)."""
    h = np.log1p(x)
    for i in range(2):
        h = dense(h, p, f"qu/Dense_{i}")
        h = conditional_normalization(h, sid, p, f"qu/ConditionalNormalization_{i}")
        h = gelu(h)
    h = h + p["qu/Embed_0/embedding"][sid]
    i = 0
    nd = "qu/NormalDistOutputNN_0"
    while f"{nd}/ResnetBlock_{i}/Dense_0/kernel" in p:
        h = resnet_block(h, p, f"{nd}/ResnetBlock_{i}", relu, relu)
        i += 1
    return h


def encoder_xu_mean(p, x, sid):
    """_EncoderXU `_module.py:182-202` + NormalDistOutputNN mean head `_components.py:86-97`.
    Returns qu.mean, i.e. u under use_mean=True (`_module.py:312-314`)."""
    return dense(_encoder_xu_trunk(p, x, sid), p, "qu/NormalDistOutputNN_0/Dense_0")


def softplus(v):
    return np.logaddexp(v, 0.0)


def encoder_xu(p, x, sid):
    """q(u|x) = Normal(mean, scale): mean head and `softplus(Dense(h)) + scale_eps` head, scale_eps = 1e-5
    (`_components.py:95-97`)."""
    h = _encoder_xu_trunk(p, x, sid)
    nd = "qu/NormalDistOutputNN_0"
    return dense(h, p, f"{nd}/Dense_0"), softplus(dense(h, p, f"{nd}/Dense_1")) + p[f"{nd}/Dense_0/bias"].dtype.type(1e-5)


def multi_head_attention(q_in, kv_in, p, prefix):
    """flax MultiHeadDotProductAttention as called at `_components.py:199-205`
    (inputs (B, E, 1); no mask; deterministic)."""
    q = np.einsum("bti,ihd->bthd", q_in, p[f"{prefix}/query/kernel"]) + p[f"{prefix}/query/bias"]
    k = np.einsum("bti,ihd->bthd", kv_in, p[f"{prefix}/key/kernel"]) + p[f"{prefix}/key/bias"]
    v = np.einsum("bti,ihd->bthd", kv_in, p[f"{prefix}/value/kernel"]) + p[f"{prefix}/value/bias"]
    depth = q.shape[-1]
    q = q / np.sqrt(depth).astype(q.dtype)
    w = np.einsum("bqhd,bkhd->bhqk", q, k)
    w = softmax(w, axis=-1)
    a = np.einsum("bhqk,bkhd->bqhd", w, v)
    return np.einsum("bqhd,hdc->bqc", a, p[f"{prefix}/out/kernel"]) + p[f"{prefix}/out/bias"]


def attention_block(u_, e_s, p, prefix="qz/AttentionBlock_0"):
    """AttentionBlock `_components.py:181-230` without MC axis."""
    q_tok = np.einsum("bl,lei->bei", u_, p[f"{prefix}/DenseGeneral_0/kernel"])  # (B,E,1)  :195-197
    kv_tok = np.einsum("bl,lei->bei", e_s, p[f"{prefix}/DenseGeneral_1/kernel"])  # (B,E,1)  :198
    o = multi_head_attention(q_tok, kv_tok, p, f"{prefix}/MultiHeadDotProductAttention_0")  # (B,E,C)
    a = o.reshape(o.shape[0], o.shape[1] * o.shape[2])  # :209-210
    eps_ = mlp(a, p, f"{prefix}/MLP_0", gelu)  # :216-221
    c = np.concatenate([u_, eps_], axis=-1)  # :222
    return mlp(c, p, f"{prefix}/MLP_1", gelu)  # :223-229


def encoder_uz(p, u, s_idx):
    """EncoderUZ `_module.py:131-170`: returns (z_base, residual)."""
    u_ = layer_norm(u, p["qz/u_ln/scale"], p["qz/u_ln/bias"])
    e = p["qz/Embed_0/embedding"][s_idx]
    e = layer_norm(e, p["qz/sample_embed_ln/scale"], p["qz/sample_embed_ln/bias"])
    residual = attention_block(u_, e, p)
    if "qz/Dense_0/kernel" in p:
        z_base = dense(u, p, "qz/Dense_0")  # raw u, `_module.py:166-168`
    else:
        z_base = u
    return z_base, residual


def inference_z(p, x, sid, cf_sample):
    """MrVAE.inference(...)["z"] with use_mean=True (`_module.py:308-343`)."""
    u = encoder_xu_mean(p, x, sid)
    z_base, eps = encoder_uz(p, u, cf_sample)
    L = z_base.shape[-1]
    if eps.shape[-1] == 2 * L:  # use_map=False: qeps.mean = first half (`_module.py:326-329`)
        eps = eps[..., :L]
    return z_base + eps


def cast_params(flat, dtype):
    return {k: np.asarray(v, dtype=dtype) for k, v in flat.items()}


def counterfactual_z(flat, x, sid, dtype=np.float64):
    """mapped_inference_fn `_model.py:326-365` (the sequential ``lax.map`` form): z[b, s, :] for every
    counterfactual sample s, output axes (cell, sample, latent) (`out_axes=-2`, `:352`)."""
    p = cast_params(flat, dtype)
    x = np.asarray(x, dtype=dtype)
    sid = np.asarray(sid).astype(np.int64).reshape(-1)
    S = p["qz/Embed_0/embedding"].shape[0]
    zs = []
    for s in range(S):
        cf = np.full(x.shape[0], s, dtype=np.int64)
        zs.append(inference_z(p, x, sid, cf))
    return np.stack(zs, axis=1)


def pairwise_distances(z, norm="l2"):
    """_compute_distances_from_representations `_model.py:545-560` (vmap over cells)."""
    delta = z[:, None, :, :] - z[:, :, None, :]  # [b, i, j] = rep[j] - rep[i]
    if norm == "l2":
        return np.sqrt((delta ** 2).sum(-1))
    if norm == "l1":
        return np.abs(delta).sum(-1)
    if norm == "linf":
        return np.abs(delta).max(-1)
    raise ValueError(f"norm {norm} not supported")


def local_sample_distances(flat, X, sample_idx, group_labels=None, keep_cell=True, batch_size=256,
                           norm="l2", dtype=np.float64, return_z=False):
    """The hot loop of ``compute_local_statistics`` (`_model.py:376-504`) on plain arrays.

    ``group_labels``: dict key -> array of labels (any hashable), one per cell.  Per batch, per
    category present, the (B,S,S) blocks are summed and added to a running sum (`:469-484`); at
    the end every running sum is divided by the category's count and categories are emitted in
    ``value_counts()`` order (`:493-504`) -- the caller passes that order in via pandas; here we
    return sums keyed by label plus counts, and ``group_means`` in descending-count order with
    first-appearance tie-break (what ``pd.Series.value_counts`` does for object dtype).
    """
    n = X.shape[0]
    out = {}
    cell_blocks, z_blocks = [], []
    sums = {k: {} for k in (group_labels or {})}
    for lo in range(0, n, batch_size):
        hi = min(lo + batch_size, n)
        z = counterfactual_z(flat, X[lo:hi], sample_idx[lo:hi], dtype=dtype)
        if return_z:
            z_blocks.append(z)
        d = pairwise_distances(z, norm)
        if keep_cell:
            cell_blocks.append(d)
        for key, labels in (group_labels or {}).items():
            lab = np.asarray(labels)[lo:hi]
            seen = []
            for c in lab:  # unique(), order of first appearance
                if c not in seen:
                    seen.append(c)
            for c in seen:
                part = d[lab == c].sum(axis=0)
                if c not in sums[key]:
                    sums[key][c] = part
                else:
                    sums[key][c] = sums[key][c] + part
    if keep_cell:
        out["cell"] = np.concatenate(cell_blocks, axis=0) if cell_blocks else np.zeros((0, 0, 0), dtype)
    if return_z:
        out["z"] = np.concatenate(z_blocks, axis=0)
    for key, labels in (group_labels or {}).items():
        lab = list(np.asarray(labels))
        cats, counts = [], {}
        for c in lab:
            if c not in counts:
                counts[c] = 0
                cats.append(c)
            counts[c] += 1
        order = sorted(range(len(cats)), key=lambda i: -counts[cats[i]])  # stable
        cats = [cats[i] for i in order]
        out[key] = {
            "labels": cats,
            "counts": np.array([counts[c] for c in cats], dtype=np.int64),
            "means": np.stack([sums[key][c] / dtype(counts[c]) for c in cats], axis=0),
        }
    return out


# --------------------------------------------------------------------------- sampled paths
# use_mean=False / normalize_distances=True.  The reference draws its noise from JAX threefry streams that
# cannot be reproduced without jax; here the standard-normal draws are INPUTS, so that the same draws can be
# handed to the CUDA path and the two compared exactly.  Shapes follow the reference's own sampling calls.


def _qz_any(p, u, s_idx, noise_eps=None):
    """z = z_base + eps for rows with arbitrary leading shape (`_module.py:320-331`); with use_map=False
    eps = loc + (softplus(raw) + 1e-3) * noise (`_module.py:326-329`)."""
    lead = u.shape[:-1]
    u2 = u.reshape(-1, u.shape[-1])
    s2 = np.broadcast_to(s_idx, lead).reshape(-1)
    z_base, eps = encoder_uz(p, u2, s2)
    L = z_base.shape[-1]
    if eps.shape[-1] == 2 * L:
        loc, raw = eps[:, :L], eps[:, L:]
        if noise_eps is None:
            raise ValueError("use_map=False needs noise_eps for the sampled path")
        eps = loc + (softplus(raw) + loc.dtype.type(1e-3)) * noise_eps.reshape(-1, L)
    return (z_base + eps).reshape(*lead, L)


def sampled_counterfactual_z(flat, x, sid, noise_u, noise_eps=None, dtype=np.float64):
    """``mapped_inference_fn(..., use_mean=False, mc_samples=mc)`` (`_model.py:326-365, 406-414`).

    The vmapped call for counterfactual sample s owns its own "u" PRNG stream (`_model.py:136-159`), so it
    draws its own ``u = qu.rsample(rng_s, (mc,))`` of shape (mc, n, Lq) (`_module.py:315-318`):
    ``noise_u[s]`` stands for those draws, ``noise_eps[s]`` (mc, n, L) for the "eps" stream when use_map=False.
    Returns (n, mc, S, L): ``out_axes=-2`` then the transpose of `_model.py:414`."""
    p = cast_params(flat, dtype)
    x = np.asarray(x, dtype=dtype)
    sid = np.asarray(sid).astype(np.int64).reshape(-1)
    noise_u = np.asarray(noise_u, dtype=dtype)
    S = p["qz/Embed_0/embedding"].shape[0]
    mean, scale = encoder_xu(p, x, sid)
    zs = []
    for s in range(S):
        u = mean[None] + scale[None] * noise_u[s]  # (mc, n, Lq)
        cf = np.full(x.shape[0], s, dtype=np.int64)
        zs.append(_qz_any(p, u, cf[None, :], None if noise_eps is None else np.asarray(noise_eps, dtype=dtype)[s]))
    z = np.stack(zs, axis=2)  # (mc, n, S, L)
    return z.transpose(1, 0, 2, 3)


def pairwise_distances_mc(z, norm="l2"):
    """``jax.vmap(jax.vmap(_compute_distance))`` over (cell, mc_sample) (`_model.py:572-573`)."""
    n, mc = z.shape[:2]
    return pairwise_distances(z.reshape(n * mc, *z.shape[2:]), norm).reshape(n, mc, z.shape[2], z.shape[2])


def local_baseline_dists(flat, x, sid, noise_base, noise_eps=None, dtype=np.float64):
    """``_compute_local_baseline_dists`` (`_model.py:508-540`): 2*mc draws of z for the cell's OWN sample,
    l2 distance between draw m and draw m+mc, mean and (population) variance over m.  ``noise_base``:
    (2*mc, n, Lq)."""
    p = cast_params(flat, dtype)
    x = np.asarray(x, dtype=dtype)
    sid = np.asarray(sid).astype(np.int64).reshape(-1)
    noise_base = np.asarray(noise_base, dtype=dtype)
    mean, scale = encoder_xu(p, x, sid)
    u = mean[None] + scale[None] * noise_base
    z = _qz_any(p, u, sid[None, :], None if noise_eps is None else np.asarray(noise_eps, dtype=dtype))
    mc = noise_base.shape[0] // 2
    a, b = z[:mc], z[mc:]
    l2 = np.sqrt(((a - b) ** 2).sum(axis=2)).T  # (n, mc)
    return l2.mean(axis=1), l2.var(axis=1)


def sampled_local_sample_distances(flat, X, sample_idx, noise_u, noise_base=None, normalize=False, dtype=np.float64):
    """The ``sampled_distances`` / ``normalized_distances`` inputs of ``compute_local_statistics``
    (`_model.py:430-450`): (n, mc, S, S) sampled distances, or their per-cell z-score against the local
    baseline averaged over mc, (n, S, S)."""
    z = sampled_counterfactual_z(flat, X, sample_idx, noise_u, dtype=dtype)
    d = pairwise_distances_mc(z)
    if not normalize:
        return d
    mu, var = local_baseline_dists(flat, X, sample_idx, noise_base, dtype=dtype)
    return ((d - mu[:, None, None, None]) / np.sqrt(var)[:, None, None, None]).mean(axis=1)


# --------------------------------------------------------------------------- DE eps stage
# `MrVI.differential_expression` (`_model.py:1012-1420`) up to its latent-space outputs (beta, effect_size, pvalue,
# padj).  The LFC / baseline-expression outputs go through the decoder and are out of the path's scope.


def sampled_counterfactual_eps(flat, x, sid, noise_u, noise_eps=None, dtype=np.float64, samples=None):
    """``inference(..., use_mean=False, mc_samples=mc)["eps"]`` vmapped over the counterfactual samples with
    ``out_axes=-2`` (`_model.py:1175-1194`): (mc, n, S_kept, L).  ``noise_u[s]`` (mc, n, Lq) are the draws of the
    vmapped call of sample s (`_module.py:315-318`); ``samples`` = the kept sample codes (`_model.py:1310-1312`)."""
    p = cast_params(flat, dtype)
    x = np.asarray(x, dtype=dtype)
    sid = np.asarray(sid).astype(np.int64).reshape(-1)
    noise_u = np.asarray(noise_u, dtype=dtype)
    S = p["qz/Embed_0/embedding"].shape[0]
    samples = np.arange(S) if samples is None else np.asarray(samples)
    mean, scale = encoder_xu(p, x, sid)
    out = []
    for s in samples:
        u = mean[None] + scale[None] * noise_u[s]  # (mc, n, Lq)
        lead = u.shape[:-1]
        cf = np.full(lead, s, dtype=np.int64).reshape(-1)
        z_base, eps = encoder_uz(p, u.reshape(-1, u.shape[-1]), cf)
        L = z_base.shape[-1]
        if eps.shape[-1] == 2 * L:  # use_map=False: eps ~ Normal(loc, softplus(raw) + 1e-3) (`_module.py:326-329`)
            loc, raw = eps[:, :L], eps[:, L:]
            eps = loc + (softplus(raw) + loc.dtype.type(1e-3)) * np.asarray(noise_eps, dtype=dtype)[s].reshape(-1, L)
        out.append(eps.reshape(*lead, L))
    return np.stack(out, axis=2)


def de_process_design_matrix(Xmat, admissible, lambd=0.0):
    """``process_design_matrix`` (`_model.py:1147-1160`): per cell, with D = diag(admissible samples),
    xtmx = X^T D X + lambd I, prefactor = real(sqrtm(xtmx)), Amat = pinv(xtmx) X^T D.
    Xmat (S_kept, K), admissible (n, S_kept) bool -> Amat (n, K, S_kept), prefactor (n, K, K)."""
    import scipy.linalg

    Xmat = np.asarray(Xmat, dtype=np.float64)
    adm = np.asarray(admissible, dtype=np.float64)
    K = Xmat.shape[1]
    xtmx = np.einsum("ak,nk,km->nam", Xmat.T, adm, Xmat) + lambd * np.eye(K)
    prefactor = np.stack([np.real(scipy.linalg.sqrtm(m)) for m in xtmx])
    inv_ = np.stack([np.linalg.pinv(m) for m in xtmx])
    Amat = np.einsum("nab,bc,nc->nac", inv_, Xmat.T, adm)
    return Amat, prefactor


def de_statistics(eps_, Amat, prefactor, n_samples_per_cell):
    """The latent-space statistics of `_model.py:1196-1215`: standardise eps over the samples (population std,
    ``1e-6 +`` in the denominator), betas = Amat . eps, whitened by ``prefactor``; effect size = mean over the mc
    draws of the squared norm over latent dims, p-value = chi2 survival with df = admissible samples of the cell;
    returned betas are rescaled by the std and averaged over the draws (`:1283`).
    eps_ (mc, n, S, L) -> beta (n, K, L), effect_size (n, K), pvalue (n, K)."""
    import scipy.stats

    eps_ = np.asarray(eps_)
    eps_std = eps_.std(axis=2, keepdims=True)
    eps_mean = eps_.mean(axis=2, keepdims=True)
    eps = (eps_ - eps_mean) / (1e-6 + eps_std)
    betas = np.einsum("nks,ansd->ankd", Amat, eps)
    betas_norm = np.einsum("ankd,nkl->anld", betas, prefactor)
    ts = (betas_norm ** 2).mean(axis=0).sum(axis=-1)
    pvals = 1 - scipy.stats.chi2.cdf(ts, df=np.asarray(n_samples_per_cell)[:, None])
    betas = betas * eps_std
    return betas.mean(0), ts, pvals


def fdr_bh(pvals):
    """Benjamini-Hochberg adjusted p-values, ``statsmodels.stats.multitest.multipletests(p, method="fdr_bh")[1]``
    (`_model.py:1363`): p_(i) * m / i, made monotone from the largest p downwards, clipped to 1."""
    p = np.asarray(pvals, dtype=np.float64).reshape(-1)
    m = p.size
    order = np.argsort(p)
    ranked = p[order] * m / np.arange(1, m + 1)
    ranked = np.minimum.accumulate(ranked[::-1])[::-1]
    out = np.empty(m)
    out[order] = np.minimum(ranked, 1.0)
    return out.reshape(np.shape(pvals))
