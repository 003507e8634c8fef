"""Duck-typed AnnData and synthetic inputs for the local-sample-distance path.

``anndata`` / ``scvi-tools`` are not installed in the build image, so the host API accepts any
object with ``.X`` (dense ndarray or scipy.sparse), ``.obs`` (pandas DataFrame), ``.obs_names``,
``.n_obs``, ``.n_vars`` -- which a real ``anndata.AnnData`` satisfies.  ``synthetic_iid`` follows the
recipe of ``scvi.data.synthetic_iid`` that the reference's tests use (`tests/test_model.py:7-24`):
NB(5, 0.3) counts times a Bernoulli(0.7) dropout mask, 2 batches, 3 labels.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class SimpleAnnData:
    """Minimal stand-in for ``anndata.AnnData`` (only what `_model.py:314-319, 400, 470-495` touches)."""

    def __init__(self, X, obs: pd.DataFrame | None = None, obs_names=None):
        self.X = X
        n = X.shape[0]
        if obs is None:
            obs = pd.DataFrame(index=pd.Index([f"cell_{i}" for i in range(n)] if obs_names is None else obs_names))
        elif obs_names is not None:
            obs = obs.copy()
            obs.index = pd.Index(obs_names)
        if len(obs) != n:
            raise ValueError(f"obs has {len(obs)} rows, X has {n}")
        self.obs = obs

    @property
    def obs_names(self):
        return self.obs.index

    @property
    def n_obs(self) -> int:
        return self.X.shape[0]

    @property
    def n_vars(self) -> int:
        return self.X.shape[1]

    @property
    def shape(self):
        return self.X.shape


def synthetic_counts(n_obs: int, n_vars: int, seed: int = 0, dtype=np.float32, chunk: int = 65536):
    """NB(n=5, p=0.3) x Bernoulli(0.7) counts as a dense float array (the synthetic_iid recipe)."""
    rng = np.random.default_rng(seed)
    X = np.empty((n_obs, n_vars), dtype=dtype)
    for lo in range(0, n_obs, chunk):
        hi = min(lo + chunk, n_obs)
        c = rng.negative_binomial(5, 0.3, size=(hi - lo, n_vars))
        c *= rng.random((hi - lo, n_vars)) < 0.7
        X[lo:hi] = c
    return X


def synthetic_iid(n_obs: int = 400, n_vars: int = 100, n_sample: int = 4, n_batches: int = 2,
                  n_labels: int = 3, seed: int = 0, group_keys=("labels", "label_2")) -> SimpleAnnData:
    """400 cells x 100 genes by default, with ``sample``, ``batch``, ``labels``, ``label_2`` columns."""
    rng = np.random.default_rng(seed + 1)
    X = synthetic_counts(n_obs, n_vars, seed)
    sample = rng.integers(0, n_sample, size=n_obs)
    obs = pd.DataFrame(
        {
            "sample": sample,
            "sample_str": pd.Categorical([_sample_name(i) for i in sample]),
            "batch": pd.Categorical([f"batch_{i}" for i in np.repeat(np.arange(n_batches), -(-n_obs // n_batches))[:n_obs]]),
            "labels": pd.Categorical([f"label_{i}" for i in rng.integers(0, n_labels, size=n_obs)]),
            "label_2": rng.integers(0, 2, size=n_obs),
        },
        index=pd.Index([str(i) for i in range(n_obs)]),
    )
    return SimpleAnnData(X, obs)


def _sample_name(i: int) -> str:
    # 'a'..'o' for 15 samples as in the reference tests; longer alphabets fall back to s<i>
    return chr(int(i) + ord("a")) if i < 26 else f"s{int(i):04d}"


def dense_float32(X, lo: int | None = None, hi: int | None = None) -> np.ndarray:
    """Rows [lo, hi) of X as a C-contiguous float32 ndarray (sparse rows are densified), the
    role of scvi's AnnDataLoader densify step (`_model.py:317-319`)."""
    sl = slice(lo, hi)
    blk = X[sl]
    if hasattr(blk, "toarray"):
        blk = blk.toarray()
    return np.ascontiguousarray(np.asarray(blk), dtype=np.float32)
