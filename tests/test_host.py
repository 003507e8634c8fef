"""CPU tests of the host-side logic: parameter schema, reduction planner, output containers, the
C-ABI library (loads, exports every symbol of include/mrvi_b200.h, fails loudly without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

import mrvi_b200
from mrvi_b200 import _capi, _params
from mrvi_b200._types import MrVIReduction
from mrvi_b200._utils import _parse_local_statistics_requirements
from mrvi_b200._xr import _LiteDataArray, _LiteDataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_schema_roundtrip(tmp_path):
    dims = mrvi_b200.MrviDims(n_input=37, n_sample=5, n_latent=12, n_latent_u=4, qz_n_layers=2, use_map=False)
    p = mrvi_b200.init_params(dims, seed=3)
    assert list(p) == list(mrvi_b200.expected_param_shapes(dims))
    assert mrvi_b200.infer_dims(p) == dims
    nested = {}
    for k, v in p.items():
        node = nested
        parts = k.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = v
    nested["px"] = {"Dense_0": {"kernel": np.zeros((3, 3))}}  # decoder params are ignored by the path
    flat = mrvi_b200.flatten_params({"params": nested})
    assert all(np.array_equal(flat[k], p[k]) for k in p)
    mrvi_b200.save_params(str(tmp_path / "m.npz"), p, dims)
    p2, d2 = mrvi_b200.load_params(str(tmp_path / "m.npz"))
    assert d2 == dims and all(np.array_equal(p2[k], p[k]) for k in p)


def test_isomorphic_dims_and_validation():
    dims = mrvi_b200.MrviDims(n_input=20, n_sample=3, n_latent=10, n_latent_u=10)
    assert dims.n_latent_u is None and dims.n_latent_q == 10
    p = mrvi_b200.init_params(dims)
    assert "qz/Dense_0/kernel" not in p
    bad = dict(p)
    bad["qz/u_ln/scale"] = np.ones(3, np.float32)
    with pytest.raises(ValueError, match="shape"):
        mrvi_b200.validate_params(bad, dims)
    del bad["qz/u_ln/scale"]
    with pytest.raises(KeyError, match="missing parameter"):
        mrvi_b200.validate_params(bad, dims)


def test_requirements_parser_truth_table():
    f = lambda x: x
    r = _parse_local_statistics_requirements([
        MrVIReduction("cell", "mean_distances", f), MrVIReduction("ct", "mean_distances", f, "ct")])
    assert r.needs_mean_representations and r.needs_mean_distances
    assert not (r.needs_sampled_representations or r.needs_sampled_distances or r.needs_normalized_distances)
    assert [x.name for x in r.ungrouped_reductions] == ["cell"] and [x.name for x in r.grouped_reductions] == ["ct"]
    r = _parse_local_statistics_requirements([MrVIReduction("n", "normalized_distances", f)])
    assert r.needs_sampled_representations and r.needs_normalized_distances and not r.needs_sampled_distances
    r = _parse_local_statistics_requirements([MrVIReduction("s", "sampled_distances", f)])
    assert r.needs_sampled_representations and r.needs_sampled_distances
    with pytest.raises(ValueError, match="Unknown reduction input"):
        _parse_local_statistics_requirements([MrVIReduction("x", "nope", f)])


def test_lite_xarray_layout():
    d = _LiteDataArray(np.arange(24.0).reshape(2, 3, 4), dims=["cell_name", "sample_x", "sample_y"],
                       coords={"cell_name": np.array(["a", "b"]), "sample_x": np.arange(3), "sample_y": np.arange(4)},
                       name="sample_distances")
    assert d.shape == (2, 3, 4) and d[0].shape == (3, 4) and d[0].dims == ("sample_x", "sample_y")
    assert np.array_equal(d.sel(cell_name=np.array(["b"])).values[0], d.values[1])
    assert np.array_equal(d.sum("cell_name").values, d.values.sum(0))
    assert np.array_equal(np.asarray(d[1].values.T), np.asarray(d[1].T.values))
    ds = _LiteDataset({"cell": d})
    assert "cell" in ds and ds.cell is d and list(ds.keys()) == ["cell"]
    with pytest.raises(ValueError):
        _LiteDataArray(np.zeros((2, 2)), dims=["a"])


def test_synthetic_iid_recipe():
    ad = mrvi_b200.synthetic_iid()
    assert ad.X.shape == (400, 100) and ad.X.dtype == np.float32 and ad.n_obs == 400 and ad.n_vars == 100
    assert (ad.X >= 0).all() and np.array_equal(ad.X, np.round(ad.X))
    assert 0.2 < (ad.X == 0).mean() < 0.4  # Bernoulli(0.7) dropout mask
    assert {"sample", "sample_str", "batch", "labels", "label_2"} <= set(ad.obs.columns)
    assert isinstance(ad.obs["labels"].dtype, pd.CategoricalDtype)


def test_capi_library_loads_and_exports_header_symbols(built_library):
    header = open(os.path.join(ROOT, "include", "mrvi_b200.h")).read()
    declared = set(re.findall(r"\b(mrvi_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(built_library, sym), sym
    assert built_library.mrvi_abi_version() == _capi.ABI_VERSION
    assert int(re.search(r"#define MRVI_B200_ABI_VERSION (\d+)", header).group(1)) == _capi.ABI_VERSION
    assert built_library.mrvi_last_error() is not None
    assert _capi.kernel_launch_count() >= 0


def test_capi_argument_validation_without_gpu(built_library):
    """Error paths that never reach a kernel: null / bad dims return a negative status with a message;
    without a B200 mrvi_create fails loudly (no CPU fallback)."""
    out = ctypes.c_void_p()
    assert built_library.mrvi_create(None, 0, None, None, None, 0, 0, ctypes.byref(out)) == -1
    assert b"null" in built_library.mrvi_last_error()
    dims = mrvi_b200.MrviDims(n_input=10, n_sample=2)
    p = mrvi_b200.init_params(dims)
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(_capi.MrviError) as e:
            _capi.Handle(dims, p)
        assert e.value.status in (-4, -3)
    bad = mrvi_b200.MrviDims(n_input=10, n_sample=2, n_latent_sample=64)
    with pytest.raises(_capi.MrviError) as e:
        _capi.Handle(bad, mrvi_b200.init_params(bad))
    assert e.value.status in (-5, -4)


def test_model_fails_loudly_without_library(monkeypatch, tmp_path):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "library_path", lambda: str(tmp_path / "missing.so"))
    ad = mrvi_b200.synthetic_iid(n_obs=20, n_vars=10, n_sample=2)
    dims = mrvi_b200.MrviDims(n_input=10, n_sample=2)
    with pytest.raises(_capi.MrviError, match="no CPU fallback"):
        mrvi_b200.MrVI(ad, mrvi_b200.init_params(dims), sample_key="sample")


def test_product_never_imports_oracle():
    """The shipped package must not route through the oracle (parity contract)."""
    pkg = os.path.join(ROOT, "mrvi_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_host_narrowing_of_count_chunks_is_lossless(built_library):
    """`mrvi_run_host` sends integer count chunks as uint8 / uint16 (host_pack.cpp): exact round trip, and every
    non-narrowable input is refused (-> sent as float32)."""
    from mrvi_b200 import _capi

    rng = np.random.default_rng(0)
    for G in (7, 32, 100, 257):
        X = rng.integers(0, 256, size=(37, G)).astype(np.float32)
        mode, out = _capi.narrow_counts(X, n_threads=3)
        assert mode == 1 and out.dtype == np.uint8 and np.array_equal(out.astype(np.float32), X)
        X[5, G // 2] = 256.0
        mode, out = _capi.narrow_counts(X)
        assert mode == 2 and out.dtype == np.uint16 and np.array_equal(out.astype(np.float32), X)
        X[6, 0] = 65535.0
        assert _capi.narrow_counts(X)[0] == 2
        for bad in (65536.0, -1.0, 0.5, np.nan, np.inf, 3e9, -0.25, 1e-30):
            Y = X.copy()
            Y[3, G - 1] = bad
            assert _capi.narrow_counts(Y)[0] == 0, bad
            Y = X.copy()
            Y[0, 0] = bad
            assert _capi.narrow_counts(Y, n_threads=1)[0] == 0, bad
    assert _capi.narrow_counts(np.zeros((0, 5), np.float32))[0] == 1


def test_group_and_sample_coding_is_vectorised_and_matches_the_loop():
    """`MrVI._group_codes` / `_sample_codes` (host side of reference `_model.py:470-504`): the vectorised coding must give
    what the per-cell dictionary look-up of round 1 gave -- value_counts() order, -1 for labels outside `adata`'s
    categories and for NaN, labels taken positionally from self.adata -- for categorical, string and integer columns."""
    import pandas as pd

    from mrvi_b200._model import MrVI

    rng = np.random.default_rng(0)
    n = 5000
    labels = rng.choice(["b", "a", "c", "d"], size=n, p=[0.5, 0.3, 0.15, 0.05])
    obs = pd.DataFrame({"cat": pd.Categorical(labels, categories=["a", "b", "c", "d", "unused"]), "str": labels,
                        "int": rng.integers(0, 7, size=n), "sample": rng.integers(0, 5, size=n)})
    obs.loc[obs.index[::97], "str"] = None
    ad = mrvi_b200.SimpleAnnData(np.zeros((n, 3), np.float32), obs)
    m = MrVI.__new__(MrVI)   # host helpers only: no library handle
    m.adata, m.sample_key, m.sample_order = ad, "sample", np.arange(5)
    for key in ("cat", "str", "int"):
        for idx in (np.arange(n), rng.permutation(n)[:777]):
            codes, cats, counts = m._group_codes(ad, key, idx)
            vc = ad.obs[key].value_counts()
            assert cats == list(vc.index) and np.array_equal(counts, vc.to_numpy())
            lookup = {c: i for i, c in enumerate(cats)}
            want = np.array([lookup.get(l, -1) for l in ad.obs[key].to_numpy()[idx].tolist()], dtype=np.int32)
            assert codes.dtype == np.int32 and np.array_equal(codes, want), key
    assert np.array_equal(m._sample_codes(ad), ad.obs["sample"].to_numpy().astype(np.int32))
    m.sample_order = np.array(["s3", "s1", "s4", "s0", "s2"])
    ad.obs["sample"] = np.array([f"s{i}" for i in obs["sample"]])
    want = pd.Categorical(ad.obs["sample"], categories=list(m.sample_order)).codes.astype(np.int32)
    assert np.array_equal(m._sample_codes(ad), want)
    ad.obs["sample"] = pd.Categorical(ad.obs["sample"])
    assert np.array_equal(m._sample_codes(ad), want)
    ad.obs.loc[ad.obs.index[3], "sample"] = np.nan
    with pytest.raises(ValueError):
        m._sample_codes(ad)


def test_host_narrowing_of_integer_matrices(built_library):
    """`narrow_counts_int` through mrvi_run_host_counts' argument checks needs a GPU; the dtype table does not."""
    assert _capi._X_DTYPES[np.dtype(np.int32)] == 3 and _capi._X_DTYPES[np.dtype(np.int64)] == 4
    header = open(os.path.join(ROOT, "include", "mrvi_b200.h")).read()
    assert "MRVI_X_INT32 = 3" in header and "MRVI_X_INT64 = 4" in header


def test_de_design_matrix_algebra_matches_oracle():
    """Host side of `differential_expression`: `process_design_matrix` (reference `_model.py:1147-1160`) restated with
    eigh / hermitian pinv per distinct admissibility pattern vs the oracle's literal per-cell sqrtm / pinv."""
    from mrvi_b200._model import MrVI
    from oracle import mrvi_oracle as orc

    rng = np.random.default_rng(0)
    Xmat = rng.random((9, 4))
    adm = rng.random((40, 9)) < 0.8
    adm[0] = True
    for lambd in (0.0, 0.3):
        A, P = MrVI._process_design_matrix(Xmat, adm, lambd)
        A2, P2 = orc.de_process_design_matrix(Xmat, adm, lambd)
        assert A.shape == (40, 4, 9) and np.abs(A - A2).max() < 1e-10
        assert np.abs(P - P2).max() < 1e-6   # scipy's Schur sqrtm itself is ~1e-8 on the ill-conditioned patterns
    A, P = MrVI._process_design_matrix(Xmat, np.ones((40, 9), dtype=bool), 0.0)
    A2, P2 = orc.de_process_design_matrix(Xmat, np.ones((1, 9), dtype=bool), 0.0)
    assert A.shape == (4, 9) and np.abs(A - A2[0]).max() < 1e-10 and np.abs(P - P2[0]).max() < 1e-6
    # least squares: Amat @ Xmat = I when every sample is admissible and X has full column rank
    assert np.abs(A @ Xmat - np.eye(4)).max() < 1e-9


def test_fdr_bh_known_answers():
    from mrvi_b200._utils import fdr_bh
    from oracle import mrvi_oracle as orc

    p = np.array([0.01, 0.04, 0.03, 0.005])
    assert np.allclose(fdr_bh(p), [0.02, 0.04, 0.04, 0.02])          # p_(i) m / i, monotone from the top
    assert np.allclose(fdr_bh(np.array([[0.5, 0.9], [1.0, 0.2]])), [[1.0, 1.0], [1.0, 0.8]])
    q = np.random.default_rng(1).random((7, 5)) ** 4
    assert np.abs(fdr_bh(q) - orc.fdr_bh(q)).max() < 1e-15 and fdr_bh(q).shape == q.shape


def test_de_construct_design_matrix_layout():
    """`_construct_design_matrix` (reference `_model.py:1422-1529`): one-hot with the first level dropped, names
    `{key}_{level}`, 0-1 normalisation, offset columns in front, sample subset."""
    import pandas as pd

    from mrvi_b200._model import MrVI

    n = 24
    sid = np.arange(n) % 6
    obs = pd.DataFrame({"sample": sid, "site": (sid // 3), "grp": pd.Categorical(np.array(["a", "b", "c"])[sid % 3]),
                        "age": (10.0 + 5.0 * sid)}, index=[f"c{i}" for i in range(n)])
    m = MrVI.__new__(MrVI)
    m.adata = mrvi_b200.SimpleAnnData(np.zeros((n, 3), np.float32), obs)
    m.sample_key, m.sample_order, m.batch_key = "sample", np.arange(6), "site"
    X, names, lfc, off = m._construct_design_matrix(["grp", "age"], np.ones(6, bool), True, False, False)
    assert list(names) == ["grp_b", "grp_c", "age"] and X.shape == (6, 3) and X.dtype == np.float32 and off is None
    assert np.allclose(X[:, 0], (sid[:6] % 3 == 1)) and np.allclose(X[:, 2], (np.arange(6) * 5.0) / (1e-6 + 25.0))
    assert not lfc.any()
    X, names, lfc, off = m._construct_design_matrix(["age"], np.array([1, 1, 0, 1, 1, 1], bool), False, True, True)
    assert list(names) == ["offset_batch_0", "offset_batch_1", "age"] and list(off) == [0, 1] and list(lfc) == [False, False, True]
    assert np.allclose(X[:, 2], [10, 15, 25, 30, 35]) and np.allclose(X[:, 0], [1, 1, 0, 0, 0])
