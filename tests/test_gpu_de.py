"""GPU parity of the differential-expression eps stage (reference `_model.py:1160-1218, 1283, 1363`)
through `mrvi_run_de_host`, against the oracle ON THE SAME DRAWS and against the fixtures produced by executing the
reference's own statements (`oracle/gen_golden.reference_de`, `tests/golden/de_*.npz`).

Tolerances: the statistics are computed from eps STANDARDISED over the samples, (eps - mean_s) / (1e-6 + std_s): an error
of the chain relative to |eps| is amplified by |eps| / std_s (the counterfactual shifts of a cell share a large common
part).  The bars are the mean path's (1e-5 fp32 mode, 1e-3 tensor-core mode) times that conditioning, measured per case
on the oracle's eps; the raw eps itself is held to the plain bar.
"""
import glob
import os

import numpy as np
import pytest

import mrvi_b200
from mrvi_b200 import _capi
from oracle import mrvi_oracle as orc
from oracle import philox

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "tc": 1e-3}
PREC = {"fp32": _capi.PRECISION_FP32, "tc": _capi.PRECISION_TC}


def _rel(got, ref):
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / np.abs(ref).max())


def _conditioning(eps):
    """max over (draw, cell, latent dim) of max_s |eps| / std_s eps: how much a relative error of eps grows in the
    standardised values."""
    sd = eps.std(axis=2)
    return float((np.abs(eps).max(axis=2) / (1e-6 + sd)).max())


CASES = [
    # name, dims, n_cells, mc, n_covariates, sample subset, per-cell admissibility, lambd
    ("S6", dict(n_input=80, n_sample=6), 40, 4, 2, None, False, 0.0),
    ("S9_subset_adm", dict(n_input=50, n_sample=9, n_latent=12, n_latent_u=6), 30, 3, 3, (0, 2, 3, 5, 6, 8), True, 0.1),
    ("S32_K9", dict(n_input=120, n_sample=32), 25, 5, 9, None, True, 0.0),
    ("S128", dict(n_input=64, n_sample=128), 10, 3, 4, None, False, 0.0),
    ("S130_L40", dict(n_input=48, n_sample=130, n_latent=40, n_latent_u=8), 6, 2, 2, tuple(range(0, 130, 2)), False, 0.0),
]


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name,dk,n,mc,K,subset,per_cell,lambd", CASES, ids=[c[0] for c in CASES])
def test_de_stage_matches_oracle(built_library, name, dk, n, mc, K, subset, per_cell, lambd, precision):
    dims = mrvi_b200.MrviDims(**dk)
    flat = mrvi_b200.init_params(dims, seed=31)
    flat["qu/NormalDistOutputNN_0/Dense_1/bias"] = flat["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    X = mrvi_b200.synthetic_counts(n, dims.n_input, seed=32)
    rng = np.random.default_rng(33)
    sid = rng.integers(0, dims.n_sample, n).astype(np.int32)
    nu = rng.standard_normal((dims.n_sample, mc, n, dims.n_latent_q)).astype(np.float32)
    kept = np.arange(dims.n_sample) if subset is None else np.asarray(subset)
    Xmat = rng.random((len(kept), K))
    adm = rng.random((n, len(kept))) < 0.85 if per_cell else np.ones((n, len(kept)), dtype=bool)
    adm[0] = True
    A, P = orc.de_process_design_matrix(Xmat, adm, lambd)
    h = _capi.Handle(dims, flat, device=0, precision=PREC[precision])
    res = h.run_de_host(X, sid, mc, A if per_cell else A[0], P if per_cell else P[0], kept_samples=None if subset is None else kept,
                        u_noise=nu, want_eps=True, chunk_cells=7 if n > 20 else 0)
    eps_ref = orc.sampled_counterfactual_eps(flat, X, sid, nu, samples=kept)         # (mc, n, S_kept, L)
    beta_ref, ts_ref, _ = orc.de_statistics(eps_ref, A, P, adm.sum(axis=1))
    eps_got = res["eps"].transpose(1, 0, 2, 3)[:, :, kept]
    assert _rel(eps_got, eps_ref) < TOL[precision]
    cond = _conditioning(eps_ref)
    assert res["beta"].shape == beta_ref.shape and res["effect_size"].shape == ts_ref.shape
    assert _rel(res["beta"], beta_ref) < TOL[precision] * cond, (cond, _rel(res["beta"], beta_ref))
    assert _rel(res["effect_size"], ts_ref) < 2 * TOL[precision] * cond, (cond, _rel(res["effect_size"], ts_ref))
    # the statistics kernel alone: the oracle's arithmetic on the DEVICE's eps must agree to float32 rounding
    beta_d, ts_d, _ = orc.de_statistics(eps_got.astype(np.float64), A, P, adm.sum(axis=1))
    assert _rel(res["beta"], beta_d) < 2e-5 and _rel(res["effect_size"], ts_d) < 2e-5
    h.close()


DE_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "de_*.npz")))


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("path", DE_GOLDEN, ids=[os.path.basename(p)[:-4] for p in DE_GOLDEN])
def test_cuda_de_matches_reference_source_goldens(built_library, path, precision):
    g = np.load(path)
    flat = {k[len("param|"):].replace("|", "/"): g[k] for k in g.files if k.startswith("param|")}
    dims = mrvi_b200.infer_dims(flat)
    h = _capi.Handle(dims, flat, device=0, precision=PREC[precision])
    kept = g["samples"]
    mc = g["noise_u"].shape[1]
    res = h.run_de_host(g["X"], g["sid"], mc, g["Amat"], g["prefactor"], kept_samples=kept if len(kept) < dims.n_sample else None,
                        u_noise=g["noise_u"], want_eps=True)
    eps_got = res["eps"].transpose(1, 0, 2, 3)[:, :, kept]
    cond = _conditioning(g["eps"])
    assert _rel(eps_got, g["eps"]) < TOL[precision]
    assert _rel(res["beta"], g["beta"]) < TOL[precision] * cond
    assert _rel(res["effect_size"], g["effect_size"]) < 2 * TOL[precision] * cond
    h.close()


def test_de_device_noise_is_chunk_and_shard_invariant(built_library):
    dims = mrvi_b200.MrviDims(n_input=60, n_sample=7)
    flat = mrvi_b200.init_params(dims, seed=5)
    X = mrvi_b200.synthetic_counts(33, 60, seed=6)
    sid = (np.arange(33) % 7).astype(np.int32)
    rng = np.random.default_rng(7)
    A, P = orc.de_process_design_matrix(rng.random((7, 2)), np.ones((33, 7), dtype=bool))
    h = _capi.Handle(dims, flat, device=0)
    a = h.run_de_host(X, sid, 5, A[0], P[0], seed=11, cell_offset=100)
    b = h.run_de_host(X, sid, 5, A[0], P[0], seed=11, cell_offset=100, chunk_cells=4)
    assert np.array_equal(a["beta"], b["beta"]) and np.array_equal(a["effect_size"], b["effect_size"])
    c = h.run_de_host(X[10:], sid[10:], 5, A[0], P[0], seed=11, cell_offset=110)
    assert np.array_equal(a["beta"][10:], c["beta"])
    d = h.run_de_host(X, sid, 5, A[0], P[0], seed=12, cell_offset=100)
    assert not np.allclose(a["beta"], d["beta"])
    h.close()


def test_de_errors(built_library):
    dims = mrvi_b200.MrviDims(n_input=30, n_sample=5)
    flat = mrvi_b200.init_params(dims, seed=0)
    h = _capi.Handle(dims, flat, device=0)
    X = mrvi_b200.synthetic_counts(8, 30, seed=0)
    sid = np.zeros(8, np.int32)
    A, P = np.zeros((2, 5), np.float32), np.eye(2, dtype=np.float32)
    with pytest.raises(_capi.MrviError, match="kept_samples"):
        h.run_de_host(X, sid, 2, A[:, :2], P, kept_samples=np.array([0, 7]))
    with pytest.raises(ValueError, match="Amat must have shape"):
        h.run_de_host(X, sid, 2, A[:, :3], P)
    with pytest.raises(_capi.MrviError, match="mc_samples"):
        h.run_de_host(X, sid, 0, A, P)
    h.close()
    dims = mrvi_b200.MrviDims(n_input=30, n_sample=5, use_map=False)
    h = _capi.Handle(dims, mrvi_b200.init_params(dims, seed=0), device=0, precision=_capi.PRECISION_TC)
    with pytest.raises(_capi.MrviError, match="use_map=False"):
        h.run_de_host(X, sid, 2, A, P)
    h.close()


def test_model_api_differential_expression(built_library):
    """`MrVI.differential_expression` through the drop-in class: the reference's output layout (`_model.py:1365-1420`:
    beta (cell_name, covariate, latent_dim), effect_size / pvalue / padj (cell_name, covariate)), covariate names of
    `_construct_design_matrix` (`:1422-1529`), values against the oracle on the draws the device generator makes."""
    import pandas as pd

    ad = mrvi_b200.synthetic_iid(n_obs=70, n_vars=50, n_sample=6, seed=4)
    dims = mrvi_b200.MrviDims(n_input=50, n_sample=6)
    params = mrvi_b200.init_params(dims, seed=8)
    params["qu/NormalDistOutputNN_0/Dense_1/bias"] = params["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    model = mrvi_b200.MrVI(ad, params, sample_key="sample_str", seed=17)
    sid = model._sample_codes(ad)
    # sample-level covariates: one categorical with 3 levels, one continuous
    ad.obs["treatment"] = pd.Categorical(np.array(["ctrl", "drugA", "drugB"])[sid % 3])
    ad.obs["dose"] = (sid * 0.5 + 1.0).astype(np.float32)
    mc = 6
    ds = model.differential_expression(sample_cov_keys=["treatment", "dose"], mc_samples=mc, lambd=0.01)
    assert ds["beta"].dims == ("cell_name", "covariate", "latent_dim") and ds["beta"].shape == (70, 3, dims.n_latent)
    assert list(np.asarray(ds["beta"].coords["covariate"])) == ["treatment_drugA", "treatment_drugB", "dose"]
    for k in ("effect_size", "pvalue", "padj"):
        assert ds[k].dims == ("cell_name", "covariate") and ds[k].shape == (70, 3)
    # oracle on the same draws
    info = ad.obs.assign(code=sid).drop_duplicates("code").set_index("code").sort_index()
    Xm = np.stack([(info["treatment"] == "drugA").to_numpy(float), (info["treatment"] == "drugB").to_numpy(float),
                   info["dose"].to_numpy(float)], axis=1).astype(np.float32)
    Xm = (Xm - Xm.min(0)) / (1e-6 + Xm.max(0) - Xm.min(0))
    adm = np.ones((70, 6), dtype=bool)
    A, P = orc.de_process_design_matrix(Xm, adm, 0.01)
    nu = philox.u_noise(70, 6, mc, dims.n_latent_q, 17)
    eps = orc.sampled_counterfactual_eps(params, ad.X, sid, nu)
    beta, ts, pv = orc.de_statistics(eps, A, P, adm.sum(1))
    cond = _conditioning(eps)
    assert _rel(np.asarray(ds["beta"].values), beta) < 3e-4 * cond        # libm vs numpy ulps in the draws
    assert _rel(np.asarray(ds["effect_size"].values), ts) < 6e-4 * cond
    assert np.abs(np.asarray(ds["pvalue"].values) - pv).max() < 1e-2
    got_padj = np.asarray(ds["padj"].values)
    assert np.abs(got_padj - orc.fdr_bh(np.asarray(ds["pvalue"].values, dtype=np.float64))).max() < 1e-6
    # sample subset + explicit admissibility -> n_samples variable
    sub = list(model.sample_order[[0, 1, 3, 4, 5]])
    adm2 = np.ones((70, 5), dtype=bool)
    adm2[::3, 2] = False
    ds2 = model.differential_expression(sample_cov_keys=["dose"], sample_subset=sub, mc_samples=3, admissible_samples=adm2)
    assert ds2["beta"].shape == (70, 1, dims.n_latent) and np.array_equal(np.asarray(ds2["n_samples"].values), adm2.sum(1))
    with pytest.raises(NotImplementedError):
        model.differential_expression(sample_cov_keys=["dose"], store_lfc=True)
    with pytest.raises(NotImplementedError):
        model.differential_expression(sample_cov_keys=["dose"], filter_inadmissible_samples=True)
    model.close()
