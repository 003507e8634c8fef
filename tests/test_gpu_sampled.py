"""GPU parity of the sampled paths: get_local_sample_distances(use_mean=False) and
(normalize_distances=True) through `mrvi_run_sampled_host`, against the oracle ON THE SAME DRAWS.

The reference's JAX threefry streams cannot be reproduced (no jax here), so the draws are explicit inputs in the
parity tests; the device Philox generator is pinned separately against its numpy mirror (`oracle/philox.py`,
itself checked against the Random123 known-answer vectors in tests/test_oracle.py).
Tolerances as on the mean path: 1e-5 (fp32 mode) / 1e-3 (tensor-core mode), normwise.
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


def _setup(dk, n, mc, seed, precision):
    dims = mrvi_b200.MrviDims(**dk)
    flat = mrvi_b200.init_params(dims, seed=seed)
    flat["qu/NormalDistOutputNN_0/Dense_1/bias"] = flat["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    X = mrvi_b200.synthetic_counts(n, dims.n_input, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    sid = rng.integers(0, dims.n_sample, n).astype(np.int32)
    nu = rng.standard_normal((dims.n_sample, mc, n, dims.n_latent_q)).astype(np.float32)
    nb = rng.standard_normal((2 * mc, n, dims.n_latent_q)).astype(np.float32)
    h = _capi.Handle(dims, flat, device=0, precision=PREC[precision])
    return dims, flat, X, sid, nu, nb, h


def _rel(got, ref):
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / np.abs(ref).max())


CASES = [
    ("S4", dict(n_input=100, n_sample=4), 60, 3),
    ("S15_iso", dict(n_input=60, n_sample=15, n_latent=10, n_latent_u=10), 40, 2),
    ("S32", dict(n_input=200, n_sample=32), 50, 4),
    ("S128", dict(n_input=64, n_sample=128), 12, 2),
    ("S130", dict(n_input=48, n_sample=130, n_latent=12, n_latent_u=6), 9, 2),   # large-S route: chain + k_dist_gram_tc, per-row z_base
]


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name,dk,n,mc", CASES, ids=[c[0] for c in CASES])
def test_sampled_distances_match_oracle(built_library, name, dk, n, mc, precision):
    dims, flat, X, sid, nu, nb, h = _setup(dk, n, mc, 21, precision)
    rng = np.random.default_rng(5)
    codes = rng.integers(-1, 3, size=n).astype(np.int32)
    res = h.run_sampled_host(X, sid, mc, u_noise=nu, group_codes=[codes], n_groups=[3], keep_cell=True, want_z=True)
    zref = orc.sampled_counterfactual_z(flat, X, sid, nu)
    dref = orc.pairwise_distances_mc(zref)
    assert res["z"].shape == zref.shape and _rel(res["z"], zref) < TOL[precision]
    assert res["cell"].shape == (n, mc, dims.n_sample, dims.n_sample)
    assert _rel(res["cell"], dref) < TOL[precision]
    assert np.array_equal(res["group_counts"][0], np.bincount(codes[codes >= 0], minlength=3))
    ref_sum = np.stack([dref[codes == g].sum(0) for g in range(3)])  # (G, mc, S, S): the mc axis is kept
    assert res["group_sums"][0].shape == ref_sum.shape and _rel(res["group_sums"][0], ref_sum) < TOL[precision]
    h.close()


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name,dk,n,mc", CASES[:3], ids=[c[0] for c in CASES[:3]])
def test_normalized_distances_match_oracle(built_library, name, dk, n, mc, precision):
    mc = max(mc, 3)
    dims, flat, X, sid, nu, nb, h = _setup(dk, n, mc, 33, precision)
    codes = (np.arange(n) % 2).astype(np.int32)
    res = h.run_sampled_host(X, sid, mc, u_noise=nu, base_noise=nb, group_codes=[codes], n_groups=[2], normalize=True,
                             keep_cell=True)
    ref = orc.sampled_local_sample_distances(flat, X, sid, nu, nb, normalize=True)
    mu, var = orc.local_baseline_dists(flat, X, sid, nb)
    # the z-score divides by sqrt(var) of only mc baseline distances: a relative error eps of the distances becomes
    # eps * (mean / std) of the score; scale the bar by that conditioning, per cell
    cond = 1.0 + (mu / np.sqrt(var)).max()
    assert res["cell"].shape == (n, dims.n_sample, dims.n_sample)
    assert _rel(res["cell"], ref) < TOL[precision] * cond
    ref_sum = np.stack([ref[codes == g].sum(0) for g in range(2)])
    assert _rel(res["group_sums"][0], ref_sum) < TOL[precision] * cond
    h.close()


def test_device_philox_matches_numpy_mirror_and_is_chunk_invariant(built_library):
    dims, flat, X, sid, _, _, h = _setup(dict(n_input=80, n_sample=6), 37, 3, 44, "fp32")
    seed, off = 1234567, 1000
    a = h.run_sampled_host(X, sid, 3, seed=seed, cell_offset=off, normalize=True, keep_cell=True)
    b = h.run_sampled_host(X, sid, 3, seed=seed, cell_offset=off, normalize=True, keep_cell=True, chunk_cells=5)
    assert np.array_equal(a["cell"], b["cell"])  # draws depend on (seed, global cell) only
    nu = philox.u_noise(37, 6, 3, dims.n_latent_q, seed, off)
    nb = philox.base_noise(37, 3, dims.n_latent_q, seed, off)
    c = h.run_sampled_host(X, sid, 3, u_noise=nu, base_noise=nb, normalize=True, keep_cell=True)
    assert np.abs(a["cell"] - c["cell"]).max() < 2e-4 * np.abs(c["cell"]).max()  # libm vs numpy log / sincos ulps
    # a shard that passes its offset reproduces its slice of the whole
    d = h.run_sampled_host(X[20:], sid[20:], 3, seed=seed, cell_offset=off + 20, normalize=True, keep_cell=True)
    assert np.array_equal(a["cell"][20:], d["cell"])
    e = h.run_sampled_host(X, sid, 3, seed=seed + 1, cell_offset=off, normalize=True, keep_cell=True)
    assert not np.allclose(a["cell"], e["cell"])
    h.close()


def test_zero_noise_reduces_to_mean_path(built_library):
    dims, flat, X, sid, nu, nb, h = _setup(dict(n_input=50, n_sample=8), 30, 2, 9, "fp32")
    res = h.run_sampled_host(X, sid, 2, u_noise=np.zeros_like(nu), keep_cell=True)
    mean = h.run_host(X, sid, keep_cell=True)
    assert np.abs(res["cell"][:, 0] - mean["cell"]).max() < 1e-5 * mean["cell"].max()
    assert np.array_equal(res["cell"][:, 0], res["cell"][:, 1])
    h.close()


SAMPLED_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sampled_*.npz")))


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("path", SAMPLED_GOLDEN, ids=[os.path.basename(p)[:-4] for p in SAMPLED_GOLDEN])
def test_cuda_sampled_matches_reference_source_goldens(built_library, path, precision):
    g = np.load(path)
    flat = {k[len("param|"):].replace("|", "/"): g[k] for k in g.files if k.startswith("param|")}
    dims = mrvi_b200.infer_dims(flat)
    h = _capi.Handle(dims, flat, device=0, precision=PREC[precision])
    mc = g["noise_u"].shape[1]
    res = h.run_sampled_host(g["X"], g["sid"], mc, u_noise=g["noise_u"], base_noise=g["noise_base"], normalize=True,
                             keep_cell=True, want_z=True)
    assert _rel(res["z"], g["z"]) < TOL[precision]
    d = orc.pairwise_distances_mc(g["z"])
    mu, sd = g["base_mean"], np.sqrt(g["base_var"])
    ref = ((d - mu[:, None, None, None]) / sd[:, None, None, None]).mean(axis=1)
    assert _rel(res["cell"], ref) < TOL[precision] * (1.0 + (mu / sd).max())
    h.close()


def test_sampled_eps_use_map_false(built_library):
    """qz_kwargs={"use_map": False} on the sampled paths (reference `_module.py:326-329`): eps ~ Normal(loc, softplus(raw) +
    1e-3) is drawn too.  fp32 mode, explicit draws vs the oracle; device Philox draws vs the numpy mirror; DE eps stage."""
    dk = dict(n_input=40, n_sample=6, use_map=False)
    n, mc = 30, 3
    dims, flat, X, sid, nu, nb, h = _setup(dk, n, mc, 51, "fp32")
    rng = np.random.default_rng(8)
    L = dims.n_latent
    ne = rng.standard_normal((dims.n_sample, mc, n, L)).astype(np.float32)
    nbe = rng.standard_normal((2 * mc, n, L)).astype(np.float32)
    res = h.run_sampled_host(X, sid, mc, u_noise=nu, eps_noise=ne, keep_cell=True, want_z=True)
    zref = orc.sampled_counterfactual_z(flat, X, sid, nu, noise_eps=ne)
    assert _rel(res["z"], zref) < 1e-5 and _rel(res["cell"], orc.pairwise_distances_mc(zref)) < 1e-5
    # the noise term matters (the same call without it differs)
    assert _rel(res["z"], orc.sampled_counterfactual_z(flat, X, sid, nu, noise_eps=np.zeros_like(ne))) > 1e-3
    resn = h.run_sampled_host(X, sid, mc, u_noise=nu, base_noise=nb, eps_noise=ne, base_eps_noise=nbe, normalize=True, keep_cell=True)
    mu, var = orc.local_baseline_dists(flat, X, sid, nb, noise_eps=nbe)
    dref = orc.pairwise_distances_mc(zref)
    nref = ((dref - mu[:, None, None, None]) / np.sqrt(var)[:, None, None, None]).mean(axis=1)
    assert _rel(resn["cell"], nref) < 1e-5 * (1.0 + (mu / np.sqrt(var)).max())
    # device generator == numpy mirror (streams 2 / 3), chunk invariant
    a = h.run_sampled_host(X, sid, mc, seed=77, cell_offset=5, normalize=True, keep_cell=True)
    b = h.run_sampled_host(X, sid, mc, seed=77, cell_offset=5, normalize=True, keep_cell=True, chunk_cells=4)
    assert np.array_equal(a["cell"], b["cell"])
    c = h.run_sampled_host(X, sid, mc, normalize=True, keep_cell=True,
                           u_noise=philox.u_noise(n, 6, mc, dims.n_latent_q, 77, 5), base_noise=philox.base_noise(n, mc, dims.n_latent_q, 77, 5),
                           eps_noise=philox.eps_noise(n, 6, mc, L, 77, 5), base_eps_noise=philox.base_eps_noise(n, mc, L, 77, 5))
    assert np.abs(a["cell"] - c["cell"]).max() < 2e-4 * np.abs(c["cell"]).max()
    # DE eps stage with sampled eps
    A, P = orc.de_process_design_matrix(rng.random((6, 2)), np.ones((n, 6), dtype=bool))
    de = h.run_de_host(X, sid, mc, A[0], P[0], u_noise=nu, eps_noise=ne, want_eps=True)
    eref = orc.sampled_counterfactual_eps(flat, X, sid, nu, noise_eps=ne)
    assert _rel(de["eps"].transpose(1, 0, 2, 3), eref) < 1e-5
    h.close()
    # tensor-core mode refuses loudly
    ht = _capi.Handle(dims, flat, device=0, precision=_capi.PRECISION_TC)
    with pytest.raises(_capi.MrviError, match="MRVI_PRECISION_FP32"):
        ht.run_sampled_host(X, sid, mc)
    ht.close()


def test_sampled_path_errors(built_library):
    X = mrvi_b200.synthetic_counts(8, 30, seed=0)
    dims = mrvi_b200.MrviDims(n_input=30, n_sample=5)
    flat = mrvi_b200.init_params(dims, seed=0)
    h = _capi.Handle(dims, {k: v for k, v in flat.items() if "NormalDistOutputNN_0/Dense_1" not in k}, device=0)
    with pytest.raises(_capi.MrviError, match="scale head"):
        h.run_sampled_host(X, np.zeros(8, np.int32), 2)
    with pytest.raises(_capi.MrviError):
        h.run_sampled_host(X, np.zeros(8, np.int32), 2, norm="l1", normalize=True)
    h.close()


def test_model_api_sampled_and_normalized(built_library):
    """`get_local_sample_distances(use_mean=False)` / `(normalize_distances=True)` through the drop-in class:
    the reference's layouts (`_model.py:572-584, 447-450, 493-504`), values against the oracle on the draws the
    device generator makes (numpy Philox mirror)."""
    ad = mrvi_b200.synthetic_iid(n_obs=90, n_vars=60, n_sample=5, seed=2)
    dims = mrvi_b200.MrviDims(n_input=60, n_sample=5)
    params = mrvi_b200.init_params(dims, seed=3)
    params["qu/NormalDistOutputNN_0/Dense_1/bias"] = params["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    model = mrvi_b200.MrVI(ad, params, sample_key="sample_str", seed=99)
    sid = model._sample_codes(ad)
    mc = 4
    nu = philox.u_noise(90, 5, mc, dims.n_latent_q, 99)
    nb = philox.base_noise(90, mc, dims.n_latent_q, 99)
    ds = model.get_local_sample_distances(use_mean=False, groupby="labels", mc_samples=mc)
    assert ds["cell"].dims == ("cell_name", "mc_sample", "sample_x", "sample_y") and ds["cell"].shape == (90, mc, 5, 5)
    assert ds["labels"].dims == ("labels_name", "mc_sample", "sample_x", "sample_y")
    dref = orc.sampled_local_sample_distances(params, ad.X, sid, nu)
    assert _rel(np.asarray(ds["cell"].values), dref) < 3e-4
    vc = ad.obs["labels"].value_counts()
    lab = ad.obs["labels"].to_numpy()
    gref = np.stack([dref[lab == c].mean(0) for c in vc.index])
    assert list(np.asarray(ds["labels"].coords["labels_name"])) == list(vc.index)
    assert _rel(np.asarray(ds["labels"].values), gref) < 3e-4
    with pytest.warns(UserWarning, match="Ignoring ``use_mean``"):
        dn = model.get_local_sample_distances(normalize_distances=True, groupby="labels", mc_samples=mc)
    assert dn["cell"].dims == ("cell_name", "sample_x", "sample_y") and dn["labels"].shape == (len(vc), 5, 5)
    nref = orc.sampled_local_sample_distances(params, ad.X, sid, nu, nb, normalize=True)
    mu, var = orc.local_baseline_dists(params, ad.X, sid, nb)
    assert _rel(np.asarray(dn["cell"].values), nref) < 3e-4 * (1 + (mu / np.sqrt(var)).max())
    with pytest.raises(ValueError, match="Norm must be 'l2'"):
        model.get_local_sample_distances(normalize_distances=True, norm="l1")
    zr = model.get_local_sample_representation(use_mean=False)
    assert zr.dims == ("cell_name", "mc_sample", "sample", "latent_dim") and zr.shape == (90, 10, 5, dims.n_latent)
    model.close()


def test_model_api_sampled_user_fn_and_latent_draws(built_library):
    """Closes the sampled-path holes of the host API: user reduction functions on the sampled inputs (the reference's batch
    loop, `_model.py:386-504`), grouped sampled representations, and `get_latent_representation(use_mean=False)`
    (`_model.py:221-271`).  Values against the oracle on the draws the device generator makes."""
    from mrvi_b200 import MrVIReduction

    ad = mrvi_b200.synthetic_iid(n_obs=50, n_vars=40, n_sample=4, seed=5)
    dims = mrvi_b200.MrviDims(n_input=40, n_sample=4)
    params = mrvi_b200.init_params(dims, seed=6)
    params["qu/NormalDistOutputNN_0/Dense_1/bias"] = params["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    model = mrvi_b200.MrVI(ad, params, sample_key="sample_str", seed=123)
    sid = model._sample_codes(ad)
    mc = 3
    nu = philox.u_noise(50, 4, mc, dims.n_latent_q, 123)
    zref = orc.sampled_counterfactual_z(params, ad.X, sid, nu)          # (n, mc, S, L)
    dref = orc.pairwise_distances_mc(zref)
    reds = [
        MrVIReduction(name="dmax", input="sampled_distances", fn=lambda x: x.mean("mc_sample"), group_by=None),
        MrVIReduction(name="zgrp", input="sampled_representations", fn=lambda x: x.mean("mc_sample"), group_by="labels"),
    ]
    out = model.compute_local_statistics(reds, batch_size=16, mc_samples=mc)   # ragged batches: 16, 16, 16, 2
    assert out["dmax"].dims == ("cell_name", "sample_x", "sample_y") and out["dmax"].shape == (50, 4, 4)
    assert _rel(np.asarray(out["dmax"].values), dref.mean(axis=1)) < 3e-4
    vc = ad.obs["labels"].value_counts()
    lab = ad.obs["labels"].to_numpy()
    gref = np.stack([zref[lab == c].mean(axis=1).mean(axis=0) for c in vc.index])
    assert out["zgrp"].dims[0] == "labels_name" and list(np.asarray(out["zgrp"].coords["labels_name"])) == list(vc.index)
    assert _rel(np.asarray(out["zgrp"].values), gref) < 3e-4
    # one draw of u per cell / z of that draw for the cell's own sample
    nu1 = philox.u_noise(50, 4, 1, dims.n_latent_q, 123)
    p64 = orc.cast_params(params, np.float64)
    mean, scale = orc.encoder_xu(p64, np.asarray(ad.X, np.float64), sid)
    u_ref = mean + scale * nu1[sid, 0, np.arange(50)]
    u = model.get_latent_representation(use_mean=False)
    assert u.shape == (50, dims.n_latent_q) and _rel(u, u_ref) < 3e-4
    z1 = orc.sampled_counterfactual_z(params, ad.X, sid, nu1)
    z = model.get_latent_representation(use_mean=False, give_z=True)
    assert z.shape == (50, dims.n_latent) and _rel(z, z1[np.arange(50), 0, sid]) < 3e-4
    model.close()
