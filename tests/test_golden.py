"""Golden vectors produced by executing the reference's own `_components.py` / `_module.py` on the numpy
flax shim (`oracle/gen_golden.py`).  They pin the oracle's composition of the network against the
reference source; the flax primitives themselves remain restated (see the oracle header)."""
import glob
import os

import numpy as np
import pytest

import mrvi_b200
from oracle import gen_golden as gg
from oracle import mrvi_oracle as orc

_ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLDEN = [p for p in _ALL if not os.path.basename(p).startswith(("sampled_", "de_"))]
SAMPLED = [p for p in _ALL if os.path.basename(p).startswith("sampled_")]
DE = [p for p in _ALL if os.path.basename(p).startswith("de_")]


def _load(path):
    z = np.load(path)
    flat = {k[len("param|"):].replace("|", "/"): z[k] for k in z.files if k.startswith("param|")}
    return flat, z["X"], z["sid"], z["z"], z["d"]


def test_golden_files_exist():
    assert {os.path.basename(p)[:-4] for p in GOLDEN} == set(gg.CASES)
    assert {os.path.basename(p)[:-4] for p in SAMPLED} == set(gg.SAMPLED_CASES)
    assert {os.path.basename(p)[:-4] for p in DE} == set(gg.DE_CASES)


def _load_sampled(path):
    z = np.load(path)
    flat = {k[len("param|"):].replace("|", "/"): z[k] for k in z.files if k.startswith("param|")}
    return flat, z


@pytest.mark.parametrize("path", SAMPLED, ids=[os.path.basename(p)[:-4] for p in SAMPLED])
def test_oracle_sampled_paths_match_reference_source_goldens(path):
    """use_mean=False / normalize_distances=True: the reference's own inference(use_mean=False,
    mc_samples) executed on the shim with explicit draws vs the oracle restatement."""
    flat, g = _load_sampled(path)
    z = orc.sampled_counterfactual_z(flat, g["X"], g["sid"], g["noise_u"])
    assert z.shape == g["z"].shape and np.abs(z - g["z"]).max() < 1e-12
    mu, var = orc.local_baseline_dists(flat, g["X"], g["sid"], g["noise_base"])
    assert np.abs(mu - g["base_mean"]).max() < 1e-12 and np.abs(var - g["base_var"]).max() < 1e-12


@pytest.mark.skipif(not gg.reference_available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("name", list(gg.SAMPLED_CASES))
def test_sampled_goldens_regenerate_from_reference_source(name):
    dims, vk, flat, X, sid, noise_u, noise_base = gg.make_sampled_case(name)
    z, mu, var = gg.reference_sampled(flat, dims, vk, X, sid, noise_u, noise_base)
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    assert np.abs(z - stored["z"]).max() < 1e-13 and np.abs(mu - stored["base_mean"]).max() < 1e-13


@pytest.mark.parametrize("path", DE, ids=[os.path.basename(p)[:-4] for p in DE])
def test_oracle_de_stage_matches_reference_source_goldens(path):
    """Differential-expression eps stage: eps from the reference's inference on the shim, design-matrix
    processing and statistics from the reference's own statements (`_model.py:1152-1158, 1196-1215`, executed from its
    source by gen_golden.reference_de) vs the oracle restatement."""
    flat, g = _load_sampled(path)
    eps = orc.sampled_counterfactual_eps(flat, g["X"], g["sid"], g["noise_u"], samples=g["samples"])
    assert eps.shape == g["eps"].shape and np.abs(eps - g["eps"]).max() < 1e-12
    A, P = orc.de_process_design_matrix(g["Xmat"], g["admissible"], float(g["lambd"]))
    assert np.abs(A - g["Amat"]).max() < 1e-12 and np.abs(P - g["prefactor"]).max() < 1e-12
    beta, ts, pv = orc.de_statistics(eps, A, P, g["admissible"].sum(axis=1))
    assert np.abs(beta - g["beta"]).max() < 1e-12 and np.abs(ts - g["effect_size"]).max() < 1e-10 * np.abs(ts).max()
    assert np.abs(pv - g["pvalue"]).max() < 1e-12


@pytest.mark.skipif(not gg.reference_available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("name", list(gg.DE_CASES))
def test_de_goldens_regenerate_from_reference_source(name):
    case = gg.make_de_case(name)
    r = gg.reference_de(case[2], case[0], case[1], *case[3:])
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    for k in ("eps", "Amat", "prefactor", "beta", "effect_size", "pvalue"):
        assert np.abs(r[k] - stored[k]).max() <= 1e-13 * max(1.0, np.abs(stored[k]).max()), k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_source_goldens(path):
    flat, X, sid, z_ref, d_ref = _load(path)
    z = orc.counterfactual_z(flat, X, sid, dtype=np.float64)
    assert np.abs(z - z_ref).max() < 1e-12
    assert np.abs(orc.pairwise_distances(z) - d_ref).max() < 1e-12
    # the reference's dtype (float32) stays within the fp32 parity bar of the goldens
    z32 = orc.counterfactual_z(flat, X, sid, dtype=np.float32)
    assert np.abs(orc.pairwise_distances(z32) - d_ref).max() / d_ref.max() < 1e-5


@pytest.mark.skipif(not gg.reference_available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("name", list(gg.CASES))
def test_goldens_regenerate_from_reference_source(name):
    """Re-execute the reference source on the shim and compare with the committed fixtures."""
    dims, vk, flat, X, sid = gg.make_case(name)
    z = gg.reference_z(flat, dims, vk, X, sid)
    stored = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    assert np.array_equal(stored["X"], X) and np.array_equal(stored["sid"], sid)
    assert np.abs(z - stored["z"]).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tc", 1e-3)])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference_source_goldens(built_library, path, precision, tol):
    from mrvi_b200 import _capi

    flat, X, sid, z_ref, d_ref = _load(path)
    dims = mrvi_b200.infer_dims(flat)
    h = _capi.Handle(dims, flat, device=0, precision=_capi.PRECISION_FP32 if precision == "fp32" else _capi.PRECISION_TC)
    res = h.run_host(X, sid, keep_cell=True, want_z=True)
    scale = d_ref.max(axis=(1, 2), keepdims=True)
    assert (np.abs(res["cell"] - d_ref) / scale).max() < tol
    assert np.abs(res["z"] - z_ref).max() / np.abs(z_ref).max() < tol
    h.close()
