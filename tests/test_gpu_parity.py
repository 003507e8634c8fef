"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle.

Tolerances (normwise, per test): fp32-FFMA mode 1e-5 of the largest distance (north_star's fp32 bar);
tensor-core mode 1e-3.  Group assignment, counts and cell order must be exact.
"""
import numpy as np
import pytest

import mrvi_b200
from mrvi_b200 import _capi
from oracle import mrvi_oracle as orc

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "tc": 1e-3}


def _model(dims, seed=0, precision="fp32", **kw):
    params = mrvi_b200.init_params(dims, seed=seed, **kw)
    return params, _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_FP32 if precision == "fp32" else _capi.PRECISION_TC)


def _inputs(n, G, S, seed=0):
    X = mrvi_b200.synthetic_counts(n, G, seed=seed)
    sid = np.random.default_rng(seed + 7).integers(0, S, size=n).astype(np.int32)
    return X, sid


def _relerr(got, ref):
    return float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max())


CASES = [
    # (name, dims kwargs, n_cells)
    ("c1_400x100_S4", dict(n_input=100, n_sample=4), 400),
    ("tests_S15", dict(n_input=100, n_sample=15), 300),
    ("isomorphic_L10", dict(n_input=100, n_sample=15, n_latent=10, n_latent_u=10), 200),
    ("latent_u5", dict(n_input=100, n_sample=6, n_latent=10, n_latent_u=5), 200),
    ("use_map_false", dict(n_input=60, n_sample=5, use_map=False), 130),
    ("S32_G2000", dict(n_input=2000, n_sample=32), 192),
    ("S128", dict(n_input=333, n_sample=128), 70),
    ("S256_L50", dict(n_input=128, n_sample=256, n_latent=50), 20),
    ("H64_skip", dict(n_input=50, n_sample=7, encoder_n_hidden=64), 100),
    ("qz_layers2", dict(n_input=50, n_sample=9, qz_n_layers=2), 100),
    ("odd_G", dict(n_input=101, n_sample=3), 65),
    ("S31_L64_bigtile", dict(n_input=64, n_sample=31, n_latent=64, n_latent_u=5), 69),   # k_dist_small above 48 KiB of smem (found by tools/fuzz_parity.py)
]


@pytest.mark.parametrize("name,dk,n", CASES, ids=[c[0] for c in CASES])
def test_fp32_matches_oracle(built_library, name, dk, n):
    dims = mrvi_b200.MrviDims(**dk)
    params, h = _model(dims, seed=3)
    X, sid = _inputs(n, dims.n_input, dims.n_sample, seed=5)
    rng = np.random.default_rng(11)
    codes = [rng.integers(0, 3, size=n).astype(np.int32), rng.integers(0, 2, size=n).astype(np.int32)]
    res = h.run_host(X, sid, codes, [3, 2], keep_cell=True, want_z=True, want_u=True)
    zref = orc.counterfactual_z(params, X, sid, dtype=np.float64)
    uref = orc.encoder_xu_mean(orc.cast_params(params, np.float64), X.astype(np.float64), sid)
    dref = orc.pairwise_distances(zref)
    assert _relerr(res["u"], uref) < 1e-5
    assert _relerr(res["z"], zref) < 1e-5
    assert _relerr(res["cell"], dref) < TOL["fp32"]
    # exact structural properties of the direct-difference kernel
    assert np.all(np.diagonal(res["cell"], axis1=1, axis2=2) == 0)
    assert np.array_equal(res["cell"], np.swapaxes(res["cell"], 1, 2))
    for k, ng in enumerate([3, 2]):
        assert np.array_equal(res["group_counts"][k], np.bincount(codes[k], minlength=ng))
        ref_sum = np.stack([dref[codes[k] == g].sum(0) for g in range(ng)])
        assert _relerr(res["group_sums"][k], ref_sum) < TOL["fp32"]
    h.close()


@pytest.mark.parametrize("norm", ["l1", "linf", "l2"])
def test_norms(built_library, norm):
    dims = mrvi_b200.MrviDims(n_input=40, n_sample=70)
    params, h = _model(dims, seed=1)
    X, sid = _inputs(50, 40, 70, seed=2)
    res = h.run_host(X, sid, norm=norm, keep_cell=True, want_z=True)
    dref = orc.pairwise_distances(res["z"].astype(np.float64), norm)
    assert _relerr(res["cell"], dref) < 1e-6
    h.close()


def test_chunking_and_ragged(built_library):
    """Results do not depend on the host chunk size; ragged final chunk; n smaller than a tile."""
    dims = mrvi_b200.MrviDims(n_input=64, n_sample=12)
    params, h = _model(dims, seed=2)
    X, sid = _inputs(1000, 64, 12, seed=9)
    codes = [np.random.default_rng(0).integers(0, 5, size=1000).astype(np.int32)]
    a = h.run_host(X, sid, codes, [5], keep_cell=True)
    b = h.run_host(X, sid, codes, [5], keep_cell=True, chunk_cells=137)
    assert np.array_equal(a["cell"], b["cell"])
    assert np.allclose(a["group_sums"][0], b["group_sums"][0], rtol=1e-6)
    c = h.run_host(X[:3], sid[:3], [codes[0][:3]], [5], keep_cell=True)
    assert np.array_equal(c["cell"], a["cell"][:3])
    e = h.run_host(X[:0], sid[:0], [codes[0][:0]], [5], keep_cell=True)
    assert e["cell"].shape == (0, 12, 12) and np.all(e["group_sums"][0] == 0)
    h.close()


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_host_narrowing_is_bit_transparent(built_library, precision):
    """Count chunks cross PCIe as uint8 / uint16 (host_pack.cpp) and are widened back on the GPU: the result must be
    BIT-identical to the float32 route, which a single non-integer value per chunk forces (only the cells that own
    the changed values may differ)."""
    dims = mrvi_b200.MrviDims(n_input=1200, n_sample=8)
    params, h = _model(dims, seed=3, precision=precision)
    n = 4000                                   # 4.8 M values: above the narrowing threshold
    X, sid = _inputs(n, 1200, 8, seed=4)
    a = h.run_host(X, sid, keep_cell=True, chunk_cells=1024)
    assert h.last_h2d_bytes() == n * 1200      # uint8
    X16 = X.copy()
    X16[::7, 5] = 40000.0                      # -> uint16 chunks
    b = h.run_host(X16, sid, keep_cell=True, chunk_cells=1024)
    assert h.last_h2d_bytes() == n * 1200 * 2
    Xf = X.copy()
    poison = np.arange(0, n, 1024)             # first cell of every chunk: every chunk goes raw
    Xf[poison, 0] += 0.5
    c = h.run_host(Xf, sid, keep_cell=True, chunk_cells=1024)
    assert h.last_h2d_bytes() == n * 1200 * 4
    keep = np.ones(n, bool)
    keep[poison] = False
    assert np.array_equal(a["cell"][keep], c["cell"][keep])
    Xf16 = X16.copy()
    Xf16[poison, 0] += 0.5
    d = h.run_host(Xf16, sid, keep_cell=True, chunk_cells=1024)
    assert np.array_equal(b["cell"][keep], d["cell"][keep])
    h.close()


def test_split_h2d_of_page_locked_input_is_bit_transparent(built_library, monkeypatch):
    """Page-locked float32 input: the leading rows of every chunk are narrowed by the host threads while the copy engine
    pulls the rest raw (run_host_impl: m_pack / m = C / (P + 3/4 C)).  Which rows go which way depends on the measured host
    rate, so the result must be bit-identical to the all-narrowed route and to pageable input, whatever the split
    (ragged last chunk, uint16 chunks, a pitched matrix)."""
    dims = mrvi_b200.MrviDims(n_input=1200, n_sample=8)
    params, h = _model(dims, seed=3, precision="tc")
    n = 9000
    X, sid = _inputs(n, 1200, 8, seed=4)
    X[::11, 7] = 30000.0                       # some uint16 chunks
    ref = h.run_host(X, sid, keep_cell=True, chunk_cells=2048)["cell"].copy()   # pageable: all narrowed
    Xp = _capi.pinned_empty((n, 1200))
    Xp[:] = X
    monkeypatch.setenv("MRVI_B200_NO_SPLIT_H2D", "1")
    a = h.run_host(Xp, sid, keep_cell=True, chunk_cells=2048)["cell"].copy()
    bytes_narrow = h.last_h2d_bytes()
    monkeypatch.delenv("MRVI_B200_NO_SPLIT_H2D")
    sent = {}
    for pcie in ("47", "5", "400"):            # assumed raw rate of the link: default / all rows narrowed / mostly raw
        monkeypatch.setenv("MRVI_B200_PCIE_GBPS", pcie)
        for _ in range(2):                     # the first call measures the host rate, the second one splits
            b = h.run_host(Xp, sid, keep_cell=True, chunk_cells=2048)["cell"]
        sent[pcie] = h.last_h2d_bytes()
        assert np.array_equal(a, b), pcie
    assert sent["400"] > sent["5"], (sent, bytes_narrow)   # more rows travel raw when the link is assumed faster: the split is live
    assert np.array_equal(a, ref)
    Xw = _capi.pinned_empty((n, 1300))         # ldx > G: pitched copies of the raw rows
    Xw[:, :1200] = X
    c = h.run_host(Xw[:, :1200], sid, keep_cell=True, chunk_cells=2048)["cell"]
    assert np.array_equal(a, c)
    h.close()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int32, np.int64])
def test_integer_count_matrix_input(built_library, dtype):
    """`mrvi_run_host_counts`: adata.X stored as integer counts gives bit-identical results to its float32 copy (what
    scvi's loader would hand the reference).  uint8 / uint16 travel as they are (1 / 2 bytes per value on the wire);
    int32 / int64 are narrowed chunk by chunk inside the library (no float32 copy of the matrix); strided rows; model API."""
    dims = mrvi_b200.MrviDims(n_input=300, n_sample=6)
    params, h = _model(dims, seed=2, precision="tc")
    X, sid = _inputs(900, 300, 6, seed=3)
    Xi = X.astype(dtype)
    if dtype == np.uint16:
        Xi[::5, 7] = 50000
    if dtype in (np.int32, np.int64):
        Xi[256:512, 3] = 70000      # one 256-cell chunk needs the float32 fall-back (>= 65536), its neighbours uint8
        Xi[700, 5] = 300            # ... and one uint16
    codes = [(np.arange(900) % 3).astype(np.int32)]
    a = h.run_host(Xi.astype(np.float32), sid, codes, [3], keep_cell=True, want_u=True)
    b = h.run_host(Xi, sid, codes, [3], keep_cell=True, want_u=True, chunk_cells=256)
    if dtype in (np.uint8, np.uint16):
        assert h.last_h2d_bytes() == 900 * 300 * Xi.itemsize
    else:  # 4 chunks of 256 (last: 132 cells): uint8, float32 fall-back, uint16 (cell 700), uint8
        assert h.last_h2d_bytes() == 300 * (256 * 1 + 256 * 4 + 256 * 2 + 132 * 1)
    assert np.array_equal(a["cell"], b["cell"]) and np.array_equal(a["u"], b["u"])
    wide = np.zeros((900, 320), dtype=dtype)
    wide[:, :300] = Xi
    c = h.run_host(wide[:, :300], sid, codes, [3], keep_cell=True)   # ldx = 320 elements
    assert np.array_equal(a["cell"], c["cell"])
    h.close()
    ad = mrvi_b200.synthetic_iid(n_obs=200, n_vars=300, n_sample=6, seed=1)
    m32 = mrvi_b200.MrVI(ad, params, sample_key="sample_str", precision="tc")
    d32 = np.asarray(m32.get_local_sample_distances(keep_cell=True)["cell"].values)
    m32.close()
    ad.X = ad.X.astype(dtype)
    mi = mrvi_b200.MrVI(ad, params, sample_key="sample_str", precision="tc")
    di = np.asarray(mi.get_local_sample_distances(keep_cell=True)["cell"].values)
    assert np.array_equal(d32, di)
    mi.close()


def test_drop_in_api_matches_reference_layout(built_library):
    """The reference's own grouped-distance test (`tests/test_model.py:221-241`): shapes, symmetry, both
    outputs in one Dataset -- plus values against the oracle and value_counts() group order."""
    ad = mrvi_b200.synthetic_iid(n_obs=400, n_vars=100, n_sample=15, seed=4)
    dims = mrvi_b200.MrviDims(n_input=100, n_sample=15, n_latent=10, n_latent_u=10)
    params = mrvi_b200.init_params(dims, seed=8)
    model = mrvi_b200.MrVI(ad, params, sample_key="sample_str")
    dists = model.get_local_sample_distances(groupby=["labels", "label_2"])
    cell = dists["cell"]
    assert cell.shape == (400, 15, 15)
    assert dists["labels"].shape == (3, 15, 15)
    assert dists["label_2"].shape == (2, 15, 15)
    for key in ("labels", "label_2"):
        g0 = np.asarray(dists[key][0].values)
        assert np.allclose(g0, g0.T, atol=1e-6)
    assert "_indices" in ad.obs
    sid = model._sample_codes(ad)
    ref = orc.local_sample_distances(params, ad.X, sid, {k: ad.obs[k].to_numpy() for k in ("labels", "label_2")},
                                     keep_cell=True, dtype=np.float64)
    assert _relerr(np.asarray(cell.values), ref["cell"]) < 1e-5
    for key in ("labels", "label_2"):
        vc = ad.obs[key].value_counts()
        assert list(np.asarray(dists[key].coords[f"{key}_name"])) == list(vc.index)
        assert _relerr(np.asarray(dists[key].values), ref[key]["means"]) < 1e-5
        assert np.asarray(dists[key].values).dtype == np.float32
    assert list(np.asarray(cell.coords["cell_name"])) == list(ad.obs_names)
    assert list(np.asarray(cell.coords["sample_x"])) == list(model.sample_order)
    # group-only call returns no "cell"
    g = model.get_local_sample_distances(groupby="labels", keep_cell=False)
    assert "cell" not in g and np.array_equal(np.asarray(g["labels"].values), np.asarray(dists["labels"].values))
    with pytest.raises(ValueError, match="Undefined computation"):
        model.get_local_sample_distances(keep_cell=False)
    z = model.get_local_sample_representation()
    assert z.shape == (400, 15, 10)
    model.close()


# ------------------------------------------------------------------------------------------------
# tensor-core mode (tcgen05): parity bar 1e-3 normwise; and agreement with the numpy emulation of the
# exact TC arithmetic (oracle/tc_emulation.py), which isolates kernel bugs from rounding effects.
TC_CASES = [
    ("S32", dict(n_input=200, n_sample=32), 300),
    ("S128", dict(n_input=200, n_sample=128), 150),
    ("S15_iso_L10", dict(n_input=100, n_sample=15, n_latent=10, n_latent_u=10), 200),
    ("S4", dict(n_input=100, n_sample=4), 400),
    ("S256_L50", dict(n_input=64, n_sample=256, n_latent=50), 12),
    ("S48_usemapF", dict(n_input=64, n_sample=48, use_map=False), 77),
    ("S200_Lu5", dict(n_input=64, n_sample=200, n_latent=10, n_latent_u=5), 9),
    ("oddG_S7", dict(n_input=101, n_sample=7), 133),
    ("G2000_S32", dict(n_input=2000, n_sample=32), 300),
]


@pytest.mark.parametrize("name,dk,n", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_tc_matches_oracle(built_library, name, dk, n):
    from oracle import tc_emulation as tce

    dims = mrvi_b200.MrviDims(**dk)
    params, h = _model(dims, seed=4, precision="tc")
    X, sid = _inputs(n, dims.n_input, dims.n_sample, seed=6)
    codes = [np.random.default_rng(3).integers(0, 4, size=n).astype(np.int32)]
    res = h.run_host(X, sid, codes, [4], keep_cell=True, want_z=True, want_u=True)
    p64 = orc.cast_params(params, np.float64)
    uref = orc.encoder_xu_mean(p64, X.astype(np.float64), sid)
    zref = orc.counterfactual_z(params, X, sid, dtype=np.float64)
    dref = orc.pairwise_distances(zref)
    L = dims.n_latent
    zb = orc.dense(uref, p64, "qz/Dense_0") if dims.n_latent_u is not None else uref
    assert _relerr(res["u"], uref) < 2e-5  # tensor-core encoder GEMM uses a 3-term f16 split (~fp32 accuracy)
    # (a) kernel == emulation of its own arithmetic ((hi, lo) f16 weights, split activations ~ exact)
    emu = tce.tc_chain_residual(params, res["u"].astype(np.float64), fmt_act=None, fmt_w="f16x2")[..., :L]
    got_res = res["z"].astype(np.float64) - zb[:, None, :]
    assert _relerr(got_res, emu) < 5e-5, "kernel deviates from the emulated tensor-core arithmetic"
    # (b) parity with the oracle
    scale = dref.max(axis=(1, 2), keepdims=True)
    err = float((np.abs(res["cell"] - dref) / scale).max())
    assert err < TOL["tc"], err
    assert np.all(np.diagonal(res["cell"], axis1=1, axis2=2) == 0)
    assert np.array_equal(res["group_counts"][0], np.bincount(codes[0], minlength=4))
    ref_sum = np.stack([dref[codes[0] == g].sum(0) for g in range(4)])
    assert _relerr(res["group_sums"][0], ref_sum) < TOL["tc"]
    h.close()


@pytest.mark.parametrize("S,n", [(32, 700), (128, 90), (15, 333), (48, 150)])
def test_tc_fused_groups_skewed_two_keys(built_library, S, n):
    """Fused kernel: Zipf-skewed groups (long and short runs, mixed tiles at the boundaries), two keys with -1
    (= in no group) codes, groupby-only output, vs the oracle distances."""
    dims = mrvi_b200.MrviDims(n_input=80, n_sample=S)
    params, h = _model(dims, seed=7, precision="tc")
    X, sid = _inputs(n, 80, S, seed=8)
    rng = np.random.default_rng(1)
    k0 = np.minimum(rng.zipf(1.6, size=n) - 1, 6).astype(np.int32)  # 7 groups, heavily skewed
    k1 = rng.integers(-1, 3, size=n).astype(np.int32)               # -1 = no group
    res = h.run_host(X, sid, [k0, k1], [7, 3], keep_cell=False)
    assert "cell" not in res
    dref = orc.pairwise_distances(orc.counterfactual_z(params, X, sid, dtype=np.float64))
    scale = dref.max()
    for k, (codes, ng) in enumerate([(k0, 7), (k1, 3)]):
        assert np.array_equal(res["group_counts"][k], np.bincount(codes[codes >= 0], minlength=ng))
        ref_sum = np.stack([dref[codes == g].sum(0) for g in range(ng)])
        cnt = np.maximum(res["group_counts"][k], 1)[:, None, None]
        assert np.abs(res["group_sums"][k] / cnt - ref_sum / cnt).max() / scale < TOL["tc"]
    # keep_cell + groups in the same call give the same sums (run accumulation vs per-cell rows)
    res2 = h.run_host(X, sid, [k0, k1], [7, 3], keep_cell=True, chunk_cells=256)
    for k in range(2):
        assert np.allclose(res2["group_sums"][k], res["group_sums"][k], rtol=1e-5, atol=1e-5 * scale)
        ref_from_cells = np.stack([res2["cell"][[k0, k1][k] == g].astype(np.float64).sum(0) for g in range([7, 3][k])])
        assert np.allclose(res2["group_sums"][k], ref_from_cells, rtol=1e-5, atol=1e-5 * scale)
    h.close()


def test_tc_fused_near_duplicate_samples(built_library):
    """Two samples with identical / nearly identical embeddings: cancellation-dominated Gram entries must take
    the exact direct-difference route (distance 0 resp. tiny, not sqrt of rounding noise)."""
    dims = mrvi_b200.MrviDims(n_input=60, n_sample=32)
    params = mrvi_b200.init_params(dims, seed=9)
    emb = params["qz/Embed_0/embedding"]
    emb[5] = emb[3]                      # exact duplicate
    emb[7] = emb[6] * (1 + 1e-3)         # near duplicate
    h = _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)
    X, sid = _inputs(200, 60, 32, seed=3)
    res = h.run_host(X, sid, keep_cell=True)
    dref = orc.pairwise_distances(orc.counterfactual_z(params, X, sid, dtype=np.float64))
    scale = dref.max(axis=(1, 2), keepdims=True)
    assert np.all(res["cell"][:, 3, 5] == 0) and np.all(res["cell"][:, 5, 3] == 0)
    assert (np.abs(res["cell"] - dref) / scale).max() < TOL["tc"]
    assert (np.abs(res["cell"][:, 6, 7] - dref[:, 6, 7]) / scale[:, 0, 0]).max() < 2e-4
    h.close()


@pytest.mark.parametrize("S,L,n", [(256, 50, 120), (200, 30, 90), (129, 10, 64), (252, 64, 40)])
def test_tc_large_S_gram_kernel(built_library, S, L, n):
    """128 < S <= 256 (BASELINE config 5): unfused chain + tensor-core Gram distance kernel (tc_dist.cuh).
    Skewed groups with -1 codes over more cells than SMs' ranges are long, keep_cell and groupby outputs, exact
    symmetry / zero diagonal, a duplicated and a near-duplicated sample (exact-recompute route), and the isolated
    stage through `mrvi_distances_device`-equivalent arithmetic (distances of the returned z)."""
    dims = mrvi_b200.MrviDims(n_input=48, n_sample=S, n_latent=L, n_latent_u=10 if L != 10 else None)
    params = mrvi_b200.init_params(dims, seed=12)
    emb = params["qz/Embed_0/embedding"]
    emb[S - 1] = emb[2]                    # duplicate across the two 128-row blocks
    emb[130 if S > 131 else 1] = emb[0] * (1 + 1e-3)
    h = _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)
    X, sid = _inputs(n, 48, S, seed=13)
    rng = np.random.default_rng(2)
    k0 = np.minimum(rng.zipf(1.7, size=n) - 1, 4).astype(np.int32)
    k1 = rng.integers(-1, 2, size=n).astype(np.int32)
    res = h.run_host(X, sid, [k0, k1], [5, 2], keep_cell=True, want_z=True)
    dref = orc.pairwise_distances(orc.counterfactual_z(params, X, sid, dtype=np.float64))
    scale = dref.max(axis=(1, 2), keepdims=True)
    assert (np.abs(res["cell"] - dref) / scale).max() < TOL["tc"]
    # (i, j) and (j, i) inside a diagonal block are separate Gram entries (different accumulation order of the
    # split products): symmetric to the Gram error, far inside the tensor-core parity bar; the off-diagonal
    # blocks are exact mirrors.  (fp32 mode is bit-symmetric, test_fp32_matches_oracle.)
    asym = np.abs(res["cell"] - np.swapaxes(res["cell"], 1, 2))
    assert (asym / scale).max() < 1e-4
    assert np.array_equal(res["cell"][:, :128, 128:], np.swapaxes(res["cell"][:, 128:, :128], 1, 2))
    assert np.all(np.diagonal(res["cell"], axis1=1, axis2=2) == 0)
    assert np.all(res["cell"][:, 2, S - 1] == 0)
    # the distance stage alone, against exact distances of the z the chain produced: Gram error only
    dz = orc.pairwise_distances(res["z"].astype(np.float64))
    # (pairs just above the 2^-12 exact-recompute threshold carry a relative error of ~2^-9 on a distance that is
    #  itself ~2^-6 of the scale)
    assert (np.abs(res["cell"] - dz) / scale).max() < 1e-4
    for k, (codes, ng) in enumerate([(k0, 5), (k1, 2)]):
        assert np.array_equal(res["group_counts"][k], np.bincount(codes[codes >= 0], minlength=ng))
        from_cells = np.stack([res["cell"][codes == g].astype(np.float64).sum(0) for g in range(ng)])
        assert np.allclose(res["group_sums"][k], from_cells, rtol=1e-5, atol=1e-5 * dref.max())
    only = h.run_host(X, sid, [k0, k1], [5, 2], keep_cell=False, chunk_cells=50)
    for k in range(2):
        assert np.allclose(only["group_sums"][k], res["group_sums"][k], rtol=1e-5, atol=1e-5 * dref.max())
    h.close()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tc", 1e-3)])
def test_latent_representation_api(built_library, precision, tol):
    """get_latent_representation (reference `_model.py:221-271`): u, and z of each cell for its own sample."""
    ad = mrvi_b200.synthetic_iid(n_obs=250, n_vars=80, n_sample=6, seed=5)
    dims = mrvi_b200.MrviDims(n_input=80, n_sample=6)
    params = mrvi_b200.init_params(dims, seed=2)
    model = mrvi_b200.MrVI(ad, params, sample_key="sample_str", precision=precision)
    sid = model._sample_codes(ad)
    p64 = orc.cast_params(params, np.float64)
    u_ref = orc.encoder_xu_mean(p64, ad.X.astype(np.float64), sid)
    z_ref = orc.inference_z(p64, ad.X.astype(np.float64), sid, sid)
    u = model.get_latent_representation()
    z = model.get_latent_representation(give_z=True)
    assert u.shape == (250, 10) and z.shape == (250, 30)
    assert _relerr(u, u_ref) < 2e-5
    assert _relerr(z, z_ref) < tol
    sub = model.get_latent_representation(indices=[5, 3, 200])
    assert np.array_equal(sub, u[[5, 3, 200]]) or _relerr(sub, u_ref[[5, 3, 200]]) < 2e-5
    model.close()


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_sparse_csr_input_equals_dense(built_library, precision):
    """scipy CSR X goes through mrvi_run_host_csr (non-zeros over PCIe, densify on the GPU): bit-identical to dense."""
    import scipy.sparse as sp

    dims = mrvi_b200.MrviDims(n_input=150, n_sample=9)
    params, h = _model(dims, seed=5, precision=precision)
    X, sid = _inputs(777, 150, 9, seed=4)
    X[5] = 0  # an all-zero cell
    codes = [np.random.default_rng(0).integers(0, 3, size=777).astype(np.int32)]
    a = h.run_host(X, sid, codes, [3], keep_cell=True, want_u=True)
    b = h.run_host(sp.csr_matrix(X), sid, codes, [3], keep_cell=True, want_u=True, chunk_cells=200)
    assert np.array_equal(a["u"], b["u"])
    assert np.array_equal(a["cell"], b["cell"])
    assert np.allclose(a["group_sums"][0], b["group_sums"][0], rtol=1e-6)
    # through the host API with a sparse AnnData-like object
    ad = mrvi_b200.SimpleAnnData(sp.csr_matrix(X), mrvi_b200.synthetic_iid(n_obs=777, n_vars=150, n_sample=9, seed=1).obs)
    ad.obs["sample"] = sid
    model = mrvi_b200.MrVI(ad, params, sample_key="sample", precision=precision)
    out = model.get_local_sample_distances(groupby="labels")
    assert np.array_equal(np.asarray(out["cell"].values), a["cell"])
    model.close()
    h.close()


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_device_entry_points_match_host_path(built_library, precision):
    """`mrvi_run_device` (device-resident buffers, caller's stream) and `mrvi_distances_device` (stage 3 alone)
    against `mrvi_run_host` / the oracle."""
    import torch

    dims = mrvi_b200.MrviDims(n_input=96, n_sample=12)
    params, h = _model(dims, seed=5, precision=precision)
    n = 700
    X, sid = _inputs(n, 96, 12, seed=6)
    codes = np.random.default_rng(0).integers(-1, 4, size=n).astype(np.int32)
    ref = h.run_host(X, sid, [codes], [4], keep_cell=True, want_z=True, want_u=True)
    dev = torch.device("cuda:0")
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        Xd = torch.from_numpy(np.pad(X, ((0, 0), (0, 4)))).to(dev)        # ldx = 100 > n_input
        sd, cd = torch.from_numpy(sid).to(dev), torch.from_numpy(codes).to(dev)
        cell = torch.empty((n, 12, 12), dtype=torch.float32, device=dev)
        sums = torch.zeros((4, 12, 12), dtype=torch.float64, device=dev)
        z = torch.empty((n, 12, dims.n_latent), dtype=torch.float32, device=dev)
        u = torch.empty((n, dims.n_latent_q), dtype=torch.float32, device=dev)
        h.run_device(n, Xd.data_ptr(), 100, sd.data_ptr(), [cd.data_ptr()], [4], cell_out_ptr=cell.data_ptr(),
                     group_sum_ptrs=[sums.data_ptr()], z_out_ptr=z.data_ptr(), u_out_ptr=u.data_ptr(), stream=st.cuda_stream)
        # sums are ADDED into: a second call doubles them
        h.run_device(n, Xd.data_ptr(), 100, sd.data_ptr(), [cd.data_ptr()], [4], group_sum_ptrs=[sums.data_ptr()],
                     stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(u.cpu().numpy(), ref["u"]) and np.array_equal(z.cpu().numpy(), ref["z"])
    assert np.array_equal(cell.cpu().numpy(), ref["cell"])
    assert np.allclose(sums.cpu().numpy(), 2 * ref["group_sums"][0], rtol=1e-6, atol=1e-9)
    # stage 3 alone on given representations, every norm
    zt = torch.from_numpy(ref["z"]).to(dev)
    for norm in ("l2", "l1", "linf"):
        out = torch.empty((n, 12, 12), dtype=torch.float32, device=dev)
        gs = torch.zeros((4, 12, 12), dtype=torch.float64, device=dev)
        h.distances_device(n, zt.data_ptr(), [cd.data_ptr()], [4], norm=norm, cell_out_ptr=out.data_ptr(), group_sum_ptrs=[gs.data_ptr()])
        torch.cuda.synchronize()
        dref = orc.pairwise_distances(ref["z"].astype(np.float64), norm)
        assert _relerr(out.cpu().numpy(), dref) < 1e-6
        gref = np.stack([dref[codes == g].sum(0) for g in range(4)])
        assert _relerr(gs.cpu().numpy(), gref) < 1e-6
    h.close()


@pytest.mark.parametrize("S,L,n", [(160, 24, 300), (128, 30, 400), (100, 50, 333), (33, 10, 500)])
def test_distances_device_large_S_tensor_core(built_library, S, L, n):
    """`mrvi_distances_device` in tensor-core mode with 32 < S <= 256 runs `k_dist_gram_tc` on caller-provided z
    (three symmetric blocks above 128 samples, one block below)."""
    import torch

    dims = mrvi_b200.MrviDims(n_input=32, n_sample=S, n_latent=L)
    params, h = _model(dims, seed=1, precision="tc")
    rng = np.random.default_rng(3)
    z = (rng.standard_normal((n, 1, L)) + 0.2 * rng.standard_normal((n, S, L))).astype(np.float32)  # common offset + spread
    z[:, 7] = z[:, 3]                                                                                 # duplicate sample
    codes = rng.integers(0, 6, size=n).astype(np.int32)
    dev = torch.device("cuda:0")
    zt, cd = torch.from_numpy(z).to(dev), torch.from_numpy(codes).to(dev)
    out = torch.empty((n, S, S), dtype=torch.float32, device=dev)
    gs = torch.zeros((6, S, S), dtype=torch.float64, device=dev)
    h.distances_device(n, zt.data_ptr(), [cd.data_ptr()], [6], cell_out_ptr=out.data_ptr(), group_sum_ptrs=[gs.data_ptr()])
    torch.cuda.synchronize()
    dref = orc.pairwise_distances(z.astype(np.float64))
    got = out.cpu().numpy()
    scale = dref.max(axis=(1, 2), keepdims=True)
    assert (np.abs(got - dref) / scale).max() < 1e-4
    assert np.all(got[:, 3, 7] == 0) and np.all(np.diagonal(got, axis1=1, axis2=2) == 0)
    gref = np.stack([dref[codes == g].sum(0) for g in range(6)])
    assert _relerr(gs.cpu().numpy(), gref) < 1e-4
    h.close()


def test_tc_kernels_are_run_to_run_deterministic(built_library):
    """Racecheck substitute (compute-sanitizer is not available on the pool): the warp-specialised kernels must give
    bit-identical u, z and per-cell distances for identical input, across runs and host chunkings.  (This is the
    check that caught the missing `fence.proxy.async` between the converters' shared-memory loads and the copy
    engine's refill of a raw stage in `k_encoder_tc`: a few hundred of 100k cells were wrong per launch.)"""
    n, G = 30000, 1200                      # several 128-cell tiles per CTA: the encoder's rings wrap many times
    dims = mrvi_b200.MrviDims(n_input=G, n_sample=8)
    params, h = _model(dims, seed=3, precision="tc")
    X, sid = _inputs(n, G, 8, seed=4)
    ref = h.run_host(X, sid, keep_cell=True, want_u=True, want_z=True)
    for it in range(6):
        r = h.run_host(X, sid, keep_cell=True, want_u=True, want_z=True, chunk_cells=[0, 4096, 777][it % 3])
        assert np.array_equal(r["u"], ref["u"]), f"encoder output differs between runs (iteration {it})"
        assert np.array_equal(r["z"], ref["z"]) and np.array_equal(r["cell"], ref["cell"])
    h.close()
    for S, L in ((128, 30), (256, 50)):
        dims = mrvi_b200.MrviDims(n_input=64, n_sample=S, n_latent=L)
        params, h = _model(dims, seed=5, precision="tc")
        X, sid = _inputs(600, 64, S, seed=6)
        codes = np.random.default_rng(2).integers(0, 5, size=600).astype(np.int32)
        ref = h.run_host(X, sid, [codes], [5], keep_cell=True)
        for it in range(4):
            r = h.run_host(X, sid, [codes], [5], keep_cell=True, chunk_cells=[0, 250][it % 2])
            assert np.array_equal(r["cell"], ref["cell"])
            assert np.allclose(r["group_sums"][0], ref["group_sums"][0], rtol=1e-5, atol=1e-5 * ref["cell"].max())
        h.close()


# ------------------------------------------------------------------------------------------------
# Regression cases from the random sweeps (tools/fuzz_parity.py).  Round 1 (gpurun_out/fuzz*.log): the eight
# configurations whose worst cell sat ABOVE the bar -- six in tensor-core mode (1.0e-3 .. 2.2e-3: single-f16 chain
# weights; fixed in round 2 by the (hi, lo) weight image, 3-term products) and two in fp32 mode (1.0e-5 / 1.4e-5 with
# one or few distances per cell).  The bar is applied PER CELL BLOCK (oracle/parity.py).
FUZZ_FAILS = [
    # name, precision, dims, n_cells
    ("tc_S300_L5", "tc", dict(n_input=100, n_sample=300, n_latent=5, n_latent_u=10, encoder_n_layers=1), 16),
    ("tc_S2_iso_L30", "tc", dict(n_input=256, n_sample=2, n_latent=30, n_latent_u=None), 1322),
    ("tc_S64_L33_usemapF", "tc", dict(n_input=256, n_sample=64, n_latent=33, n_latent_u=10, use_map=False), 44),
    ("tc_S4_L5", "tc", dict(n_input=100, n_sample=4, n_latent=5, n_latent_u=10, encoder_n_layers=1), 1658),
    ("tc_S16_L10_Lu8_usemapF", "tc", dict(n_input=8, n_sample=16, n_latent=10, n_latent_u=8, use_map=False), 185),
    ("tc_S255_L5_Lu16_usemapF", "tc", dict(n_input=31, n_sample=255, n_latent=5, n_latent_u=16, use_map=False, encoder_n_layers=1), 34),
    ("fp32_S8_L5_Lu16_qz2", "fp32", dict(n_input=31, n_sample=8, n_latent=5, n_latent_u=16, qz_n_layers=2, encoder_n_layers=3), 347),
    ("fp32_S2_L31_Lu5_qz2", "fp32", dict(n_input=256, n_sample=2, n_latent=31, n_latent_u=5, qz_n_layers=2), 9948),
    # round 2, second sweep on the two-sample centre: n_latent > 32 keeps ONE reference row per cell (shared memory) and the
    # rows were still indexed with the two-row stride -- with many cells per tile (small S) they ran past the buffer
    # (NaN blocks, wrong group sums, an illegal address)
    ("tc_S5_L33_Lu8_usemapF", "tc", dict(n_input=31, n_sample=5, n_latent=33, n_latent_u=8, use_map=False, encoder_n_layers=1), 97),
    ("tc_S5_L50_Lu8", "tc", dict(n_input=101, n_sample=5, n_latent=50, n_latent_u=8), 541),
    ("tc_S4_L33_Lu8", "tc", dict(n_input=8, n_sample=4, n_latent=33, n_latent_u=8), 2477),
]


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("name,precision,dk,n", FUZZ_FAILS, ids=[c[0] for c in FUZZ_FAILS])
def test_round1_fuzz_failures(built_library, name, precision, dk, n, seed):
    from oracle.parity import distance_error_stats

    dims = mrvi_b200.MrviDims(**dk)
    params, h = _model(dims, seed=100 + seed, precision=precision)
    X, sid = _inputs(n, dims.n_input, dims.n_sample, seed=200 + seed)
    rng = np.random.default_rng(300 + seed)
    codes = rng.integers(-1, 3, size=n).astype(np.int32)
    res = h.run_host(X, sid, [codes], [3], keep_cell=True, want_z=True, chunk_cells=int(rng.choice([0, 1, 7, 64])))
    g_only = h.run_host(X, sid, [codes], [3], keep_cell=False)
    h.close()
    zref = orc.counterfactual_z(params, X, sid, dtype=np.float64)
    dref = orc.pairwise_distances(zref)
    st = distance_error_stats(res["cell"], dref)
    bar = TOL[precision]
    if precision == "fp32":
        # With one or few distances per cell the per-cell norm degenerates to an element-wise relative error, and the
        # reference's own dtype (float32) is that far from the exact result too: the bar is 1e-5 or twice the float32
        # oracle's own distance from the float64 one, whichever is larger (both numbers are printed).
        d32 = orc.pairwise_distances(orc.counterfactual_z(params, X, sid, dtype=np.float32).astype(np.float64))
        floor = distance_error_stats(d32, dref)["max_norm_err"]
        bar = max(bar, 2.0 * floor)
        print(f"\n[fuzz-regression] {name} seed {seed}: float32 oracle vs float64 oracle {floor:.2e}")
    print(f"\n[fuzz-regression] {name} seed {seed}: max_norm_err {st['max_norm_err']:.2e} (bar {bar:.1e}), "
          f"p99 elementwise rel {st['p99_elementwise_rel']:.2e}")
    assert st["max_norm_err"] < bar, st
    assert _relerr(res["z"], zref) < bar
    gref = np.stack([dref[codes == g].sum(0) for g in range(3)])
    gs = max(np.abs(gref).max(), 1e-30)
    assert np.abs(res["group_sums"][0] - gref).max() / gs < bar
    assert np.abs(g_only["group_sums"][0] - gref).max() / gs < bar
    assert np.array_equal(res["group_counts"][0], np.bincount(codes[codes >= 0], minlength=3))
    assert np.all(np.diagonal(res["cell"], axis1=1, axis2=2) == 0)


def test_pinned_output_buffers(built_library):
    """`mrvi_alloc_pinned`: the big outputs of run_host live in page-locked memory (views keep it alive; freed with the
    last view) and caller-provided slices of one pinned array receive the shards' blocks."""
    a = _capi.pinned_empty((1000, 8, 8))
    a[:] = 1.0
    v = a[10:20]
    del a
    assert float(v.sum()) == 10 * 64
    dims = mrvi_b200.MrviDims(n_input=50, n_sample=8)
    params, h = _model(dims, seed=1, precision="tc")
    X, sid = _inputs(300, 50, 8, seed=2)
    whole = h.run_host(X, sid, keep_cell=True, want_z=True)
    out = _capi.pinned_empty((300, 8, 8))
    zo = _capi.pinned_empty((300, 8, dims.n_latent))
    for lo, hi in ((0, 100), (100, 300)):
        h.run_host(X[lo:hi], sid[lo:hi], keep_cell=True, want_z=True, cell_out=out[lo:hi], z_out=zo[lo:hi])
    assert np.array_equal(out, whole["cell"]) and np.array_equal(zo, whole["z"])
    h.close()


def test_multi_device_api_matches_single_device(built_library):
    """`MrVI(devices=[...])`: cells sharded over the visible GPUs (contiguous ranges, one handle + host thread each), group
    sums added on the host, keep_cell blocks written by every device into its slice of one output: counts and cell order
    bit-exact, per-cell blocks bit-identical (same kernels on the same cells), group means equal up to float64
    re-association.  Needs >= 2 GPUs (`gpurun --gpus 2`); with one GPU the sharding logic runs over a single device."""
    ndev = _capi.device_count()
    ad = mrvi_b200.synthetic_iid(n_obs=1003, n_vars=120, n_sample=16, seed=5)
    dims = mrvi_b200.MrviDims(n_input=120, n_sample=16)
    params = mrvi_b200.init_params(dims, seed=6)
    single = mrvi_b200.MrVI(ad, params, sample_key="sample_str", precision="tc", device=0)
    ref = single.get_local_sample_distances(groupby=["labels", "label_2"], keep_cell=True)
    zref = single.get_local_sample_representation()
    uref = single.get_latent_representation()
    single.close()
    devices = list(range(min(ndev, 4))) if ndev >= 2 else [0]
    multi = mrvi_b200.MrVI(ad, params, sample_key="sample_str", precision="tc", devices=devices)
    if ndev < 2:   # one GPU: force the sharded code path over two handles on the same device
        multi._handles.append(_capi.Handle(multi.dims, multi.params, device=0, precision=_capi.PRECISION_TC))
    got = multi.get_local_sample_distances(groupby=["labels", "label_2"], keep_cell=True)
    assert np.array_equal(np.asarray(got["cell"].values), np.asarray(ref["cell"].values))
    assert list(got["cell"].coords["cell_name"]) == list(ref["cell"].coords["cell_name"])
    for key in ("labels", "label_2"):
        assert list(got[key].coords[f"{key}_name"]) == list(ref[key].coords[f"{key}_name"])
        np.testing.assert_allclose(np.asarray(got[key].values), np.asarray(ref[key].values), rtol=2e-6, atol=0)
    assert np.array_equal(np.asarray(multi.get_local_sample_representation().values), np.asarray(zref.values))
    assert np.array_equal(multi.get_latent_representation(), uref)
    multi.close()
