"""CPU tests of the oracle: the two independent restatements agree, and the invariants of
SURVEY.md section 8(c) hold.  (No numeric golden vectors exist in the reference's own tests --
tests/test_model.py:233-241 assert shapes and symmetry only -- see the oracle header.)"""
import numpy as np
import pytest
import torch

import mrvi_b200
from oracle import mrvi_oracle as orc
from oracle import tc_emulation as tce
from oracle.mrvi_oracle_torch import TorchOracle

CASES = [
    dict(n_input=100, n_sample=4),
    dict(n_input=60, n_sample=15, n_latent=10, n_latent_u=10),
    dict(n_input=40, n_sample=6, n_latent=10, n_latent_u=5),
    dict(n_input=30, n_sample=5, use_map=False),
    dict(n_input=30, n_sample=7, encoder_n_hidden=64, qz_n_layers=2),
]


def _setup(dk, n=48, seed=0):
    dims = mrvi_b200.MrviDims(**dk)
    p = mrvi_b200.init_params(dims, seed=seed + 1)
    X = mrvi_b200.synthetic_counts(n, dims.n_input, seed)
    sid = np.random.default_rng(seed).integers(0, dims.n_sample, n)
    return dims, p, X, sid


@pytest.mark.parametrize("dk", CASES)
def test_numpy_and_torch_restatements_agree(dk):
    dims, p, X, sid = _setup(dk)
    z_np = orc.counterfactual_z(p, X, sid, dtype=np.float64)
    t = TorchOracle(p, torch.float64)
    z_t = t.z(torch.as_tensor(X).double(), torch.as_tensor(sid)).numpy()
    assert z_np.shape == (X.shape[0], dims.n_sample, dims.n_latent)
    assert np.abs(z_np - z_t).max() < 1e-12
    d_np = orc.pairwise_distances(z_np)
    d_t = TorchOracle.distances(torch.as_tensor(z_t)).numpy()
    assert np.abs(d_np - d_t).max() < 1e-12
    # float32 (the reference's dtype) stays within 1e-5 of float64
    z32 = TorchOracle(p, torch.float32).z(torch.as_tensor(X), torch.as_tensor(sid)).numpy()
    assert np.abs(z32 - z_np).max() / np.abs(z_np).max() < 1e-5


def test_distance_invariants():
    dims, p, X, sid = _setup(dict(n_input=50, n_sample=9), n=40)
    z = orc.counterfactual_z(p, X, sid)
    for norm in ("l2", "l1", "linf"):
        d = orc.pairwise_distances(z, norm)
        assert np.array_equal(d, np.swapaxes(d, 1, 2))  # symmetric
        assert np.all(np.diagonal(d, axis1=1, axis2=2) == 0)  # exact zero diagonal
        assert np.all(d >= 0)
        tri = d[:, :, None, :] <= d[:, :, :, None] + np.swapaxes(d, 1, 2)[:, None, :, :] + 1e-12  # d(i,k) <= d(i,j) + d(j,k)
        assert tri.all()
    with pytest.raises(ValueError, match="not supported"):
        orc.pairwise_distances(z, "l3")


def test_z_base_cancels_in_distances():
    dims, p, X, sid = _setup(dict(n_input=50, n_sample=6), n=30)
    z = orc.counterfactual_z(p, X, sid)
    p2 = dict(p)
    p2["qz/Dense_0/kernel"] = np.zeros_like(p["qz/Dense_0/kernel"])
    p2["qz/Dense_0/bias"] = np.zeros_like(p["qz/Dense_0/bias"])
    z0 = orc.counterfactual_z(p2, X, sid)
    assert np.abs(orc.pairwise_distances(z) - orc.pairwise_distances(z0)).max() < 1e-12


def test_own_sample_changes_u_cf_sample_permutes():
    dims, p, X, sid = _setup(dict(n_input=50, n_sample=5), n=20)
    p64 = orc.cast_params(p, np.float64)
    u0 = orc.encoder_xu_mean(p64, X.astype(np.float64), sid)
    u1 = orc.encoder_xu_mean(p64, X.astype(np.float64), (sid + 1) % 5)
    assert np.abs(u0 - u1).max() > 1e-6  # the encoder reads the cell's own sample (`_module.py:190-199`)
    # permuting the sample embedding table permutes rows/cols of D
    perm = np.array([3, 0, 4, 1, 2])
    p2 = dict(p)
    p2["qz/Embed_0/embedding"] = p["qz/Embed_0/embedding"][perm]
    d = orc.pairwise_distances(orc.counterfactual_z(p, X, sid))
    d2 = orc.pairwise_distances(orc.counterfactual_z(p2, X, sid))
    assert np.abs(d2 - d[:, perm][:, :, perm]).max() < 1e-12


def test_use_map_false_keeps_first_half():
    dims, p, X, sid = _setup(dict(n_input=30, n_sample=4, use_map=False), n=16)
    z = orc.counterfactual_z(p, X, sid)
    p2 = dict(p)
    L = dims.n_latent
    p2["qz/AttentionBlock_0/MLP_1/Dense_0/kernel"] = p["qz/AttentionBlock_0/MLP_1/Dense_0/kernel"][:, :L]
    p2["qz/AttentionBlock_0/MLP_1/Dense_0/bias"] = p["qz/AttentionBlock_0/MLP_1/Dense_0/bias"][:L]
    assert np.abs(z - orc.counterfactual_z(p2, X, sid)).max() < 1e-12


def test_group_means_equal_mean_of_cell_blocks_and_value_counts_order():
    import pandas as pd

    dims, p, X, sid = _setup(dict(n_input=40, n_sample=4), n=300)
    labels = np.random.default_rng(5).choice(["b", "a", "c"], size=300, p=[0.2, 0.5, 0.3])
    out = orc.local_sample_distances(p, X, sid, {"ct": labels}, keep_cell=True, batch_size=64)
    vc = pd.Series(labels).value_counts()
    assert out["ct"]["labels"] == list(vc.index)
    assert np.array_equal(out["ct"]["counts"], vc.to_numpy())
    for g, lab in enumerate(out["ct"]["labels"]):
        assert np.allclose(out["ct"]["means"][g], out["cell"][labels == lab].mean(0), rtol=1e-12, atol=1e-14)
    # batch size (vmap chunking) does not change the result
    out2 = orc.local_sample_distances(p, X, sid, {"ct": labels}, keep_cell=True, batch_size=300)
    assert np.abs(out["cell"] - out2["cell"]).max() < 1e-12


def test_torch_port_batch_loop_matches_numpy():
    dims, p, X, sid = _setup(dict(n_input=40, n_sample=5), n=130)
    codes = np.random.default_rng(2).integers(0, 3, size=130)
    ref = orc.local_sample_distances(p, X, sid, {"k": codes}, keep_cell=True, batch_size=50, dtype=np.float64)
    got = TorchOracle(p, torch.float64).local_sample_distances(X, sid, [codes], [3], keep_cell=True, batch_size=50)
    assert np.abs(got["cell"] - ref["cell"]).max() < 1e-12
    order = ref["k"]["labels"]
    sums = np.stack([ref["k"]["means"][order.index(g)] * ref["k"]["counts"][order.index(g)] for g in range(3)])
    assert np.abs(got["sums"][0] - sums).max() < 1e-10


@pytest.mark.parametrize("dk", [dict(n_input=50, n_sample=8), dict(n_input=50, n_sample=5, n_latent=10, n_latent_u=10),
                                dict(n_input=50, n_sample=6, use_map=False)])
def test_tensor_core_algebra_is_exact_without_rounding(dk):
    """The folded algebra of the tcgen05 chain (attention collapse, bias / skip / eps_ folding) equals the
    oracle when no operand rounding is applied."""
    dims, p, X, sid = _setup(dk, n=24)
    p64 = orc.cast_params(p, np.float64)
    u = orc.encoder_xu_mean(p64, X.astype(np.float64), sid)
    z = orc.counterfactual_z(p, X, sid)
    zb = orc.dense(u, p64, "qz/Dense_0") if dims.n_latent_u is not None else u
    res = tce.tc_chain_residual(p, u, None, None)[..., : dims.n_latent]
    assert np.abs(res - (z - zb[:, None, :])).max() / np.abs(z).max() < 5e-6


def test_philox_mirror_known_answers():
    """Random123 known-answer vectors for philox4x32-10 (the generator of the device noise, oracle/philox.py)."""
    from oracle import philox

    kat = [((0, 0, 0, 0), 0, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, 0xFFFFFFFFFFFFFFFF, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0x299F31D0 << 32) | 0xA4093822,
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        got = tuple(int(v) for v in philox.philox4x32_10(*ctr, key))
        assert got == want
    x = philox.u_noise(3000, 4, 2, 10, seed=7)
    assert abs(float(x.mean())) < 5e-3 and abs(float(x.std()) - 1.0) < 5e-3


def test_sampled_oracle_properties():
    """Sampled paths: zero noise == mean path; the baseline statistics follow `_model.py:535-540`."""
    dims = mrvi_b200.MrviDims(n_input=40, n_sample=5)
    flat = mrvi_b200.init_params(dims, seed=2)
    X = mrvi_b200.synthetic_counts(10, 40, seed=3)
    sid = (np.arange(10) % 5).astype(np.int32)
    z0 = orc.sampled_counterfactual_z(flat, X, sid, np.zeros((5, 2, 10, dims.n_latent_q)))
    assert np.abs(z0[:, 0] - orc.counterfactual_z(flat, X, sid)).max() < 1e-13
    rng = np.random.default_rng(0)
    nb = rng.standard_normal((6, 10, dims.n_latent_q))
    mu, var = orc.local_baseline_dists(flat, X, sid, nb)
    assert mu.shape == (10,) and (mu > 0).all() and (var >= 0).all()
    nu = rng.standard_normal((5, 3, 10, dims.n_latent_q))
    d = orc.sampled_local_sample_distances(flat, X, sid, nu)
    assert d.shape == (10, 3, 5, 5) and np.allclose(d, np.swapaxes(d, 2, 3))
    dn = orc.sampled_local_sample_distances(flat, X, sid, nu, nb, normalize=True)
    assert dn.shape == (10, 5, 5)
    assert np.allclose(np.diagonal(dn, axis1=1, axis2=2), (-mu / np.sqrt(var))[:, None])
