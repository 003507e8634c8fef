"""Randomised parity sweep of the sampled / normalised paths (explicit noise) against the oracle.
python tools/fuzz_sampled.py [n_configs] [seed]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
from oracle import mrvi_oracle as orc

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
fails = 0
worst = {"fp32": 0.0, "tc": 0.0}
for c in range(n_cfg):
    S = int(rng.choice([2, 4, 5, 8, 16, 31, 32, 33, 64, 100, 128, 129, 150]))
    L = int(rng.choice([5, 10, 16, 30, 40]))
    Lu = None if rng.random() < 0.25 else int(rng.choice([5, 8, 10, 16]))
    if Lu == L:
        Lu = None
    G = int(rng.choice([16, 33, 100, 256]))
    mc = int(rng.choice([1, 2, 3, 5]))
    n = int(rng.integers(1, max(2, 6000 // (S * S * mc) + 12)))
    dims = mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=L, n_latent_u=Lu)
    flat = mrvi_b200.init_params(dims, seed=int(rng.integers(1 << 30)))
    flat["qu/NormalDistOutputNN_0/Dense_1/bias"] = flat["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    X = mrvi_b200.synthetic_counts(n, G, seed=int(rng.integers(1 << 30)))
    sid = rng.integers(0, S, n).astype(np.int32)
    ng = int(rng.integers(1, 4))
    codes = rng.integers(-1, ng, n).astype(np.int32)
    Lq = dims.n_latent_q
    nu = rng.standard_normal((S, mc, n, Lq)).astype(np.float32)
    nb = rng.standard_normal((2 * max(mc, 2), n, Lq)).astype(np.float32)
    dref = orc.sampled_local_sample_distances(flat, X, sid, nu)
    mcn = max(mc, 2)
    nun = rng.standard_normal((S, mcn, n, Lq)).astype(np.float32)
    nref = orc.sampled_local_sample_distances(flat, X, sid, nun, nb, normalize=True)
    mu, var = orc.local_baseline_dists(flat, X, sid, nb)
    cond = 1.0 + float((mu / np.sqrt(var)).max())
    for prec, tol in (("fp32", 1e-5), ("tc", 1e-3)):
        try:
            h = _capi.Handle(dims, flat, device=0, precision=_capi.PRECISION_FP32 if prec == "fp32" else _capi.PRECISION_TC)
            r = h.run_sampled_host(X, sid, mc, u_noise=nu, group_codes=[codes], n_groups=[ng], keep_cell=True, chunk_cells=int(rng.choice([0, 1, 5])))
            rn = h.run_sampled_host(X, sid, mcn, u_noise=nun, base_noise=nb, group_codes=[codes], n_groups=[ng], normalize=True, keep_cell=True)
            h.close()
        except _capi.MrviError as e:
            if e.status != -5:
                fails += 1
                print(f"cfg {c} {prec}: S={S} L={L} Lu={Lu} G={G} n={n} mc={mc}: {e}")
            continue
        scale = np.maximum(dref.max(axis=(2, 3), keepdims=True), 1e-30)
        e1 = float((np.abs(r["cell"] - dref) / scale).max())
        gref = np.stack([dref[codes == g].sum(0) for g in range(ng)])
        e2 = float(np.abs(r["group_sums"][0] - gref).max() / max(np.abs(gref).max(), 1e-30))
        e3 = float(np.abs(rn["cell"] - nref).max() / np.abs(nref).max()) / cond
        gnref = np.stack([nref[codes == g].sum(0) for g in range(ng)])
        e4 = float(np.abs(rn["group_sums"][0] - gnref).max() / max(np.abs(gnref).max(), 1e-30)) / cond
        worst[prec] = max(worst[prec], e1, e2, e3, e4)
        ok = max(e1, e2, e3, e4) < tol
        fails += 0 if ok else 1
        print(f"cfg {c:2d} {prec:4s} S={S:3d} L={L:2d} Lu={Lu} G={G:3d} n={n:3d} mc={mc}: sampled {e1:.1e}/{e2:.1e} normalised {e3:.1e}/{e4:.1e} (cond {cond:.0f}) {'ok' if ok else 'FAIL'}", flush=True)
print("worst:", worst, "failures:", fails)
