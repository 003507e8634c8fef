"""Edge-case calls through the C-ABI (empty inputs, S = 1, tiny G, one cell, every chunking): must not crash and must
match the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
from oracle import mrvi_oracle as orc
bad = 0
def check(name, cond):
    global bad
    print(("ok   " if cond else "FAIL ") + name, flush=True)
    bad += 0 if cond else 1
for prec in (_capi.PRECISION_FP32, _capi.PRECISION_TC):
    tol = 1e-5 if prec == _capi.PRECISION_FP32 else 1e-3
    for (G, S, n) in ((1, 1, 3), (3, 1, 1), (5, 2, 1), (7, 3, 2), (4, 4, 1), (9, 129, 1), (8, 256, 2), (33, 5, 129)):
        dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
        flat = mrvi_b200.init_params(dims, seed=G + S)
        flat["qu/NormalDistOutputNN_0/Dense_1/bias"] = flat["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
        h = _capi.Handle(dims, flat, device=0, precision=prec)
        X = mrvi_b200.synthetic_counts(n, G, seed=1)
        sid = (np.arange(n) % S).astype(np.int32)
        codes = (np.arange(n) % 2).astype(np.int32)
        r = h.run_host(X, sid, [codes], [2], keep_cell=True, want_z=True, want_u=True)
        dref = orc.pairwise_distances(orc.counterfactual_z(flat, X, sid))
        sc = max(dref.max(), 1e-30)
        check(f"prec={prec} G={G} S={S} n={n} mean path", r["cell"].shape == (n, S, S) and np.abs(r["cell"] - dref).max() / sc < tol + (S == 1))
        e = h.run_host(X[:0], sid[:0], [codes[:0]], [2], keep_cell=True, want_z=True)
        check(f"prec={prec} G={G} S={S} empty", e["cell"].shape == (0, S, S) and np.all(e["group_sums"][0] == 0))
        u8 = h.run_host(X.astype(np.uint8), sid, [codes], [2], keep_cell=True)
        check(f"prec={prec} G={G} S={S} uint8 input", np.array_equal(u8["cell"], r["cell"]))
        rng = np.random.default_rng(0)
        mc = 2
        nu = rng.standard_normal((S, mc, n, dims.n_latent_q)).astype(np.float32)
        nb = rng.standard_normal((2 * mc, n, dims.n_latent_q)).astype(np.float32)
        s = h.run_sampled_host(X, sid, mc, u_noise=nu, base_noise=nb, group_codes=[codes], n_groups=[2], normalize=False, keep_cell=True)
        sref = orc.sampled_local_sample_distances(flat, X, sid, nu)
        check(f"prec={prec} G={G} S={S} n={n} sampled", np.abs(s["cell"] - sref).max() / max(sref.max(), 1e-30) < tol + (S == 1))
        s0 = h.run_sampled_host(X[:0], sid[:0], mc, group_codes=[codes[:0]], n_groups=[2], keep_cell=True)
        check(f"prec={prec} G={G} S={S} sampled empty", s0["cell"].shape == (0, mc, S, S))
        h.close()
print("failures:", bad)
