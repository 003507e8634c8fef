"""Repeatability stress of the TC chain / distance kernels: the same input many times must give bit-identical
per-cell distance matrices (a substitute for racecheck, which is not available on this pool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
total_bad = 0
for S, L, n in ((32, 30, 6000), (128, 30, 1500), (256, 50, 600), (200, 10, 700)):
    dims = mrvi_b200.MrviDims(n_input=64, n_sample=S, n_latent=L, n_latent_u=(None if L == 10 else 10))
    flat = mrvi_b200.init_params(dims, seed=3)
    h = _capi.Handle(dims, flat, device=0, precision=_capi.PRECISION_TC)
    X = mrvi_b200.synthetic_counts(n, 64, seed=4)
    rng = np.random.default_rng(1)
    sid = rng.integers(0, S, n).astype(np.int32)
    codes = rng.integers(0, 7, n).astype(np.int32)
    ref = h.run_host(X, sid, [codes], [7], keep_cell=True, want_z=True)
    bad = 0
    for it in range(25):
        r = h.run_host(X, sid, [codes], [7], keep_cell=True, want_z=True, chunk_cells=[0, 333, 1024][it % 3])
        if not (np.array_equal(r["cell"], ref["cell"]) and np.array_equal(r["z"], ref["z"])):
            bad += 1
            if bad <= 3:
                dc = np.abs(r["cell"] - ref["cell"]); dz = np.abs(r["z"] - ref["z"])
                cells = np.argwhere(dc.reshape(n, -1).max(1) > 0).ravel()
                print(f"  it={it} chunk={[0, 333, 1024][it % 3]}: cell max diff {dc.max():.3e} in {len(cells)} cells {cells[:6]}, z max diff {dz.max():.3e} in {int((dz.reshape(n,-1).max(1)>0).sum())} cells; scale {ref['cell'].max():.3f}")
        # group sums: fp32 run accumulators whose run boundaries depend on scheduling -> equal to fp32 rounding only
        if not np.allclose(r["group_sums"][0], ref["group_sums"][0], rtol=1e-5, atol=1e-5 * ref["cell"].max()):
            bad += 1
    print(f"S={S} L={L}: mismatching iterations {bad} of 25", flush=True)
    total_bad += bad
    h.close()
print("TOTAL", total_bad)
