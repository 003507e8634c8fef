"""Quick check of an opt-in fused kernel variant (MRVI_B200_FUSED=<n>, default 4 = tc_fused4.cuh) against the second version
and the oracle, small shapes first (a deadlock shows up in seconds: run under `timeout`), then a many-tiles-per-CTA timing run.
  timeout 120 python tools/v4_smoke.py [variant]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
from oracle import mrvi_oracle as orc


def handle(dims, params, v):
    if v == NEW:
        os.environ["MRVI_B200_FUSED"] = str(NEW)
    else:
        os.environ.pop("MRVI_B200_FUSED", None)
    return _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)


NEW = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cases = [(64, 50, 100, 64), (8, 40, 257, 64), (32, 30, 300, 200), (128, 30, 150, 64), (48, 30, 150, 64), (4, 30, 400, 64), (15, 10, 333, 64), (100, 16, 90, 40), (256, 50, 12, 64), (200, 10, 9, 64)]
for S, L, n, G in cases:
    dims = mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=L, n_latent_u=10 if L != 10 else None)
    params = mrvi_b200.init_params(dims, seed=4)
    X = mrvi_b200.synthetic_counts(n, G, seed=6)
    rng = np.random.default_rng(3)
    sid = rng.integers(0, S, n).astype(np.int32)
    codes = [np.minimum(rng.zipf(1.6, size=n) - 1, 3).astype(np.int32)]
    dref = orc.pairwise_distances(orc.counterfactual_z(params, X, sid, dtype=np.float64))
    gref = np.stack([dref[codes[0] == g].sum(0) for g in range(4)])
    out = {}
    for v in (NEW, 2):
        h = handle(dims, params, v)
        t0 = time.time()
        r = h.run_host(X, sid, codes, [4], keep_cell=True, want_z=True)
        g = h.run_host(X, sid, codes, [4], keep_cell=False)
        out[v] = (r, g)
        h.close()
        err = float((np.abs(r["cell"] - dref) / dref.max(axis=(1, 2), keepdims=True)).max())
        eg = float(np.abs(g["group_sums"][0] - gref).max() / np.abs(gref).max())
        print(f"S={S} L={L} n={n} v{v}: cell err {err:.2e} group-only err {eg:.2e} ({time.time() - t0:.2f}s)", flush=True)
    d32 = float(np.abs(out[NEW][0]["cell"] - out[2][0]["cell"]).max())
    dz = float(np.abs(out[NEW][0]["z"] - out[2][0]["z"]).max())
    print(f"   new vs v2: max |dD| {d32:.2e}  max |dz| {dz:.2e}", flush=True)

# many tiles per CTA: timing of the fused stage (C2 / C4-shard shapes)
for S, n, ng in ((32, 100000, 20), (128, 20000, 50)):
    G = 256
    dims = mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=30)
    params = mrvi_b200.init_params(dims, seed=0)
    X = mrvi_b200.synthetic_counts(n, G, seed=1)
    rng = np.random.default_rng(5)
    sid = rng.integers(0, S, n).astype(np.int32)
    codes = [rng.integers(0, ng, n).astype(np.int32)]
    res = {}
    for v in (NEW, 2):
        h = handle(dims, params, v)
        for _ in range(3):
            r = h.run_host(X, sid, codes, [ng], keep_cell=False, chunk_cells=n)
        ms = h.last_stage_ms()
        res[v] = r["group_sums"][0]
        print(f"S={S} n={n} groupby {ng} v{v}: stage ms {ms}", flush=True)
        if S == 128:
            for _ in range(2):
                r = h.run_host(X, sid, [], [], keep_cell=True, chunk_cells=n)
            print(f"S={S} n={n} keep_cell v{v}: stage ms {h.last_stage_ms()}", flush=True)
        h.close()
    print(f"   group sums new vs v2: rel {np.abs(res[NEW] - res[2]).max() / np.abs(res[2]).max():.2e}", flush=True)
print("ok")
