"""End-to-end rate with adata.X stored as uint8 counts (mrvi_run_host_counts) vs float32 (mrvi_run_host), pinned host memory."""
import numpy as np, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200, torch
from mrvi_b200 import _capi
dims = mrvi_b200.MrviDims(n_input=2000, n_sample=32)
h = _capi.Handle(dims, mrvi_b200.init_params(dims, seed=0), device=0, precision=_capi.PRECISION_TC)
n = 100000
X = mrvi_b200.synthetic_counts(n, 2000, seed=1)
sid = np.random.default_rng(0).integers(0, 32, n).astype(np.int32); codes = [np.random.default_rng(1).integers(0, 20, n).astype(np.int32)]
for name, arr in (("uint8", X.astype(np.uint8)), ("uint16", X.astype(np.uint16)), ("float32", X)):
    t_ = torch.from_numpy(arr).pin_memory(); Xp = t_.numpy()
    for chunk in (0, 8192, 32768):
        for _ in range(2): h.run_host(Xp, sid, codes, [20], keep_cell=False, chunk_cells=chunk)
        t = time.perf_counter()
        for _ in range(5): h.run_host(Xp, sid, codes, [20], keep_cell=False, chunk_cells=chunk)
        dt = (time.perf_counter() - t) / 5
        print(f"{name:8s} chunk={chunk:6d}: {dt*1e3:6.2f} ms per 100k cells = {n/dt/1e6:5.1f} M cells/s e2e, wire {h.last_h2d_bytes()/1e6:.0f} MB", flush=True)
