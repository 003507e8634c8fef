"""End-to-end rate of mrvi_run_host on the C4 shard as a function of the pipeline chunk (cells per chunk)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mrvi_b200
from mrvi_b200 import _capi

n, G, S, ng = 125_000, 2000, 128, 50
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
params = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, G, seed=1000)).pin_memory().numpy()
rng = np.random.default_rng(1)
sid = rng.integers(0, S, n).astype(np.int32)
codes = rng.integers(0, ng, n).astype(np.int32)
X8 = torch.from_numpy(X.astype(np.uint8)).pin_memory().numpy()
for name, A in (("float32", X), ("uint8", X8)):
    for ch in (0, 4096, 8192, 12288, 16384, 24576, 32768):
        for _ in range(2):
            h.run_host(A, sid, [codes], [ng], keep_cell=False, chunk_cells=ch)
        t = time.perf_counter()
        for _ in range(5):
            h.run_host(A, sid, [codes], [ng], keep_cell=False, chunk_cells=ch)
        dt = (time.perf_counter() - t) / 5
        print(f"{name} chunk {ch:6d}: {dt * 1e3:7.2f} ms  {n / dt / 1e6:6.2f} M cells/s  h2d {h.last_h2d_bytes() / 1e6:.0f} MB", flush=True)
