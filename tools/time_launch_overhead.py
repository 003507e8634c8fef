"""Device-resident time of mrvi_run_device on a C4-shaped shard as a function of the number of cells per call and of the
number of groups: t(n) = a + b n separates the per-launch cost (prologue, accumulator flush atomics, sort) from the rate."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mrvi_b200
from mrvi_b200 import _capi

G, S = 2000, 128
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
params = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)
dev = torch.device("cuda", 0)
nmax = 125_000
X = torch.from_numpy(mrvi_b200.synthetic_counts(nmax, G, seed=1000)).to(dev)
rng = np.random.default_rng(1)
sid = torch.from_numpy(rng.integers(0, S, nmax).astype(np.int32)).to(dev)
st = torch.cuda.current_stream(dev)
for ng in (50, 1):
    codes = torch.from_numpy(rng.integers(0, ng, nmax).astype(np.int32)).to(dev)
    sums = torch.zeros((ng, S, S), dtype=torch.float64, device=dev)
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 125000):
        def run():
            h.run_device(n, X.data_ptr(), G, sid.data_ptr(), [codes.data_ptr()], [ng], "l2", 0, [sums.data_ptr()], 0, 0, st.cuda_stream)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(50, 400000 // n))
        e0.record(st)
        for _ in range(reps):
            run()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        stg = h.last_stage_ms()
        print(f"groups {ng:3d} n {n:7d}: {ms * 1e3:9.1f} us per call, {ms * 1e6 / n:7.1f} ns/cell, stages {dict((k, round(v * 1e3)) for k, v in stg.items())} us", flush=True)
