"""Wall-clock of the sampled paths on the C2 shape (device Philox noise, host buffers in / out).
usage: python tools/time_sampled.py [n_cells] [mc_samples]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
mc = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for S, G in ((32, 2000), (128, 2000)):
    dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
    flat = mrvi_b200.init_params(dims, seed=0)
    X = mrvi_b200.synthetic_counts(n, G, seed=1)
    rng = np.random.default_rng(2)
    sid = rng.integers(0, S, n).astype(np.int32)
    codes = rng.integers(0, 20, n).astype(np.int32)
    for prec, name in ((_capi.PRECISION_TC, "tc"), (_capi.PRECISION_FP32, "fp32")):
        if name == "fp32" and S > 32:
            continue
        h = _capi.Handle(dims, flat, device=0, precision=prec)
        for normalize in (False, True):
            h.run_sampled_host(X[:2048], sid[:2048], mc, group_codes=[codes[:2048]], n_groups=[20], normalize=normalize, keep_cell=False)
            t = time.perf_counter()
            h.run_sampled_host(X, sid, mc, group_codes=[codes], n_groups=[20], normalize=normalize, keep_cell=False)
            dt = time.perf_counter() - t
            print(f"S={S} {name} normalize={normalize} mc={mc}: {n} cells in {dt*1e3:.1f} ms = {n/dt/1e3:.1f} k cells/s "
                  f"({n*mc*S/dt/1e6:.1f} M (cell,draw,sample) rows/s)", flush=True)
        t = time.perf_counter()
        h.run_host(X, sid, [codes], [20], keep_cell=False)
        dt = time.perf_counter() - t
        print(f"S={S} {name} mean path: {n/dt/1e3:.1f} k cells/s", flush=True)
        h.close()
