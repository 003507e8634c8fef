"""Host narrowing throughput (float32 bytes consumed per second) on this machine."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
import ctypes as C
X = mrvi_b200.synthetic_counts(100000, 2000, seed=0)
out = np.empty(X.size * 2, dtype=np.uint8); out[:] = 0   # pre-faulted staging, as in the library
lib = _capi.load_library()
for nt in (8, 16, 32):
    best = 0
    for _ in range(5):
        t = time.perf_counter()
        m = lib.mrvi_narrow_counts(X.ctypes.data, 2000, 100000, 2000, out.ctypes.data, nt)
        dt = time.perf_counter() - t
        best = max(best, X.nbytes / dt / 1e9)
    print("threads", nt, "mode", m, f"best {best:.1f} GB/s", "cpus", os.cpu_count())
