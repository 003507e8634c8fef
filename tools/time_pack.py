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
for nt in (1, 2, 4, 8, 12, 16, 24, 32):
    best = 0
    for _ in range(5):
        t = time.perf_counter()
        m = lib.mrvi_narrow_counts(X.ctypes.data, 2000, 100000, 2000, out.ctypes.data, nt)
        dt = time.perf_counter() - t
        best = max(best, X.nbytes / dt / 1e9)
    print("threads", nt, "mode", m, f"best {best:.1f} GB/s", "cpus", os.cpu_count())

# reference points: plain multi-threaded reads of the same matrix (numpy sum over row blocks in a thread pool) and lscpu
from concurrent.futures import ThreadPoolExecutor
for nt in (1, 4, 8, 16):
    blocks = np.array_split(np.arange(X.shape[0]), nt * 4)
    with ThreadPoolExecutor(nt) as ex:
        t = time.perf_counter(); list(ex.map(lambda b: float(X[b[0]:b[-1] + 1].sum()), blocks)); dt = time.perf_counter() - t
    print("numpy sum, threads", nt, f"{X.nbytes / dt / 1e9:.1f} GB/s")
os.system("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz' ; numactl -H 2>/dev/null | head -8; nproc")
