import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mrvi_b200
from mrvi_b200 import _capi
_capi.library_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg.so")
dims = mrvi_b200.MrviDims(n_input=2000, n_sample=32)
p = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, p, precision=_capi.PRECISION_TC)
n = 128 * 148 * 5
X = mrvi_b200.synthetic_counts(n, 2000, 0); sid = np.zeros(n, np.int32)
for _ in range(1): h.run_host(X, sid, keep_cell=False, want_u=True, chunk_cells=n)
lib = _capi.load_library()
buf = (ctypes.c_longlong * 512)()
lib.mrvi_dbg_read(buf, 512)
a = np.array(buf[:512]); t0 = a[0]
print("producer tile starts:", [int(v - t0) for v in a[0:6]])
print("mma1 (before acc_free wait, after):", [(int(a[16+2*i]-t0), int(a[17+2*i]-t0)) for i in range(6)])
for ti in range(5):
    b = a[64+ti*32: 64+ti*32+32]
    print(f"tile {ti}: ep wait_start {int(b[0]-t0)} acc_full {int(b[1]-t0)} ep0_done {int(b[2]-t0)} | layers (d_ready, done):",
          [(int(b[4+2*l]-t0), int(b[5+2*l]-t0)) for l in range(5)], "head_done", int(b[20]-t0))
