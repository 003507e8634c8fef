"""clock64 timeline of k_chain_fused4_tc (CTA 0).  Build the debug library first:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DMRVI_DBG_F4 \
       -o tools/dbg/libdbg_f4.so mrvi_b200/csrc/mrvi_b200.cu mrvi_b200/csrc/host_pack.cpp
  python tools/dbg/run_dbg_f4.py [S] [keep_cell 0/1]"""
import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["MRVI_B200_FUSED"] = "4"
import mrvi_b200
from mrvi_b200 import _capi
_capi.library_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg_f4.so")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
keep = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dims = mrvi_b200.MrviDims(n_input=200, n_sample=S)
p = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, p, precision=_capi.PRECISION_TC)
n = (128 // S) * 148 * 2 * 16
X = mrvi_b200.synthetic_counts(n, 200, 0); sid = np.zeros(n, np.int32); codes = [np.zeros(n, np.int32)]
for _ in range(2):
    h.run_host(X, sid, [] if keep else codes, [] if keep else [3], keep_cell=keep, chunk_cells=n)
lib = _capi.load_library()
N = 2 * 16 * 7 * 2 * 4 + 16 * 6 * 2 * 2
buf = (ctypes.c_longlong * N)()
lib.mrvi_dbg_f4_read(buf, N)
a = np.array(buf[:N])
ep = a[:1792].reshape(2, 16, 7, 2, 4); iss = a[1792:].reshape(16, 6, 2, 2)
names = ["att", "ep1", "ep2", "ep3", "ep4", "d5", "dist"]
for w in range(2):
    for pr in range(4, 8):
        t0 = ep[w, pr, 0, 0, 0]
        print(f"warp{15 * w} pair{pr} total {int(ep[w, pr + 1, 0, 0, 0] - t0)}")
        for ph in range(7):
            print("   " + "  ".join(f"{names[ph]}({'AB'[sl]}): @{int(ep[w, pr, ph, sl, 0] - t0):6d} wait {int(ep[w, pr, ph, sl, 1] - ep[w, pr, ph, sl, 0]):5d} work {int(ep[w, pr, ph, sl, 2] - ep[w, pr, ph, sl, 1]):5d}" for sl in range(2)))
print("issuer: per (pair, step, slot): ready seen @ (relative to warp 0's pair start), issue duration")
for pr in range(4, 7):
    t0 = ep[0, pr, 0, 0, 0]
    for st in range(6):
        print(f"  pair{pr} G{st + 1}: " + "  ".join(f"{'AB'[s]} @{int(iss[pr, st, s, 0] - t0):6d} issue {int(iss[pr, st, s, 1] - iss[pr, st, s, 0]):5d}" for s in range(2)))
