"""clock64 timeline of k_chain_fused3_tc (CTA 0).  Build the debug library first:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DMRVI_DBG_F3 \
       -o tools/dbg/libdbg_f3.so mrvi_b200/csrc/mrvi_b200.cu mrvi_b200/csrc/host_pack.cpp
  python tools/dbg/run_dbg_f3.py [S] [keep_cell 0/1]"""
import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mrvi_b200
from mrvi_b200 import _capi
_capi.library_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg_f3.so")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
keep = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dims = mrvi_b200.MrviDims(n_input=200, n_sample=S)
p = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, p, precision=_capi.PRECISION_TC)
n = (128 // S) * 148 * 2 * 10
X = mrvi_b200.synthetic_counts(n, 200, 0); sid = np.zeros(n, np.int32); codes = [np.zeros(n, np.int32)]
for _ in range(2):
    h.run_host(X, sid, [] if keep else codes, [] if keep else [3], keep_cell=keep, chunk_cells=n)
lib = _capi.load_library()
N = 2 * 16 * 32 + 2 * 16 * 16 + 4 * 64 * 2
buf = (ctypes.c_longlong * N)()
lib.mrvi_dbg_f3_read(buf, N)
a = np.array(buf[:N])
ep = a[:1024].reshape(2, 16, 32); iss = a[1024:1536].reshape(2, 16, 8, 2); pr = a[1536:].reshape(4, 64, 2)
t0 = ep[0, 0, 0]
names = ["A3static+E0", "wait G1", "ep1", "wait G2", "ep2", "wait G3", "ep3", "wait G4", "ep4", "wait G5", "z+operands+E5", "wait Gram", "bookkeep/lock", "dist", "release+sync"]
for wg in range(2):
    for t in range(2, 6):
        row = ep[wg, t]
        d = [int(row[i + 1] - row[i]) for i in range(15)]
        print(f"wg{wg} tile{t} start {int(row[0]-t0)} total {int(row[15]-row[0])}: " + " ".join(f"{names[i]}={d[i]}" for i in range(15)))
print("issuer (per slot, tile, step): ready->issued clk, and time since tile start of the epilogue warpgroup")
for s in range(2):
    for t in range(2, 5):
        print(f" slot{s} tile{t}: " + " ".join(f"G{st+1}:{int(iss[s,t,st,1]-iss[s,t,st,0])}@{int(iss[s,t,st,0]-ep[s,t,0])}" for st in range(6)))
print("producers: task durations (clk) and gaps")
for w in range(3):
    d = [int(pr[w, t, 1] - pr[w, t, 0]) for t in range(4, 14)]
    g = [int(pr[w, t + 1, 0] - pr[w, t, 1]) for t in range(4, 14)]
    print(f" warp{w}: dur {d} gap {g}")
