"""clock64 timeline of k_chain_fused_tc (CTA 0, thread 0 of each warpgroup).  Build the debug library first:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DMRVI_DBG_FUSED \
       -o tools/dbg/libdbg_fused.so mrvi_b200/csrc/mrvi_b200.cu mrvi_b200/csrc/host_pack.cpp
  python tools/dbg/run_dbgf.py [S] [keep_cell 0/1] [n_groups]"""
import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["MRVI_B200_LIBRARY"] = os.environ.get("MRVI_B200_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg_fused.so")
import mrvi_b200
from mrvi_b200 import _capi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
keep = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
ng = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dims = mrvi_b200.MrviDims(n_input=200, n_sample=S)
p = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, p, precision=_capi.PRECISION_TC)
n = (128 // S) * 148 * 2 * 16
rng = np.random.default_rng(0)
X = mrvi_b200.synthetic_counts(n, 200, 0); sid = rng.integers(0, S, n).astype(np.int32); codes = [rng.integers(0, ng, n).astype(np.int32)]
for _ in range(2):
    h.run_host(X, sid, [] if keep else codes, [] if keep else [ng], keep_cell=keep, chunk_cells=n)
lib = _capi.load_library()
buf = (ctypes.c_longlong * 2048)()
lib.mrvi_dbg_fused_read(buf, 2048)
a = np.array(buf[:2048]).reshape(2, 32, 32)
names = ["att", "sync", "G1", "wait1", "ep1", "sync", "G2iss", "wait2", "ep2", "sync", "G3iss", "wait3", "ep3", "sync", "G4iss", "wait4", "ep4",
         "sync", "G5iss", "wait5", "d5", "sync", "split", "sync", "Gram_iss", "waitG", "bookkeep", "dist", "unlock+sync"]
print("S", S, "keep_cell", keep, "groups", ng, "cells", n)
for wg in range(2):
    for t in range(4, 8):
        row = a[wg, t]
        d = [int(row[i + 1] - row[i]) for i in range(29)]
        print(f"wg{wg} tile{t} total {int(row[29] - row[0])} next_start {int(a[wg, t + 1, 0] - row[0])}:", " ".join(f"{names[i]}={d[i]}" for i in range(29)))
m = np.diff(a[:, 4:12, :30].astype(np.int64), axis=2).mean(axis=(0, 1))
print("mean over tiles 4-11:", " ".join(f"{names[i]}={m[i]:.0f}" for i in range(29)), "| total", m.sum())
