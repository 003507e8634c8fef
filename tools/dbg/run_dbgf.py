import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mrvi_b200
from mrvi_b200 import _capi
_capi.library_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg.so")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dims = mrvi_b200.MrviDims(n_input=200, n_sample=S)
p = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, p, precision=_capi.PRECISION_TC)
n = (128 // S) * 148 * 2 * 8
X = mrvi_b200.synthetic_counts(n, 200, 0); sid = np.zeros(n, np.int32); codes = [np.zeros(n, np.int32)]
h.run_host(X, sid, codes, [3], keep_cell=False)
lib = _capi.load_library()
buf = (ctypes.c_longlong * 2048)()
lib.mrvi_dbg_read(buf, 2048)
a = np.array(buf[:2048]).reshape(2, 32, 32)
names = ["start","attn_done","sync","G1_done","ep1_done","sync","G2_done","ep2_done","sync","G3_done","ep3_done","sync","G4_done","ep4_done","sync","G5_done","z_done","sync","ctr_done","Gram_done","bookkeep","dist_done","sync"]
for wg in range(2):
    for t in range(1, 4):
        row = a[wg, t]
        d = [int(row[i+1]-row[i]) for i in range(22)]
        print(f"wg{wg} tile{t} total {int(row[22]-row[0])}:", " ".join(f"{names[i+1]}={d[i]}" for i in range(22)))
