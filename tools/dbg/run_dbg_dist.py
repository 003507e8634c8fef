"""clock64 timeline of k_dist_gram_tc (CTA 0): build with
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DMRVI_DBG_DIST -o tools/dbg/libdbg_dist.so mrvi_b200/csrc/mrvi_b200.cu"""
import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mrvi_b200
from mrvi_b200 import _capi
_capi.library_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg_dist.so")
S, L = 256, 50
G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NG = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=L)
p = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, p, precision=_capi.PRECISION_TC)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 20
X = mrvi_b200.synthetic_counts(n, G, 0); rng = np.random.default_rng(0); sid = rng.integers(0, S, n).astype(np.int32); codes = [rng.integers(0, NG, n).astype(np.int32)]
h.run_host(X, sid, codes, [NG], keep_cell=False)
lib = _capi.load_library()
buf = (ctypes.c_longlong * 2048)()
lib.mrvi_dbg_dist_read(buf, 2048)
a = np.array(buf[:2048]).reshape(64, 32)
names = {0: "iss_b0_go", 1: "iss_b0_done", 2: "iss_b1_go", 3: "iss_b1_done", 4: "iss_b2_go", 5: "iss_b2_done", 8: "w_top", 9: "w_stage", 10: "w_prepdone",
         11: "w_b0_wait", 12: "w_b0_got", 13: "w_b0_arr", 14: "w_b1_wait", 15: "w_b1_got", 16: "w_b1_arr", 17: "w_b2_wait", 18: "w_b2_got", 19: "w_b2_arr",
         20: "w_end", 21: "w_synced", 22: "w_looptop", 23: "w_flushed"}
print("tmin<0 chunk events:", int(buf[2046]), "exact recomputes:", int(buf[2047]), "of", n * S * S, "entries")
tops = a[:, 8]
print("per-cell top-to-top:", [int(tops[i+1]-tops[i]) for i in range(2, 40) if tops[i+1] > 0])
print("flush durations:", [int(a[i,23]-a[i,10]) for i in range(2, 40) if a[i,23] > 0])
for c in range(3, 6):
    t0 = a[c, 8]
    ev = sorted((int(a[c, k] - t0), v) for k, v in names.items() if a[c, k])
    print(f"cell {c}: next_top={int(a[c+1,8]-t0)}", " ".join(f"{v}={t}" for t, v in ev))

full = np.array(buf[:2048])
for c in range(4, 8):
    ends = full[1024 + c * 16: 1024 + c * 16 + 16] - a[c, 8]
    preps = full[1536 + c * 16: 1536 + c * 16 + 16] - a[c, 8]
    print(f"cell {c} per-warp prep_done:", [int(v) for v in preps])
    print(f"cell {c} per-warp end:      ", [int(v) for v in ends])

print("exact calls in cell 5 (dt, i, j, t0-top):", [(int(full[2000+4*k]), int(full[2001+4*k]), int(full[2002+4*k]), int(full[2003+4*k]-a[5,8])) for k in range(int(min(full[2040], 8)))])
print("counters:", int(full[2046]), int(full[2047]))
