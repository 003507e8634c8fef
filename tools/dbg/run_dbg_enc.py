"""clock64 timeline of k_encoder_tc<true> (CTA 0, K blocks 128..191 of its stream): build with
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -DMRVI_DBG_ENC -o tools/dbg/libdbg_enc.so mrvi_b200/csrc/mrvi_b200.cu mrvi_b200/csrc/host_pack.cpp"""
import ctypes, numpy as np, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mrvi_b200
from mrvi_b200 import _capi
_capi.library_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbg_enc.so")
G, n, S = 2000, 100000, 32
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
h = _capi.Handle(dims, mrvi_b200.init_params(dims, seed=0), precision=_capi.PRECISION_TC)
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, G, seed=4)).cuda()
sid = torch.zeros(n, dtype=torch.int32, device="cuda")
u = torch.empty((n, dims.n_latent_q), dtype=torch.float32, device="cuda")
for _ in range(3):
    h.run_device(n, X.data_ptr(), G, sid.data_ptr(), u_out_ptr=u.data_ptr())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
_capi.load_library().mrvi_dbg_enc_read(buf, 1024)
a = np.array(buf[:1024]).reshape(64, 16)
names = ["c_top", "c_xfull", "c_conv", "c_xrel", "c_aempty", "c_afull", "m_afull", "m_wfull", "m_issued", "l_top", "l_xempty"]
t0 = a[0, 0]
for g in range(8, 20):
    print(f"g={128+g}: " + " ".join(f"{names[k]}={int(a[g,k]-t0)}" for k in range(11)))
print("per-K-block period (c_top):", [int(a[g+1,0]-a[g,0]) for g in range(2, 40)])
