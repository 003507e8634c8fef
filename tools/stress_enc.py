"""Repeatability stress of the TC encoder on DEVICE-resident input: the same X many times must give bit-identical u."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
dims = mrvi_b200.MrviDims(n_input=G, n_sample=8)
flat = mrvi_b200.init_params(dims, seed=3)
h = _capi.Handle(dims, flat, device=0, precision=_capi.PRECISION_TC)
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, G, seed=4)).cuda()
sid = torch.from_numpy(np.random.default_rng(1).integers(0, 8, n).astype(np.int32)).cuda()
u = torch.empty((n, dims.n_latent_q), dtype=torch.float32, device="cuda")
def run():
    h.run_device(n, X.data_ptr(), G, sid.data_ptr(), u_out_ptr=u.data_ptr(), z_out_ptr=0)
    torch.cuda.synchronize()
    return u.cpu().numpy().copy()
ref = run()
bad = 0
for it in range(40):
    v = run()
    if not np.array_equal(v, ref):
        bad += 1
        d = np.argwhere((v != ref).any(axis=1)).ravel()
        if bad <= 4: print("iteration", it, "differs in", len(d), "cells, first", d[:6], "max abs", np.abs(v - ref).max())
print("mismatching iterations:", bad, "of 40")
