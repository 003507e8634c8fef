"""Encoder-only timing (run_device with u_out only) vs the encoder stage inside a full step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
G, n, S = 2000, 100000, 32
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
flat = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, flat, device=0, precision=_capi.PRECISION_TC)
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, G, seed=4)).cuda()
sid = torch.from_numpy(np.random.default_rng(1).integers(0, S, n).astype(np.int32)).cuda()
codes = torch.from_numpy(np.random.default_rng(2).integers(0, 20, n).astype(np.int32)).cuda()
sums = torch.zeros((20, S, S), dtype=torch.float64, device="cuda")
u = torch.empty((n, dims.n_latent_q), dtype=torch.float32, device="cuda")
def enc_only():
    h.run_device(n, X.data_ptr(), G, sid.data_ptr(), u_out_ptr=u.data_ptr())
def full():
    h.run_device(n, X.data_ptr(), G, sid.data_ptr(), [codes.data_ptr()], [20], group_sum_ptrs=[sums.data_ptr()])
for name, fn in (("encoder only", enc_only), ("full step", full), ("encoder only", enc_only)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 20, 4), "ms per call; last stage ms:", h.last_stage_ms())
