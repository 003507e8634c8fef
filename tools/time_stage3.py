"""Stage 3 alone (`mrvi_distances_device`, tensor-core mode): S x S blocks written per second = the keep_cell WRITE PATH
without the chain in front of it (north_star: ">= 70 % of HBM peak on the keep_cell write path")."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
for S, L, n in ((128, 30, 60000), (256, 50, 15000), (64, 30, 200000)):
    dims = mrvi_b200.MrviDims(n_input=32, n_sample=S, n_latent=L)
    h = _capi.Handle(dims, mrvi_b200.init_params(dims, seed=0), device=0, precision=_capi.PRECISION_TC)
    z = (torch.randn((n, 1, L), device="cuda") + 0.3 * torch.randn((n, S, L), device="cuda")).contiguous()
    out = torch.empty((n, S, S), dtype=torch.float32, device="cuda")
    for _ in range(3):
        h.distances_device(n, z.data_ptr(), cell_out_ptr=out.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        h.distances_device(n, z.data_ptr(), cell_out_ptr=out.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    wr = n * S * S * 4 / (ms * 1e-3) / 1e9
    rd = n * S * L * 4 / (ms * 1e-3) / 1e9
    print(f"S={S} L={L}: {n} cells in {ms:.3f} ms = {n/ms/1e3:.1f} M cells/s; written {wr:.0f} GB/s (+ read {rd:.0f} GB/s) of 6544 GB/s measured peak = {100*(wr+rd)/6544:.0f} %")
    h.close()
