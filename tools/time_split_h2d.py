"""mrvi_run_host from page-locked float32 input: chunk rows split between host narrowing and raw DMA (default) against
all-narrowed (MRVI_B200_NO_SPLIT_H2D=1) and all-raw (MRVI_B200_NO_HOST_PACK=1, separate process).  Checks that the split is
bit-transparent, then times the C4 shard.   python tools/time_split_h2d.py   [MRVI_B200_PCIE_GBPS=<assumed raw rate>]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mrvi_b200
from mrvi_b200 import _capi


def toggle(split):
    if split: os.environ.pop("MRVI_B200_NO_SPLIT_H2D", None)
    else: os.environ["MRVI_B200_NO_SPLIT_H2D"] = "1"


# ---- bit transparency (keep_cell output is deterministic)
dims = mrvi_b200.MrviDims(n_input=2000, n_sample=32)
params = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)
n = 40_000
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, 2000, seed=5)).pin_memory().numpy()
sid = np.random.default_rng(1).integers(0, 32, n).astype(np.int32)
outs = []
for split in (False, True, True):
    toggle(split)
    r = h.run_host(X, sid, [], [], keep_cell=True)
    outs.append((r["cell"].copy(), h.last_h2d_bytes()))
print("bit-identical:", np.array_equal(outs[0][0], outs[2][0]), "| h2d MB all-narrowed", outs[0][1] / 1e6, "split", outs[2][1] / 1e6, flush=True)
h.close()

# ---- C4 shard
n, G, S, ng = 125_000, 2000, 128, 50
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
params = mrvi_b200.init_params(dims, seed=0)
h = _capi.Handle(dims, params, device=0, precision=_capi.PRECISION_TC)
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, G, seed=1000)).pin_memory().numpy()
rng = np.random.default_rng(1)
sid = rng.integers(0, S, n).astype(np.int32)
codes = rng.integers(0, ng, n).astype(np.int32)
for rep in range(2):
    for split in (False, True):
        toggle(split)
        for _ in range(3):
            h.run_host(X, sid, [codes], [ng], keep_cell=False)
        t = time.perf_counter()
        for _ in range(8):
            h.run_host(X, sid, [codes], [ng], keep_cell=False)
        dt = (time.perf_counter() - t) / 8
        print(f"C4 shard split={int(split)} pcie_gbps={os.environ.get('MRVI_B200_PCIE_GBPS', '47')} no_pack={os.environ.get('MRVI_B200_NO_HOST_PACK', '0')}: "
              f"{dt * 1e3:7.2f} ms  {n / dt / 1e6:6.2f} M cells/s  h2d {h.last_h2d_bytes() / 1e6:.0f} MB", flush=True)
