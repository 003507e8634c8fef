"""Does a kernel run faster when launched in isolation than in a back-to-back loop?  (ncu reports 1.13 ms for the fused
kernel of the C2 step and 0.28 ms for the encoder; CUDA events inside the bench loop say 1.74 / 0.43.)"""
import os, sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mrvi_b200
from mrvi_b200 import _capi
G, n, S = 2000, 100000, 32
dims = mrvi_b200.MrviDims(n_input=G, n_sample=S)
h = _capi.Handle(dims, mrvi_b200.init_params(dims, seed=0), device=0, precision=_capi.PRECISION_TC)
X = torch.from_numpy(mrvi_b200.synthetic_counts(n, G, seed=4)).cuda()
sid = torch.from_numpy(np.random.default_rng(1).integers(0, S, n).astype(np.int32)).cuda()
codes = torch.from_numpy(np.random.default_rng(2).integers(0, 20, n).astype(np.int32)).cuda()
sums = torch.zeros((20, S, S), dtype=torch.float64, device="cuda")
def full():
    h.run_device(n, X.data_ptr(), G, sid.data_ptr(), [codes.data_ptr()], [20], group_sum_ptrs=[sums.data_ptr()])
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
for _ in range(3): full()
torch.cuda.synchronize()
print("idle:", clocks())
iso = []
for i in range(8):
    time.sleep(0.2)
    full(); torch.cuda.synchronize()
    iso.append(h.last_stage_ms())
print("isolated launches (200 ms idle between):", [(round(d["encoder"], 3), round(d["qz_chain"], 3)) for d in iso])
for reps in (5, 50, 500):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): full()
    e1.record(); torch.cuda.synchronize()
    d = h.last_stage_ms()
    print(f"back-to-back x{reps}: {e0.elapsed_time(e1)/reps:.3f} ms per step; last step stages", round(d["encoder"], 3), round(d["qz_chain"], 3), "|", clocks())
