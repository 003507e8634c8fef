#!/usr/bin/env python
"""Benchmark of the counterfactual local-sample-distance path (BASELINE.json metric: cells/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|tc]
                  [--workload c2|c3|c4|c5] [--cells-per-gpu M] [--no-extras]

One "step" = one full pass of the hot path (encoder -> q(z|u,s) for all S samples -> S x S distances -> per-group
sums) over this rank's shard of the workload.  The default workload is the configuration BASELINE.json quotes its
target on: C4 = configs[3], 1 M cells x 2 k genes x 128 samples, groupby 50, sharded over the 8 GPUs of a box, i.e.
125 000 cells per GPU (weak scaling: every rank owns `cells-per-gpu` cells; at --gpus 8 the job IS the full C4).  The
per-group sums are combined by one NCCL all-reduce per step (6.5 MB of float64: `mrvi_allreduce_groups`).  At N = 1 the
other BASELINE configurations (C2, a C3 shard, a C5 shard) are measured as `extra` sub-lines of the same run.

`value`   : cells/s with X already resident in HBM (mrvi_run_device), CUDA-event timed, max over ranks.
`e2e`     : cells/s through the host-buffer C-ABI call (mrvi_run_host) from pinned host memory, H2D of X / sample /
            group codes and D2H of the group sums inside the timed region.
`e2e_api` : the same through the public Python API, `MrVI.get_local_sample_distances(groupby=...)` on an
            AnnData-like object (pandas group coding, value_counts, output assembly included); `e2e_csr`: with
            adata.X as a 90 %-sparse scipy CSR matrix.
`roofline`: the dominant kernel (fused q(z|u,s) chain + distances + group sums), algorithmic FLOPs / its
            CUDA-event duration, against the measured bf16 GEMM peak.
`parity`  : after the timed region, the output of the timed path is checked: a 512-cell stratified sample of per-cell
            blocks against the float64 oracle, and the group sums of the timed (groupby-only) launch against float64
            sums of the per-cell blocks over ALL cells.  The run fails (exit 1) if the bar is exceeded.
`cpu_baseline`: the torch-CPU restatement of the reference (oracle "port"; the JAX reference cannot be installed here) on
            a bounded sample of the same workload, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cells per GPU, genes, samples, n_latent, n_groups, keep_cell, description)
    "c2": (100_000, 2000, 32, 30, 20, False,
           "C2: 100k cells x 2k genes x 32 samples, n_latent=30, groupby 20 cell types (BASELINE.json configs[1])"),
    "c3": (62_500, 2000, 128, 30, 0, True,
           "C3 shard: 62.5k of 500k cells x 2k genes x 128 samples, keep_cell=True (configs[2] / 8 GPUs)"),
    "c4": (125_000, 2000, 128, 30, 50, False,
           "C4 shard: 125k of 1M cells x 2k genes x 128 samples, groupby 50 (configs[3] / 8 GPUs: the north_star target)"),
    "c5": (250_000, 5000, 256, 50, 50, False,
           "C5 shard: 250k of 2M cells x 5k genes x 256 samples, n_latent=50, groupby (configs[4] / 8 GPUs)"),
}
EXTRA_CELLS = {"c2": 100_000, "c3": 62_500, "c4": 125_000, "c5": 20_000}   # cells of the `extra` sub-lines
PARITY_BAR = {"tc": 1e-3, "fp32": 1e-5}


def flops_per_cell(G, S, L, Lu=10, H=128, R=128, E=16, C=4, nh=2, M=32):
    """SURVEY.md section 8(d) / BASELINE.md section 3."""
    Lq = Lu if Lu != L else L
    enc = G * H + H * H + 2 * (H * R + R * H) + H * Lq
    percell = Lq * E + (Lq * L if Lu != L else 0)
    att = E * 8 + 2 * E * E * nh * C + E * 8 * C
    mlp0 = 2 * (E * C) * R + R * M + M * E
    mlp1 = 2 * (Lq + E) * R + R * M + M * L
    chain = 2 * S * (att + mlp0 + mlp1)
    return {"total": 2 * (enc + percell) + chain + 3 * S * S * L, "chain": chain, "encoder": 2 * enc,
            "dist": 3 * S * S * L}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(workload, n_cells):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    (profiles/r2_fused_<workload>_ncu.json, written by tools/ncu_summary.py from the .ncu-rep raw page); None when no
    capture of this launch size is on file."""
    p = os.path.join(ROOT, "profiles", f"r2_fused_{workload}_ncu.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    if int(d.get("cells", -1)) != int(n_cells):
        return None, f"capture on file is of a {d.get('cells')}-cell launch"
    return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), os.path.basename(p)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(n, G, S, n_groups, seed):
    import mrvi_b200

    X = mrvi_b200.synthetic_counts(n, G, seed=seed)
    rng = np.random.default_rng(seed + 101)
    sid = rng.integers(0, S, size=n).astype(np.int32)
    codes = rng.integers(0, n_groups, size=n).astype(np.int32) if n_groups else None
    return X, sid, codes


def cpu_reference_run(params, G, S, n_groups, keep_cell, sample_cells, steps, warmup, seed=0, min_seconds=0.0):
    """Time the torch-CPU restatement (oracle port) on `sample_cells` cells per step, all host threads."""
    import torch

    from oracle.mrvi_oracle_torch import TorchOracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = TorchOracle(params)
    X, sid, codes = make_inputs(sample_cells, G, S, n_groups, seed)
    kw = dict(group_codes=[codes] if n_groups else [], n_groups=[n_groups] if n_groups else [], keep_cell=keep_cell)
    for _ in range(warmup):
        orc.local_sample_distances(X[:512], sid[:512], **({**kw, "group_codes": [codes[:512]]} if n_groups else kw))
    t0 = time.perf_counter()
    done = 0
    while done < steps or (min_seconds and time.perf_counter() - t0 < min_seconds and done < 200):
        orc.local_sample_distances(X, sid, **kw)
        done += 1
    dt = time.perf_counter() - t0
    return sample_cells * done / dt, cores, dt / done, done


def cpu_sample_cells(G, S, n_per):
    """Cells per CPU pass: ~2-4 s of work on 16 cores at any BASELINE shape (the torch port does ~33 k cells/s at S = 32,
    G = 2000); the timed loop repeats passes until >= 20 s so that the denominator of every ratio is stable."""
    return max(1024, min(n_per, int(40000 * 32 / S * 2000 / G) // 256 * 256))


class Job:
    """One workload on this rank's GPU: handle, device-resident inputs, the step functions."""

    def __init__(self, wl, n_per, rank, local_rank, world, precision, torch, dist, unique_id=None):
        import mrvi_b200
        from mrvi_b200 import _capi

        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank
        self.wl = wl
        _, self.G, self.S, self.L, self.n_groups, self.keep_cell, self.desc = WORKLOADS[wl]
        self.n = n_per
        self.dev = torch.device("cuda", local_rank)
        self.dims = mrvi_b200.MrviDims(n_input=self.G, n_sample=self.S, n_latent=self.L)
        self.params = mrvi_b200.init_params(self.dims, seed=0)
        if precision == "auto":
            try:
                self.h = _capi.Handle(self.dims, self.params, device=local_rank, precision=_capi.PRECISION_TC)
                precision = "tc"
            except _capi.MrviError as e:
                if e.status != -5:
                    raise
                self.h = _capi.Handle(self.dims, self.params, device=local_rank, precision=_capi.PRECISION_FP32)
                precision = "fp32"
        else:
            self.h = _capi.Handle(self.dims, self.params, device=local_rank,
                                  precision=_capi.PRECISION_TC if precision == "tc" else _capi.PRECISION_FP32)
        self.precision = precision
        self.own_comm = False
        if world > 1 and self.n_groups and unique_id is not None:
            self.h.comm_create(unique_id, world, rank)   # the library's own NCCL communicator (mrvi_allreduce_groups)
            self.own_comm = True
        self.X, self.sid, self.codes = make_inputs(self.n, self.G, self.S, self.n_groups, seed=1000 + rank)
        n, S, ng = self.n, self.S, self.n_groups
        self.Xp = torch.from_numpy(self.X).pin_memory()
        self.sidp = torch.from_numpy(self.sid).pin_memory()
        self.codesp = torch.from_numpy(self.codes).pin_memory() if ng else None
        self.Xd, self.sidd = self.Xp.to(self.dev), self.sidp.to(self.dev)
        self.codesd = self.codesp.to(self.dev) if ng else None
        self.sums_d = torch.zeros((ng, S, S), dtype=torch.float64, device=self.dev) if ng else None
        self.cell_d = torch.empty((n, S, S), dtype=torch.float32, device=self.dev) if self.keep_cell else None
        self.stream = torch.cuda.current_stream(self.dev)
        self.cell_h = _capi.pinned_empty((n, S, S)) if self.keep_cell else None

    def step_device(self):
        ng = self.n_groups
        if ng:
            self.sums_d.zero_()
        self.h.run_device(self.n, self.Xd.data_ptr(), self.G, self.sidd.data_ptr(),
                          [self.codesd.data_ptr()] if ng else [], [ng] if ng else [], "l2",
                          self.cell_d.data_ptr() if self.keep_cell else 0, [self.sums_d.data_ptr()] if ng else [], 0, 0,
                          self.stream.cuda_stream)
        if self.world > 1 and ng:
            if self.own_comm:
                self.h.allreduce_groups([self.sums_d.data_ptr()], [ng], stream=self.stream.cuda_stream)
            else:
                self.dist.all_reduce(self.sums_d)

    def step_host(self, X=None):
        ng = self.n_groups
        res = self.h.run_host(self.Xp.numpy() if X is None else X, self.sidp.numpy(), [self.codesp.numpy()] if ng else [],
                              [ng] if ng else [], keep_cell=self.keep_cell, cell_out=self.cell_h)
        if self.world > 1 and ng:
            t = self.torch.from_numpy(res["group_sums"][0]).to(self.dev)
            self.dist.all_reduce(t)
            res["group_sums"][0] = t.cpu().numpy()
        return res

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed_device(self, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            self.step_device()
        e1.record(self.stream)
        self.barrier()
        return e0.elapsed_time(e1) * 1e-3

    def timed_wall(self, fn, steps):
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return time.perf_counter() - t0

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity of the timed path (oracle = checker only)
    def parity(self):
        """(a) 512 stratified cells' S x S blocks from a keep_cell launch of the same kernels vs the float64 oracle;
        (b) group sums of the TIMED launch type (groupby only) vs float64 sums of the per-cell blocks over all cells."""
        from oracle import mrvi_oracle as orc
        from oracle.parity import distance_error_stats, stratified_indices

        torch = self.torch
        n, S, ng, G = self.n, self.S, self.n_groups, self.G
        out = {"bar": PARITY_BAR[self.precision], "sample_cells": 0}
        cell = self.cell_d
        sums_timed = None
        if ng:
            sums_timed = torch.zeros((ng, S, S), dtype=torch.float64, device=self.dev)
            self.h.run_device(n, self.Xd.data_ptr(), G, self.sidd.data_ptr(), [self.codesd.data_ptr()], [ng], "l2", 0,
                              [sums_timed.data_ptr()], 0, 0, self.stream.cuda_stream)
        if cell is None:
            # per-cell blocks in slabs (n x S x S float32 of the big workloads does not have to fit next to everything else)
            slab = max(1024, min(n, (8 << 30) // (S * S * 4)))
        else:
            slab = n
            self.h.run_device(n, self.Xd.data_ptr(), G, self.sidd.data_ptr(), [], [], "l2", cell.data_ptr(), [], 0, 0,
                              self.stream.cuda_stream)
        idx = stratified_indices(n, 512)
        got = np.empty((len(idx), S, S), dtype=np.float32)
        ref_sums = torch.zeros((ng, S * S), dtype=torch.float64, device=self.dev) if ng else None
        buf = cell if cell is not None else torch.empty((slab, S, S), dtype=torch.float32, device=self.dev)
        for lo in range(0, n, slab):
            hi = min(n, lo + slab)
            if cell is None:
                self.h.run_device(hi - lo, self.Xd[lo:hi].data_ptr(), G, self.sidd[lo:hi].data_ptr(), [], [], "l2", buf.data_ptr(),
                                  [], 0, 0, self.stream.cuda_stream)
            blk = buf[: hi - lo] if cell is None else buf[lo:hi]
            sel = np.nonzero((idx >= lo) & (idx < hi))[0]
            if len(sel):
                got[sel] = blk[torch.from_numpy(idx[sel] - (lo if cell is None else 0)).to(self.dev)].cpu().numpy()
            if ng:
                for a in range(0, hi - lo, 8192):
                    b = min(hi - lo, a + 8192)
                    ref_sums.index_add_(0, self.codesd[lo + a:lo + b].to(torch.int64), blk[a:b].reshape(-1, S * S).to(torch.float64))
        torch.cuda.synchronize(self.dev)
        zref = orc.counterfactual_z(self.params, self.X[idx], self.sid[idx], dtype=np.float64)
        st = distance_error_stats(got, orc.pairwise_distances(zref))
        out.update({"sample_cells": st["n_cells"], "max_norm_err": st["max_norm_err"], "p99_elementwise_rel": st["p99_elementwise_rel"],
                    "max_elementwise_rel": st["max_elementwise_rel"]})
        if ng:
            ref_sums = ref_sums.reshape(ng, S, S)
            out["group_sum_err"] = float((sums_timed - ref_sums).abs().max() / ref_sums.abs().max())
        out["ok"] = bool(out["max_norm_err"] < out["bar"] and out["p99_elementwise_rel"] < out["bar"]
                         and out.get("group_sum_err", 0.0) < 1e-5)
        del buf
        return out

    def roofline(self, stage, clocks):
        fl = flops_per_cell(self.G, self.S, self.L)
        pk = peaks()
        chain_ms = stage["qz_chain"]
        fused = self.precision == "tc" and 4 <= self.S <= 128 and self.L <= 64
        alg_flop = (fl["chain"] + (fl["dist"] if fused else 0)) * self.n
        achieved = alg_flop / (chain_ms * 1e-3) / 1e12 if chain_ms > 0 else None
        # a kernel timed in a loop that keeps the SM clock at its maximum runs in the burst regime
        burst = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
        peak_tf = pk["bf16_tflops"] if burst else pk["bf16_tflops_sustained"]
        rows = self.n * self.S
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        clk_per_row = chain_ms * 1e-3 * 148 * sm_mhz * 1e6 / rows if chain_ms > 0 else None
        traffic, src = ncu_traffic(self.wl, self.n) if fused else (None, None)
        return {"bound": "tensor",
                "kernel": "k_qz_fp32" if self.precision == "fp32" else ("k_chain_fused_tc" if fused else "k_chain_fused_tc (chain only; distances in k_dist_gram_tc)"),
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                "traffic": traffic, "traffic_source": src,
                "peak_source": f"MEASURED_PEAKS.json bf16 {'burst' if burst else 'sustained'} ({pk['source']})",
                "stage_ms_last_step": stage,
                "algorithmic_flop_per_cell": (fl["chain"] + (fl["dist"] if fused else 0)),
                "cuda_core_bound": {"sm_clk_per_row": clk_per_row, "issue_floor_clk_per_row": 46.0, "mufu_floor_clk_per_row": 30.0,
                                    "note": "the fused chain is bound by its fp32 epilogues, not the tensor pipe: per (cell,sample) row "
                                            "183 warp-instructions (ncu inst_executed / rows, profiles/r2_fused_c4_ncu.json) = 46 issue clk "
                                            "on 4 schedulers, ~480 MUFU ops (320 tanh, 128 sqrt, 32 rcp/rsq) at the measured 16/clk/SM = "
                                            "30 clk, ~28 clk of tensor pipe with the 3-term (hi, lo) products; measured time = "
                                            "instructions x dependency stalls at 4 warps per scheduler (DESIGN.md section 4)"},
                "note": "achieved = algorithmic FLOP/cell (SURVEY 8d: 2*S*(att+mlp0+mlp1) [+ 3*S^2*L fused]) x cells / kernel CUDA-event time"}

    def config(self):
        return {"workload": self.desc, "cells_per_gpu": self.n, "n_genes": self.G, "n_sample": self.S, "n_latent": self.L,
                "n_groups": self.n_groups, "keep_cell": self.keep_cell, "qz_flavor": "attention (reference EncoderUZ)",
                "weights": "random-init (reference initialisers, seed 0)", "sharding": f"cells x{self.world}",
                "l2_policy": f"inputs {self.n * self.G * 4 / 1e9:.2f} GB per GPU >> 126 MB L2 (no flush needed)"}

    def close(self):
        self.h.close()


def api_adata(job, sparse=False):
    """AnnData-like object over this rank's cells for the public-API legs (labels as a pandas Categorical, as scanpy
    / scvi keep them)."""
    import pandas as pd

    import mrvi_b200

    X = job.X
    if sparse:
        import scipy.sparse as sp

        rng = np.random.default_rng(7)
        keep = rng.random(X.shape, dtype=np.float32) < 0.125   # with the 30 % dropout of the recipe: ~91 % zeros
        X = sp.csr_matrix(np.where(keep, X, 0).astype(np.float32))
    obs = pd.DataFrame({"sample": job.sid.astype(np.int64)})
    if job.n_groups:
        obs["cell_type"] = pd.Categorical.from_codes(job.codes, categories=[f"type_{i:02d}" for i in range(job.n_groups)])
    return mrvi_b200.SimpleAnnData(X, obs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MRVI_BENCH_PRECISION", "auto"), choices=["auto", "fp32", "tc"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--cells-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-sample-cells", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C2 / C3 / C5 sub-lines (N = 1 only)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-api", action="store_true", help="skip the e2e_api / e2e_csr legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE line, the JSON line of rank 0: everything else any library prints (NCCL's version
    # banner, warnings) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    all_cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None
    if local_world > 1 and hasattr(os, "sched_setaffinity"):
        # one process per GPU: a contiguous slice of the host cores each, so that the packing threads and the pages they
        # first touch stay on one socket (GPUs and cores are enumerated socket by socket on the HGX boards)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // local_world)
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per] or cores)
        except OSError:
            pass
    n_per, G, S, L, n_groups, keep_cell, desc = WORKLOADS[args.workload]
    if args.cells_per_gpu:
        n_per = args.cells_per_gpu

    import mrvi_b200

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return
        dims = mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=L)
        params = mrvi_b200.init_params(dims, seed=0)
        config = {"workload": desc, "cells_per_gpu": n_per, "n_genes": G, "n_sample": S, "n_latent": L, "n_groups": n_groups,
                  "keep_cell": keep_cell, "qz_flavor": "attention (reference EncoderUZ)",
                  "weights": "random-init (reference initialisers, seed 0)", "sharding": f"cells x{world}",
                  "l2_policy": f"inputs {n_per * G * 4 / 1e9:.2f} GB per GPU >> 126 MB L2 (no flush needed)"}
        sample = args.cpu_sample_cells or cpu_sample_cells(G, S, n_per)
        # a fixed >= 20 s of CPU work whatever --steps says: the denominator of the driver's ratio must not wander
        val, cores, per_step, steps = cpu_reference_run(params, G, S, n_groups, keep_cell, sample, max(1, min(args.steps, 3)),
                                                        min(args.warmup, 1), min_seconds=20.0)
        line = {"metric": "cells/sec", "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample} cells/step of the workload x {steps} steps ({per_step * steps:.1f} s of CPU work); "
                                           "torch-CPU fp32 restatement of the reference path, batch_size=256 (JAX reference not installable here)"},
                "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    from mrvi_b200 import _capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    unique_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        try:  # the library's own communicator for the group-sum all-reduce; torch.distributed stays the plumbing
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(_capi.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            unique_id = bytes(idt.cpu().numpy().tobytes())
        except _capi.MrviError:
            unique_id = None

    job = Job(args.workload, n_per, rank, local_rank, world, args.precision, torch, dist, unique_id)
    precision = job.precision

    for _ in range(args.warmup):
        job.step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _capi.kernel_launch_count()
    dev_s = job.timed_device(args.steps)
    launches = _capi.kernel_launch_count() - launches0
    stage = job.h.last_stage_ms()   # per-stage device times of the last step (events recorded inside the library)
    # ---- e2e: host buffers through the C-ABI
    for _ in range(min(args.warmup, 2)):
        job.step_host()
    e2e_steps = max(1, min(args.steps, 5))
    e2e_wall = job.timed_wall(job.step_host, e2e_steps)
    x_wire = job.h.last_h2d_bytes()  # bytes of X that actually crossed PCIe in the last step (count chunks are narrowed)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the same call when adata.X is stored as uint8 raw counts (reported next to e2e, not as the headline)
    e2e_u8 = None
    if rank == 0 and world == 1 and float(job.Xp.max()) < 256:
        Xu8 = torch.from_numpy(job.X.astype(np.uint8)).pin_memory().numpy()
        job.step_host(Xu8)
        w8 = job.timed_wall(lambda: job.step_host(Xu8), e2e_steps)
        e2e_u8 = {"value": n_per * e2e_steps / w8, "unit": "cells/s",
                  "h2d_bytes_per_step": job.h.last_h2d_bytes() + n_per * (4 + (4 if n_groups else 0)),
                  "input": "adata.X as uint8 counts in pinned host memory (mrvi_run_host_counts)"}
        del Xu8

    # ---- the public API: MrVI.get_local_sample_distances on an AnnData-like object (Python + pandas included)
    e2e_api = e2e_csr = None
    if not args.no_api and world == 1:
        for sparse in (False, True):
            ad = api_adata(job, sparse=sparse)
            model = mrvi_b200.MrVI(ad, job.params, sample_key="sample", sample_order=np.arange(S), precision=precision,
                                   device=local_rank, dims=job.dims)
            kw = dict(groupby="cell_type", keep_cell=False) if n_groups else dict(keep_cell=True)

            def call():
                return model.get_local_sample_distances(**kw)

            call()
            steps_api = max(1, min(args.steps, 3))
            w = job.timed_wall(call, steps_api)
            w = job.max_over_ranks(w)
            leg = {"value": n_per * world * steps_api / w, "unit": "cells/s", "steps": steps_api, "ms_per_step": w / steps_api * 1e3,
                   "api": "mrvi_b200.MrVI.get_local_sample_distances(" + ", ".join(f"{k}={v!r}" for k, v in kw.items()) + ")"}
            if sparse:
                leg["input"] = f"adata.X = scipy CSR, {100.0 * (1.0 - ad.X.nnz / (ad.X.shape[0] * ad.X.shape[1])):.1f} % zeros (mrvi_run_host_csr)"
                e2e_csr = leg
            else:
                leg["input"] = "adata.X = float32 ndarray (pageable), obs['cell_type'] categorical"
                e2e_api = leg
            model.close()
            del ad, model

    # ---- parity of the timed path
    parity = None
    if not args.no_parity:
        parity = job.parity()
        worst = job.max_over_ranks(max(parity["max_norm_err"], parity["p99_elementwise_rel"]))
        parity["worst_over_ranks"] = worst
        parity["ok"] = bool(parity["ok"] and worst < parity["bar"])

    dev_s = job.max_over_ranks(dev_s)
    e2e_wall = job.max_over_ranks(e2e_wall)
    total_cells = n_per * world
    value = total_cells * args.steps / dev_s
    e2e_value = total_cells * e2e_steps / e2e_wall
    line = None
    if rank == 0:
        h2d = x_wire + n_per * (4 + (4 if n_groups else 0))
        d2h = (n_groups * S * S * 8 if n_groups else 0) + (n_per * S * S * 4 if keep_cell else 0)
        line = {"metric": "cells/sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if precision == "fp32" else "f16 (hi, lo) operand pairs, f32 accumulate",
                "data": "synthetic", "impl": "ours", "precision_mode": precision, "config": job.config(),
                "cell_S2_per_sec": value * S * S,
                "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3,
                        "host_input_bytes_per_step": n_per * (G * 4 + 4 + (4 if n_groups else 0)),
                        "api": "mrvi_run_host (C-ABI, pinned host float32 buffers; inside the timed region the host threads narrow "
                               "the leading rows of every chunk losslessly to uint8/uint16 while the copy engine pulls the remaining "
                               "rows raw; widened back on the GPU, bit-identical)"},
                "gpu_launches": int(launches), "roofline": job.roofline(stage, clocks), "clocks": clocks,
                "collective": (("mrvi_allreduce_groups (NCCL, library communicator)" if job.own_comm else "torch.distributed all_reduce (NCCL)")
                               + f", {n_groups * S * S * 8} bytes per step") if (world > 1 and n_groups) else None}
        if e2e_u8 is not None:
            line["e2e_uint8_input"] = e2e_u8
        if e2e_api is not None:
            line["e2e_api"] = e2e_api
        if e2e_csr is not None:
            line["e2e_csr"] = e2e_csr
        if parity is not None:
            line["parity"] = parity
    job.close()
    del job
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations as sub-lines (N = 1: they are parity-test cases and profile targets, not the headline)
    if rank == 0 and world == 1 and not args.no_extras:
        extras = []
        for wl in ("c2", "c3", "c5"):
            if wl == args.workload:
                continue
            ej = Job(wl, EXTRA_CELLS[wl], 0, local_rank, 1, precision, torch, dist)
            for _ in range(2):
                ej.step_device()
            st_n = max(2, min(args.steps, 5))
            s = ej.timed_device(st_n)
            stg = ej.h.last_stage_ms()
            rf = ej.roofline(stg, clocks)
            e = {"workload": ej.desc, "cells": ej.n, "value": ej.n * st_n / s, "unit": "cells/s", "ms_per_step": s / st_n * 1e3,
                 "steps": st_n, "stage_ms_last_step": stg, "roofline_frac": rf["frac"], "sm_clk_per_row": rf["cuda_core_bound"]["sm_clk_per_row"]}
            if ej.keep_cell:
                e["keep_cell_write_gbs"] = ej.n * ej.S * ej.S * 4 / (s / st_n) / 1e9
            if not args.no_parity:
                e["parity"] = ej.parity()
            extras.append(e)
            ej.close()
            del ej
            torch.cuda.empty_cache()
        line["extra"] = extras

    if rank == 0:
        if not args.no_cpu_baseline:
            if all_cores and hasattr(os, "sched_setaffinity"):
                os.sched_setaffinity(0, all_cores)   # the CPU baseline gets every host core again
            sample = args.cpu_sample_cells or cpu_sample_cells(G, S, n_per)
            cpu_params = mrvi_b200.init_params(mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=L), seed=0)
            val, cores, per_step, passes = cpu_reference_run(cpu_params, G, S, n_groups, keep_cell, sample, 1, 1, min_seconds=20.0)
            line["cpu_baseline"] = {"value": val, "unit": "cells/s", "cores": cores, "kind": "port",
                                    "sample": f"{sample} cells of the workload x {passes} passes ({per_step * passes:.1f} s of CPU work); "
                                              "torch-CPU fp32 restatement of the reference path, batch_size=256, all host threads"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ok = True
    if rank == 0 and line.get("parity") is not None:
        ok = line["parity"]["ok"] and all(e.get("parity", {"ok": True})["ok"] for e in line.get("extra", []))
    if not ok:
        print("bench.py: PARITY CHECK FAILED (see 'parity' in the line above)", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
