#!/usr/bin/env python
"""Benchmark of the counterfactual local-sample-distance path (BASELINE.json metric: cells/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|tc]
                  [--workload c2|c3|c4|c5] [--cells-per-gpu M]

One "step" = one full pass of the hot path (encoder -> q(z|u,s) for all S samples -> S x S
distances -> per-group sums) over this rank's shard of the workload.  Default workload is
BASELINE.json configs[1] ("C2": 100k cells x 2k genes x 32 samples, n_latent=30, groupby 20 cell
types; attention q(z|u,s), the only flavor the reference implements).  Weak scaling: every rank
owns `cells-per-gpu` cells; the per-group sums are combined with one NCCL all-reduce per step.

`value`  : cells/s with X already resident in HBM (mrvi_run_device), CUDA-event timed, max over ranks.
`e2e`    : cells/s through the host-buffer C-ABI call (mrvi_run_host) from pinned host memory,
           H2D of X/sample/group codes and D2H of the group sums inside the timed region.
`roofline`: the dominant kernel (q(z|u,s) chain), algorithmic FLOPs / its CUDA-event duration.
`cpu_baseline`: the torch-CPU restatement of the reference (oracle "port"; the JAX reference cannot
           be installed here) on a bounded sample of the same workload, all host threads.
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
           "C4 shard: 125k of 1M cells x 2k genes x 128 samples, groupby 50 (configs[3] / 8 GPUs)"),
    "c5": (250_000, 5000, 256, 50, 50, False,
           "C5 shard: 250k of 2M cells x 5k genes x 256 samples, n_latent=50, groupby (configs[4] / 8 GPUs)"),
}


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
    while done < steps or (min_seconds and time.perf_counter() - t0 < min_seconds and done < 40):
        orc.local_sample_distances(X, sid, **kw)
        done += 1
    dt = time.perf_counter() - t0
    return sample_cells * done / dt, cores, dt / done, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MRVI_BENCH_PRECISION", "auto"), choices=["auto", "fp32", "tc"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cells-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-sample-cells", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_per, G, S, L, n_groups, keep_cell, desc = WORKLOADS[args.workload]
    if args.cells_per_gpu:
        n_per = args.cells_per_gpu
    fl = flops_per_cell(G, S, L)

    import mrvi_b200

    dims = mrvi_b200.MrviDims(n_input=G, n_sample=S, n_latent=L)
    params = mrvi_b200.init_params(dims, seed=0)
    config = {"workload": desc, "cells_per_gpu": n_per, "n_genes": G, "n_sample": S, "n_latent": L,
              "n_groups": n_groups, "keep_cell": keep_cell, "qz_flavor": "attention (reference EncoderUZ)",
              "weights": "random-init (reference initialisers, seed 0)", "sharding": f"cells x{world}",
              "l2_policy": f"inputs {n_per * G * 4 / 1e9:.2f} GB per GPU >> 126 MB L2 (no flush needed)"}

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample_cells or max(2048, min(n_per, int(40000 * 32 / S * 2000 / G) // 256 * 256))
        steps = max(1, min(args.steps, 20))
        val, cores, per_step, steps = cpu_reference_run(params, G, S, n_groups, keep_cell, sample, steps, min(args.warmup, 1))
        line = {"metric": "cells/sec", "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 1), "ms_per_step": per_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample} cells/step of the workload; torch-CPU fp32 restatement of the reference "
                                           "path, batch_size=256 (JAX reference not installable here)"},
                "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist

    from mrvi_b200 import _capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision
    if precision == "auto":
        try:
            h = _capi.Handle(dims, params, device=local_rank, precision=_capi.PRECISION_TC)
            precision = "tc"
        except _capi.MrviError as e:
            if e.status != -5:
                raise
            h = _capi.Handle(dims, params, device=local_rank, precision=_capi.PRECISION_FP32)
            precision = "fp32"
    else:
        h = _capi.Handle(dims, params, device=local_rank,
                         precision=_capi.PRECISION_TC if precision == "tc" else _capi.PRECISION_FP32)

    X, sid, codes = make_inputs(n_per, G, S, n_groups, seed=1000 + rank)
    Xp = torch.from_numpy(X).pin_memory()
    sidp = torch.from_numpy(sid).pin_memory()
    codesp = torch.from_numpy(codes).pin_memory() if n_groups else None
    Xd, sidd = Xp.to(dev), sidp.to(dev)
    codesd = codesp.to(dev) if n_groups else None
    sums_d = torch.zeros((n_groups, S, S), dtype=torch.float64, device=dev) if n_groups else None
    cell_d = torch.empty((n_per, S, S), dtype=torch.float32, device=dev) if keep_cell else None
    stream = torch.cuda.current_stream(dev)

    def step_device():
        if n_groups:
            sums_d.zero_()
        h.run_device(n_per, Xd.data_ptr(), G, sidd.data_ptr(),
                     [codesd.data_ptr()] if n_groups else [], [n_groups] if n_groups else [], "l2",
                     cell_d.data_ptr() if keep_cell else 0, [sums_d.data_ptr()] if n_groups else [], 0, 0,
                     stream.cuda_stream)
        if world > 1 and n_groups:
            dist.all_reduce(sums_d)

    cell_h = np.empty((n_per, S, S), dtype=np.float32) if keep_cell else None
    if keep_cell:
        cell_h = torch.from_numpy(cell_h).pin_memory().numpy()

    def step_host():
        res = h.run_host(Xp.numpy(), sidp.numpy(), [codesp.numpy()] if n_groups else [], [n_groups] if n_groups else [],
                         keep_cell=keep_cell, cell_out=cell_h)
        if world > 1 and n_groups:
            t = torch.from_numpy(res["group_sums"][0]).to(dev)
            dist.all_reduce(t)
            res["group_sums"][0] = t.cpu().numpy()
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps bracketed by barrier+sync, CUDA events on the launching stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        return e0.elapsed_time(e1) * 1e-3, wall

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _capi.kernel_launch_count()
    stage = {"encoder": 0.0, "qz_chain": 0.0, "distances": 0.0, "total": 0.0}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    dev_s = e0.elapsed_time(e1) * 1e-3
    launches = _capi.kernel_launch_count() - launches0
    # per-stage device times of the last step (events recorded inside the library on the same stream)
    stage = h.last_stage_ms()
    # e2e: host buffers through the C-ABI
    for _ in range(min(args.warmup, 2)):
        step_host()
    e2e_steps = max(1, min(args.steps, 5))
    _, e2e_wall = timed(step_host, e2e_steps)
    x_wire = h.last_h2d_bytes()  # bytes of X that actually crossed PCIe in the last step (count chunks are narrowed)
    clocks = sampler.stop() if rank == 0 else None
    # the same end-to-end call when adata.X is stored as uint8 raw counts (mrvi_run_host_counts): reported next to
    # e2e, NOT as the headline -- the reference's loader hands float32
    e2e_u8 = None
    if rank == 0 and world == 1 and float(Xp.max()) < 256:
        Xu8 = torch.from_numpy(Xp.numpy().astype(np.uint8)).pin_memory()

        def step_host_u8():
            return h.run_host(Xu8.numpy(), sidp.numpy(), [codesp.numpy()] if n_groups else [], [n_groups] if n_groups else [],
                              keep_cell=keep_cell, cell_out=cell_h)

        step_host_u8()
        _, w8 = timed(step_host_u8, e2e_steps)
        e2e_u8 = {"value": n_per * e2e_steps / w8, "unit": "cells/s", "h2d_bytes_per_step": h.last_h2d_bytes() + n_per * (4 + (4 if n_groups else 0)),
                  "input": "adata.X as uint8 counts in pinned host memory (mrvi_run_host_counts)"}

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_s = max_over_ranks(dev_s)
    e2e_wall = max_over_ranks(e2e_wall)
    total_cells = n_per * world
    value = total_cells * args.steps / dev_s
    e2e_value = total_cells * e2e_steps / e2e_wall
    if rank == 0:
        pk = peaks()
        chain_ms = stage["qz_chain"]
        fused = precision == "tc" and 4 <= S <= 128
        # the dominant kernel is launched once per device chunk; the stage time is the sum over this step's
        # launches (CUDA events recorded by the library on the launching stream).  In tensor-core mode the
        # q(z|u,s) chain, the S x S distances and the group sums are ONE kernel (k_chain_fused_tc).
        alg_flop = (fl["chain"] + (fl["dist"] if fused else 0)) * n_per
        achieved = alg_flop / (chain_ms * 1e-3) / 1e12 if chain_ms > 0 else None
        peak_tf = pk["bf16_tflops_sustained"]
        rows = n_per * S
        clk_per_row = chain_ms * 1e-3 * 148 * 1.965e9 / rows if chain_ms > 0 else None
        roofline = {"bound": "tensor", "kernel": "k_qz_fp32" if precision == "fp32" else ("k_chain_fused_tc" if fused else "k_chain_fused_tc (chain only; distances in k_dist_gram_tc)"),
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": (achieved / peak_tf) if achieved else None,
                    # DRAM bytes of the dominant kernel per launch, from the ncu --set full capture of this workload's
                    # step (profiles/r1_final_tc_fused_c2_ncu_raw.csv: dram__bytes_read 11.56 MB + dram__bytes_write 0 for
                    # the 100k-cell launch = 115.6 B per cell: the per-cell chain inputs; z and D stay on chip)
                    "traffic": (115.6 * n_per) if (fused and args.workload == "c2" and not keep_cell) else None,
                    "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({pk['source']})",
                    "stage_ms_last_step": stage,
                    "algorithmic_flop_per_cell": (fl["chain"] + (fl["dist"] if fused else 0)),
                    "cuda_core_bound": {"sm_clk_per_row": clk_per_row, "mufu_floor_clk_per_row": 62.0,
                                        "note": "the fused chain is bound by its fp32 epilogues, not the tensor pipe: per (cell,sample) row "
                                                "~992 MUFU ops (512 attention ex2, 320 tanh, 128 sqrt, 32 rcp) at the measured 16/clk/SM "
                                                "= 62 clk; tensor work is ~12 clk/row (DESIGN.md section 4)"},
                    "note": "achieved = algorithmic FLOP/cell (SURVEY 8d: 2*S*(att+mlp0+mlp1) [+ 3*S^2*L fused]) x cells / kernel CUDA-event time"}
        h2d = x_wire + n_per * (4 + (4 if n_groups else 0))
        d2h = (n_groups * S * S * 8 if n_groups else 0) + (n_per * S * S * 4 if keep_cell else 0)
        line = {"metric": "cells/sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if precision == "fp32" else "f16/tf32 operands, f32 accumulate",
                "data": "synthetic", "impl": "ours", "precision_mode": precision, "config": config,
                "cell_S2_per_sec": value * S * S,
                "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "ms_per_step": e2e_wall / e2e_steps * 1e3,
                        "host_input_bytes_per_step": n_per * (G * 4 + 4 + (4 if n_groups else 0)),
                        "api": "mrvi_run_host (C-ABI, pinned host float32 buffers; integer count chunks are narrowed "
                               "losslessly to uint8/uint16 on the host, inside the timed region, before the copy)"},
                "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks}
        if e2e_u8 is not None:
            line["e2e_uint8_input"] = e2e_u8
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample_cells or max(2048, min(n_per, int(40000 * 32 / S * 2000 / G) // 256 * 256))
            val, cores, per_step, passes = cpu_reference_run(params, G, S, n_groups, keep_cell, sample, 1, 1, min_seconds=12.0)
            line["cpu_baseline"] = {"value": val, "unit": "cells/s", "cores": cores, "kind": "port",
                                    "sample": f"{sample} cells of the workload x {passes} passes ({per_step * passes:.1f} s of CPU work); "
                                              "torch-CPU fp32 restatement of the reference path, batch_size=256, all host threads"}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
