"""Multi-GPU sharding of the local-sample-distance path (one process per GPU, ``torch.distributed``).

Every cell is independent through stages (1)-(3); only the per-group mean couples cells.  So cells
are sharded in contiguous ranges ``[r*N/P, (r+1)*N/P)`` (keeps ``keep_cell`` output a plain
concatenation in the original order), weights are replicated, and the only exchange step is ONE
all-reduce of the concatenated per-group sums (float64) and counts (int64) -- NCCL over NVLink on
GPUs, gloo in the CPU tests.  ``keep_cell`` blocks need no collective: each rank owns its slice.
The reference has no multi-device code at all; this is the B200 layer above
its single-device loop (`_model.py:376-504`).
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous cell range of ``rank``: balanced to within one cell, ordered by rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    return (rank * n) // world, ((rank + 1) * n) // world


def allreduce_group_stats(sums: Sequence[np.ndarray], counts: Sequence[np.ndarray], group=None, device=None):
    """Sum the per-key group sums / counts over all ranks with a single all-reduce each.

    sums[k]: float64 [n_groups_k, S, S]; counts[k]: int64 [n_groups_k].  Returns new lists.
    ``device``: torch device to stage on (cuda for NCCL, None/cpu for gloo)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [np.array(s, copy=True) for s in sums], [np.array(c, copy=True) for c in counts]
    if not sums:
        return [], []
    flat = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1) for s in sums]))
    cnt = torch.from_numpy(np.concatenate([np.asarray(c, dtype=np.int64).reshape(-1) for c in counts]))
    if device is not None:
        flat, cnt = flat.to(device), cnt.to(device)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    flat, cnt = flat.cpu().numpy(), cnt.cpu().numpy()
    out_s, out_c, o, oc = [], [], 0, 0
    for s, c in zip(sums, counts):
        out_s.append(flat[o:o + s.size].reshape(s.shape).copy())
        out_c.append(cnt[oc:oc + c.size].copy())
        o += s.size
        oc += c.size
    return out_s, out_c


def sharded_group_means(compute_shard: Callable[[int, int], dict], n_cells: int, rank: int, world: int, group=None,
                        device=None) -> dict:
    """Run ``compute_shard(lo, hi)`` on this rank's cell range and combine the group statistics.

    ``compute_shard`` returns ``{"group_sums": [...], "group_counts": [...], "cell": optional}`` for cells
    [lo, hi) -- in the product it is ``Handle.run_host`` / ``run_device`` on this rank's GPU.  Returns
    ``{"group_means": [float32 arrays], "group_counts": [...], "cell": this rank's slice or None, "range": (lo, hi)}``.
    Groups with zero cells get NaN means (the host API raises for them like the reference does)."""
    lo, hi = shard_range(n_cells, rank, world)
    res = compute_shard(lo, hi)
    sums, counts = allreduce_group_stats(res.get("group_sums", []), res.get("group_counts", []), group=group, device=device)
    means = []
    for s, c in zip(sums, counts):
        with np.errstate(invalid="ignore", divide="ignore"):
            means.append((s / c[:, None, None].astype(np.float64)).astype(np.float32))
    return {"group_means": means, "group_counts": counts, "cell": res.get("cell"), "range": (lo, hi)}
