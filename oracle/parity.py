"""Error statistics for the parity checks (TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and the parity
spot-check of bench.py, never by the product).

The north_star bar is "distances within 1e-3 relative (tensor-core mode) / 1e-5 (fp32 mode)" of the reference.  Three
readings of "relative" are reported side by side so that none hides behind another:

* ``max_norm_err``      max over cells of  max|D - D_ref| / max(D_ref of that cell)  -- the bar is applied to THIS one
                        (per cell block: a cell with small distances is not helped by other cells' large ones);
* ``global_norm_err``   max|D - D_ref| / max(D_ref) over the whole tensor (what round 1's fp32 tests used);
* ``p99_elementwise_rel`` / ``max_elementwise_rel``  |D - D_ref| / D_ref over the off-diagonal entries.  Element-wise
                        relative error of a SMALL distance is bounded below by the precision of z itself in any
                        float32 pipeline, the reference's included (error of z ~ eps * |z|, not eps * d): it is
                        reported, and bounded by the bar only at the 99th percentile.
"""
from __future__ import annotations

import numpy as np


def distance_error_stats(got: np.ndarray, ref: np.ndarray) -> dict:
    """got, ref: (n, S, S) distance blocks (ref float64)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    diff = np.abs(got - ref)
    cell_scale = np.maximum(ref.max(axis=(1, 2)), 1e-300)
    per_cell = diff.max(axis=(1, 2)) / cell_scale
    S = ref.shape[-1]
    off = ~np.eye(S, dtype=bool)
    r = ref[:, off]
    d = diff[:, off]
    ok = r > 0
    rel = d[ok] / r[ok] if ok.any() else np.zeros(1)
    return {
        "max_norm_err": float(per_cell.max()) if per_cell.size else 0.0,
        "global_norm_err": float(diff.max() / max(ref.max(), 1e-300)) if diff.size else 0.0,
        "p99_elementwise_rel": float(np.quantile(rel, 0.99)) if rel.size else 0.0,
        "max_elementwise_rel": float(rel.max()) if rel.size else 0.0,
        "worst_cell": int(per_cell.argmax()) if per_cell.size else -1,
        "n_cells": int(ref.shape[0]),
    }


def stratified_indices(n: int, k: int) -> np.ndarray:
    """k cell indices spread evenly over [0, n) (first and last cell included): a sample that touches every region of a
    device chunk -- the first tiles, the tail tile, both ends of every CTA's range."""
    k = min(k, n)
    if k <= 1:
        return np.zeros(k, dtype=np.int64)
    return np.unique(np.round(np.linspace(0, n - 1, k)).astype(np.int64))
