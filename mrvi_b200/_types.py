"""Reduction descriptors for ``compute_local_statistics``.

Same names, fields and defaults as the reference's ``src/mrvi/_types.py:12-52`` so that
caller code which builds ``MrVIReduction`` objects keeps working unchanged.
"""
from __future__ import annotations

from collections.abc import Iterable
from dataclasses import dataclass
from typing import Callable, Literal, Optional


def _default_fn(x):
    # reference: ``lambda x: xr.DataArray(x)`` (identity copy), `_types.py:38`
    from ._xr import DataArray

    return x if isinstance(x, DataArray) else DataArray(x)


@dataclass(frozen=True)
class MrVIReduction:
    """One reduction over the per-cell counterfactual statistics (reference `_types.py:12-39`)."""

    name: str
    input: Literal[
        "mean_representations",
        "mean_distances",
        "sampled_representations",
        "sampled_distances",
        "normalized_distances",
    ]
    fn: Callable = _default_fn
    group_by: Optional[str] = None


@dataclass(frozen=True)
class _ComputeLocalStatisticsRequirements:
    """Summarised needs of a list of reductions (reference `_types.py:42-52`)."""

    needs_mean_representations: bool
    needs_mean_distances: bool
    needs_sampled_representations: bool
    needs_sampled_distances: bool
    needs_normalized_distances: bool
    ungrouped_reductions: Iterable[MrVIReduction]
    grouped_reductions: Iterable[MrVIReduction]
