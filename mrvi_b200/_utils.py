"""Host-side planner for ``compute_local_statistics`` (reference `src/mrvi/_utils.py:9-51`)."""
from __future__ import annotations

from ._types import MrVIReduction, _ComputeLocalStatisticsRequirements

_INPUT_NEEDS = {
    # input kind -> (mean_rep, sampled_rep, mean_dists, sampled_dists, normalized_dists)
    "mean_representations": (True, False, False, False, False),
    "sampled_representations": (False, True, False, False, False),
    "mean_distances": (True, False, True, False, False),
    "sampled_distances": (False, True, False, True, False),
    "normalized_distances": (False, True, False, False, True),
}


def _parse_local_statistics_requirements(
    reductions: list[MrVIReduction],
) -> _ComputeLocalStatisticsRequirements:
    """Which intermediates a list of reductions needs; same truth table as the reference."""
    flags = [False] * 5
    ungrouped, grouped = [], []
    for r in reductions:
        if r.input not in _INPUT_NEEDS:
            raise ValueError(f"Unknown reduction input: {r.input}")
        flags = [a or b for a, b in zip(flags, _INPUT_NEEDS[r.input])]
        (ungrouped if r.group_by is None else grouped).append(r)
    return _ComputeLocalStatisticsRequirements(
        needs_mean_representations=flags[0],
        needs_sampled_representations=flags[1],
        needs_mean_distances=flags[2],
        needs_sampled_distances=flags[3],
        needs_normalized_distances=flags[4],
        grouped_reductions=grouped,
        ungrouped_reductions=ungrouped,
    )


def fdr_bh(pvals):
    """Benjamini-Hochberg adjusted p-values: what ``multipletests(p, method="fdr_bh")[1]`` returns in the reference
    (`_model.py:1363`; statsmodels is not a dependency here).  p_(i) * m / i over the sorted p-values, made monotone from
    the largest downwards, clipped at 1, returned in the input's order and shape."""
    import numpy as np

    p = np.asarray(pvals, dtype=np.float64)
    flat = p.reshape(-1)
    m = flat.size
    if m == 0:
        return p.copy()
    order = np.argsort(flat, kind="stable")
    adj = flat[order] * (m / np.arange(1, m + 1, dtype=np.float64))
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    out = np.empty(m, dtype=np.float64)
    out[order] = np.minimum(adj, 1.0)
    return out.reshape(p.shape)
