"""Host-side mirror of the reference's ``MrVI`` for the local-sample-distance path.

Same method names, arguments, validation, warnings, output layout and quirks as
``src/mrvi/_model.py`` (``get_local_sample_distances`` :627-714, ``compute_local_statistics``
:273-506, ``get_local_sample_representation`` :586-625), with the device work
(``mapped_inference_fn`` :326-365, ``_compute_distances_from_representations`` :542-584 and the
per-category sums :469-484) replaced by one call into the CUDA library through the C-ABI
(``mrvi_b200._capi``).  pandas / xarray semantics (group order = ``value_counts()`` order, counts,
coords, errors) stay on the host, exactly where the reference has them.

There is no CPU fallback: constructing a model without ``libmrvi_b200.so`` and a B200 raises.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from typing import Mapping, Optional, Sequence

import numpy as np
import pandas as pd

from . import _capi
from ._data import dense_float32
from ._params import MrviDims, flatten_params, infer_dims, validate_params
from ._types import MrVIReduction, _default_fn
from ._utils import _parse_local_statistics_requirements, fdr_bh
from ._xr import DataArray, Dataset

_PRECISIONS = {"fp32": _capi.PRECISION_FP32, "tc": _capi.PRECISION_TC}


_SAMPLED_INPUTS = ("sampled_representations", "sampled_distances", "normalized_distances")


def _identity(x):
    return x


def _is_identity_fn(fn) -> bool:
    return fn is _identity or fn is _default_fn or getattr(fn, "_mrvi_identity", False)


class MrVI:
    """Trained MrVI model restricted to the counterfactual local-sample-distance path.

    Parameters
    ----------
    adata
        AnnData-like object the model was trained on (``.X``, ``.obs``, ``.obs_names``, ``.n_obs``).
    params
        The trained flax parameters, nested (``module.params``) or already flat
        (``{"qu/Dense_0/kernel": ...}``); see :mod:`mrvi_b200._params`.
    sample_key
        ``obs`` column holding the sample label of each cell (reference ``setup_anndata(sample_key=...)``).
    sample_order
        Category labels in code order (reference ``self.sample_order``, `_model.py:113-115`).  Default:
        sorted unique labels of ``adata.obs[sample_key]`` -- what scvi's CategoricalObsField registers.
    precision
        ``"fp32"`` (FFMA, parity 1e-5) or ``"tc"`` (tcgen05 tensor cores, parity 1e-3).
    device
        CUDA device index (single-GPU use).
    devices
        CUDA device indices to shard over, e.g. ``devices=range(8)`` or ``"all"``: cells are split into contiguous
        ranges, one per device (one library handle and one host thread each); the per-group sums are added up on the
        host in float64 and every device streams its ``keep_cell`` blocks into its slice of ONE page-locked output.
        Results are the single-GPU ones (counts and cell order bit-exact; sums up to float64 re-association).
    seed
        Seed of the sampled paths' noise generator (see ``self.seed``).
    """

    def __init__(self, adata, params: Mapping, sample_key: str = "sample", sample_order: Optional[Sequence] = None,
                 precision: str = "fp32", device: int = 0, dims: Optional[MrviDims] = None, seed: int = 0,
                 devices=None, batch_key: Optional[str] = None):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, not {precision!r}")
        self.adata = adata
        self.sample_key = sample_key
        #: ``obs`` column of the batch (site) of each cell; only ``differential_expression(add_batch_specific_offsets=True)`` reads it
        self.batch_key = batch_key
        flat = flatten_params(params)
        self.dims = infer_dims(flat) if dims is None else dims
        self.params = validate_params(flat, self.dims)
        if sample_order is None:
            col = adata.obs[sample_key]
            if isinstance(col.dtype, pd.CategoricalDtype):
                sample_order = np.asarray(col.cat.categories)
            else:
                sample_order = np.sort(pd.unique(col))
        self.sample_order = np.asarray(sample_order)
        if len(self.sample_order) != self.dims.n_sample:
            raise ValueError(f"{len(self.sample_order)} sample categories but the parameters were trained with "
                             f"n_sample={self.dims.n_sample}")
        if adata.n_vars != self.dims.n_input:
            raise ValueError(f"adata has {adata.n_vars} genes, parameters expect {self.dims.n_input}")
        self.precision = precision
        if devices is None:
            devices = [int(device)]
        elif isinstance(devices, str):
            if devices != "all":
                raise ValueError("devices must be a sequence of CUDA device indices or 'all'")
            devices = list(range(_capi.device_count()))
        self.devices = [int(d) for d in devices]
        if not self.devices or len(set(self.devices)) != len(self.devices):
            raise ValueError(f"devices must be a non-empty list of distinct CUDA device indices, not {devices!r}")
        self.device = self.devices[0]
        #: seed of the device Philox generator of the sampled paths (use_mean=False / normalize_distances=True).
        #: Draws depend on (seed, cell index, draw, sample) only, so results are reproducible and independent of
        #: chunking / sharding; the reference's JAX threefry streams are NOT reproduced (statistical parity).
        self.seed = int(seed)
        self.is_trained_ = True
        self._handles = [_capi.Handle(self.dims, self.params, device=d, precision=_PRECISIONS[precision]) for d in self.devices]
        self._handle = self._handles[0]
        self.last_stage_ms: dict = {}

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_reference(cls, ref_model, precision: str = "fp32", device: int = 0) -> "MrVI":
        """Build from a trained reference ``mrvi.MrVI`` instance (reads ``module.params``, ``adata``,
        ``sample_key``, ``sample_order``)."""
        return cls(ref_model.adata, ref_model.module.params, sample_key=ref_model.sample_key,
                   sample_order=ref_model.sample_order, precision=precision, device=device)

    def to_device(self, device):  # reference `_model.py:132-134` is a no-op as well
        pass

    def close(self):
        for h in self._handles:
            h.close()

    # ------------------------------------------------------------------ host helpers
    def _check_if_trained(self, warn: bool = False):
        if not self.is_trained_:
            raise RuntimeError("Trying to query inferred values from an untrained model. Please train the model first.")

    def _sample_codes(self, adata) -> np.ndarray:
        """Sample label -> code through the registered mapping (scvi ``_validate_anndata`` + CategoricalObsField)."""
        col = adata.obs[self.sample_key]
        order = self.sample_order
        vals = col.to_numpy() if not isinstance(col.dtype, pd.CategoricalDtype) else None
        if (vals is not None and vals.dtype.kind in "iu" and order.dtype.kind in "iu"
                and np.array_equal(order, np.arange(len(order)))):
            codes = vals.astype(np.int64, copy=False)   # labels ARE the codes: one range check instead of a hash pass
            bad_mask = (codes < 0) | (codes >= len(order))
        else:
            if vals is None:   # categorical column: translate the category table, then one table look-up per cell
                lut = pd.Index(order).get_indexer(col.cat.categories)
                old = col.cat.codes.to_numpy()
                codes = np.where(old >= 0, lut[np.maximum(old, 0)], -1)
            else:
                codes = pd.Index(order).get_indexer(vals)
            bad_mask = codes < 0
        if bad_mask.any():
            bad = pd.unique(np.asarray(col)[bad_mask])[:5]
            raise ValueError(f"Category {list(bad)} not found in source registry for sample key {self.sample_key!r}.")
        return np.asarray(codes, dtype=np.int32)

    def _group_codes(self, adata, key: str, indices: np.ndarray):
        """Per-cell int32 group code in ``value_counts()`` order plus (labels, counts).  Vectorised: one hash pass over
        the labels (or one table look-up of the category codes), no Python loop over cells."""
        vc = adata.obs[key].value_counts()  # reference `_model.py:494-495`
        cats = list(vc.index)
        counts = vc.to_numpy().astype(np.int64)
        # labels come from self.adata, positionally (reference quirk, `_model.py:470`; pandas>=3 needs to_numpy)
        col = self.adata.obs[key]
        whole = len(indices) == len(col) and (len(indices) == 0 or (indices[0] == 0 and indices[-1] == len(col) - 1
                                                                    and np.all(np.diff(indices) == 1)))
        if isinstance(col.dtype, pd.CategoricalDtype):
            # category code -> position in value_counts() order (-1: not in `adata`'s categories / NaN)
            lut = pd.Categorical(col.cat.categories, categories=cats).codes.astype(np.int32)
            old = col.cat.codes.to_numpy()
            if not whole:
                old = old[indices]
            codes = np.where(old >= 0, lut[np.maximum(old, 0)], -1).astype(np.int32)
        else:
            labels = col.to_numpy() if whole else col.to_numpy()[indices]
            codes = pd.Categorical(labels, categories=cats).codes.astype(np.int32)
        return codes, cats, counts

    def _host_matrix(self, X, idx: np.ndarray, whole: bool):
        """``adata.X`` (or its rows ``idx``) in a form the C-ABI takes WITHOUT a float32 copy of the matrix where
        possible: CSR stays sparse (densified on the GPU), integer count matrices (uint8 / uint16 / int32 / int64) are
        narrowed chunk by chunk inside the library, C-contiguous float32 is passed as it is."""
        if hasattr(X, "tocsr") and not isinstance(X, np.ndarray):
            return X.tocsr() if whole else X.tocsr()[idx]
        if isinstance(X, np.ndarray) and X.dtype in _capi._X_DTYPES:
            return X if whole else X[idx]
        if whole and isinstance(X, np.ndarray) and X.dtype == np.float32 and X.ndim == 2 and X.strides[1] == 4 and X.strides[0] > 0:
            return X
        return dense_float32(X) if whole else dense_float32(X[idx])

    def _run_sharded(self, Xh, sample_codes, group_codes, n_groups, norm, keep_cell, want_z, want_u=False):
        """One ``run_host`` per device over contiguous cell ranges (host threads; ctypes releases the GIL), group sums
        added on the host (float64: the path's only exchange step), keep_cell / z blocks written by
        every device straight into its slice of one page-locked output."""
        n = Xh.shape[0]
        P = min(len(self._handles), max(1, n))
        if P <= 1:
            res = self._handle.run_host(Xh, sample_codes, group_codes, n_groups, norm=norm, keep_cell=keep_cell, want_z=want_z,
                                        want_u=want_u)
            self.last_stage_ms = self._handle.last_stage_ms()
            return res
        from concurrent.futures import ThreadPoolExecutor

        from ._dist import shard_range

        d = self.dims
        S, L = d.n_sample, d.n_latent
        cell = _capi._out_empty((n, S, S)) if keep_cell else None
        z = _capi._out_empty((n, S, L)) if want_z else None
        u = np.empty((n, d.n_latent_q), dtype=np.float32) if want_u else None

        def work(r):
            lo, hi = shard_range(n, r, P)
            res = self._handles[r].run_host(Xh[lo:hi], sample_codes[lo:hi], [c[lo:hi] for c in group_codes], n_groups, norm=norm,
                                            keep_cell=keep_cell, want_z=want_z, want_u=want_u,
                                            cell_out=cell[lo:hi] if keep_cell else None, z_out=z[lo:hi] if want_z else None)
            if want_u:
                u[lo:hi] = res["u"]
            return res

        with ThreadPoolExecutor(max_workers=P) as ex:
            parts = list(ex.map(work, range(P)))
        out = {"group_sums": [np.sum([p["group_sums"][k] for p in parts], axis=0) for k in range(len(n_groups))],
               "group_counts": [np.sum([p["group_counts"][k] for p in parts], axis=0) for k in range(len(n_groups))]}
        if keep_cell:
            out["cell"] = cell
        if want_z:
            out["z"] = z
        if want_u:
            out["u"] = u
        self.last_stage_ms = self._handle.last_stage_ms()
        return out

    # ------------------------------------------------------------------ reference API
    def compute_local_statistics(self, reductions: list, adata=None, indices: Optional[Sequence[int]] = None,
                                 batch_size: Optional[int] = None, use_vmap: bool = True, norm: str = "l2",
                                 mc_samples: int = 10):
        """Reference `_model.py:273-506`.  ``batch_size`` / ``use_vmap`` only change memory use in the
        reference; here ``batch_size`` is accepted for signature compatibility and the device chunking is
        chosen by the library (results do not depend on either)."""
        if not reductions or len(reductions) == 0:
            raise ValueError("At least one reduction must be provided.")
        adata = self.adata if adata is None else adata
        self._check_if_trained(warn=False)
        # Hack to ensure new AnnDatas have indices.  (reference `_model.py:314`)
        adata.obs["_indices"] = np.arange(adata.n_obs).astype(int)
        if norm not in _capi.NORMS:
            raise ValueError(f"norm {norm} not supported")

        reqs = _parse_local_statistics_requirements(reductions)
        sampled = [r for r in reductions if r.input in _SAMPLED_INPUTS]
        if sampled:
            if reqs.needs_normalized_distances and norm != "l2":
                raise ValueError(f"Norm must be 'l2' when using normalized distances. Got {norm}.")
            return self._compute_with_sampled(reductions, reqs, adata, indices, batch_size, use_vmap, norm, mc_samples)

        idx = np.arange(adata.n_obs) if indices is None else np.asarray(indices).astype(int).reshape(-1)
        sample_codes = self._sample_codes(adata)[idx]
        whole = indices is None
        Xh = self._host_matrix(adata.X, idx, whole)

        generic = [r for r in reductions if not _is_identity_fn(r.fn)]
        if generic:
            return self._compute_generic(reductions, adata, idx, Xh, sample_codes, batch_size, norm)

        # ---- fused device path: one C-ABI call for everything
        keys = [r.group_by for r in reqs.grouped_reductions]
        plans = [self._group_codes(adata, k, idx) for k in keys]
        keep_cell = any(r.input == "mean_distances" for r in reqs.ungrouped_reductions)
        want_z = any(r.input == "mean_representations" for r in reductions)
        if any(r.input == "mean_representations" for r in reqs.grouped_reductions):
            return self._compute_generic(reductions, adata, idx, Xh, sample_codes, batch_size, norm)
        res = self._run_sharded(Xh, sample_codes, [p[0] for p in plans], [len(p[1]) for p in plans], norm, keep_cell, want_z)
        # (cell coordinates only when a per-cell output exists: indexing 10^6 string labels is milliseconds of pure overhead)
        cell_names = None
        if reqs.ungrouped_reductions:
            cell_names = self.adata.obs_names.values if (whole and self.adata.n_obs == len(idx)) else self.adata.obs_names[idx].values
        final = OrderedDict()
        for ur in reqs.ungrouped_reductions:
            if ur.input == "mean_distances":
                final[ur.name] = self._wrap_dists(res["cell"], cell_names)
            else:
                final[ur.name] = self._wrap_reps(res["z"], cell_names)
        for gi, gr in enumerate(reqs.grouped_reductions):
            codes, cats, counts = plans[gi]
            present = np.bincount(codes[codes >= 0], minlength=len(cats))
            for c, p in zip(cats, present):
                if p == 0:
                    raise KeyError(c)  # reference: grouped_data_arrs[gr.name][cat] missing (`_model.py:499`)
            means = (res["group_sums"][gi] / counts[:, None, None].astype(np.float64)).astype(np.float32)
            final[gr.name] = DataArray(
                means,
                dims=[f"{gr.group_by}_name", "sample_x", "sample_y"],
                coords={f"{gr.group_by}_name": np.asarray(cats), "sample_x": self.sample_order,
                        "sample_y": self.sample_order},
                name="sample_distances",
            )
        return Dataset(final)

    def _compute_with_sampled(self, reductions, reqs, adata, indices, batch_size, use_vmap, norm, mc_samples):
        """Reductions over the sampled inputs (reference `_model.py:405-450`): ``mc_samples`` draws of u per
        (cell, counterfactual sample) on the device (`mrvi_run_sampled_host`).  Mean-path reductions in the same
        request go through the usual call."""
        sampled = [r for r in reductions if r.input in _SAMPLED_INPUTS]
        if (any(not _is_identity_fn(r.fn) for r in sampled)
                or any(r.input == "sampled_representations" and r.group_by is not None for r in sampled)):
            # user reduction functions (or grouped representations): the reference's batch loop on the host, every batch's
            # sampled z / distances from the GPU (draws keyed by the global cell index: independent of the batching)
            idx = np.arange(adata.n_obs) if indices is None else np.asarray(indices).astype(int).reshape(-1)
            Xh = dense_float32(adata.X) if indices is None else dense_float32(adata.X[idx])
            return self._compute_generic(reductions, adata, idx, Xh, self._sample_codes(adata)[idx], batch_size, norm,
                                         mc_samples=int(mc_samples))
        results = {}
        mean_reds = [r for r in reductions if r.input not in _SAMPLED_INPUTS]
        if mean_reds:
            ds = self.compute_local_statistics(mean_reds, adata=adata, indices=indices, batch_size=batch_size,
                                               use_vmap=use_vmap, norm=norm, mc_samples=mc_samples)
            results.update(ds.data_vars)
        idx = np.arange(adata.n_obs) if indices is None else np.asarray(indices).astype(int).reshape(-1)
        sample_codes = self._sample_codes(adata)[idx]
        Xh = dense_float32(adata.X) if indices is None else dense_float32(adata.X[idx])
        cell_names = self.adata.obs_names[idx].values
        mc = int(mc_samples)
        want_z = any(r.input == "sampled_representations" for r in sampled)
        for normalize in (False, True):
            fam = "normalized_distances" if normalize else "sampled_distances"
            ung = [r for r in sampled if r.input == fam and r.group_by is None]
            grp = [r for r in sampled if r.input == fam and r.group_by is not None]
            z_here = want_z and not normalize
            if not ung and not grp and not z_here:
                continue
            plans = [self._group_codes(adata, r.group_by, idx) for r in grp]
            res = self._handle.run_sampled_host(Xh, sample_codes, mc, seed=self.seed, group_codes=[p[0] for p in plans],
                                                n_groups=[len(p[1]) for p in plans], norm=norm, normalize=normalize,
                                                keep_cell=bool(ung), want_z=z_here)
            mc_dims = [] if normalize else ["mc_sample"]
            mc_coords = {} if normalize else {"mc_sample": np.arange(mc)}
            for r in ung:
                results[r.name] = DataArray(
                    res["cell"], dims=["cell_name", *mc_dims, "sample_x", "sample_y"],
                    coords={"cell_name": cell_names, **mc_coords, "sample_x": self.sample_order,
                            "sample_y": self.sample_order}, name="sample_distances")
            for gi, r in enumerate(grp):
                codes, cats, counts = plans[gi]
                present = np.bincount(codes[codes >= 0], minlength=len(cats))
                for c, p in zip(cats, present):
                    if p == 0:
                        raise KeyError(c)
                shape = (-1,) + (1,) * (res["group_sums"][gi].ndim - 1)
                means = (res["group_sums"][gi] / counts.reshape(shape).astype(np.float64)).astype(np.float32)
                results[r.name] = DataArray(
                    means, dims=[f"{r.group_by}_name", *mc_dims, "sample_x", "sample_y"],
                    coords={f"{r.group_by}_name": np.asarray(cats), **mc_coords, "sample_x": self.sample_order,
                            "sample_y": self.sample_order}, name="sample_distances")
            if z_here:
                for r in sampled:
                    if r.input == "sampled_representations":
                        results[r.name] = DataArray(
                            res["z"], dims=["cell_name", "mc_sample", "sample", "latent_dim"],
                            coords={"cell_name": cell_names, "sample": self.sample_order}, name="sample_representations")
        final = OrderedDict()
        for r in list(reqs.ungrouped_reductions) + list(reqs.grouped_reductions):
            final[r.name] = results[r.name]
        return Dataset(final)

    def _wrap_dists(self, d, cell_names):
        return DataArray(d, dims=["cell_name", "sample_x", "sample_y"],
                         coords={"cell_name": cell_names, "sample_x": self.sample_order, "sample_y": self.sample_order},
                         name="sample_distances")

    def _wrap_reps(self, z, cell_names):
        return DataArray(z, dims=["cell_name", "sample", "latent_dim"],
                         coords={"cell_name": cell_names, "sample": self.sample_order}, name="sample_representations")

    def _compute_generic(self, reductions, adata, idx, Xh, sample_codes, batch_size, norm, mc_samples: int = 10):
        """Reductions with a user ``fn``: the reference's batch loop (`_model.py:376-504`) on the host,
        with each batch's z / distances (mean and sampled inputs alike, `_model.py:386-450`) computed on the GPU."""
        bs = 128 if batch_size is None else int(batch_size)  # scvi.settings.batch_size default
        reqs = _parse_local_statistics_requirements(reductions)
        ungrouped = {r.name: [] for r in reqs.ungrouped_reductions}
        grouped = {r.name: OrderedDict() for r in reqs.grouped_reductions}
        inputs = {r.input for r in reductions}
        need_mean = bool(inputs & {"mean_representations", "mean_distances"})
        need_sampled = bool(inputs & {"sampled_representations", "sampled_distances"})
        need_norm = "normalized_distances" in inputs
        mc = int(mc_samples)
        for lo in range(0, len(idx), bs):
            hi = min(lo + bs, len(idx))
            names = self.adata.obs_names[idx[lo:hi]].values
            avail = {}
            if need_mean:
                res = self._handle.run_host(Xh[lo:hi], sample_codes[lo:hi], norm=norm, keep_cell=reqs.needs_mean_distances, want_z=True)
                avail["mean_representations"] = self._wrap_reps(res["z"], names)
                if reqs.needs_mean_distances:
                    avail["mean_distances"] = self._wrap_dists(res["cell"], names)
            if need_sampled:
                res = self._handle.run_sampled_host(Xh[lo:hi], sample_codes[lo:hi], mc, seed=self.seed, cell_offset=lo, norm=norm,
                                                    keep_cell="sampled_distances" in inputs, want_z="sampled_representations" in inputs)
                if "sampled_representations" in inputs:
                    avail["sampled_representations"] = DataArray(
                        res["z"], dims=["cell_name", "mc_sample", "sample", "latent_dim"],
                        coords={"cell_name": names, "sample": self.sample_order}, name="sample_representations")
                if "sampled_distances" in inputs:
                    avail["sampled_distances"] = DataArray(
                        res["cell"], dims=["cell_name", "mc_sample", "sample_x", "sample_y"],
                        coords={"cell_name": names, "mc_sample": np.arange(mc), "sample_x": self.sample_order,
                                "sample_y": self.sample_order}, name="sample_distances")
            if need_norm:
                res = self._handle.run_sampled_host(Xh[lo:hi], sample_codes[lo:hi], mc, seed=self.seed, cell_offset=lo, norm=norm,
                                                    normalize=True, keep_cell=True)
                avail["normalized_distances"] = self._wrap_dists(res["cell"], names)
            for r in reductions:
                outputs = r.fn(avail[r.input])
                if r.group_by is not None:
                    group_by = self.adata.obs[r.group_by].to_numpy()[idx[lo:hi]]
                    for cat in pd.unique(group_by):
                        sel = outputs.sel(cell_name=names[group_by == cat]).sum(dim="cell_name")
                        if cat not in grouped[r.name]:
                            grouped[r.name][cat] = sel
                        else:
                            prev = grouped[r.name][cat]
                            grouped[r.name][cat] = DataArray(np.asarray(prev.values) + np.asarray(sel.values),
                                                             dims=prev.dims, coords=dict(prev.coords), name=prev.name)
                else:
                    ungrouped[r.name].append(outputs)
        final = OrderedDict()
        for name, outs in ungrouped.items():
            first = outs[0]
            final[name] = DataArray(
                np.concatenate([np.asarray(o.values) for o in outs], axis=0), dims=first.dims,
                coords={**{k: v for k, v in dict(first.coords).items() if k != "cell_name"},
                        "cell_name": np.concatenate([np.asarray(dict(o.coords)["cell_name"]) for o in outs])},
                name=first.name)
        for gr in reqs.grouped_reductions:
            vc = adata.obs[gr.group_by].value_counts()
            arrs, cats = [], []
            for cat, count in vc.items():
                g = grouped[gr.name][cat]  # KeyError for an empty category, as in the reference
                arrs.append(np.asarray(g.values) / count)
                cats.append(cat)
            g0 = grouped[gr.name][cats[0]]
            final[gr.name] = DataArray(np.stack(arrs, axis=0), dims=[f"{gr.group_by}_name", *g0.dims],
                                       coords={**dict(g0.coords), f"{gr.group_by}_name": np.asarray(cats)}, name=g0.name)
        return Dataset(final)

    # ------------------------------------------------------------------ differential expression (latent-space stage)
    @property
    def sample_info(self) -> pd.DataFrame:
        """One row of ``adata.obs`` per sample, indexed by sample code (reference `_model.py:107-109`)."""
        obs_df = self.adata.obs.copy()
        obs_df["_scvi_sample"] = self._sample_codes(self.adata)
        if self.batch_key is not None:
            obs_df["_scvi_batch"] = pd.Categorical(obs_df[self.batch_key]).codes
        obs_df = obs_df.loc[~obs_df._scvi_sample.duplicated("first")]
        return obs_df.set_index("_scvi_sample").sort_index()

    def _construct_design_matrix(self, sample_cov_keys, sample_mask, normalize_design_matrix, add_batch_specific_offsets,
                                 store_lfc, store_lfc_metadata_subset=None):
        """Reference `_model.py:1422-1529`: samples x covariates design matrix (categorical covariates one-hot encoded with
        the first level dropped, optional 0-1 normalisation, optional per-batch offset columns in front).  Returns
        (Xmat float32, column names, covariates_require_lfc mask, offset_indices or None)."""
        Xmat, Xmat_names, Xmat_dim_to_key = [], [], []
        sample_info = self.sample_info.iloc[sample_mask]
        for key in sample_cov_keys:
            cov = sample_info[key]
            if (cov.dtype == str) or isinstance(cov.dtype, pd.CategoricalDtype) or cov.dtype == object or pd.api.types.is_string_dtype(cov.dtype):
                cov = cov.astype("category").cat.remove_unused_categories()
                cov = pd.get_dummies(cov, drop_first=True)
                cov_names = np.array([f"{key}_{col}" for col in cov.columns])
                cov = cov.values
            else:
                cov_names = np.array([key])
                cov = cov.values[:, None]
            Xmat.append(cov)
            Xmat_names.append(cov_names)
            Xmat_dim_to_key.append([key] * cov.shape[1])
        Xmat_names = np.concatenate(Xmat_names)
        Xmat = np.concatenate(Xmat, axis=1).astype(np.float32)
        Xmat_dim_to_key = np.concatenate(Xmat_dim_to_key)
        if normalize_design_matrix:
            Xmat = (Xmat - Xmat.min(axis=0)) / (1e-6 + Xmat.max(axis=0) - Xmat.min(axis=0))
        offset_indices = None
        if add_batch_specific_offsets:
            if "_scvi_batch" not in sample_info:
                raise ValueError("add_batch_specific_offsets=True needs the model to be built with batch_key=...")
            n_batch = int(pd.Categorical(self.adata.obs[self.batch_key]).categories.size)
            cov = sample_info["_scvi_batch"]
            if cov.nunique() == n_batch:
                onehot = np.eye(n_batch)[cov.values]
                cov_names = ["offset_batch_" + str(i) for i in range(n_batch)]
                Xmat = np.concatenate([onehot, Xmat], axis=1).astype(np.float32)
                Xmat_names = np.concatenate([np.array(cov_names), Xmat_names])
                Xmat_dim_to_key = np.concatenate([np.array(cov_names), Xmat_dim_to_key])
                offset_indices = pd.Series(np.arange(len(Xmat_names)), index=Xmat_names).loc[cov_names].values
            else:
                warnings.warn(
                    "Number of batches in sample_info does not match number of batches in summary_stats. "
                    "`add_batch_specific_offsets=True` assumes that samples are not shared across batches. "
                    "Setting `add_batch_specific_offsets=False`...", stacklevel=2)
        if store_lfc:
            covariates_require_lfc = (np.isin(Xmat_dim_to_key, store_lfc_metadata_subset) if store_lfc_metadata_subset is not None
                                      else np.isin(Xmat_dim_to_key, sample_cov_keys))
        else:
            covariates_require_lfc = np.zeros(len(Xmat_names), dtype=bool)
        return Xmat, Xmat_names, covariates_require_lfc, offset_indices

    @staticmethod
    def _process_design_matrix(Xmat, admissible, lambd):
        """``process_design_matrix`` of the reference (`_model.py:1147-1160`) in float64 on the host: per distinct
        admissibility pattern D = diag(admissible), xtmx = X^T D X + lambd I; prefactor = real(sqrtm(xtmx)) and
        Amat = pinv(xtmx) X^T D.  Returns (Amat, prefactor) as (K, S) / (K, K) when every cell admits every sample,
        else per cell (n, K, S) / (n, K, K)."""
        X = np.asarray(Xmat, dtype=np.float64)
        K = X.shape[1]

        def solve(mask):   # (P, S) bool -> (P, K, S), (P, K, K)
            D = mask.astype(np.float64)
            xtmx = np.einsum("sk,ps,sm->pkm", X, D, X) + lambd * np.eye(K)
            w, V = np.linalg.eigh(xtmx)    # symmetric PSD: sqrtm = V sqrt(w) V^T (what the real part of the Schur sqrtm is)
            pref = np.einsum("pik,pk,pjk->pij", V, np.sqrt(np.clip(w, 0.0, None)), V)
            inv = np.linalg.pinv(xtmx, hermitian=True)
            return np.einsum("pab,sb,ps->pas", inv, X, D), pref

        adm = np.asarray(admissible, dtype=bool)
        if adm.all():
            A, P = solve(np.ones((1, X.shape[0]), dtype=bool))
            return A[0], P[0]
        pats, inverse = np.unique(adm, axis=0, return_inverse=True)
        A, P = solve(pats)
        inverse = np.asarray(inverse).reshape(-1)
        return A[inverse], P[inverse]

    def differential_expression(self, adata=None, sample_cov_keys: Optional[Sequence[str]] = None,
                                sample_subset: Optional[Sequence] = None, batch_size: int = 128, use_vmap: bool = True,
                                normalize_design_matrix: bool = True, add_batch_specific_offsets: bool = False,
                                mc_samples: int = 100, store_lfc: bool = False, store_lfc_metadata_subset=None,
                                store_baseline: bool = False, eps_lfc: float = 1e-4, filter_inadmissible_samples: bool = False,
                                lambd: float = 0.0, delta: Optional[float] = 0.3, admissible_samples=None,
                                **filter_samples_kwargs):
        """Reference `_model.py:1012-1420`, latent-space outputs: per cell, the counterfactual shifts eps_d of every kept
        sample (``mc_samples`` draws) are regressed on the sample covariates; returns ``beta`` (cell_name, covariate,
        latent_dim), ``effect_size``, ``pvalue`` (chi2, df = admissible samples of the cell) and ``padj``
        (Benjamini-Hochberg over all cells x covariates).  The device work (draws, q(z|u,s) chain, standardisation, the
        two einsums, `_model.py:1160-1218`) is one ``mrvi_run_de_host`` call; design-matrix algebra and the tests stay on
        the host, where the reference has them.

        Not part of this path (raise): ``store_lfc`` / ``store_baseline`` (decoder), ``filter_inadmissible_samples=True``
        (aggregated-posterior outlier detection, `_model.py:1121-1127`) -- pass its result as ``admissible_samples``
        (n_obs, n_kept_samples) bool instead.  ``batch_size`` / ``use_vmap`` only change memory use in the reference."""
        if store_lfc or store_baseline:
            raise NotImplementedError("store_lfc / store_baseline evaluate the decoder, which is outside the B200 path")
        if filter_inadmissible_samples:
            raise NotImplementedError("filter_inadmissible_samples=True runs get_outlier_cell_sample_pairs (outside the B200 "
                                      "path): pass its 'is_admissible' matrix as admissible_samples=")
        if not sample_cov_keys:
            raise ValueError("sample_cov_keys must name at least one sample covariate")
        self._check_if_trained(warn=False)
        if adata is not None:
            adata.obs["_indices"] = np.arange(adata.n_obs).astype(int)   # reference `_model.py:1102-1103`
        adata = self.adata if adata is None else adata
        n_sample = self.dims.n_sample
        sample_mask = np.isin(self.sample_order, sample_subset) if sample_subset is not None else np.ones(n_sample, dtype=bool)
        sample_mask = np.array(sample_mask)
        kept = np.where(sample_mask)[0]
        if admissible_samples is None:
            admissible = np.ones((adata.n_obs, len(kept)), dtype=bool)
        else:
            admissible = np.asarray(getattr(admissible_samples, "values", admissible_samples)).astype(bool)
            if admissible.shape != (adata.n_obs, len(kept)):
                raise ValueError(f"admissible_samples must have shape {(adata.n_obs, len(kept))}, not {admissible.shape}")
        n_admissible = admissible.sum(axis=1)
        Xmat, Xmat_names, _, _ = self._construct_design_matrix(
            sample_cov_keys=list(sample_cov_keys), sample_mask=sample_mask, normalize_design_matrix=normalize_design_matrix,
            add_batch_specific_offsets=add_batch_specific_offsets, store_lfc=False, store_lfc_metadata_subset=store_lfc_metadata_subset)
        Amat, prefactor = self._process_design_matrix(Xmat, admissible, float(lambd))
        sample_codes = self._sample_codes(adata)
        Xh = dense_float32(adata.X)
        res = self._handle.run_de_host(Xh, sample_codes, int(mc_samples), Amat, prefactor,
                                       kept_samples=None if sample_subset is None else kept, seed=self.seed)
        from scipy.stats import chi2

        effect = res["effect_size"]
        pvalue = (1.0 - chi2.cdf(effect.astype(np.float64), df=n_admissible[:, None])).astype(np.float32)   # `_model.py:1208`
        padj = fdr_bh(pvalue.reshape(-1)).reshape(pvalue.shape)                                           # `_model.py:1363`
        names = adata.obs_names.values
        cov = np.asarray(Xmat_names)
        data_vars = OrderedDict()
        data_vars["beta"] = DataArray(res["beta"], dims=["cell_name", "covariate", "latent_dim"],
                                      coords={"cell_name": names, "covariate": cov, "latent_dim": np.arange(res["beta"].shape[2])})
        for name, arr in (("effect_size", effect), ("pvalue", pvalue), ("padj", padj)):
            data_vars[name] = DataArray(arr, dims=["cell_name", "covariate"], coords={"cell_name": names, "covariate": cov})
        if admissible_samples is not None:
            data_vars["n_samples"] = DataArray(n_admissible, dims=["cell_name"], coords={"cell_name": names})
        return Dataset(data_vars)

    def get_latent_representation(self, adata=None, indices: Optional[Sequence[int]] = None, batch_size: Optional[int] = None,
                                  use_mean: bool = True, give_z: bool = False) -> np.ndarray:
        """Reference `_model.py:221-271`: ``u`` (n_obs, n_latent_u) or, with ``give_z``, ``z`` of every cell for
        its OWN sample (n_obs, n_latent).  A by-product of stages (1)+(2): ``u`` comes
        straight from the encoder; ``z`` is row ``sample_idx[cell]`` of the counterfactual tile."""
        self._check_if_trained(warn=False)
        adata = self.adata if adata is None else adata
        idx = np.arange(adata.n_obs) if indices is None else np.asarray(indices).astype(int).reshape(-1)
        sample_codes = self._sample_codes(adata)[idx]
        if not use_mean:
            # one draw u ~ q(u|x) per cell and z of that draw for the cell's own sample (reference `_module.py:315-331` with
            # mc_samples=None, cf_sample=None): row (cell, draw 0, sample = the cell's own) of the sampled path.  The draws
            # come from the device generator (seed = self.seed), not from JAX's streams: statistical parity.
            Xs = dense_float32(adata.X) if indices is None else dense_float32(adata.X[idx])
            res = self._handle.run_sampled_host(Xs, sample_codes, 1, seed=self.seed, keep_cell=False, want_z=give_z, want_u=not give_z)
            rows = np.arange(len(idx))
            return np.ascontiguousarray((res["z"] if give_z else res["u"])[rows, 0, sample_codes])
        Xh = self._host_matrix(adata.X, idx, indices is None)
        res = self._run_sharded(Xh, sample_codes, [], [], "l2", False, give_z, want_u=not give_z)
        if give_z:
            return np.ascontiguousarray(res["z"][np.arange(len(idx)), sample_codes])
        return res["u"]

    def get_local_sample_representation(self, adata=None, indices: Optional[Sequence[int]] = None, batch_size: int = 256,
                                        use_mean: bool = True, use_vmap: bool = True):
        """Reference `_model.py:586-625`: (cell_name, sample, latent_dim) counterfactual representations."""
        reductions = [
            MrVIReduction(
                name="sample_representations",
                input="mean_representations" if use_mean else "sampled_representations",
                fn=_identity,
                group_by=None,
            )
        ]
        return self.compute_local_statistics(reductions, adata=adata, indices=indices, batch_size=batch_size,
                                             use_vmap=use_vmap).sample_representations

    def get_local_sample_distances(self, adata=None, batch_size: int = 256, use_mean: bool = True,
                                   normalize_distances: bool = False, use_vmap: bool = True, groupby=None,
                                   keep_cell: bool = True, norm: str = "l2", mc_samples: int = 10):
        """Reference `_model.py:627-714`: per-cell (n_sample x n_sample) distances between counterfactual
        sample representations as variable ``cell`` and, per ``groupby`` key, their per-category means."""
        input = "mean_distances" if use_mean else "sampled_distances"
        if normalize_distances:
            if use_mean:
                warnings.warn(
                    "Normalizing distances uses sampled distances. Ignoring ``use_mean``.",
                    UserWarning,
                    stacklevel=2,
                )
            input = "normalized_distances"
        if groupby and not isinstance(groupby, list):
            groupby = [groupby]

        reductions = []
        if not keep_cell and not groupby:
            raise ValueError("Undefined computation because not keep_cell and no groupby.")
        if keep_cell:
            reductions.append(MrVIReduction(name="cell", input=input, fn=_identity))
        if groupby:
            for groupby_key in groupby:
                reductions.append(MrVIReduction(name=groupby_key, input=input, group_by=groupby_key))
        return self.compute_local_statistics(reductions, adata=adata, batch_size=batch_size, use_vmap=use_vmap,
                                             norm=norm, mc_samples=mc_samples)
