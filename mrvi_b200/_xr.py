"""Output containers with the reference's xarray layout.

The reference returns ``xr.Dataset`` / ``xr.DataArray`` (`_model.py:396-404, 561-570, 506`).  When
xarray is importable the real classes are used, so the result is a drop-in.  xarray is not
installed in the build image, so a minimal same-layout stand-in (dims, coords, name, values,
label-based ``sel``, ``sum``/``mean`` over a dim) is provided; it holds numpy arrays without copying.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - exercised only where xarray exists
    import xarray as _xarray

    HAVE_XARRAY = True
except Exception:  # ModuleNotFoundError in the build image
    _xarray = None
    HAVE_XARRAY = False


class _LiteDataArray:
    """Tiny subset of ``xarray.DataArray``: labelled numpy array."""

    def __init__(self, data, dims=None, coords=None, name=None):
        if isinstance(data, _LiteDataArray):
            dims = data.dims if dims is None else dims
            coords = data.coords if coords is None else coords
            name = data.name if name is None else name
            data = data.values
        self.values = np.asarray(data)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(self.values.ndim))
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError(f"dims {self.dims} do not match array of ndim {self.values.ndim}")
        self.coords = OrderedDict()
        for k, v in (coords or {}).items():
            v = np.asarray(v)
            if k in self.dims and v.ndim == 1 and v.shape[0] != self.values.shape[self.dims.index(k)]:
                raise ValueError(f"coordinate {k!r} has length {v.shape[0]}, dimension has "
                                 f"{self.values.shape[self.dims.index(k)]}")
            self.coords[k] = v
        self.name = name

    # numpy interop
    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def ndim(self):
        return self.values.ndim

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __len__(self):
        return self.values.shape[0]

    def _sub_coords(self, dim, idx):
        c = OrderedDict()
        for k, v in self.coords.items():
            if k == dim:
                if np.ndim(idx) == 0 and not isinstance(idx, slice):
                    continue
                c[k] = v[idx]
            else:
                c[k] = v
        return c

    def isel(self, **indexers):
        out = self
        for dim, idx in indexers.items():
            ax = out.dims.index(dim)
            sl = [slice(None)] * out.values.ndim
            sl[ax] = idx
            vals = out.values[tuple(sl)]
            scalar = np.ndim(idx) == 0 and not isinstance(idx, slice)
            dims = tuple(d for d in out.dims if not (scalar and d == dim))
            out = _LiteDataArray(vals, dims, out._sub_coords(dim, idx), out.name)
        return out

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        out = self
        # apply from the last axis so earlier positions stay valid when scalars drop dims
        for ax in reversed(range(len(key))):
            k = key[ax]
            if isinstance(k, slice) and k == slice(None):
                continue
            out = out.isel(**{self.dims[ax]: k})
        return out

    def sel(self, **indexers):
        out = self
        for dim, labels in indexers.items():
            coord = out.coords[dim]
            lookup = {}
            for i, c in enumerate(coord.tolist()):
                lookup.setdefault(c, i)
            if np.ndim(labels) == 0:
                idx = lookup[labels.item() if hasattr(labels, "item") else labels]
            else:
                idx = np.array([lookup[l] for l in np.asarray(labels).tolist()], dtype=np.int64)
            out = out.isel(**{dim: idx})
        return out

    def _reduce(self, fn, dim):
        ax = self.dims.index(dim)
        dims = tuple(d for d in self.dims if d != dim)
        coords = OrderedDict((k, v) for k, v in self.coords.items() if k != dim)
        return _LiteDataArray(fn(self.values, axis=ax), dims, coords, self.name)

    def sum(self, dim):
        return self._reduce(np.sum, dim)

    def mean(self, dim):
        return self._reduce(np.mean, dim)

    @property
    def T(self):
        return _LiteDataArray(self.values.T, self.dims[::-1], self.coords, self.name)

    def __repr__(self):
        dims = ", ".join(f"{d}: {n}" for d, n in zip(self.dims, self.shape))
        return f"<mrvi_b200 DataArray {self.name!r} ({dims}) {self.dtype}>"


class _LiteDataset:
    """Tiny subset of ``xarray.Dataset``: ordered mapping name -> DataArray."""

    def __init__(self, data_vars=None):
        self.data_vars = OrderedDict(data_vars or {})

    def __getitem__(self, k):
        return self.data_vars[k]

    def __getattr__(self, k):
        dv = self.__dict__.get("data_vars", {})
        if k in dv:
            return dv[k]
        raise AttributeError(k)

    def __contains__(self, k):
        return k in self.data_vars

    def __iter__(self):
        return iter(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def items(self):
        return self.data_vars.items()

    def __repr__(self):
        return "<mrvi_b200 Dataset\n  " + "\n  ".join(repr(v) for v in self.data_vars.values()) + ">"


if HAVE_XARRAY:  # pragma: no cover
    DataArray = _xarray.DataArray
    Dataset = _xarray.Dataset
else:
    DataArray = _LiteDataArray
    Dataset = _LiteDataset
