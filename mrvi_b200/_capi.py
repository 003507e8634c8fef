"""ctypes binding of ``include/mrvi_b200.h`` (the C-ABI shared library ``libmrvi_b200.so``).

There is no CPU fallback: if the library is missing or no sm_100 GPU is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Mapping, Optional, Sequence

import numpy as np

from ._params import MrviDims

ABI_VERSION = 5
PRECISION_FP32 = 0
PRECISION_TC = 1
NORMS = {"l2": 0, "l1": 1, "linf": 2}

_LIB_NAME = "libmrvi_b200.so"
_lib = None

#: every symbol declared in include/mrvi_b200.h (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = (
    "mrvi_abi_version",
    "mrvi_last_error",
    "mrvi_kernel_launch_count",
    "mrvi_device_count",
    "mrvi_create",
    "mrvi_destroy",
    "mrvi_run_device",
    "mrvi_run_host",
    "mrvi_run_host_csr",
    "mrvi_run_host_counts",
    "mrvi_run_sampled_host",
    "mrvi_run_de_host",
    "mrvi_distances_device",
    "mrvi_last_stage_ms",
    "mrvi_last_h2d_bytes",
    "mrvi_attention_table_info",
    "mrvi_narrow_counts",
    "mrvi_comm_unique_id",
    "mrvi_comm_create",
    "mrvi_allreduce_groups",
    "mrvi_alloc_pinned",
    "mrvi_free_pinned",
)

#: MrviXDtype of include/mrvi_b200.h: integer count matrices the host entry points take as they are
_X_DTYPES = {np.dtype(np.uint8): 1, np.dtype(np.uint16): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 4}


class MrviError(RuntimeError):
    """A non-zero status from the C-ABI (message from ``mrvi_last_error``)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"mrvi_b200 C-ABI error {status}: {message}")
        self.status = status


class _CDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_input", "n_sample", "n_latent", "n_latent_u", "encoder_n_hidden", "encoder_n_layers",
        "n_latent_sample", "n_channels", "n_heads", "qz_n_hidden", "qz_n_layers", "use_map")]


def library_path() -> str:
    # MRVI_B200_LIBRARY: an alternative build of the same library (kernel A/B experiments, debug builds)
    return os.environ.get("MRVI_B200_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """dlopen the in-tree ``libmrvi_b200.so``; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise MrviError(-100, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(mrvi_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.mrvi_abi_version.restype = C.c_int
    lib.mrvi_last_error.restype = C.c_char_p
    lib.mrvi_kernel_launch_count.restype = i64
    lib.mrvi_device_count.restype = C.c_int
    lib.mrvi_create.restype = C.c_int
    lib.mrvi_create.argtypes = [C.POINTER(_CDims), i32, C.POINTER(C.c_char_p), C.POINTER(vp),
                                C.POINTER(i64), i32, i32, C.POINTER(vp)]
    lib.mrvi_destroy.restype = C.c_int
    lib.mrvi_destroy.argtypes = [vp]
    lib.mrvi_run_device.restype = C.c_int
    lib.mrvi_run_device.argtypes = [vp, i64, vp, i64, vp, i32, C.POINTER(vp), C.POINTER(i32), i32, vp,
                                    C.POINTER(vp), vp, vp, vp]
    lib.mrvi_run_host.restype = C.c_int
    lib.mrvi_run_host.argtypes = [vp, i64, vp, i64, vp, i32, C.POINTER(vp), C.POINTER(i32), i32, vp,
                                  C.POINTER(vp), C.POINTER(vp), vp, vp, i64]
    lib.mrvi_run_host_csr.restype = C.c_int
    lib.mrvi_run_host_csr.argtypes = [vp, i64, vp, vp, vp, vp, i32, C.POINTER(vp), C.POINTER(i32), i32, vp,
                                      C.POINTER(vp), C.POINTER(vp), vp, vp, i64]
    lib.mrvi_run_host_counts.restype = C.c_int
    lib.mrvi_run_host_counts.argtypes = [vp, i64, vp, i32, i64, vp, i32, C.POINTER(vp), C.POINTER(i32), i32, vp,
                                         C.POINTER(vp), C.POINTER(vp), vp, vp, i64]
    lib.mrvi_run_sampled_host.restype = C.c_int
    lib.mrvi_run_sampled_host.argtypes = [vp, i64, vp, i64, vp, i32, C.c_uint64, i64, vp, vp, vp, vp, i32, C.POINTER(vp),
                                          C.POINTER(i32), i32, i32, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, i64]
    lib.mrvi_run_de_host.restype = C.c_int
    lib.mrvi_run_de_host.argtypes = [vp, i64, vp, i64, vp, i32, C.c_uint64, i64, vp, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, i64]
    lib.mrvi_distances_device.restype = C.c_int
    lib.mrvi_distances_device.argtypes = [vp, i64, vp, i32, C.POINTER(vp), C.POINTER(i32), i32, vp,
                                          C.POINTER(vp), vp]
    lib.mrvi_attention_table_info.restype = C.c_int
    lib.mrvi_attention_table_info.argtypes = [vp, vp, vp, vp, vp]
    lib.mrvi_last_h2d_bytes.restype = i64
    lib.mrvi_last_h2d_bytes.argtypes = [vp]
    lib.mrvi_narrow_counts.restype = C.c_int
    lib.mrvi_narrow_counts.argtypes = [vp, i64, i64, i32, vp, i32]
    lib.mrvi_last_stage_ms.restype = C.c_int
    lib.mrvi_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.mrvi_comm_unique_id.restype = C.c_int
    lib.mrvi_comm_unique_id.argtypes = [vp]
    lib.mrvi_comm_create.restype = C.c_int
    lib.mrvi_comm_create.argtypes = [vp, vp, i32, i32]
    lib.mrvi_allreduce_groups.restype = C.c_int
    lib.mrvi_allreduce_groups.argtypes = [vp, vp, i32, C.POINTER(vp), C.POINTER(i32), vp]
    lib.mrvi_alloc_pinned.restype = C.c_int
    lib.mrvi_alloc_pinned.argtypes = [i64, C.POINTER(vp)]
    lib.mrvi_free_pinned.restype = C.c_int
    lib.mrvi_free_pinned.argtypes = [vp]
    if lib.mrvi_abi_version() != ABI_VERSION:
        raise MrviError(-101, f"ABI version mismatch: library {lib.mrvi_abi_version()}, binding {ABI_VERSION}")
    _lib = lib
    return lib


def _check(status: int):
    if status != 0:
        raise MrviError(status, load_library().mrvi_last_error().decode("utf-8", "replace"))


def _ptr_array(ptrs: Sequence[int]):
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def _i32_array(vals: Sequence[int]):
    arr = (C.c_int32 * max(len(vals), 1))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def narrow_counts(X: np.ndarray, n_threads: int = 4):
    """The host-side lossless narrowing ``mrvi_run_host`` applies to dense count chunks (no GPU needed).
    Returns ``(mode, array)``: mode 1 -> uint8, 2 -> uint16, 0 -> not narrowable (array is None)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, G = X.shape
    out = np.empty(n * G * 2, dtype=np.uint8)
    mode = load_library().mrvi_narrow_counts(X.ctypes.data, G, n, G, out.ctypes.data, int(n_threads))
    if mode < 0:
        _check(mode)
    if mode == 1:
        return 1, out[: n * G].reshape(n, G)
    if mode == 2:
        return 2, out.view(np.uint16)[: n * G].reshape(n, G)
    return 0, None


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """An uninitialised ndarray in page-locked host memory (``mrvi_alloc_pinned``): device->host copies into it run at
    PCIe speed and overlap the kernels.  Freed when the array (and every view of it) is garbage collected."""
    import weakref

    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if len(shape) else 1
    nbytes = max(n * dtype.itemsize, 1)
    lib = load_library()
    ptr = C.c_void_p()
    _check(lib.mrvi_alloc_pinned(nbytes, C.byref(ptr)))
    buf = (C.c_uint8 * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    weakref.finalize(buf, lib.mrvi_free_pinned, C.c_void_p(ptr.value))
    return arr


def comm_unique_id() -> bytes:
    """128 bytes identifying a new NCCL communicator (``mrvi_comm_unique_id``); distribute them to every rank."""
    buf = (C.c_uint8 * 128)()
    _check(load_library().mrvi_comm_unique_id(buf))
    return bytes(buf)


def device_count() -> int:
    """Number of CUDA devices visible to the library (``mrvi_device_count``)."""
    return int(load_library().mrvi_device_count())


def kernel_launch_count() -> int:
    return int(load_library().mrvi_kernel_launch_count())


def _out_empty(shape) -> np.ndarray:
    """Output buffer of ``run_host``: page-locked above 16 MiB (the D2H stream then runs at PCIe speed), else plain."""
    if int(np.prod(shape)) * 4 >= (16 << 20):
        try:
            return pinned_empty(shape, np.float32)
        except MrviError:
            pass
    return np.empty(shape, dtype=np.float32)


class Handle:
    """Owns one ``MrviHandle`` (weights + workspaces on one GPU)."""

    def __init__(self, dims: MrviDims, flat_params: Mapping[str, np.ndarray], device: int = 0,
                 precision: int = PRECISION_FP32):
        lib = load_library()
        self.dims = dims
        self.device = int(device)
        self.precision = int(precision)
        cd = _CDims(dims.n_input, dims.n_sample, dims.n_latent, dims.n_latent_u or 0, dims.encoder_n_hidden,
                    dims.encoder_n_layers, dims.n_latent_sample, dims.n_channels, dims.n_heads,
                    dims.qz_n_hidden, dims.qz_n_layers, int(dims.use_map))
        names = list(flat_params.keys())
        arrays = [np.ascontiguousarray(flat_params[k], dtype=np.float32) for k in names]
        c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
        c_vals = _ptr_array([a.ctypes.data for a in arrays])
        c_numels = (C.c_int64 * len(names))(*[a.size for a in arrays])
        out = C.c_void_p()
        self._h = None
        _check(lib.mrvi_create(C.byref(cd), len(names), c_names, c_vals, c_numels, self.device,
                               self.precision, C.byref(out)))
        self._h = out

    def close(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.mrvi_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- device-resident pass: integer device addresses (e.g. torch.Tensor.data_ptr())
    def run_device(self, n_cells: int, X_ptr: int, ldx: int, sample_ptr: int, group_code_ptrs: Sequence[int] = (),
                   n_groups: Sequence[int] = (), norm: str = "l2", cell_out_ptr: int = 0,
                   group_sum_ptrs: Sequence[int] = (), z_out_ptr: int = 0, u_out_ptr: int = 0, stream: int = 0):
        if norm not in NORMS:
            raise ValueError(f"norm {norm} not supported")
        _check(load_library().mrvi_run_device(
            self._h, int(n_cells), X_ptr, int(ldx), sample_ptr, len(n_groups), _ptr_array(group_code_ptrs),
            _i32_array(n_groups), NORMS[norm], cell_out_ptr or None, _ptr_array(group_sum_ptrs),
            z_out_ptr or None, u_out_ptr or None, stream or None))

    def distances_device(self, n_cells: int, z_ptr: int, group_code_ptrs: Sequence[int] = (),
                         n_groups: Sequence[int] = (), norm: str = "l2", cell_out_ptr: int = 0,
                         group_sum_ptrs: Sequence[int] = (), stream: int = 0):
        if norm not in NORMS:
            raise ValueError(f"norm {norm} not supported")
        _check(load_library().mrvi_distances_device(
            self._h, int(n_cells), z_ptr, len(n_groups), _ptr_array(group_code_ptrs), _i32_array(n_groups),
            NORMS[norm], cell_out_ptr or None, _ptr_array(group_sum_ptrs), stream or None))

    # ---- host-buffer pass: numpy arrays in, numpy arrays out
    def run_host(self, X: np.ndarray, sample_idx: np.ndarray, group_codes: Sequence[np.ndarray] = (),
                 n_groups: Sequence[int] = (), norm: str = "l2", keep_cell: bool = True, want_z: bool = False,
                 want_u: bool = False, chunk_cells: int = 0, cell_out: Optional[np.ndarray] = None,
                 z_out: Optional[np.ndarray] = None):
        if norm not in NORMS:
            raise ValueError(f"norm {norm} not supported")
        d = self.dims
        csr = None
        if hasattr(X, "tocsr") and not isinstance(X, np.ndarray):  # scipy sparse: ship the non-zeros, densify on the GPU
            X = X.tocsr()
            if not X.has_canonical_format:
                X = X.copy()
                X.sum_duplicates()
            csr = (np.ascontiguousarray(X.indptr, dtype=np.int64), np.ascontiguousarray(X.indices, dtype=np.int32),
                   np.ascontiguousarray(X.data, dtype=np.float32))
        elif X.dtype in _X_DTYPES and X.ndim == 2:
            # raw counts stored as integers: uint8 / uint16 cross PCIe as they are, int32 / int64 are narrowed chunk by
            # chunk inside the library (no float32 copy of the matrix); widened to float32 on the GPU
            if X.strides[1] != X.itemsize or X.strides[0] % X.itemsize != 0 or X.strides[0] < 0:
                X = np.ascontiguousarray(X)
        elif X.dtype != np.float32 or X.ndim != 2 or X.strides[1] != 4 or X.strides[0] % 4 != 0:
            X = np.ascontiguousarray(X, dtype=np.float32)
        n = X.shape[0]
        if X.shape[1] != d.n_input:
            raise ValueError(f"X has {X.shape[1]} genes, model expects {d.n_input}")
        ldx = d.n_input if (csr is not None or n <= 1) else X.strides[0] // X.itemsize
        sid = np.ascontiguousarray(sample_idx, dtype=np.int32).reshape(-1)
        if sid.shape[0] != n:
            raise ValueError("sample_idx length does not match X")
        if n and (sid.min() < 0 or sid.max() >= d.n_sample):
            raise ValueError("sample_idx outside [0, n_sample)")
        codes = [np.ascontiguousarray(c, dtype=np.int32).reshape(-1) for c in group_codes]
        S, L, Lq = d.n_sample, d.n_latent, d.n_latent_q
        out = {}
        if keep_cell:
            if cell_out is None:
                cell_out = _out_empty((n, S, S))
            elif cell_out.shape != (n, S, S) or cell_out.dtype != np.float32 or not cell_out.flags.c_contiguous:
                raise ValueError("cell_out must be a C-contiguous float32 array of shape (n, S, S)")
            out["cell"] = cell_out
        sums = [np.zeros((g, S, S), dtype=np.float64) for g in n_groups]
        counts = [np.zeros((g,), dtype=np.int64) for g in n_groups]
        z = (z_out if z_out is not None else _out_empty((n, S, L))) if want_z else None
        if want_z and (z.shape != (n, S, L) or z.dtype != np.float32 or not z.flags.c_contiguous):
            raise ValueError("z_out must be a C-contiguous float32 array of shape (n, S, L)")
        u = np.empty((n, Lq), dtype=np.float32) if want_u else None
        tail = (sid.ctypes.data, len(codes), _ptr_array([c.ctypes.data for c in codes]),
                _i32_array(n_groups), NORMS[norm], out["cell"].ctypes.data if keep_cell else None,
                _ptr_array([s.ctypes.data for s in sums]), _ptr_array([c.ctypes.data for c in counts]),
                z.ctypes.data if want_z else None, u.ctypes.data if want_u else None, int(chunk_cells))
        if csr is not None:
            _check(load_library().mrvi_run_host_csr(self._h, n, csr[0].ctypes.data, csr[1].ctypes.data, csr[2].ctypes.data, *tail))
        elif X.dtype in _X_DTYPES:
            _check(load_library().mrvi_run_host_counts(self._h, n, X.ctypes.data, _X_DTYPES[X.dtype], ldx, *tail))
        else:
            _check(load_library().mrvi_run_host(self._h, n, X.ctypes.data, ldx, *tail))
        out["group_sums"] = sums
        out["group_counts"] = counts
        if want_z:
            out["z"] = z
        if want_u:
            out["u"] = u
        return out

    # ---- sampled paths (use_mean=False / normalize_distances=True), host buffers
    def run_sampled_host(self, X: np.ndarray, sample_idx: np.ndarray, mc_samples: int, seed: int = 0,
                         cell_offset: int = 0, u_noise: Optional[np.ndarray] = None,
                         base_noise: Optional[np.ndarray] = None, group_codes: Sequence[np.ndarray] = (),
                         n_groups: Sequence[int] = (), norm: str = "l2", normalize: bool = False,
                         keep_cell: bool = True, want_z: bool = False, chunk_cells: int = 0,
                         eps_noise: Optional[np.ndarray] = None, base_eps_noise: Optional[np.ndarray] = None,
                         want_u: bool = False):
        """``mrvi_run_sampled_host``: distances on ``mc_samples`` draws of u per (cell, counterfactual sample).
        ``u_noise`` (S, mc, n, Lq) / ``base_noise`` (2 mc, n, Lq): explicit standard-normal draws (tests); default
        is the device Philox generator keyed by ``seed`` and the global cell index ``cell_offset + i``.
        ``eps_noise`` (S, mc, n, L) / ``base_eps_noise`` (2 mc, n, L): draws of eps for ``use_map=False`` models."""
        if norm not in NORMS:
            raise ValueError(f"norm {norm} not supported")
        d = self.dims
        if hasattr(X, "toarray") and not isinstance(X, np.ndarray):
            X = X.toarray()
        if X.dtype != np.float32 or X.ndim != 2 or X.strides[1] != 4 or X.strides[0] % 4 != 0:
            X = np.ascontiguousarray(X, dtype=np.float32)
        n = X.shape[0]
        if X.shape[1] != d.n_input:
            raise ValueError(f"X has {X.shape[1]} genes, model expects {d.n_input}")
        ldx = d.n_input if n <= 1 else X.strides[0] // 4
        sid = np.ascontiguousarray(sample_idx, dtype=np.int32).reshape(-1)
        if sid.shape[0] != n:
            raise ValueError("sample_idx length does not match X")
        if n and (sid.min() < 0 or sid.max() >= d.n_sample):
            raise ValueError("sample_idx outside [0, n_sample)")
        mc, S, L, Lq = int(mc_samples), d.n_sample, d.n_latent, d.n_latent_q
        if u_noise is not None:
            u_noise = np.ascontiguousarray(u_noise, dtype=np.float32)
            if u_noise.shape != (S, mc, n, Lq):
                raise ValueError(f"u_noise must have shape {(S, mc, n, Lq)}")
        if base_noise is not None:
            base_noise = np.ascontiguousarray(base_noise, dtype=np.float32)
            if base_noise.shape != (2 * mc, n, Lq):
                raise ValueError(f"base_noise must have shape {(2 * mc, n, Lq)}")
        if eps_noise is not None:
            eps_noise = np.ascontiguousarray(eps_noise, dtype=np.float32)
            if eps_noise.shape != (S, mc, n, L):
                raise ValueError(f"eps_noise must have shape {(S, mc, n, L)}")
        if base_eps_noise is not None:
            base_eps_noise = np.ascontiguousarray(base_eps_noise, dtype=np.float32)
            if base_eps_noise.shape != (2 * mc, n, L):
                raise ValueError(f"base_eps_noise must have shape {(2 * mc, n, L)}")
        codes = [np.ascontiguousarray(c, dtype=np.int32).reshape(-1) for c in group_codes]
        out = {}
        cell_shape = (n, S, S) if normalize else (n, mc, S, S)
        cell = np.empty(cell_shape, dtype=np.float32) if keep_cell else None
        sums = [np.zeros((g, S, S) if normalize else (g, mc, S, S), dtype=np.float64) for g in n_groups]
        counts = [np.zeros((g,), dtype=np.int64) for g in n_groups]
        z = np.empty((n, mc, S, L), dtype=np.float32) if want_z else None
        u = np.empty((n, mc, S, Lq), dtype=np.float32) if want_u else None
        _check(load_library().mrvi_run_sampled_host(
            self._h, n, X.ctypes.data, ldx, sid.ctypes.data, mc, int(seed) & (2 ** 64 - 1), int(cell_offset),
            u_noise.ctypes.data if u_noise is not None else None,
            base_noise.ctypes.data if base_noise is not None else None,
            eps_noise.ctypes.data if eps_noise is not None else None,
            base_eps_noise.ctypes.data if base_eps_noise is not None else None, len(codes),
            _ptr_array([c.ctypes.data for c in codes]), _i32_array(n_groups), NORMS[norm], int(bool(normalize)),
            cell.ctypes.data if keep_cell else None, _ptr_array([s.ctypes.data for s in sums]),
            _ptr_array([c.ctypes.data for c in counts]), z.ctypes.data if want_z else None,
            u.ctypes.data if want_u else None, int(chunk_cells)))
        if keep_cell:
            out["cell"] = cell
        out["group_sums"] = sums
        out["group_counts"] = counts
        if want_z:
            out["z"] = z
        if want_u:
            out["u"] = u
        return out

    def run_de_host(self, X: np.ndarray, sample_idx: np.ndarray, mc_samples: int, Amat: np.ndarray, prefactor: np.ndarray,
                    kept_samples: Optional[np.ndarray] = None, seed: int = 0, cell_offset: int = 0,
                    u_noise: Optional[np.ndarray] = None, want_eps: bool = False, chunk_cells: int = 0,
                    eps_noise: Optional[np.ndarray] = None):
        """``mrvi_run_de_host``: the latent-space stage of ``differential_expression`` (reference `_model.py:1160-1218`).
        ``Amat`` (K, S_kept) or per cell (n, K, S_kept); ``prefactor`` (K, K) or (n, K, K); returns ``beta`` (n, K, L)
        and ``effect_size`` (n, K) (and the raw residuals ``eps`` (n, mc, S, L) with ``want_eps``)."""
        d = self.dims
        if hasattr(X, "toarray") and not isinstance(X, np.ndarray):
            X = X.toarray()
        if X.dtype != np.float32 or X.ndim != 2 or X.strides[1] != 4 or X.strides[0] % 4 != 0:
            X = np.ascontiguousarray(X, dtype=np.float32)
        n = X.shape[0]
        if X.shape[1] != d.n_input:
            raise ValueError(f"X has {X.shape[1]} genes, model expects {d.n_input}")
        ldx = d.n_input if n <= 1 else X.strides[0] // 4
        sid = np.ascontiguousarray(sample_idx, dtype=np.int32).reshape(-1)
        if sid.shape[0] != n:
            raise ValueError("sample_idx length does not match X")
        mc, S, L, Lq = int(mc_samples), d.n_sample, d.n_latent, d.n_latent_q
        kept = None if kept_samples is None else np.ascontiguousarray(kept_samples, dtype=np.int32).reshape(-1)
        Sk = S if kept is None else kept.shape[0]
        Amat = np.ascontiguousarray(Amat, dtype=np.float32)
        prefactor = np.ascontiguousarray(prefactor, dtype=np.float32)
        per_cell = Amat.ndim == 3
        K = Amat.shape[-2]
        if Amat.shape != ((n, K, Sk) if per_cell else (K, Sk)):
            raise ValueError(f"Amat must have shape (n_covariates, {Sk}) or ({n}, n_covariates, {Sk}), not {Amat.shape}")
        if prefactor.shape != ((n, K, K) if per_cell else (K, K)):
            raise ValueError(f"prefactor must have shape {((n, K, K) if per_cell else (K, K))}, not {prefactor.shape}")
        if u_noise is not None:
            u_noise = np.ascontiguousarray(u_noise, dtype=np.float32)
            if u_noise.shape != (S, mc, n, Lq):
                raise ValueError(f"u_noise must have shape {(S, mc, n, Lq)}")
        if eps_noise is not None:
            eps_noise = np.ascontiguousarray(eps_noise, dtype=np.float32)
            if eps_noise.shape != (S, mc, n, L):
                raise ValueError(f"eps_noise must have shape {(S, mc, n, L)}")
        beta = np.empty((n, K, L), dtype=np.float32)
        effect = np.empty((n, K), dtype=np.float32)
        eps = np.empty((n, mc, S, L), dtype=np.float32) if want_eps else None
        _check(load_library().mrvi_run_de_host(
            self._h, n, X.ctypes.data, ldx, sid.ctypes.data, mc, int(seed) & (2 ** 64 - 1), int(cell_offset),
            u_noise.ctypes.data if u_noise is not None else None, eps_noise.ctypes.data if eps_noise is not None else None, Sk,
            kept.ctypes.data if kept is not None else None, K,
            Amat.ctypes.data, prefactor.ctypes.data, int(per_cell), beta.ctypes.data, effect.ctypes.data,
            eps.ctypes.data if want_eps else None, int(chunk_cells)))
        out = {"beta": beta, "effect_size": effect}
        if want_eps:
            out["eps"] = eps
        return out

    def attention_table_info(self) -> dict:
        """``mrvi_attention_table_info``: intervals (0 = direct softmax), tabulated t range, verified interpolation error."""
        n = C.c_int32(0)
        t0, t1, err = C.c_float(0), C.c_float(0), C.c_float(0)
        _check(load_library().mrvi_attention_table_info(self._h, C.byref(n), C.byref(t0), C.byref(t1), C.byref(err)))
        return {"n_intervals": int(n.value), "t_min": float(t0.value), "t_max": float(t1.value), "max_rel_err": float(err.value)}

    # ---- multi-GPU: the one exchange step of the path
    def comm_create(self, unique_id: bytes, n_ranks: int, rank: int):
        """``ncclCommInitRank`` on this handle's device (``unique_id`` from :func:`comm_unique_id` of one rank)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(load_library().mrvi_comm_create(self._h, buf, int(n_ranks), int(rank)))

    def allreduce_groups(self, group_sum_ptrs: Sequence[int], n_groups: Sequence[int], stream: int = 0, nccl_comm: int = 0):
        """In-place NCCL all-reduce (sum, float64) of the per-key group sums in DEVICE memory."""
        _check(load_library().mrvi_allreduce_groups(self._h, nccl_comm or None, len(n_groups), _ptr_array(group_sum_ptrs),
                                                    _i32_array(n_groups), stream or None))

    def last_h2d_bytes(self) -> int:
        """Bytes of X the last ``run_host`` copied host->device (count chunks are narrowed to uint8 / uint16)."""
        return int(load_library().mrvi_last_h2d_bytes(self._h))

    def last_stage_ms(self):
        buf = (C.c_float * 4)()
        _check(load_library().mrvi_last_stage_ms(self._h, buf))
        return {"encoder": buf[0], "qz_chain": buf[1], "distances": buf[2], "total": buf[3]}
