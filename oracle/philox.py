"""numpy mirror of the device noise generator of the sampled paths.  TEST INFRASTRUCTURE ONLY.

`mrvi_b200/csrc/sampled_kernels.cuh` draws standard normals with Philox4x32-10 (Salmon et al., SC'11; the same
round function as cuRAND's / numpy's Philox) keyed by the user seed, with the counter
``(global cell, a, b, (l // 4) | stream << 16)`` and a Box-Muller transform on 24-bit uniforms.  This file restates
it so that a test can hand the very same draws to the oracle.  (The reference's JAX threefry streams are a
different generator; they are not reproduced anywhere.)
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, seed: int):
    c = [np.asarray(v, dtype=np.uint64) & _MASK for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK, p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c


def normal_draws(gcell, a, b, n_lq: int, seed: int, stream: int):
    """Standard normals [..., n_lq] for broadcastable integer arrays (gcell, a, b)."""
    gcell, a, b = np.broadcast_arrays(np.asarray(gcell), np.asarray(a), np.asarray(b))
    nblk = (n_lq + 3) // 4
    out = np.empty(gcell.shape + (nblk * 4,), dtype=np.float32)
    for blk in range(nblk):
        x = philox4x32_10(gcell, a, b, np.uint64(blk | (stream << 16)), seed)
        u = [((v >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0) for v in x]
        r0 = np.sqrt(np.float32(-2.0) * np.log(u[0]))
        r1 = np.sqrt(np.float32(-2.0) * np.log(u[2]))
        t0, t1 = np.float32(2.0 * np.pi) * u[1], np.float32(2.0 * np.pi) * u[3]
        out[..., 4 * blk + 0] = r0 * np.cos(t0)
        out[..., 4 * blk + 1] = r0 * np.sin(t0)
        out[..., 4 * blk + 2] = r1 * np.cos(t1)
        out[..., 4 * blk + 3] = r1 * np.sin(t1)
    return out[..., :n_lq]


def u_noise(n_cells: int, n_sample: int, mc: int, n_lq: int, seed: int, cell_offset: int = 0):
    """(S, mc, n, Lq): the draws `mrvi_run_sampled_host` generates for the counterfactual rows (stream 0)."""
    s, m, c = np.meshgrid(np.arange(n_sample), np.arange(mc), np.arange(n_cells) + cell_offset, indexing="ij")
    return normal_draws(c, m, s, n_lq, seed, 0)


def base_noise(n_cells: int, mc: int, n_lq: int, seed: int, cell_offset: int = 0):
    """(2 mc, n, Lq): the draws of the local baseline (stream 1; b = 0)."""
    a, c = np.meshgrid(np.arange(2 * mc), np.arange(n_cells) + cell_offset, indexing="ij")
    return normal_draws(c, a, np.zeros_like(a), n_lq, seed, 1)


def eps_noise(n_cells: int, n_sample: int, mc: int, n_latent: int, seed: int, cell_offset: int = 0):
    """(S, mc, n, L): the draws of eps for ``use_map=False`` models on the counterfactual rows (stream 2)."""
    s, m, c = np.meshgrid(np.arange(n_sample), np.arange(mc), np.arange(n_cells) + cell_offset, indexing="ij")
    return normal_draws(c, m, s, n_latent, seed, 2)


def base_eps_noise(n_cells: int, mc: int, n_latent: int, seed: int, cell_offset: int = 0):
    """(2 mc, n, L): the draws of eps on the baseline rows (stream 3; b = 0)."""
    a, c = np.meshgrid(np.arange(2 * mc), np.arange(n_cells) + cell_offset, indexing="ij")
    return normal_draws(c, a, np.zeros_like(a), n_latent, seed, 3)
