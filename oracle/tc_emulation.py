"""Numpy emulation of the tensor-core q(z|u,s) chain (design study; TEST INFRASTRUCTURE ONLY).

Emulates exactly the algebra the tcgen05 kernel executes (DESIGN.md "TC chain"): the attention block
collapsed to softmax-weighted means m[h,i] of the per-sample kv tokens, attention-output / bias /
skip-connection folding into the neighbouring GEMMs, 16-bit operand rounding with fp32
accumulation, fp32 LayerNorm / gelu epilogues.  Used to choose the operand format (f16 vs bf16) and
to verify offline that the folded algebra equals the oracle (with rounding switched off it must
agree to ~1e-6).
"""
from __future__ import annotations

import numpy as np

from . import mrvi_oracle as orc

R = 128


def _round(a, fmt):
    if fmt is None:
        return a.astype(np.float32)
    if fmt == "f16":
        return a.astype(np.float16).astype(np.float32)
    if fmt == "f16x2":  # (hi, lo) pair of f16: what the kernel's weight image holds since round 2 (3-term products)
        a32 = np.asarray(a, dtype=np.float64)
        hi = a32.astype(np.float16).astype(np.float64)
        lo = (a32 - hi).astype(np.float16).astype(np.float64)  # subnormal for the smallest weights, as on the device
        return (hi + lo).astype(np.float32)
    if fmt == "bf16":
        b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        b = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
        return b.view(np.float32)
    if fmt == "tf32":
        b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        b = ((b + 0xFFF + ((b >> 13) & 1)) & 0xFFFFE000).astype(np.uint32)
        return b.view(np.float32)
    raise ValueError(fmt)


def fold_weights(flat):
    """All host-side weight folding (float64).  Returns the GEMM operand matrices [K, N] and the
    fp32 epilogue vectors.  K layouts:  A1 = [m(h*E+i) | 1],  A3 = [t2 | u_ln | 1],  A5 = [t4 | 1]
    (the kernel stores A3 as [t2 | 1 | u_ln]; the order of the K columns does not change the arithmetic)."""
    p = {k: np.asarray(v, dtype=np.float64) for k, v in flat.items()}
    ab = "qz/AttentionBlock_0"
    mha = f"{ab}/MultiHeadDotProductAttention_0"
    wq, bq = p[f"{mha}/query/kernel"][0], p[f"{mha}/query/bias"]  # (nh, hd)
    wk, bk = p[f"{mha}/key/kernel"][0], p[f"{mha}/key/bias"]
    wv, bv = p[f"{mha}/value/kernel"][0], p[f"{mha}/value/bias"]
    wo, bo = p[f"{mha}/out/kernel"], p[f"{mha}/out/bias"]  # (nh, hd, C), (C)
    nh, hd = wq.shape
    C = bo.shape[0]
    E = p[f"{ab}/DenseGeneral_1/kernel"].shape[0]
    inv = 1.0 / np.sqrt(hd)
    f = {}
    # score[h,i,j] = t[h,i] * kv[j] + const(h,i);  t[h,i] = alpha[h] * qtok[i] + beta[h]
    f["alpha"] = (wq * wk).sum(1) * inv
    f["beta"] = (bq * wk).sum(1) * inv
    pv = np.einsum("hd,hdc->hc", wv, wo)  # o[i,c] = sum_h m[h,i] pv[h,c] + ro[c]
    ro = np.einsum("hd,hdc->c", bv, wo) + bo
    b0 = f"{ab}/MLP_0/ResnetBlock_0"
    W0, bb0 = p[f"{b0}/Dense_0/kernel"], p[f"{b0}/Dense_0/bias"]  # (E*C, R)
    Ws, bs = p[f"{b0}/Dense_1/kernel"], p[f"{b0}/Dense_1/bias"]
    Wout0, bout0 = p[f"{b0}/Dense_2/kernel"], p[f"{b0}/Dense_2/bias"]  # (R, M)
    W0r = W0.reshape(E, C, -1)
    Wsr = Ws.reshape(E, C, -1)
    W0f = np.einsum("hc,icn->hin", pv, W0r).reshape(nh * E, -1)
    b0f = bb0 + np.einsum("c,icn->n", ro, W0r)
    Wsf = np.einsum("hc,icn->hin", pv, Wsr).reshape(nh * E, -1)
    bsf = bs + np.einsum("c,icn->n", ro, Wsr)
    # the kernel's features are log2(e) * m (its softmax runs in base 2), so the feature rows carry ln 2
    W0f = W0f * np.log(2.0)
    Wsf = Wsf * np.log(2.0)
    f["B1"] = np.vstack([W0f, b0f[None]])  # K1 = nh*E + 1
    # GEMM2: y2 = t1 @ Wout0 + A1 @ ([Wsf; bsf] @ Wout0) + bout0
    f["B2t"] = Wout0
    B2a = np.vstack([Wsf, bsf[None]]) @ Wout0
    B2a[-1] += bout0
    f["B2a"] = B2a
    f["ln"] = {}
    for nm, key in (("ln1", f"{b0}/LayerNorm_0"), ("ln2", f"{b0}/LayerNorm_1")):
        f["ln"][nm] = (p[f"{key}/scale"], p[f"{key}/bias"])
    Wfin0, bfin0 = p[f"{ab}/MLP_0/Dense_0/kernel"], p[f"{ab}/MLP_0/Dense_0/bias"]  # (M, E)
    b1 = f"{ab}/MLP_1/ResnetBlock_0"
    W1, bb1 = p[f"{b1}/Dense_0/kernel"], p[f"{b1}/Dense_0/bias"]  # (Lq+E, R)
    Ws1, bs1 = p[f"{b1}/Dense_1/kernel"], p[f"{b1}/Dense_1/bias"]
    Wout1, bout1 = p[f"{b1}/Dense_2/kernel"], p[f"{b1}/Dense_2/bias"]
    Lq = W1.shape[0] - E
    # A3 = [t2 (M) | u_ln (Lq) | 1]
    f["B3"] = np.vstack([Wfin0 @ W1[Lq:], W1[:Lq], (bfin0 @ W1[Lq:] + bb1)[None]])
    S3 = np.vstack([Wfin0 @ Ws1[Lq:], Ws1[:Lq], (bfin0 @ Ws1[Lq:] + bs1)[None]])
    B4a = S3 @ Wout1
    B4a[-1] += bout1
    f["B4t"] = Wout1
    f["B4a"] = B4a
    for nm, key in (("ln3", f"{b1}/LayerNorm_0"), ("ln4", f"{b1}/LayerNorm_1")):
        f["ln"][nm] = (p[f"{key}/scale"], p[f"{key}/bias"])
    Wfin1, bfin1 = p[f"{ab}/MLP_1/Dense_0/kernel"], p[f"{ab}/MLP_1/Dense_0/bias"]
    f["B5"] = np.vstack([Wfin1, bfin1[None]])  # A5 = [t4 | 1]
    f["Lq"], f["E"], f["nh"] = Lq, E, nh
    return f


def _ln_gelu(h, scale, bias):
    h = h.astype(np.float32)
    y = orc.layer_norm(h, scale.astype(np.float32), bias.astype(np.float32))
    return orc.gelu(y)


def tc_chain_residual(flat, u, fmt_act="f16", fmt_w="f16"):
    """Residual eps[b, s, :out_dim] of q(z|u,s) for every sample.  `fmt_act` / `fmt_w`: rounding of the
    A (activation) / B (weight) operands of every GEMM (None = fp32; the kernel's hi/lo split of the
    activations is equivalent to fmt_act=None, its (hi, lo) weight image to fmt_w="f16x2"; round 1 used
    fmt_w="f16").  u: (B, Lq) float (the encoder output)."""
    f = fold_weights(flat)
    if not isinstance(fmt_w, dict):  # per-operand formats: {"B1": ..., "B2t": ..., ...} (design studies)
        fmt_w = {k: fmt_w for k in ("B1", "B2t", "B2a", "B3", "B4t", "B4a", "B5")}
    p32 = {k: np.asarray(v, dtype=np.float32) for k, v in flat.items()}
    Lq, E, nh = f["Lq"], f["E"], f["nh"]
    B = u.shape[0]
    u = u.astype(np.float32)
    u_ln = orc.layer_norm(u, p32["qz/u_ln/scale"], p32["qz/u_ln/bias"])
    qtok = u_ln @ p32["qz/AttentionBlock_0/DenseGeneral_0/kernel"][:, :, 0]  # (B, E)
    e = orc.layer_norm(p32["qz/Embed_0/embedding"], p32["qz/sample_embed_ln/scale"], p32["qz/sample_embed_ln/bias"])
    kv = e @ p32["qz/AttentionBlock_0/DenseGeneral_1/kernel"][:, :, 0]  # (S, E)
    S = kv.shape[0]
    t = f["alpha"].astype(np.float32)[None, :, None] * qtok[:, None, :] + f["beta"].astype(np.float32)[None, :, None]  # (B, nh, E)
    sc = t[:, None, :, :, None] * kv[None, :, None, None, :]  # (B, S, nh, E_i, E_j)
    sc = sc - sc.max(-1, keepdims=True)
    w = np.exp(sc)
    m = (w * kv[None, :, None, None, :]).sum(-1) / w.sum(-1)  # (B, S, nh, E)
    ones = np.ones((B, S, 1), np.float32)
    m = m * np.float32(1.4426950408889634)
    A1 = _round(np.concatenate([m.reshape(B, S, nh * E), ones], -1), fmt_act)
    h1 = A1 @ _round(f["B1"], fmt_w["B1"])
    t1 = _ln_gelu(h1, *f["ln"]["ln1"])
    y2 = _round(t1, fmt_act) @ _round(f["B2t"], fmt_w["B2t"]) + A1 @ _round(f["B2a"], fmt_w["B2a"])
    t2 = _ln_gelu(y2, *f["ln"]["ln2"])
    A3 = _round(np.concatenate([t2, np.broadcast_to(u_ln[:, None, :], (B, S, Lq)), ones], -1), fmt_act)
    h3 = A3 @ _round(f["B3"], fmt_w["B3"])
    t3 = _ln_gelu(h3, *f["ln"]["ln3"])
    y4 = _round(t3, fmt_act) @ _round(f["B4t"], fmt_w["B4t"]) + A3 @ _round(f["B4a"], fmt_w["B4a"])
    t4 = _ln_gelu(y4, *f["ln"]["ln4"])
    A5 = _round(np.concatenate([t4, ones], -1), fmt_act)
    return A5 @ _round(f["B5"], fmt_w["B5"])
