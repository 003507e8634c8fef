"""Second, independent CPU restatement of the reference hot path -- torch, float32, vectorised
over the sample axis the way ``jax.vmap`` runs it.  TEST / BASELINE INFRASTRUCTURE ONLY.

Used (a) to cross-check ``oracle/mrvi_oracle.py`` (sequential-over-samples numpy), and (b) as the
timed CPU baseline of ``bench.py`` ("port": the reference's own JAX-CPU path cannot run here, see
the header of ``mrvi_oracle.py`` -- PARITY UNPINNED at the flax/jax boundary).  It is driven the
way the reference drives its path: sequential minibatches of ``batch_size`` cells
(`_model.py:376`), all S samples per minibatch (`:351-355`), direct-difference distances
(`:545-550`), per-batch per-category sums (`:469-484`), final division by the counts (`:493-504`).
"""
from __future__ import annotations

import math

import numpy as np
import torch

LN_EPS = 1e-6
R_HIDDEN = 128


def _ln(v, scale=None, bias=None):
    mu = v.mean(-1, keepdim=True)
    var = ((v * v).mean(-1, keepdim=True) - mu * mu).clamp_min(0)
    mul = torch.rsqrt(var + LN_EPS)
    if scale is not None:
        mul = mul * scale
    y = (v - mu) * mul
    return y if bias is None else y + bias


def _gelu(v):
    return torch.nn.functional.gelu(v, approximate="tanh")


class TorchOracle:
    """Holds the flat parameters as torch tensors; ``z(x, sid)`` and ``distances`` mirror the jitted fns."""

    def __init__(self, flat, dtype=torch.float32):
        self.p = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in flat.items()}
        self.dtype = dtype
        self.S = self.p["qz/Embed_0/embedding"].shape[0]

    def _dense(self, v, name):
        return v @ self.p[f"{name}/kernel"] + self.p[f"{name}/bias"]

    def _resnet(self, h, prefix, act):
        p = self.p
        t = act(_ln(self._dense(h, f"{prefix}/Dense_0"), p[f"{prefix}/LayerNorm_0/scale"], p[f"{prefix}/LayerNorm_0/bias"]))
        if h.shape[-1] != R_HIDDEN:
            t = t + self._dense(h, f"{prefix}/Dense_1")
            k = 2
        else:
            t = t + h
            k = 1
        t = _ln(self._dense(t, f"{prefix}/Dense_{k}"), p[f"{prefix}/LayerNorm_1/scale"], p[f"{prefix}/LayerNorm_1/bias"])
        return act(t)

    def _mlp(self, h, prefix, act):
        i = 0
        while f"{prefix}/ResnetBlock_{i}/Dense_0/kernel" in self.p:
            h = self._resnet(h, f"{prefix}/ResnetBlock_{i}", act)
            i += 1
        return self._dense(h, f"{prefix}/Dense_0")

    def u(self, x, sid):
        """_EncoderXU mean (`_module.py:182-202`)."""
        p = self.p
        h = torch.log1p(x)
        for i in range(2):
            h = self._dense(h, f"qu/Dense_{i}")
            cn = f"qu/ConditionalNormalization_{i}"
            h = p[f"{cn}/gamma_conditional/embedding"][sid] * _ln(h) + p[f"{cn}/beta_conditional/embedding"][sid]
            h = _gelu(h)
        h = h + p["qu/Embed_0/embedding"][sid]
        i = 0
        nd = "qu/NormalDistOutputNN_0"
        while f"{nd}/ResnetBlock_{i}/Dense_0/kernel" in p:
            h = self._resnet(h, f"{nd}/ResnetBlock_{i}", torch.relu)
            i += 1
        return self._dense(h, f"{nd}/Dense_0")

    def z(self, x, sid):
        """(B, S, L): vmap over counterfactual samples of inference(...)["z"] (`_model.py:336-355`)."""
        p = self.p
        B, S = x.shape[0], self.S
        u = self.u(x, sid)  # (B, Lq): sample-independent, computed once under vmap
        u_ = _ln(u, p["qz/u_ln/scale"], p["qz/u_ln/bias"])
        e = _ln(p["qz/Embed_0/embedding"], p["qz/sample_embed_ln/scale"], p["qz/sample_embed_ln/bias"])  # (S, E)
        ab = "qz/AttentionBlock_0"
        q_tok = torch.einsum("bl,lei->bei", u_, p[f"{ab}/DenseGeneral_0/kernel"])  # (B,E,1)
        kv_tok = torch.einsum("sl,lei->sei", e, p[f"{ab}/DenseGeneral_1/kernel"])  # (S,E,1)
        mha = f"{ab}/MultiHeadDotProductAttention_0"
        q = torch.einsum("bti,ihd->bthd", q_tok, p[f"{mha}/query/kernel"]) + p[f"{mha}/query/bias"]
        k = torch.einsum("sti,ihd->sthd", kv_tok, p[f"{mha}/key/kernel"]) + p[f"{mha}/key/bias"]
        v = torch.einsum("sti,ihd->sthd", kv_tok, p[f"{mha}/value/kernel"]) + p[f"{mha}/value/bias"]
        q = q / math.sqrt(q.shape[-1])
        w = torch.softmax(torch.einsum("bqhd,skhd->bshqk", q, k), dim=-1)
        a = torch.einsum("bshqk,skhd->bsqhd", w, v)
        o = torch.einsum("bsqhd,hdc->bsqc", a, p[f"{mha}/out/kernel"]) + p[f"{mha}/out/bias"]
        o = o.reshape(B, S, -1)
        eps_ = self._mlp(o, f"{ab}/MLP_0", _gelu)
        c = torch.cat([u_[:, None, :].expand(B, S, -1), eps_], dim=-1)
        res = self._mlp(c, f"{ab}/MLP_1", _gelu)
        if "qz/Dense_0/kernel" in p:
            z_base = self._dense(u, "qz/Dense_0")
        else:
            z_base = u
        L = z_base.shape[-1]
        return z_base[:, None, :] + res[..., :L]

    @staticmethod
    def distances(z, norm="l2"):
        delta = z[:, None, :, :] - z[:, :, None, :]
        if norm == "l2":
            return torch.sqrt((delta ** 2).sum(-1))
        if norm == "l1":
            return delta.abs().sum(-1)
        if norm == "linf":
            return delta.abs().amax(-1)
        raise ValueError(f"norm {norm} not supported")

    @torch.no_grad()
    def local_sample_distances(self, X, sample_idx, group_codes=None, n_groups=None, keep_cell=False, batch_size=256,
                               norm="l2"):
        """The reference batch loop on integer group codes; returns dict(cell=..., sums=[...], counts=[...])."""
        n = X.shape[0]
        group_codes = group_codes or []
        sums = [torch.zeros((g, self.S, self.S), dtype=self.dtype) for g in (n_groups or [])]
        cells = []
        for lo in range(0, n, batch_size):
            hi = min(lo + batch_size, n)
            x = torch.as_tensor(np.ascontiguousarray(X[lo:hi])).to(self.dtype)
            sid = torch.as_tensor(np.asarray(sample_idx[lo:hi]).astype(np.int64))
            d = self.distances(self.z(x, sid), norm)
            if keep_cell:
                cells.append(d)
            for k, codes in enumerate(group_codes):
                c = torch.as_tensor(np.asarray(codes[lo:hi]).astype(np.int64))
                for g in torch.unique(c).tolist():
                    if g >= 0:
                        sums[k][g] += d[c == g].sum(0)
        out = {"sums": [s.numpy() for s in sums],
               "counts": [np.bincount(np.asarray(c)[np.asarray(c) >= 0], minlength=g) for c, g in zip(group_codes, n_groups or [])]}
        if keep_cell:
            out["cell"] = torch.cat(cells, 0).numpy() if cells else np.zeros((0, self.S, self.S), np.float32)
        return out
