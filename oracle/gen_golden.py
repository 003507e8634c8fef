"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE (TEST INFRASTRUCTURE ONLY).

`/root/reference/src/mrvi/_components.py` and `_module.py` are imported unmodified on top of the numpy
shim of flax / jax / numpyro / scvi in `oracle/flax_shim` (the real packages are not installable here).
`MrVAE(...).apply({"params": tree}, method=inference, x, sample_index, cf_sample=s, use_mean=True)["z"]`
is evaluated for every counterfactual sample s, exactly as `mapped_inference_fn` does
(`_model.py:336-365`), in float64.  The outputs, the inputs and the seeds go to `tests/golden/*.npz`;
`tests/test_golden.py` checks the oracle (and, on the GPU box, the CUDA path) against them and, when
/root/reference is present, re-runs this comparison live.

  python -m oracle.gen_golden            # rewrites tests/golden/
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "flax_shim")
REF_SRC = "/root/reference/src/mrvi"

CASES = {
    # name: (MrviDims kwargs, MrVAE kwargs of the reference, n_cells, seeds)
    "default_S4": (dict(n_input=100, n_sample=4), dict(), 48, (0, 1)),
    "iso_L10_S15": (dict(n_input=60, n_sample=15, n_latent=10, n_latent_u=10), dict(n_latent=10), 32, (2, 3)),
    "latent_u5": (dict(n_input=40, n_sample=6, n_latent=10, n_latent_u=5), dict(n_latent=10, n_latent_u=5), 32, (4, 5)),
    "use_map_false": (dict(n_input=30, n_sample=5, use_map=False), dict(qz_kwargs={"use_map": False}), 24, (6, 7)),
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "_module.py"))


def load_reference_modules():
    """Import the reference's _components / _module as package `mrvi_ref` with the shim on sys.path."""
    if "mrvi_ref._module" in sys.modules:
        return sys.modules["mrvi_ref._components"], sys.modules["mrvi_ref._module"]
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("jax", "flax", "numpyro", "scvi")}
    sys.path.insert(0, SHIM)
    try:
        pkg = types.ModuleType("mrvi_ref")
        pkg.__path__ = [REF_SRC]
        sys.modules["mrvi_ref"] = pkg
        mods = []
        for name in ("_constants", "_components", "_module"):
            spec = importlib.util.spec_from_file_location(f"mrvi_ref.{name}", os.path.join(REF_SRC, f"{name}.py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[f"mrvi_ref.{name}"] = m
            spec.loader.exec_module(m)
            mods.append(m)
        return mods[1], mods[2]
    finally:
        sys.path[:] = saved_path
        # keep the shim modules importable only under their own names inside mrvi_ref; drop them from the
        # global table if they were not there before, so the rest of the test session is unaffected
        for k, v in saved_mods.items():
            if v is None:
                for name in [n for n in sys.modules if n == k or n.startswith(k + ".")]:
                    sys.modules.pop(name, None)


def nest(flat):
    tree = {}
    for k, v in flat.items():
        node = tree
        parts = k.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = np.asarray(v, dtype=np.float64)
    return tree


def _tree_with_scale_head(flat):
    tree = nest(flat)
    # the reference also evaluates the (unused under use_mean) scale head of q(u|x): it needs parameters
    H, Lq = flat["qu/NormalDistOutputNN_0/Dense_0/kernel"].shape
    tree["qu"]["NormalDistOutputNN_0"].setdefault("Dense_1", {"kernel": np.zeros((H, Lq)), "bias": np.zeros(Lq)})
    return tree


def reference_sampled(flat, dims, vae_kwargs, X, sid, noise_u, noise_base):
    """Sampled paths through the reference's own ``MrVAE.inference(use_mean=False, mc_samples=...)``
    (`_module.py:308-343`), with the draws of ``qu.rsample`` supplied explicitly (numpyro shim NOISE_PROVIDER):

    * z_s: for each counterfactual sample s, ``inference(x, sample_index, mc_samples=mc, cf_sample=s)["z"]``
      with draws ``noise_u[s]`` (mc, n, Lq); stacked on axis -2 and transposed as `_model.py:352, 414`
      -> (n, mc, S, L);
    * baseline: ``inference(x, sample_index, mc_samples=2*mc)["z"]`` (own sample) with draws ``noise_base``,
      then the arithmetic of `_model.py:535-540` (l2 between the two halves, mean / var over mc)."""
    components, module = load_reference_modules()
    npd = components.dist  # the shim's numpyro.distributions module object the reference source holds
    assert module.dist is npd

    vae = module.MrVAE(n_input=dims.n_input, n_sample=dims.n_sample, n_batch=2, n_labels=1, n_continuous_cov=0,
                       training=False, **vae_kwargs)
    variables = {"params": _tree_with_scale_head(flat)}
    x = np.asarray(X, dtype=np.float64)
    sample_index = np.asarray(sid, dtype=np.float64).reshape(-1, 1)
    mc = noise_u.shape[1]
    zs = []
    try:
        for s in range(dims.n_sample):
            npd.NOISE_PROVIDER = lambda key, shape, s=s: {"u": noise_u[s]}[key]
            cf = np.full((x.shape[0], 1), s)
            out = vae.apply(variables, method=vae.inference, x=x, sample_index=sample_index, cf_sample=cf,
                            use_mean=False, mc_samples=mc)
            zs.append(np.asarray(out["z"]))  # (mc, n, L)
        z = np.stack(zs, axis=-2).transpose(1, 0, 2, 3)  # (n, mc, S, L)
        npd.NOISE_PROVIDER = lambda key, shape: {"u": noise_base}[key]
        zb = np.asarray(vae.apply(variables, method=vae.inference, x=x, sample_index=sample_index,
                                  use_mean=False, mc_samples=2 * mc)["z"])
    finally:
        npd.NOISE_PROVIDER = None
    first, second = zb[:mc], zb[mc:]
    l2 = np.sqrt(np.sum((first - second) ** 2, axis=2)).T
    return z, l2.mean(axis=1), l2.var(axis=1)


def reference_z(flat, dims, vae_kwargs, X, sid):
    """z[b, s, :] by running the reference's MrVAE.inference for every counterfactual sample."""
    _, module = load_reference_modules()
    vae = module.MrVAE(n_input=dims.n_input, n_sample=dims.n_sample, n_batch=2, n_labels=1, n_continuous_cov=0,
                       training=False, **vae_kwargs)
    variables = {"params": _tree_with_scale_head(flat)}
    x = np.asarray(X, dtype=np.float64)
    sample_index = np.asarray(sid, dtype=np.float64).reshape(-1, 1)  # scvi loader hands floats, (B, 1)
    zs = []
    for s in range(dims.n_sample):
        cf = np.full((x.shape[0], 1), s)
        out = vae.apply(variables, method=vae.inference, x=x, sample_index=sample_index, cf_sample=cf, use_mean=True)
        zs.append(np.asarray(out["z"]))
    return np.stack(zs, axis=1)  # out_axes=-2  ->  (cell, sample, latent)


def _reference_source_block(first_marker, last_marker):
    """Dedented source lines of `_model.py` from the line containing ``first_marker`` to the one containing
    ``last_marker`` (inclusive): the reference's own statements, executed below, never copied into this repo."""
    import textwrap

    lines = open(os.path.join(REF_SRC, "_model.py")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if first_marker in l)
    i1 = next(i for i in range(i0, len(lines)) if last_marker in lines[i])
    return textwrap.dedent("\n".join(lines[i0:i1 + 1]))


def reference_de(flat, dims, vae_kwargs, X, sid, noise_u, Xmat, admissible, lambd, samples):
    """The differential-expression eps stage through the reference's own code: eps_ from
    ``MrVAE.inference(use_mean=False, mc_samples)["eps"]`` on the flax shim for every kept counterfactual sample
    (`_model.py:1175-1194`, ``out_axes=-2``), then the reference's own statements for the design matrix
    (`_model.py:1152-1158`) and for the statistics (`_model.py:1196-1215`) EXECUTED from its source file with
    ``jnp`` -> numpy, ``jax.vmap`` -> a python loop, ``jax.scipy`` -> scipy."""
    import scipy.linalg
    import scipy.stats

    components, module = load_reference_modules()
    npd = components.dist
    vae = module.MrVAE(n_input=dims.n_input, n_sample=dims.n_sample, n_batch=2, n_labels=1, n_continuous_cov=0,
                       training=False, **vae_kwargs)
    variables = {"params": _tree_with_scale_head(flat)}
    x = np.asarray(X, dtype=np.float64)
    sample_index = np.asarray(sid, dtype=np.float64).reshape(-1, 1)
    mc = noise_u.shape[1]
    eps_list = []
    try:
        for s in samples:
            npd.NOISE_PROVIDER = lambda key, shape, s=s: {"u": noise_u[s]}[key]
            cf = np.full((x.shape[0], 1), s)
            out = vae.apply(variables, method=vae.inference, x=x, sample_index=sample_index, cf_sample=cf,
                            use_mean=False, mc_samples=mc)
            eps_list.append(np.asarray(out["eps"]))  # (mc, n, L)
    finally:
        npd.NOISE_PROVIDER = None
    eps_ = np.stack(eps_list, axis=-2)  # (mc, n, S_kept, L)

    class _NS:
        pass

    jnp = _NS()
    jnp.einsum, jnp.eye, jnp.real = np.einsum, np.eye, np.real
    jnp.linalg = _NS()
    jnp.linalg.pinv = np.linalg.pinv
    jax = _NS()
    jax.vmap = lambda f: (lambda a: np.stack([f(ai) for ai in a]))
    jax.scipy = _NS()
    jax.scipy.linalg, jax.scipy.stats = scipy.linalg, scipy.stats
    adm = np.asarray(admissible)
    ns = {"jnp": jnp, "jax": jax, "lambd": lambd, "n_covariates": Xmat.shape[1], "Xmat": np.asarray(Xmat, dtype=np.float64),
          "admissible_samples_dmat": np.stack([np.diag(a) for a in adm]).astype(float),   # `_model.py:1330-1332`
          "n_samples_per_cell": adm.sum(axis=1), "eps_": eps_}
    exec(_reference_source_block("xtmx = jnp.einsum(", "Amat = jnp.einsum("), ns)
    exec(_reference_source_block("eps_std = eps_.std(axis=2, keepdims=True)", "betas = betas * eps_std"), ns)
    return {"eps": eps_, "Amat": ns["Amat"], "prefactor": ns["prefactor"], "beta": ns["betas"].mean(0),   # "beta": `_model.py:1283`
            "effect_size": ns["ts"], "pvalue": ns["pvals"]}


DE_CASES = {
    # name: (MrviDims kwargs, MrVAE kwargs, n_cells, mc_samples, n_covariates, sample subset or None, lambd, seeds)
    "de_S6_mc4": (dict(n_input=80, n_sample=6), dict(), 20, 4, 2, None, 0.0, (20, 21)),
    "de_S9_subset_mc3": (dict(n_input=50, n_sample=9, n_latent=12, n_latent_u=6), dict(n_latent=12, n_latent_u=6), 16, 3, 3,
                         (0, 2, 3, 5, 6, 8), 0.1, (22, 23)),
}


def make_de_case(name):
    sys.path.insert(0, ROOT)
    import mrvi_b200

    dk, vk, n, mc, K, subset, lambd, (pseed, xseed) = DE_CASES[name]
    dims = mrvi_b200.MrviDims(**dk)
    flat = mrvi_b200.init_params(dims, seed=pseed)
    flat["qu/NormalDistOutputNN_0/Dense_1/bias"] = flat["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    X = mrvi_b200.synthetic_counts(n, dims.n_input, seed=xseed)
    rng = np.random.default_rng(xseed)
    sid = rng.integers(0, dims.n_sample, n).astype(np.int32)
    noise_u = rng.standard_normal((dims.n_sample, mc, n, dims.n_latent_q))
    samples = np.arange(dims.n_sample) if subset is None else np.asarray(subset)
    Xmat = rng.random((len(samples), K))
    Xmat[:, 0] = (np.arange(len(samples)) % 2).astype(float)   # one binary covariate, as a one-hot encoded category
    admissible = rng.random((n, len(samples))) < 0.85           # some inadmissible (cell, sample) pairs
    admissible[0] = True
    return dims, vk, flat, X, sid, noise_u, Xmat, admissible, lambd, samples


SAMPLED_CASES = {
    # name: (MrviDims kwargs, MrVAE kwargs, n_cells, mc_samples, seeds)
    "sampled_S4_mc3": (dict(n_input=100, n_sample=4), dict(), 24, 3, (10, 11)),
    "sampled_iso_S6_mc2": (dict(n_input=40, n_sample=6, n_latent=10, n_latent_u=10), dict(n_latent=10), 16, 2, (12, 13)),
}


def make_sampled_case(name):
    sys.path.insert(0, ROOT)
    import mrvi_b200

    dk, vk, n, mc, (pseed, xseed) = SAMPLED_CASES[name]
    dims = mrvi_b200.MrviDims(**dk)
    flat = mrvi_b200.init_params(dims, seed=pseed)
    # shrink the scale head's bias so that q(u|x) has a trained-model-like spread (softplus(-2) ~ 0.13)
    flat["qu/NormalDistOutputNN_0/Dense_1/bias"] = flat["qu/NormalDistOutputNN_0/Dense_1/bias"] - 2.0
    X = mrvi_b200.synthetic_counts(n, dims.n_input, seed=xseed)
    rng = np.random.default_rng(xseed)
    sid = rng.integers(0, dims.n_sample, n).astype(np.int32)
    noise_u = rng.standard_normal((dims.n_sample, mc, n, dims.n_latent_q))
    noise_base = rng.standard_normal((2 * mc, n, dims.n_latent_q))
    return dims, vk, flat, X, sid, noise_u, noise_base


def make_case(name):
    sys.path.insert(0, ROOT)
    import mrvi_b200

    dk, vk, n, (pseed, xseed) = CASES[name]
    dims = mrvi_b200.MrviDims(**dk)
    flat = mrvi_b200.init_params(dims, seed=pseed)
    X = mrvi_b200.synthetic_counts(n, dims.n_input, seed=xseed)
    sid = np.random.default_rng(xseed).integers(0, dims.n_sample, n).astype(np.int32)
    return dims, vk, flat, X, sid


def main():
    if not reference_available():
        raise SystemExit("/root/reference is not present: golden vectors can only be regenerated in the build container")
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in CASES:
        dims, vk, flat, X, sid = make_case(name)
        z = reference_z(flat, dims, vk, X, sid)
        delta = z[:, None, :, :] - z[:, :, None, :]
        d = np.sqrt((delta ** 2).sum(-1))  # `_model.py:547-550`
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), X=X, sid=sid, z=z.astype(np.float64), d=d.astype(np.float64),
                            **{"param|" + k.replace("/", "|"): v for k, v in flat.items()})
        print(f"{name}: z {z.shape}, max|z| {np.abs(z).max():.4f}, median d {np.median(d):.4f}")
    for name in SAMPLED_CASES:
        dims, vk, flat, X, sid, noise_u, noise_base = make_sampled_case(name)
        z, mu, var = reference_sampled(flat, dims, vk, X, sid, noise_u, noise_base)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), X=X, sid=sid, noise_u=noise_u, noise_base=noise_base,
                            z=z, base_mean=mu, base_var=var,
                            **{"param|" + k.replace("/", "|"): v for k, v in flat.items()})
        print(f"{name}: sampled z {z.shape}, baseline mean {mu.mean():.4f} var {var.mean():.4f}")
    for name in DE_CASES:
        dims, vk, flat, X, sid, noise_u, Xmat, admissible, lambd, samples = make_de_case(name)
        r = reference_de(flat, dims, vk, X, sid, noise_u, Xmat, admissible, lambd, samples)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), X=X, sid=sid, noise_u=noise_u, Xmat=Xmat, admissible=admissible,
                            lambd=lambd, samples=samples, **r,
                            **{"param|" + k.replace("/", "|"): v for k, v in flat.items()})
        print(f"{name}: eps {r['eps'].shape}, beta {r['beta'].shape}, effect_size median {np.median(r['effect_size']):.3f}")


if __name__ == "__main__":
    main()
