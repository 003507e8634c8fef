"""World-size-2 gloo test of the sharding / all-reduce logic (the N>1 host path), on CPU.

The per-shard compute is the oracle here (tests may use it); in the product it is the CUDA handle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import mrvi_b200
    from mrvi_b200._dist import shard_range, sharded_group_means
    from oracle import mrvi_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dims = mrvi_b200.MrviDims(n_input=30, n_sample=5)
    p = mrvi_b200.init_params(dims, seed=2)
    n = 101  # not divisible by the world size
    X = mrvi_b200.synthetic_counts(n, 30, seed=1)
    sid = np.random.default_rng(1).integers(0, 5, n)
    codes = np.random.default_rng(2).integers(0, 3, n)

    def compute_shard(lo, hi):
        d = orc.pairwise_distances(orc.counterfactual_z(p, X[lo:hi], sid[lo:hi]))
        sums = np.stack([d[codes[lo:hi] == g].sum(0) for g in range(3)])
        return {"group_sums": [sums], "group_counts": [np.bincount(codes[lo:hi], minlength=3).astype(np.int64)], "cell": d}

    res = sharded_group_means(compute_shard, n, rank, world)
    assert res["range"] == shard_range(n, rank, world)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), means=res["group_means"][0], counts=res["group_counts"][0],
             cell=res["cell"], lo=res["range"][0], hi=res["range"][1])
    dist.destroy_process_group()


def test_world2_sharding_allreduce(tmp_path):
    import mrvi_b200
    from mrvi_b200._dist import shard_range
    from oracle import mrvi_oracle as orc

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    dims = mrvi_b200.MrviDims(n_input=30, n_sample=5)
    p = mrvi_b200.init_params(dims, seed=2)
    n = 101
    X = mrvi_b200.synthetic_counts(n, 30, seed=1)
    sid = np.random.default_rng(1).integers(0, 5, n)
    codes = np.random.default_rng(2).integers(0, 3, n)
    d = orc.pairwise_distances(orc.counterfactual_z(p, X, sid))
    ref_means = np.stack([d[codes == g].mean(0) for g in range(3)])
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for o in outs:  # every rank holds the same global means and counts
        assert np.array_equal(o["counts"], np.bincount(codes, minlength=3))
        assert np.allclose(o["means"], ref_means, rtol=1e-6)
    # keep_cell slices concatenate to the single-process result in the original order
    assert [int(o["lo"]) for o in outs] == [shard_range(n, r, world)[0] for r in range(world)]
    # (BLAS results depend on the batch shape at the 1e-16 level, hence allclose and not array_equal)
    assert np.allclose(np.concatenate([o["cell"] for o in outs]), d, rtol=0, atol=1e-13)


def test_shard_range_partition():
    from mrvi_b200._dist import shard_range

    for n in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
