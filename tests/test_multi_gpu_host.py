"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: shard ranges, the rank-ordered
observable reduction, the in-place position all-gather layout and the ensemble reduction."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mg = importlib.import_module(PKG + ".multi_gpu")
    n = 16
    b, e = mg.shard_range(n, rank, world)
    # positions: every rank fills only its own range of x|y|z, then gathers in place
    buf = torch.full((3 * n,), -1.0, dtype=torch.float64)
    truth = torch.arange(3 * n, dtype=torch.float64) * 0.5
    for c in range(3):
        buf[c * n + b:c * n + e] = truth[c * n + b:c * n + e]
    mg.allgather_positions(buf, n, rank, world)
    ok_gather = bool(torch.equal(buf, truth))
    # observables: rank shares + tails
    part = torch.tensor([1.0 + rank, 2.0 * (rank + 1), -3.0 - rank, 0.25 * rank], dtype=torch.float64)
    obs = mg.reduce_observables(part, (10.0, 20.0, 30.0))
    # ensemble: rank r owns replicas r, r+world, ...
    reps = mg.replicas_of_rank(5, rank, world)
    vals = np.array([[1.0 + r, 10.0 * r] for r in reps])
    mean, std, cnt = mg.ensemble_reduce(vals.sum(0), (vals ** 2).sum(0), len(reps))
    out[rank] = (ok_gather, obs.tolist(), mean.tolist(), std.tolist(), cnt, (b, e), reps)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok_gather, obs, mean, std, cnt, rng, reps = out[rank]
        assert ok_gather
        assert rng == (rank * 8, rank * 8 + 8)
        # sum over ranks in rank order, tails added once to epot, d_epot, dd_epot
        assert obs == [1.0 + 2.0 + 10.0, 2.0 + 4.0, -3.0 - 4.0 + 20.0, 0.0 + 0.25 + 30.0]
        all_vals = np.array([[1.0 + r, 10.0 * r] for r in range(5)])
        assert np.allclose(mean, all_vals.mean(0)) and np.allclose(std, all_vals.std(0)) and cnt == 5
    assert out[0][1] == out[1][1]  # identical on every rank
    assert sorted(out[0][6] + out[1][6]) == list(range(5))


def test_shard_range_requires_divisibility():
    mg = importlib.import_module(PKG + ".multi_gpu")
    assert mg.shard_range(1048576, 3, 8) == (393216, 524288)
    with pytest.raises(ValueError):
        mg.shard_range(10, 0, 4)


def test_slab_layers_tile_the_box():
    """Cell-layer ownership of the slab-sharded neighbour-list path: contiguous, disjoint, complete,
    balanced to within one layer (same rule as slab_layers() in csrc/lj_nlist.cu)."""
    import importlib

    import ljmd_b200 as lj
    mg = importlib.import_module(lj.__name__ + ".multi_gpu")
    for nc in (3, 12, 39, 128):
        for world in (1, 2, 3, 4, 8):
            if world > nc:
                continue
            spans = [mg.slab_layers(nc, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nc
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
