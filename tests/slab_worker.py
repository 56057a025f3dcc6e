"""Worker of tests/test_gpu_multi_process.py: one rank of a REAL multi-process slab-sharded run
(torchrun, NCCL, CUDA IPC peer memory, concurrently running step kernels that wait on each other's
per-step flags).  Every rank also runs the single-context reference on its own GPU and compares its
final state BITWISE; rank 0 writes the verdict."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_path, k, nsteps, skin = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    os.environ["LJMD_NL_SKIN"] = skin
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    import ljmd_b200 as lj
    mg = importlib.import_module(lj.__name__ + ".multi_gpu")

    L = k * (4.0 / 0.8) ** (1.0 / 3.0)  # rho = 0.8
    w = dict(k=k, L=L, rc=2.5, dt=1e-3, method="nlist", T=1.0)
    n, arrs = bench.synthetic_state(w, seed=5)
    p = lj.init_params(n, L, 1e-3, 2.5, k)
    state = lj.SimState.from_arrays(arrs)

    with lj.Context(p, method=lj.METHOD_NLIST, device=local) as ref:
        ref.upload(state)
        ref.compute_forces()
        ref_obs = ref.step(nsteps)
        ref_state = ref.download()
        ref_rebuilds = ref.nlist_info()["rebuilds"]

    ctx = lj.Context(p, method=lj.METHOD_NLIST, device=local)
    ctx.upload(state)
    ctx.compute_forces()
    runner = mg.SlabRunner(ctx, rank, world, chunk=32)
    mom_ok = True
    share = None
    left = nsteps
    while left > 0:  # uneven call lengths on purpose: chunks of 32 plus ragged ends, across rebuilds
        m = min(left, 45)
        share = runner.step(m)
        left -= m
        mom = torch.tensor(ctx.nlist_momentum(), dtype=torch.float64, device="cuda")
        dist.all_reduce(mom)
        mom_ok = mom_ok and float(mom.abs().max()) < 1e-7
    obs = mg.reduce_observables(torch.tensor(share, dtype=torch.float64, device="cuda"), lj.tail_corrections(p))
    runner.finish()
    torch.cuda.synchronize()
    got = ctx.download()
    same = all(np.array_equal(a, b) for a, b in zip(got.arrays(), ref_state.arrays()))
    worst = max(float(np.max(np.abs(a - b))) for a, b in zip(got.arrays(), ref_state.arrays()))
    res = {"rank": rank, "bitwise_equal_to_single_context": bool(same), "max_abs_diff": worst, "momentum_ok": bool(mom_ok),
           "rebuilds_sharded": int(runner.rebuilds), "rebuilds_single": int(ref_rebuilds),
           "obs_rel_err": float(np.max(np.abs(obs - ref_obs) / np.abs(ref_obs))), "n": n, "steps": nsteps}
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(allres, f)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
