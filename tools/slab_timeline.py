"""Per-step device timeline of a slab-sharded run (torchrun, one rank per GPU): where a step's time goes.
   LJMD_SLAB_TIMELINE=1 python -m torch.distributed.run --nproc-per-node N tools/slab_timeline.py [steps]
Per rank, medians over the steps without a rebuild: wait = first CTA start -> all peers' flags seen,
compute = flags seen -> last CTA end, tail = last CTA end -> flag published, gap = published -> next step's
first CTA."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LJMD_SLAB_TIMELINE"] = "1"


def main():
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    import ljmd_b200 as lj
    mg = importlib.import_module(lj.__name__ + ".multi_gpu")
    w = bench.WORKLOADS["cfg4_nlist"]
    n, arrs = bench.synthetic_state(w)
    p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
    ctx = lj.Context(p, method=lj.METHOD_NLIST, device=local)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()
    runner = mg.SlabRunner(ctx, rank, world, chunk=32)
    runner.step(32)
    L = lj.load_library()
    buf = np.zeros((1024, 4), dtype=np.uint64)
    L.ljmd_ctx_debug_slab_timeline(ctx._h, buf.ctypes.data_as(C.c_void_p), 1024)  # re-arm
    s0 = runner.steps
    runner.step(nsteps)
    m = L.ljmd_ctx_debug_slab_timeline(ctx._h, buf.ctypes.data_as(C.c_void_p), 1024)
    t = buf[s0:s0 + nsteps].astype(np.float64)
    ok = (t[:, 1] > 0) & (t[:, 3] > 0) & (t[:, 0] < 1e19)
    wait, comp, tail = t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2]
    gap = np.r_[t[1:, 0] - t[:-1, 3], np.nan]
    period = np.r_[t[1:, 0] - t[:-1, 0], np.nan]
    good = ok & np.r_[ok[1:], False] & (period < 5e5)  # drop the steps next to a rebuild
    line = ("rank %d: steps %d (used %d)  period %.1f us = wait %.1f + compute %.1f + tail %.1f + gap %.1f   "
            "(p90 period %.1f, compute %.1f)" % (rank, nsteps, good.sum(), np.median(period[good]) / 1e3,
            np.median(wait[good]) / 1e3, np.median(comp[good]) / 1e3, np.median(tail[good]) / 1e3,
            np.median(gap[good]) / 1e3, np.percentile(period[good], 90) / 1e3, np.percentile(comp[good], 90) / 1e3))
    lines = [None] * world
    dist.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
