"""Diagnostic: per-atom difference between a slab-sharded run and the single-GPU run after `steps`."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
import ljmd_b200 as lj

steps = int(sys.argv[1]); chunk = int(sys.argv[2])
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mg = importlib.import_module(lj.__name__ + ".multi_gpu")
w = bench.WORKLOADS["cfg4_nlist"]
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
ctx = lj.Context(p, method=lj.METHOD_NLIST, device=local)
ctx.upload(lj.SimState.from_arrays(arrs))
ctx.compute_forces()
runner = mg.SlabRunner(ctx, rank, world, chunk=chunk)
log = []
done = 0
while done < steps:
    k = min(chunk, steps - done)
    r0 = runner.rebuilds
    runner.step(k)
    done += k
    if runner.rebuilds != r0: log.append(done)
runner.finish()
d = ctx.download()
if rank == 0:
    c2 = lj.Context(p, method=lj.METHOD_NLIST, device=local)
    c2.upload(lj.SimState.from_arrays(arrs))
    c2.compute_forces()
    c2.step(steps)
    r = c2.download()
    bad = {}
    for name in ("rx", "vx", "ax"):
        diff = np.nonzero(getattr(d, name) != getattr(r, name))[0]
        bad[name] = int(len(diff))
        if len(diff):
            bad[name + "_idx"] = [int(x) for x in diff[:12]]
            bad[name + "_xpos"] = [float(r.rx[x]) for x in diff[:12]]
            bad[name + "_maxabs"] = float(np.abs(getattr(d, name) - getattr(r, name)).max())
    Lc = w["L"] / 39
    if bad["vx"]:
        idx = np.nonzero(d.vx != r.vx)[0]
        layers = np.floor(r.rx[idx] / Lc).astype(int)
        bad["vx_layers_hist"] = {int(k): int(v) for k, v in zip(*np.unique(layers, return_counts=True))}
    print(json.dumps({"steps": steps, "chunk": chunk, "rebuild_chunk_ends": log, "bad": bad}), flush=True)
dist.barrier()
dist.destroy_process_group()
