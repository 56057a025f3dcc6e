"""Diagnostic: per-step observable series of a slab-sharded run vs the single-GPU run.
torchrun --nproc-per-node N tools/slab_trace.py [steps] [chunk]"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
import ljmd_b200 as lj

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 720
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mg = importlib.import_module(lj.__name__ + ".multi_gpu")
w = bench.WORKLOADS[os.environ.get("WL", "cfg4_nlist")]
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
ctx = lj.Context(p, method=lj.METHOD_NLIST, device=local)
ctx.upload(lj.SimState.from_arrays(arrs))
ctx.compute_forces()
runner = mg.SlabRunner(ctx, rank, world, chunk=chunk)
tails = lj.tail_corrections(p)
ctx.series_arm(steps)
rebuild_steps = []
done = 0
while done < steps:
    r0 = runner.rebuilds
    k = min(chunk, steps - done)
    runner.step(k)
    done += k
    if runner.rebuilds != r0:
        rebuild_steps.append((done - k, done))
ser = torch.tensor(ctx.series_fetch(steps), dtype=torch.float64, device="cuda")
allser = [torch.empty_like(ser) for _ in range(world)]
dist.all_gather(allser, ser)
tot = sum(a.cpu().numpy() for a in allser)
tot[:, 0] += tails[0]; tot[:, 2] += tails[1]; tot[:, 3] += tails[2]
if rank == 0:
    c2 = lj.Context(p, method=lj.METHOD_NLIST, device=local)
    c2.upload(lj.SimState.from_arrays(arrs))
    c2.compute_forces()
    ref = c2.step_series(steps)
    rel = np.abs(tot - ref) / np.abs(ref)
    bad = np.nonzero(rel.max(axis=1) > 1e-12)[0]
    out = {"world": world, "steps": steps, "chunk": chunk, "rebuild_chunks": rebuild_steps, "n_bad": int(len(bad)),
           "first_bad_step": int(bad[0]) + 1 if len(bad) else None}
    if len(bad):
        b = int(bad[0])
        out["around"] = [{"step": s + 1, "dE": float(tot[s, 0] - ref[s, 0]), "dK": float(tot[s, 1] - ref[s, 1])}
                         for s in range(max(0, b - 2), min(steps, b + 4))]
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
