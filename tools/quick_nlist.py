"""Quick timing of the large-N paths (tuning aid): python tools/quick_nlist.py [workload ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ljmd_b200 as lj

for name in sys.argv[1:] or ["cfg3_nlist", "cfg4_nlist"]:
    w = bench.WORKLOADS[name]
    n, arrs = bench.synthetic_state(w)
    p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
    method = {"cells": lj.METHOD_CELLS, "nlist": lj.METHOD_NLIST}.get(w["method"], lj.METHOD_ALLPAIRS)
    ctx = lj.Context(p, method=method)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()
    nst = int(os.environ.get("STEPS", "320"))
    ctx.step(64)
    r0 = ctx.nlist_info()["rebuilds"] if method == lj.METHOD_NLIST else 0
    obs = ctx.step(nst)
    ms = ctx.last_step_ms
    out = {"workload": name, "n": n, "step_us": 1e3 * ms / nst, "atom_steps_per_s": n * nst / (ms * 1e-3), "obs": [float(x) for x in obs]}
    if method == lj.METHOD_NLIST:
        info = ctx.nlist_info()
        out.update(rebuilds_in_run=info["rebuilds"] - r0, entries_per_atom=info["entries"] / n, cap=info["capacity"], nc=info["cells_per_edge"], skin=info["skin"])
    ctx.force_timing(True)
    ctx.step(20)
    msf, nev = ctx.force_timing(False)
    out["force_us"] = 1e3 * msf / nev
    print(json.dumps(out), flush=True)
    ctx.close()
