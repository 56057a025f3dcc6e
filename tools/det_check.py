"""Run-to-run determinism of the single-GPU neighbour-list path at N=2^20."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ljmd_b200 as lj
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
w = bench.WORKLOADS["cfg4_nlist"]
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
sers = []
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    ctx = lj.Context(p, method=lj.METHOD_NLIST)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()
    sers.append(ctx.step_series(steps))
    ctx.close()
out = []
for k in range(1, len(sers)):
    bad = np.nonzero((sers[k] != sers[0]).any(axis=1))[0]
    out.append(None if len(bad) == 0 else int(bad[0]) + 1)
print(json.dumps({"mode": os.environ.get("LJMD_NL_NOGRAPH", "graph"), "first_diff_step_vs_run0": out}))
