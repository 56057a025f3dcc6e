import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import ljmd_b200 as lj
steps = 120
w = bench.WORKLOADS["cfg4_nlist"]
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
def run(b2b):
    os.environ["LJMD_NL_NOGRAPH"] = "1"
    if b2b: os.environ["LJMD_NL_B2B"] = "1"
    else: os.environ.pop("LJMD_NL_B2B", None)
    ctx = lj.Context(p, method=lj.METHOD_NLIST)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()
    ser = ctx.step_series(steps)
    info = ctx.nlist_info()
    ctx.close()
    return ser, info["rebuilds"]
ref, rb = run(False)
out = []
for k in range(10):
    s, r = run(True)
    bad = np.nonzero((s != ref).any(axis=1))[0]
    out.append(None if len(bad) == 0 else (int(bad[0]) + 1, float(s[bad[0], 1] - ref[bad[0], 1])))
print(json.dumps({"ref_rebuilds": rb, "b2b_first_diff": out}))
