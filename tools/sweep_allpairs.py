"""Tuning sweep of the symmetric all-pairs kernel geometry (run on the GPU box)."""
import itertools, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np, bench, ljmd_b200 as lj
w = bench.WORKLOADS[sys.argv[1]]
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
ctx = lj.Context(p, method=lj.METHOD_ALLPAIRS)
ctx.upload(lj.SimState.from_arrays(arrs)); ctx.compute_forces()
nst = int(sys.argv[2])
ctx.step(nst); ctx.step(nst); t = ctx.last_step_ms / nst
ctx.force_timing(True); ctx.step(max(4, nst // 10)); ms, ne = ctx.force_timing(False)
print(json.dumps(dict(step_us=1e3*t, force_us=1e3*ms/ne)))
''' % ROOT
wl, nst = sys.argv[1], sys.argv[2]
# (RI, NW, CTAs, ITEMS)
cfgs = [(2,8,444,3552),(2,8,444,2664),(2,8,444,4736),(2,8,296,2368),(2,8,296,3552),(2,10,296,2960),(2,20,148,2960),
        (4,8,296,2368),(4,16,148,2368),(4,8,296,1776)]
for ri, nw, ctas, items in cfgs:
    env = dict(os.environ, LJMD_SYM_RI=str(ri), LJMD_SYM_W="1", LJMD_SYM_NW=str(nw), LJMD_SYM_CTAS=str(ctas), LJMD_SYM_ITEMS=str(items))
    r = subprocess.run([sys.executable, "-c", code, wl, nst], env=env, capture_output=True, text=True)
    print(wl, "RI", ri, "NW", nw, "CTAS", ctas, "ITEMS", items, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
