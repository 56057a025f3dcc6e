"""Force-kernel time vs N for the dense all-pairs case (fit fixed overhead vs per-pair cost)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np, bench, ljmd_b200 as lj
k = int(sys.argv[1])
w = dict(k=k, L=2.0*k, rc=0.98*k, dt=1e-4, method="allpairs", T=1.416)
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
ctx = lj.Context(p, method=lj.METHOD_ALLPAIRS)
ctx.upload(lj.SimState.from_arrays(arrs)); ctx.compute_forces()
nst = 320
ctx.step(nst); ctx.step(nst); t = ctx.last_step_ms / nst
ctx.force_timing(True); ctx.step(40); ms, ne = ctx.force_timing(False)
pairs = n*(n-1)/2
print(json.dumps(dict(n=n, step_us=round(1e3*t,2), force_us=round(1e3*ms/ne,2), ps_per_pair=round(1e9*t/pairs*1e-3*1e3,3))))
''' % ROOT
for label, env in [("default", {}), ("W4 barrier", dict(LJMD_SYM_RI="4", LJMD_SYM_W="4", LJMD_SYM_INDEP="0")), ("RI4 indep NW16", dict(LJMD_SYM_RI="4", LJMD_SYM_W="1", LJMD_SYM_NW="16"))]:
    for k in [4, 6, 8, 10, 13, 16, 20]:
        r = subprocess.run([sys.executable, "-c", code, str(k)], env=dict(os.environ, **env), capture_output=True, text=True)
        print(label, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
