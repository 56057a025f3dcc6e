"""Per-warp start/end timestamps of the independent-warp all-pairs kernel (tuning aid)."""
import ctypes as C, os, sys
import numpy as np
os.environ["LJMD_SYM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ljmd_b200 as lj
w = bench.WORKLOADS["cfg2"]
n, arrs = bench.synthetic_state(w)
p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
ctx = lj.Context(p, method=lj.METHOD_ALLPAIRS)
ctx.upload(lj.SimState.from_arrays(arrs)); ctx.compute_forces(); ctx.step(20)
L = lj.load_library()
buf = np.zeros((8192, 4), dtype=np.int64)
m = L.ljmd_ctx_debug_warp_times(ctx._h, buf.ctypes.data_as(C.c_void_p), 8192)
d = buf[:m]; d = d[d[:, 3] > 0]
t0 = d[:, 0].min()
start, end, dur = d[:, 0] - t0, d[:, 1] - t0, d[:, 1] - d[:, 0]
print("warps with work", len(d), "of", m, "items/warp hist", np.bincount(d[:, 3]))
print("start: min %d med %d max %d" % (start.min(), np.median(start), start.max()))
print("end:   min %d med %d max %d" % (end.min(), np.median(end), end.max()))
for k in np.unique(d[:, 3]):
    s = dur[d[:, 3] == k]
    print("items=%d: n=%d duration min %d med %d max %d" % (k, len(s), s.min(), np.median(s), s.max()))
per_sm = np.bincount(d[:, 2], minlength=148)
print("warps per SM: min %d max %d" % (per_sm[per_sm > 0].min(), per_sm.max()), " SMs used", (per_sm > 0).sum())
