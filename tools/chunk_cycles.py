import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ljmd_b200 as lj
for dense in (1, 0):
    for ri, rj in ((1, 1), (2, 1), (4, 1), (1, 2), (2, 2), (1, 4), (2, 4), (4, 2)):
        for w in (1, 4, 8, 16):
            c = lj.bench_chunk_cycles(ri, rj, dense, w)
            bound = max(1, w / 4) * ri * (35 if dense else 15) * 2
            print("dense", dense, "RI", ri, "RJ", rj, "warps/SM", w, "cyc/(tile/32) %.1f" % c, "fp64-issue-bound %.0f" % bound,
                  "eff %.2f" % (bound / c), flush=True)
