"""Generates tests/golden/*.npz from the CPU oracle (oracle/ljmd_oracle.c).

The reference ships no golden vectors and cannot be compiled here, so these
fixtures are REGRESSION pins of the oracle itself plus the SURVEY.md Appendix-B anchors (a numpy
emulation made by the surveyor, not Fortran output).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # config 1: shipped input (inputs/input_simulation_parameters.txt:21,35,45)
    p = O.params_init(500, 10.0, 1e-4, 4.9, 5)
    st, epot0, ekin0 = O.initial_config(p, -12345, -555.0)
    ax, ay, az, e, d, dd = O.compute_lj(p, *st[:3])
    s1 = [a.copy() for a in st]
    obs1 = O.run_steps(p, 1, s1)
    s10 = [a.copy() for a in s1]
    obs10 = O.run_steps(p, 9, s10)
    s200 = [a.copy() for a in s10]
    obs200 = O.run_steps(p, 190, s200)
    a200 = O.compute_lj(p, *s200[:3])
    g = O.Rng(-12345)
    draws = np.array([g.uniform() for _ in range(1500)])
    np.savez_compressed(
        os.path.join(HERE, "cfg1_n500.npz"),
        state0=np.array(st), epot0=epot0, ekin0=ekin0, obs_t0=np.array([e, d, dd]),
        state1=np.array(s1), obs1=np.array(obs1), obs10=np.array(obs10),
        state200=np.array(s200), obs200=np.array(obs200), obs_f200=np.array(a200[3:]),
        rng_first=draws[:6], rng_sum1500=draws.sum(),
        tails=np.array(O.tail_corrections(p)),
        pairs_in_cutoff_t0=O.count_pairs_in_cutoff(p, *st[:3]),
        pairs_in_cutoff_t200=O.count_pairs_in_cutoff(p, *s200[:3]),
    )
    # small rho=0.8, rc=2.5 system for the cell list (N=864, 4 cells per edge)
    L = (864 / 0.8) ** (1.0 / 3.0)
    p = O.params_init(864, L, 1e-3, 2.5, 6)
    st, _, _ = O.initial_config(p, -12345, 0.0, 1.0)
    s = [a.copy() for a in st]
    O.run_steps(p, 300, s)
    f = O.compute_lj(p, *s[:3])
    fc = O.compute_lj_cells(p, *s[:3])
    cell = O.cell_assign(p, *s[:3])
    start, order = O.cell_sort(p, cell)
    np.savez_compressed(
        os.path.join(HERE, "cells_n864.npz"),
        L=L, state300=np.array(s), acc=np.array(f[:3]), obs=np.array(f[3:]),
        obs_cells=np.array(fc[3:6]), n_cand=fc[6], n_cut=fc[7], cell=cell, start=start,
        order=order,
    )
    # neighbour-list fixture of the same state: cells with edge >= rc + skin (skin 0.3) and the full
    # Verlet list within 2.8 (original numbering, ascending); does NOT rewrite cells_n864.npz's state
    g = np.load(os.path.join(HERE, "cells_n864.npz"))
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    rng = 2.5 + 0.3
    cell = O.cell_assign_range(p, rng, *s[:3])
    start, order = O.cell_sort_range(p, rng, cell)
    counts, nbr = O.neighbor_list(p, rng, *s[:3], cap=144)
    np.savez_compressed(
        os.path.join(HERE, "nlist_n864.npz"),
        range=rng, cell=cell, start=start, order=order, counts=counts, nbr=nbr.astype(np.int16),
    )
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
