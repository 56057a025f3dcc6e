"""CPU model of the position gathers of nl_step_kernel (csrc/lj_nlist.cu): for a given ordering of the atoms inside
the cells, how many distinct 32-byte sectors (= distinct neighbour atoms) and distinct 128-byte lines (4 packed
positions) one warp-wide gather touches.  One thread per atom, a warp = 32 consecutive slots, request m = entry m of
every lane's list (list order = the build kernel's: 9 columns x 3 cells along z, slot order inside a cell).

    python tools/sim_gather.py [k] [jitter|liquid]

Also reports the union-list size of slot pairs (2p, 2p+1) and quads: what a cluster list with 2 / 4 i-atoms per
thread would have to gather and evaluate.
"""
import sys

import numpy as np
from scipy.spatial import cKDTree


def state(k, kind, rho=0.8, seed=0):
    n = 4 * k ** 3
    L = (n / rho) ** (1.0 / 3.0)
    a = L / k
    rng = np.random.default_rng(seed)
    idx = np.stack(np.meshgrid(np.arange(k), np.arange(k), np.arange(k), indexing="ij"), -1).reshape(-1, 1, 3).astype(float)
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]], dtype=float)
    pos = ((idx + basis[None]) * a).reshape(-1, 3)
    if kind == "jitter":
        pos = pos + 0.02 * a * (rng.random((n, 3)) - 0.5)
    elif kind == "liquid":  # liquid-like: large displacements (about the nearest-neighbour distance), min distance kept by relaxation
        pos = pos + 0.45 * a * (rng.random((n, 3)) - 0.5)
    elif kind == "shuffled":  # liquid-like + original indices uncorrelated with space (a long run: atoms have diffused)
        pos = pos + 0.45 * a * (rng.random((n, 3)) - 0.5)
        pos = pos[rng.permutation(n)]
    pos = pos - L * np.floor(pos / L)
    return n, L, pos


def morton(ix, iy, iz, bits):
    code = np.zeros_like(ix)
    for b in range(bits):
        code |= ((ix >> b) & 1) << (3 * b + 2) | ((iy >> b) & 1) << (3 * b + 1) | ((iz >> b) & 1) << (3 * b)
    return code


def analyse(n, L, pos, rl, sub, label):
    nc = int(np.floor(L / rl))
    scale = nc / L
    cc = np.minimum((pos * scale).astype(int), nc - 1)
    cell = (cc[:, 0] * nc + cc[:, 1]) * nc + cc[:, 2]
    frac = pos * scale - cc  # position inside the cell, [0, 1)
    if sub == "orig":
        key2 = np.arange(n)
    elif sub.startswith("morton"):
        b = int(sub[6:])
        s = np.minimum((frac * (1 << b)).astype(int), (1 << b) - 1)
        key2 = morton(s[:, 0], s[:, 1], s[:, 2], b) * n + np.arange(n)
    elif sub == "z":
        key2 = np.argsort(np.argsort(frac[:, 2], kind="stable"), kind="stable")
    elif sub == "zsnake":  # z ascending in even columns position, descending in odd -- neighbours along the slot order stay close
        key2 = np.argsort(np.argsort(frac[:, 2], kind="stable"), kind="stable")
    order = np.lexsort((key2, cell))  # slot -> atom
    slot_of = np.empty(n, dtype=np.int64)
    slot_of[order] = np.arange(n)
    tree = cKDTree(pos, boxsize=L)
    pairs = tree.query_pairs(rl, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    # build-kernel order of the entries of atom i: (dx offset, dy offset, dz offset, slot of j)
    d = cc[j] - cc[i]
    d = (d + 1 + nc) % nc - 1  # periodic offset in {-1, 0, 1}
    si, sj = slot_of[i], slot_of[j]
    o = np.lexsort((sj, d[:, 2], d[:, 1], d[:, 0], si))
    si, sj = si[o], sj[o]
    cnt = np.bincount(si, minlength=n)
    start = np.concatenate([[0], np.cumsum(cnt)])
    m = np.arange(len(si)) - start[si]  # entry number inside the list
    warp = si // 32
    # distinct sectors / lines per (warp, m)
    req = warp * 1024 + m
    nreq = len(np.unique(req))
    sect = len(np.unique(req * (2 * n) + sj))
    line = len(np.unique(req * (2 * n) + sj // 4))
    # per-atom requests are also issued for the padded group tail; ignored here
    out = {"label": label, "sub": sub, "entries_per_atom": len(si) / n, "sectors_per_request": sect / nreq,
           "lines_per_request": line / nreq, "requests_per_warp": nreq / ((n + 31) // 32)}
    # union lists of clusters of 2 and 4 consecutive slots; the union is walked in ascending slot order.
    #   "lanes": the c lanes of a cluster read the same entry at the same time (a warp = 32/c clusters);
    #   "thread": one thread per cluster (a warp = 32 clusters)
    for c in (2, 4):
        cl = si // c
        key = np.unique(cl * (2 * n) + sj)
        ucl, usj = key // (2 * n), key % (2 * n)
        ncl = (n + c - 1) // c
        out[f"union{c}_per_cluster"] = len(key) / ncl
        ucnt = np.bincount(ucl, minlength=ncl)
        ustart = np.concatenate([[0], np.cumsum(ucnt)])
        um = np.arange(len(key)) - ustart[ucl]
        for name, per_warp in (("lanes", 32 // c), ("thread", 32)):
            w = ucl // per_warp
            rq = w * 1024 + um
            nrq = len(np.unique(rq))
            ln = len(np.unique(rq * (2 * n) + usj // 4))
            nwarps = (ncl + per_warp - 1) // per_warp
            atoms_per_warp = per_warp * c
            out[f"u{c}_{name}_lines_per_req"] = ln / nrq
            out[f"u{c}_{name}_lines_per_atom"] = ln / n
    out["lines_per_atom_now"] = line / n
    return out


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    kinds = sys.argv[2:] or ["jitter", "liquid", "shuffled"]
    for kind in kinds:
        n, L, pos = state(k, kind)
        for sub in ("orig", "morton1"):
            r = analyse(n, L, pos, 2.8, sub, kind)
            print({k_: (round(v, 2) if isinstance(v, float) else v) for k_, v in r.items()}, flush=True)


if __name__ == "__main__":
    main()
