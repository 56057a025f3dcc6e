#!/usr/bin/env python
"""Measurement of the trajectory consumers on one GPU, with the numpy
oracle timed beside them on a bounded sample.  Synthetic trajectories of the shipped production
run's shape: n = 500 atoms, 5000 snapshots (md_simulation_program: 500,000 steps sampled every 100);
the script's defaults are max_lag = min(n_snap-1, 3000), origin stride 1 up to 5000 snapshots
(md_one_run_analysis.py:756-767) and an RDF over 200 snapshots x 500 atoms (:555-561).

Prints one JSON line per kernel:
  rdf : algorithmic work = pairs distance-tested; FP64-bound (4 IEEE divisions + 1 sqrt per pair)
  msd / vacf : algorithmic bytes = 2 rows x 3 components x 8 B per (origin, lag, atom) -- L2-resident
               (the trajectory is 60 MB per vector field), so the bound is L2/L1 bandwidth, not HBM
Usage: python tools/bench_analysis.py [--snapshots 5000] [--atoms 500] [--cpu-seconds 10]
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--snapshots", type=int, default=5000)
    ap.add_argument("--atoms", type=int, default=500)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--repeat", type=int, default=5)
    args = ap.parse_args()

    import ljmd_b200 as lj
    an = importlib.import_module(lj.__name__ + ".analysis")
    from oracle import analysis_oracle as A

    if lj.device_count() < 1:
        raise SystemExit("bench_analysis: no CUDA device (no CPU fallback)")
    rng = np.random.default_rng(0)
    n_snap, n, L = args.snapshots, args.atoms, 10.0
    v = rng.normal(size=(3, n_snap, n))
    ru = np.cumsum(v, axis=1) * 0.01 + rng.random((3, 1, n)) * L
    r = ru - L * np.floor(ru / L)
    max_lag = min(n_snap - 1, 3000)
    stride = 20 if n_snap > 20000 else (5 if n_snap > 5000 else 1)  # md_one_run_analysis.py:762-767

    # ---- RDF ----
    snaps = np.arange(n_snap) if n_snap <= 200 else np.linspace(0, n_snap - 1, 200, dtype=int)
    parts = np.arange(n) if n <= 800 else np.linspace(0, n - 1, 800, dtype=int)
    pick = np.ix_(snaps, parts)
    x, y, z = [np.ascontiguousarray(c[pick]) for c in r]
    pairs = len(snaps) * len(parts) * (len(parts) - 1) // 2
    ms = []
    for _ in range(args.repeat + 1):
        t0 = time.perf_counter()
        counts = an.rdf_pair_counts(x, y, z, L, 200, 0.5 * L)
        wall = time.perf_counter() - t0
        ms.append((an.last_kernel_ms["rdf"], wall * 1e3))
    kms, wms = np.median([m[0] for m in ms[1:]]), np.median([m[1] for m in ms[1:]])
    k = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < args.cpu_seconds and k < len(snaps):
        c1 = A.rdf_pair_counts(x[k], y[k], z[k], L, 200, 0.5 * L)
        k += 1
    cpu_s = time.perf_counter() - t0
    cpu_pairs = k * len(parts) * (len(parts) - 1) // 2
    print(json.dumps({"kernel": "rdf_hist_kernel", "metric": "pair_distances_per_s", "value": pairs / (kms * 1e-3),
                      "kernel_ms": kms, "e2e_ms": wms, "pairs": pairs, "in_range": int(counts.sum()),
                      "cpu_baseline": {"value": cpu_pairs / cpu_s, "kind": "port", "cores": 1,
                                       "sample": f"{k} snapshots of {len(parts)} atoms, numpy oracle"}}))

    # ---- MSD / VACF ----
    origins = len(range(0, n_snap - 1, stride))
    evals = sum(min(max_lag, n_snap - 1 - t0) + 1 for t0 in range(0, n_snap - 1, stride))
    for kind, arrs, fn, ofn in (("msd", ru, an.compute_msd_tau_timeorig, A.compute_msd_tau_timeorig),
                                ("vacf", v, an.compute_vacf_tau_timeorig, A.compute_vacf_tau_timeorig)):
        ms = []
        for _ in range(args.repeat + 1):
            t0 = time.perf_counter()
            out = fn(*arrs, max_lag=max_lag, origin_stride=stride)
            ms.append((an.last_kernel_ms[kind], (time.perf_counter() - t0) * 1e3))
        kms, wms = np.median([m[0] for m in ms[1:]]), np.median([m[1] for m in ms[1:]])
        # CPU sample: the same call on the first `m` snapshots, grown until it fills the time budget
        m, cpu_s, cpu_evals = 64, 0.0, 0
        while m <= n_snap:
            t0 = time.perf_counter()
            ofn(*[a[:m] for a in arrs], max_lag=min(max_lag, m - 1), origin_stride=stride)
            cpu_s = time.perf_counter() - t0
            cpu_evals = sum(min(min(max_lag, m - 1), m - 1 - t0) + 1 for t0 in range(0, m - 1, stride))
            if cpu_s > args.cpu_seconds / 4 or m == n_snap:
                break
            m = min(2 * m, n_snap)
        print(json.dumps({"kernel": f"time_corr_kernel<{kind}>", "metric": "origin_lag_rows_per_s",
                          "value": evals / (kms * 1e-3), "kernel_ms": kms, "e2e_ms": wms, "origins": origins,
                          "max_lag": max_lag, "stride": stride, "rows": evals,
                          "algorithmic_GBps": evals * n * 48 / (kms * 1e-3) / 1e9, "checksum": float(out.sum()),
                          "cpu_baseline": {"value": cpu_evals / cpu_s, "kind": "port", "cores": 1,
                                           "sample": f"first {m} snapshots, numpy oracle"}}))


if __name__ == "__main__":
    main()
