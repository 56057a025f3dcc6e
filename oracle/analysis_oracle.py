"""CPU oracle for the trajectory consumers -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the analysis loops of the reference's post-processing script (file:line
relative to /root/reference/scripts/md_one_run_analysis.py):
  :245-326  FortranSequentialReader (4- or 8-byte record markers, auto-detected)
  :345-397  read_rva
  :404-441  compute_msd_tau_timeorig      :444-489  compute_vacf_tau_timeorig
  :492-513  diffusion_from_msd_linear_fit :516-534  diffusion_from_vacf_integral
  :537-598  compute_rdf

PARITY PINNED AGAINST THE REFERENCE ITSELF: the reference here is Python, so its own functions were run
in the build container by tests/golden/make_golden_analysis.py (module imported with matplotlib stubbed;
np.trapz aliased to np.trapezoid for numpy 2.x) and their outputs committed as
tests/golden/analysis_ref.npz together with the inputs.  tests/test_analysis_golden.py checks every
function below against that fixture: bit-identical on the numpy it was made with (asserted: g(r), the
bin centres and the rva.dat arrays exactly, MSD / VACF within 1e-15 relative).  The restatement keeps the
reference's numpy expressions (operation order, np.rint, true division, np.mean's pairwise sum,
astype(int) truncation); tests/test_analysis_oracle.py adds known answers (lattice shell counts,
ballistic MSD, ideal-gas g(r) = 1).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module.
"""
from __future__ import annotations

import math
import struct
from typing import Optional, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------
# rva.dat (:245-397)
# ---------------------------------------------------------------------------------------------
def _marker_size(head: bytes) -> int:
    """:262-291 -- the first 4 (or 8) bytes are a plausible record length whose trailing copy matches."""
    for size, code in ((4, "<i"), (8, "<q")):
        if len(head) >= 2 * size:
            (nb,) = struct.unpack(code, head[:size])
            if 0 < nb < 1_000_000_000 and len(head) >= size + nb + size:
                (tail,) = struct.unpack(code, head[size + nb:size + nb + size])
                if tail == nb:
                    return size
    return 4


def _records(path):
    with open(path, "rb") as f:
        blob = f.read()
    ms = _marker_size(blob[:64])
    code = "<i" if ms == 4 else "<q"
    pos = 0
    while pos < len(blob):
        if pos + ms > len(blob):
            raise IOError("Truncated record header")
        (nb,) = struct.unpack(code, blob[pos:pos + ms])
        if nb < 0:
            raise IOError(f"Negative record length: {nb}")
        data = blob[pos + ms:pos + ms + nb]
        if len(data) != nb:
            raise IOError("Truncated record payload")
        foot = blob[pos + ms + nb:pos + 2 * ms + nb]
        if len(foot) != ms:
            raise IOError("Truncated record footer")
        if struct.unpack(code, foot)[0] != nb:
            raise IOError("Record marker mismatch")
        yield data
        pos += 2 * ms + nb


def read_rva(path):
    """:345-397.  Returns a dict with n, L, dt, output_interval, n_snapshots and the (n_snap, n)
    arrays rx..rz, rux..ruz, vx..vz (accelerations are skipped like in the reference)."""
    rec = _records(path)
    head = next(rec)
    if len(head) != struct.calcsize("<i d d i i"):
        raise IOError("Record size mismatch")
    n, L, dt, interval, n_snap = struct.unpack("<i d d i i", head)
    out = {k: np.empty((n_snap, n)) for k in ("rx", "ry", "rz", "rux", "ruy", "ruz", "vx", "vy", "vz")}
    for s in range(n_snap):
        for names in (("rx", "ry", "rz"), ("rux", "ruy", "ruz"), ("vx", "vy", "vz"), None):
            a = np.frombuffer(next(rec), dtype=np.float64, count=3 * n)
            if names:
                for q, k in enumerate(names):
                    out[k][s, :] = a[q * n:(q + 1) * n]
    out.update(n=int(n), L=float(L), dt=float(dt), output_interval=int(interval), n_snapshots=int(n_snap))
    return out


# ---------------------------------------------------------------------------------------------
# MSD / VACF (:404-489)
# ---------------------------------------------------------------------------------------------
def _time_origin_average(a, b, c, max_lag, origin_stride, per_origin):
    n_snap = a.shape[0]
    max_lag = n_snap - 1 if max_lag is None else max_lag
    max_lag = int(min(max_lag, n_snap - 1))              # :422 / :463
    origin_stride = max(1, int(origin_stride))           # :423 / :464
    acc = np.zeros(max_lag + 1)
    cnt = np.zeros(max_lag + 1, dtype=np.int64)
    for t0 in range(0, n_snap - 1, origin_stride):       # :428 / :469
        lmax = min(max_lag, (n_snap - 1) - t0)
        if lmax <= 0:
            continue
        acc[:lmax + 1] += per_origin(a, b, c, t0, lmax)  # means over particles, one per lag
        cnt[:lmax + 1] += 1
    used = cnt > 0
    acc[used] /= cnt[used]                               # :440-441 / :488-489
    return acc


def _msd_origin(x, y, z, t0, lmax):
    dx = x[t0:t0 + lmax + 1, :] - x[t0, :][None, :]      # :432-434
    dy = y[t0:t0 + lmax + 1, :] - y[t0, :][None, :]
    dz = z[t0:t0 + lmax + 1, :] - z[t0, :][None, :]
    return np.mean(dx * dx + dy * dy + dz * dz, axis=1)  # :435


def _vacf_origin(x, y, z, t0, lmax):
    x0, y0, z0 = x[t0, :][None, :], y[t0, :][None, :], z[t0, :][None, :]   # :473-475
    return np.mean(x[t0:t0 + lmax + 1, :] * x0 + y[t0:t0 + lmax + 1, :] * y0 + z[t0:t0 + lmax + 1, :] * z0,
                   axis=1)                                                 # :477-482


def compute_msd_tau_timeorig(rux, ruy, ruz, max_lag: Optional[int] = None, origin_stride: int = 1):
    if rux.shape[0] < 2:
        return np.array([0.0])                           # :416-417
    return _time_origin_average(rux, ruy, ruz, max_lag, origin_stride, _msd_origin)


def compute_vacf_tau_timeorig(vx, vy, vz, max_lag: Optional[int] = None, origin_stride: int = 1):
    if vx.shape[0] < 2:
        return np.array([np.mean(vx[0] * vx[0] + vy[0] * vy[0] + vz[0] * vz[0])])   # :456-458
    return _time_origin_average(vx, vy, vz, max_lag, origin_stride, _vacf_origin)


def diffusion_from_msd_linear_fit(t, msd, fit_frac: float = 0.5) -> float:
    """:492-513: slope of the least-squares line through the last `fit_frac` of MSD(t), over 6."""
    if len(t) < 10:
        return float("nan")
    fit_frac = max(0.05, min(float(fit_frac), 1.0))
    i0 = int((1.0 - fit_frac) * len(t))
    i0 = max(0, min(i0, len(t) - 5))
    design = np.vstack([t[i0:], np.ones_like(t[i0:])]).T
    slope = np.linalg.lstsq(design, msd[i0:], rcond=None)[0][0]
    return float(slope / 6.0)


def diffusion_from_vacf_integral(t, vacf, tmax_frac: float = 1.0) -> float:
    """:516-534: (1/3) x trapezoid integral of VACF up to tmax_frac of the available lags."""
    if len(t) < 2:
        return float("nan")
    tmax_frac = max(0.05, min(float(tmax_frac), 1.0))
    imax = int(math.floor(tmax_frac * (len(t) - 1)))
    imax = max(1, min(imax, len(t) - 1))
    return float((1.0 / 3.0) * np.trapezoid(vacf[:imax + 1], t[:imax + 1]))   # np.trapz in the reference


# ---------------------------------------------------------------------------------------------
# RDF (:537-598)
# ---------------------------------------------------------------------------------------------
def rdf_subsample(n_snap: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """:555-561: at most 200 snapshots and 800 particles, evenly spaced integer indices."""
    snap_idx = np.arange(n_snap) if n_snap <= 200 else np.linspace(0, n_snap - 1, 200, dtype=int)
    part_idx = np.arange(n) if n <= 800 else np.linspace(0, n - 1, 800, dtype=int)
    return snap_idx, part_idx


def rdf_pair_counts(x, y, z, L: float, nbins: int, rmax: float) -> np.ndarray:
    """Unordered-pair histogram of ONE snapshot (:566-584), int64.  All unique pairs at once; the
    per-element arithmetic is the reference's (d = r_j - r_i, np.rint(d / L), sqrt, r / dr truncated)."""
    dr = rmax / nbins
    i, j = np.triu_indices(len(x), 1)
    dx, dy, dz = x[j] - x[i], y[j] - y[i], z[j] - z[i]
    dx -= L * np.rint(dx / L)
    dy -= L * np.rint(dy / L)
    dz -= L * np.rint(dz / L)
    r = np.sqrt(dx * dx + dy * dy + dz * dz)
    bins = (r[r < rmax] / dr).astype(int)
    if bins.size and bins.max() >= nbins:
        raise IndexError("bin index out of range (np.add.at in the reference raises here)")
    return np.bincount(bins, minlength=nbins).astype(np.int64)


def rdf_normalise(pair_counts, n_samples: int, n_eff: int, L: float, nbins: int, rmax: float):
    """:563-564,586-597 with hist = 2 x unordered counts (:584)."""
    hist = 2.0 * np.asarray(pair_counts, dtype=np.float64)
    vol = L ** 3
    rho = n_eff / vol
    r_edges = np.linspace(0.0, rmax, nbins + 1)
    r_centers = 0.5 * (r_edges[:-1] + r_edges[1:])
    shell_vol = (4.0 / 3.0) * math.pi * (r_edges[1:] ** 3 - r_edges[:-1] ** 3)
    norm = n_samples * n_eff * rho * shell_vol
    g = np.zeros_like(r_centers)
    ok = norm > 0
    g[ok] = hist[ok] / norm[ok]
    return r_centers, g


def compute_rdf(rx, ry, rz, L: float, nbins: int = 200, rmax: Optional[float] = None, return_counts: bool = False):
    n_snap, n = rx.shape
    if rmax is None:
        rmax = 0.5 * L                                   # :549-550
    snap_idx, part_idx = rdf_subsample(n_snap, n)
    if len(part_idx) < 2:
        raise ValueError("Not enough particles for RDF after subsampling.")
    counts = np.zeros(nbins, dtype=np.int64)
    for s in snap_idx:
        counts += rdf_pair_counts(rx[s, part_idx], ry[s, part_idx], rz[s, part_idx], L, nbins, rmax)
    r, g = rdf_normalise(counts, len(snap_idx), len(part_idx), L, nbins, rmax)
    return (r, g, counts) if return_counts else (r, g)
