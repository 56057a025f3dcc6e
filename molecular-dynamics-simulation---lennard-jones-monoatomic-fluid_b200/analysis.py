"""Trajectory consumers on the GPU: the same function names, argument
meaning and return values as the reference's post-processing script, with the two heavy loops --
the O(n^2)-per-snapshot pair histogram behind g(r) and the O(n_snap^2 n) time-origin averages --
running in libljmd.so (csrc/md_analysis.cu).  What stays here is host work on O(nbins) / O(n_snap)
numbers: subsampling indices, normalisation, the two diffusion estimates.

Reference (file:line relative to /root/reference/scripts/md_one_run_analysis.py):
  :404-441 compute_msd_tau_timeorig        :444-489 compute_vacf_tau_timeorig
  :492-513 diffusion_from_msd_linear_fit   :516-534 diffusion_from_vacf_integral
  :537-598 compute_rdf                     :345-397 read_rva (see initial_config.read_rva / RvaWriter)

No CPU fallback: without the CUDA library or a GPU every compute call raises LjmdError.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import numpy as np

from . import LjmdError, _chk, load_library

CORR_MSD, CORR_VACF = 0, 1
last_kernel_ms = {"rdf": None, "msd": None, "vacf": None}  # CUDA-event time of the most recent launch

_DP = C.POINTER(C.c_double)
_I64P = C.POINTER(C.c_int64)


def _dev_array(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _bind():
    L = load_library()
    if not getattr(L, "_analysis_bound", False):
        L.ljmd_rdf_histogram.argtypes = [_DP, _DP, _DP, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double,
                                         _I64P, C.c_int32, _DP]
        L.ljmd_time_correlation.argtypes = [_DP, _DP, _DP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            _DP, C.c_int32, _DP]
        L._analysis_bound = True
    return L


# ---------------------------------------------------------------------------------------------
# g(r)
# ---------------------------------------------------------------------------------------------
def rdf_pair_counts(rx, ry, rz, L: float, nbins: int, rmax: float, device: int = -1) -> np.ndarray:
    """Unordered-pair histogram of all snapshots given (arrays (n_snap, n)), int64, bit-exact
    with the reference's pair loop (:566-584).  An input for which the reference raises IndexError
    (r < rmax but int(r/dr) == nbins) raises IndexError here too."""
    x, y, z = _dev_array(rx), _dev_array(ry), _dev_array(rz)
    if x.ndim != 2 or y.shape != x.shape or z.shape != x.shape:
        raise ValueError("rdf_pair_counts(): rx, ry, rz must be (n_snap, n) arrays of one shape")
    hist = np.zeros(int(nbins), dtype=np.int64)
    ms = C.c_double(0.0)
    lib = _bind()
    code = lib.ljmd_rdf_histogram(x.ctypes.data_as(_DP), y.ctypes.data_as(_DP), z.ctypes.data_as(_DP),
                                  x.shape[0], x.shape[1], float(L), int(nbins), float(rmax),
                                  hist.ctypes.data_as(_I64P), int(device), C.byref(ms))
    if code == 9 and b"IndexError" in lib.ljmd_last_error():
        raise IndexError(lib.ljmd_last_error().decode())
    _chk(code)
    last_kernel_ms["rdf"] = ms.value
    return hist


def compute_rdf(rx, ry, rz, L: float, nbins: int = 200, rmax: Optional[float] = None,
                device: int = -1) -> Tuple[np.ndarray, np.ndarray]:
    """compute_rdf (:537-598): returns (r_centers, g_r).  At most 200 evenly spaced snapshots and
    800 evenly spaced particles enter, as in the reference (:555-561)."""
    rx, ry, rz = np.asarray(rx), np.asarray(ry), np.asarray(rz)
    n_snap, n = rx.shape
    L = float(L)
    rmax = 0.5 * L if rmax is None else float(rmax)
    snaps = np.arange(n_snap) if n_snap <= 200 else np.linspace(0, n_snap - 1, 200, dtype=int)
    parts = np.arange(n) if n <= 800 else np.linspace(0, n - 1, 800, dtype=int)
    n_eff, n_samples = len(parts), len(snaps)
    if n_eff < 2:
        raise ValueError("Not enough particles for RDF after subsampling.")
    pick = np.ix_(snaps, parts)
    counts = rdf_pair_counts(rx[pick], ry[pick], rz[pick], L, nbins, rmax, device)
    # normalisation :563-564,586-597; every pair enters the reference's histogram twice (:584)
    edges = np.linspace(0.0, rmax, nbins + 1)
    centers = 0.5 * (edges[:-1] + edges[1:])
    shell = (4.0 / 3.0) * math.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    density = n_eff / L ** 3
    norm = n_samples * n_eff * density * shell
    g = np.zeros_like(centers)
    np.divide(2.0 * counts, norm, out=g, where=norm > 0)
    return centers, g


# ---------------------------------------------------------------------------------------------
# MSD(tau), VACF(tau)
# ---------------------------------------------------------------------------------------------
def _time_correlation(a, b, c, max_lag, origin_stride, kind, device):
    a, b, c = _dev_array(a), _dev_array(b), _dev_array(c)
    if a.ndim != 2 or b.shape != a.shape or c.shape != a.shape:
        raise ValueError("time correlation: the three arrays must be (n_snap, n) and of one shape")
    n_snap, n = a.shape
    lag = n_snap - 1 if max_lag is None else int(min(int(max_lag), n_snap - 1))   # :420-422
    if lag < 0:
        raise ValueError("max_lag must be >= 0")
    if n_snap < 2:
        lag = 0
    out = np.zeros(lag + 1)
    ms = C.c_double(0.0)
    _chk(_bind().ljmd_time_correlation(a.ctypes.data_as(_DP), b.ctypes.data_as(_DP), c.ctypes.data_as(_DP),
                                       n_snap, n, lag, max(1, int(origin_stride)), kind,
                                       out.ctypes.data_as(_DP), int(device), C.byref(ms)))
    last_kernel_ms["msd" if kind == CORR_MSD else "vacf"] = ms.value
    return out


def compute_msd_tau_timeorig(rux, ruy, ruz, max_lag: Optional[int] = None, origin_stride: int = 1,
                             device: int = -1) -> np.ndarray:
    """MSD(tau) = < |ru(t+tau) - ru(t)|^2 > over particles and time origins (:404-441)."""
    return _time_correlation(rux, ruy, ruz, max_lag, origin_stride, CORR_MSD, device)


def compute_vacf_tau_timeorig(vx, vy, vz, max_lag: Optional[int] = None, origin_stride: int = 1,
                              device: int = -1) -> np.ndarray:
    """VACF(tau) = < v(t) . v(t+tau) > over particles and time origins (:444-489)."""
    return _time_correlation(vx, vy, vz, max_lag, origin_stride, CORR_VACF, device)


# ---------------------------------------------------------------------------------------------
# diffusion estimates: O(n_snap) host arithmetic
# ---------------------------------------------------------------------------------------------
def diffusion_from_msd_linear_fit(t, msd, fit_frac: float = 0.5) -> float:
    """Einstein relation MSD ~ 6 D t on the last `fit_frac` of the lags (:492-513): least-squares
    slope / 6 (closed form of the straight-line fit; the reference calls lstsq)."""
    t, msd = np.asarray(t, dtype=np.float64), np.asarray(msd, dtype=np.float64)
    if len(t) < 10:
        return float("nan")
    frac = min(max(float(fit_frac), 0.05), 1.0)
    first = min(max(int((1.0 - frac) * len(t)), 0), len(t) - 5)
    x, y = t[first:], msd[first:]
    xc = x - x.mean()
    return float(np.dot(xc, y - y.mean()) / np.dot(xc, xc) / 6.0)


def diffusion_from_vacf_integral(t, vacf, tmax_frac: float = 1.0) -> float:
    """Green-Kubo D = (1/3) int VACF dt, trapezoids up to `tmax_frac` of the lags (:516-534)."""
    t, vacf = np.asarray(t, dtype=np.float64), np.asarray(vacf, dtype=np.float64)
    if len(t) < 2:
        return float("nan")
    frac = min(max(float(tmax_frac), 0.05), 1.0)
    last = min(max(int(math.floor(frac * (len(t) - 1))), 1), len(t) - 1)
    x, y = t[:last + 1], vacf[:last + 1]
    return float((1.0 / 3.0) * np.sum(np.diff(x) * (y[1:] + y[:-1]) / 2.0))


__all__ = ["compute_rdf", "rdf_pair_counts", "compute_msd_tau_timeorig", "compute_vacf_tau_timeorig",
           "diffusion_from_msd_linear_fit", "diffusion_from_vacf_integral", "LjmdError"]
