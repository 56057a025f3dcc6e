import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (checker only)."""
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def lj():
    """The product package; the CUDA library must already be built (no fallback)."""
    import __graft_entry__ as g

    lib = os.path.join(g.PKG, "lib", "libljmd.so")
    if not os.path.exists(lib):
        g.build()
    import ljmd_b200

    ljmd_b200.load_library()
    return ljmd_b200


# ---- the benchmark / parity configurations of SURVEY.md section 8d ------------------------
CONFIGS = {
    # name: (k, L, rc, dt, E_target or None, T_target or None)
    "cfg1": dict(k=5, L=10.0, rc=4.9, dt=1e-4, e_target=-555.0),
    "cfg2": dict(k=10, L=20.0, rc=9.8, dt=1e-4, e_target=-4440.0),
    "cfg3": dict(k=20, L=34.19951893353394, rc=2.5, dt=1e-3, t_target=1.0),
    "small_cells": dict(k=6, L=10.259855680060184, rc=2.5, dt=1e-3, t_target=1.0),  # N=864, rho=0.8
}


def make_config(O, name, seed=-12345):
    """Initial configuration as md_initial_config_program builds it (no warm-up)."""
    c = CONFIGS[name]
    n = 4 * c["k"] ** 3
    p = O.params_init(n, c["L"], c["dt"], c["rc"], c["k"])
    cells = O.cells_per_edge(p) >= 3 and n > 2000
    st, _, _ = O.initial_config(p, seed, c.get("e_target", 0.0) or 0.0,
                                c.get("t_target", 0.0) or 0.0, cells=cells)
    return p, st


def melt(O, p, st, nsteps, cells=False):
    """Advance a copy of the state with the oracle to get a disordered configuration."""
    s = [a.copy() for a in st]
    for _ in range(nsteps):
        O.verlet_step(p, s, cells=cells)
    return s


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = np.max(np.abs(want))
    if scale == 0.0:
        return float(np.max(np.abs(got - want)))
    return float(np.max(np.abs(got - want)) / scale)


def force_scale(L, rc, x, y, z, return_sums=False):
    """Per-atom cancellation-aware scale of the acceleration: s_i = 24 * sum_j |g_ij| * |d_ij| over the
    in-cutoff neighbours (g = (2 r^-12 - r^-6) / r^2, lj_potential_energy.f90:143-150), i.e. the sum of
    the MAGNITUDES of the pair forces that add up to a_i.  "Forces within 1e-12 relative" is tested per
    atom against this scale (|da_i| <= tol * s_i) next to the array-normalised rel_err(): an atom whose
    pair forces nearly cancel must still be right to 1e-12 of what was summed.  Independent of the oracle:
    scipy's periodic k-d tree finds the pairs."""
    from scipy.spatial import cKDTree

    pos = np.stack([x, y, z], axis=1)
    pos = pos - L * np.floor(pos / L)
    pos[pos >= L] -= L
    pairs = cKDTree(pos, boxsize=L).query_pairs(rc, output_type="ndarray")
    d = pos[pairs[:, 0]] - pos[pairs[:, 1]]
    d -= L * np.rint(d / L)
    r2 = (d * d).sum(axis=1)
    ir2 = 1.0 / r2
    ir6 = ir2 ** 3
    mag = 24.0 * np.abs((2.0 * ir6 * ir6 - ir6) * ir2) * np.sqrt(r2)
    s = np.zeros(len(x))
    np.add.at(s, pairs[:, 0], mag)
    np.add.at(s, pairs[:, 1], mag)
    if return_sums:
        return s, float(np.sum(ir6 * ir6)), float(np.sum(ir6))
    return s


def scalar_scales(s12, s6, tails):
    """Cancellation-aware scales of (epot, d_epot, dd_epot): the sums of the MAGNITUDES of what is added up,
    4 (S12 + S6), 24 (2 S12 + S6), 24 (26 S12 + 7 S6) plus |tail| (lj_potential_energy.f90:140,143,178,
    188-223).  d_epot = 24 (-2 S12 + S6) + tail passes through zero near the zero-pressure density, where a
    bound relative to d_epot itself would ask for more digits than a 1e6-term FP64 sum has in ANY order."""
    return (4.0 * (s12 + s6) + abs(tails[0]), 24.0 * (2.0 * s12 + s6) + abs(tails[1]),
            24.0 * (26.0 * s12 + 7.0 * s6) + abs(tails[2]))


def assert_forces_close(got_xyz, want_xyz, scale, tol=1e-12, what="a"):
    """Per-atom bound |got_i - want_i| <= tol * scale_i (see force_scale) on every component."""
    for c, (g, w) in enumerate(zip(got_xyz, want_xyz)):
        err = np.abs(np.asarray(g) - np.asarray(w))
        bad = np.nonzero(err > tol * scale)[0]
        assert len(bad) == 0, (f"{what}{'xyz'[c]}: {len(bad)} atoms outside {tol:g} of their pair-force scale; worst atom "
                               f"{bad[np.argmax(err[bad] / scale[bad])]} err/scale "
                               f"{np.max(err[bad] / scale[bad]):.3e}")
