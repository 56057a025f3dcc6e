"""GPU parity tests of the trajectory consumers (csrc/md_analysis.cu through
the C ABI ljmd_rdf_histogram / ljmd_time_correlation and the analysis.py mirror).

Bars: the RDF pair histogram is integer work => BIT-EXACT against the oracle (and g(r), computed
from it by the same host expressions, equal to the last bit); MSD / VACF are floating-point means whose
summation order differs from numpy's pairwise sum => 1e-12 relative.

The file name sorts last on purpose: these kernels were written at the very end of round 1 (11 of the
tests below had one hardware run, all green), so they run after the long-validated suites."""
import importlib
import os

import numpy as np
import pytest

from conftest import make_config, melt

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
RTOL = 1e-12


@pytest.fixture(scope="module")
def A():
    from oracle import analysis_oracle

    return analysis_oracle


@pytest.fixture(scope="module")
def an(lj):
    return importlib.import_module(lj.__name__ + ".analysis")


def _oracle_counts(A, x, y, z, L, nbins, rmax):
    c = np.zeros(nbins, dtype=np.int64)
    for s in range(x.shape[0]):
        c += A.rdf_pair_counts(x[s], y[s], z[s], L, nbins, rmax)
    return c


@pytest.mark.parametrize("n,n_snap,L,nbins,rmax_frac", [
    (300, 3, 7.3, 200, 0.5),      # three atom tiles, ragged last tile
    (2, 1, 3.0, 10, 0.5),         # smallest input
    (129, 2, 11.7, 64, 0.6),      # rmax beyond L/2 is allowed by the reference
    (257, 1, 5.1, 9000, 0.5),     # more bins than fit the shared-memory histogram
])
def test_rdf_pair_histogram_bit_exact(an, A, n, n_snap, L, nbins, rmax_frac):
    rng = np.random.default_rng(1000 + n)
    x, y, z = rng.random((3, n_snap, n)) * L
    rmax = rmax_frac * L
    got = an.rdf_pair_counts(x, y, z, L, nbins, rmax)
    want = _oracle_counts(A, x, y, z, L, nbins, rmax)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    assert an.last_kernel_ms["rdf"] > 0.0


def test_rdf_fcc_lattice_distances_on_bin_edges(an, A):
    # perfect lattice: many distances coincide, some sit exactly on bin edges -- the truncation
    # int(r/dr) must take the same side as numpy's
    k, a = 4, 2.0
    g = np.arange(k) * a
    x0, y0, z0 = [c.reshape(-1, 1) for c in np.meshgrid(g, g, g, indexing="ij")]
    h = 0.5 * a
    x = np.concatenate([x0, x0, x0 + h, x0 + h], axis=1).reshape(1, -1)
    y = np.concatenate([y0, y0 + h, y0, y0 + h], axis=1).reshape(1, -1)
    z = np.concatenate([z0, z0 + h, z0 + h, z0], axis=1).reshape(1, -1)
    for nbins, rmax in ((390, 3.9), (200, 4.0), (8, 4.0)):
        got = an.rdf_pair_counts(x, y, z, 8.0, nbins, rmax)
        assert np.array_equal(got, _oracle_counts(A, x, y, z, 8.0, nbins, rmax))
    assert got.sum() > 0


def test_compute_rdf_matches_oracle_on_a_liquid_with_subsampling(an, A, O):
    # N = 864 > 800: the evenly spaced 800-particle subsample of the reference (:559-561)
    p, st = make_config(O, "small_cells")
    s = melt(O, p, st, 300, cells=False)
    snaps = []
    for _ in range(2):
        O.run_steps(p, 20, s)
        snaps.append([a.copy() for a in s[:3]])
    rx, ry, rz = [np.stack([sn[q] for sn in snaps]) for q in range(3)]
    L = 10.259855680060184
    r, g = an.compute_rdf(rx, ry, rz, L)
    r0, g0, c0 = A.compute_rdf(rx, ry, rz, L, return_counts=True)
    assert np.array_equal(r, r0) and np.array_equal(g, g0)
    assert g[:20].max() == 0.0 and 2.0 < g.max() < 4.0   # excluded core, first liquid peak
    # many snapshots: 205 > 200 are thinned to 200 (:555-557)
    rng = np.random.default_rng(5)
    x, y, z = rng.random((3, 205, 12)) * 4.0
    r, g = an.compute_rdf(x, y, z, 4.0, nbins=25)
    r0, g0 = A.compute_rdf(x, y, z, 4.0, nbins=25)
    assert np.array_equal(g, g0)


def test_rdf_full_reference_size_properties(an):
    # 200 snapshots x 800 particles = the most the reference ever histograms (6.4e7 pairs)
    rng = np.random.default_rng(21)
    L, nbins = 10.0, 200
    x, y, z = rng.random((3, 200, 800)) * L
    c = an.rdf_pair_counts(x, y, z, L, nbins, 0.5 * L)
    assert 0 < c.sum() <= 200 * 800 * 799 // 2
    frac = c.sum() / (200 * 800 * 799 / 2)
    assert abs(frac - (4.0 / 3.0) * np.pi * 0.125) < 2e-3        # volume fraction of the r < L/2 sphere
    perm = rng.permutation(800)
    assert np.array_equal(c, an.rdf_pair_counts(x[:, perm], y[:, perm], z[:, perm], L, nbins, 0.5 * L))
    r, g = an.compute_rdf(x, y, z, L)
    assert np.all(np.abs(g[20:] - 1.0) < 0.05)                   # ideal gas


@pytest.mark.parametrize("n_snap,n,max_lag,stride", [(40, 77, None, 1), (40, 77, 7, 3), (9, 500, 100, 2),
                                                      (2, 33, None, 1), (1, 33, None, 1), (25, 1, 0, 1)])
def test_msd_vacf_match_oracle(an, A, n_snap, n, max_lag, stride):
    rng = np.random.default_rng(n_snap * 1000 + n)
    v = rng.normal(size=(3, n_snap, n))
    ru = np.cumsum(v, axis=1) * 0.01 + rng.random((3, 1, n)) * 5.0
    msd = an.compute_msd_tau_timeorig(*ru, max_lag=max_lag, origin_stride=stride)
    vacf = an.compute_vacf_tau_timeorig(*v, max_lag=max_lag, origin_stride=stride)
    msd0 = A.compute_msd_tau_timeorig(*ru, max_lag=max_lag, origin_stride=stride)
    vacf0 = A.compute_vacf_tau_timeorig(*v, max_lag=max_lag, origin_stride=stride)
    assert msd.shape == msd0.shape and vacf.shape == vacf0.shape
    assert np.allclose(msd, msd0, rtol=RTOL, atol=0.0)
    assert np.allclose(vacf, vacf0, rtol=RTOL, atol=1e-15)       # VACF crosses zero: tiny absolute floor
    if n_snap >= 2 and (max_lag is None or max_lag > 0):
        assert msd[0] == 0.0 and msd[1] > 0.0


def test_msd_ballistic_known_answer_at_production_size(an):
    # size of the shipped production run's trajectory: 2000 x 500, lags to 1000.  r = r0 + v t
    # gives MSD(tau) = <v^2> tau^2 and VACF(tau) = <v^2> for every origin
    rng = np.random.default_rng(8)
    n_snap, n, max_lag = 2000, 500, 1000
    v = rng.normal(size=(3, 1, n))
    t = np.arange(n_snap, dtype=np.float64)[None, :, None] * 0.01
    ru = rng.random((3, 1, n)) * 10.0 + v * t
    vv = np.broadcast_to(v, (3, n_snap, n))
    v2 = float(np.mean(v[0, 0] ** 2 + v[1, 0] ** 2 + v[2, 0] ** 2))
    tau = np.arange(max_lag + 1) * 0.01
    msd = an.compute_msd_tau_timeorig(*ru, max_lag=max_lag, origin_stride=5)
    vacf = an.compute_vacf_tau_timeorig(*vv, max_lag=max_lag, origin_stride=5)
    assert np.allclose(msd, v2 * tau ** 2, rtol=1e-9, atol=1e-12)
    assert np.allclose(vacf, v2, rtol=1e-12)
    assert an.last_kernel_ms["msd"] > 0.0 and an.last_kernel_ms["vacf"] > 0.0
    d = an.diffusion_from_vacf_integral(tau, vacf)
    assert abs(d - v2 * tau[-1] / 3.0) < 1e-9 * d


@pytest.mark.parametrize("n_snap,n,max_lag,stride", [(70, 77, None, 1), (70, 77, 33, 1), (41, 130, 9, 3),
                                                      (2, 5, None, 1), (300, 500, 200, 5)])
def test_tiled_time_correlation_is_bit_identical_to_the_default_kernel(an, A, n_snap, n, max_lag, stride):
    rng = np.random.default_rng(n_snap + n)
    v = rng.normal(size=(3, n_snap, n))
    ru = np.cumsum(v, axis=1) * 0.01 + rng.random((3, 1, n)) * 5.0
    # the register-tiled kernel is the default; LJMD_CORR_TILED=0 selects the one-warp-per-lag kernel
    got_m = an.compute_msd_tau_timeorig(*ru, max_lag=max_lag, origin_stride=stride)
    got_v = an.compute_vacf_tau_timeorig(*v, max_lag=max_lag, origin_stride=stride)
    os.environ["LJMD_CORR_TILED"] = "0"
    try:
        want_m = an.compute_msd_tau_timeorig(*ru, max_lag=max_lag, origin_stride=stride)
        want_v = an.compute_vacf_tau_timeorig(*v, max_lag=max_lag, origin_stride=stride)
    finally:
        del os.environ["LJMD_CORR_TILED"]
    assert np.array_equal(got_m, want_m) and np.array_equal(got_v, want_v)
    assert np.allclose(got_m, A.compute_msd_tau_timeorig(*ru, max_lag=max_lag, origin_stride=stride), rtol=RTOL, atol=0.0)


# ---------------------------------------------------------------------------------------------------
# Against OUTPUT OF THE REFERENCE ITSELF: tests/golden/analysis_ref.npz (made by
# tests/golden/make_golden_analysis.py, which runs md_one_run_analysis.py's own functions)
# ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "analysis_ref.npz"))


def test_rdf_kernel_equals_the_reference_output(an, G):
    # g(r) = 2 c / norm with integer pair counts c from the kernel: equal to the reference's g TO THE BIT
    L = float(G["liquid_header"][1])
    pos = [G["liquid_r" + c] for c in "xyz"]
    r, g = an.compute_rdf(*pos, L)
    assert np.array_equal(r, G["liquid_rdf_r"]) and np.array_equal(g, G["liquid_rdf_g"])
    r, g = an.compute_rdf(*pos, L, nbins=64, rmax=2.0)
    assert np.array_equal(r, G["liquid_rdf64_r"]) and np.array_equal(g, G["liquid_rdf64_g"])
    r, g = an.compute_rdf(G["gas864_x"], G["gas864_y"], G["gas864_z"], float(G["gas864_L"]), nbins=50)
    assert np.array_equal(r, G["gas864_rdf_r"]) and np.array_equal(g, G["gas864_rdf_g"])
    r, g = an.compute_rdf(G["many_x"], G["many_y"], G["many_z"], float(G["many_L"]), nbins=25)
    assert np.array_equal(r, G["many_rdf_r"]) and np.array_equal(g, G["many_rdf_g"])


@pytest.mark.parametrize("traj,tag,kw", [
    ("liquid", "full", {}), ("liquid", "lag7_stride3", dict(max_lag=7, origin_stride=3)),
    ("liquid", "lag100_stride2", dict(max_lag=100, origin_stride=2)),
    ("walk", "full", {}), ("walk", "lag20_stride4", dict(max_lag=20, origin_stride=4)), ("walk", "lag0", dict(max_lag=0)),
])
def test_msd_vacf_kernels_equal_the_reference_output(an, G, traj, tag, kw):
    if traj == "liquid":
        ru, v = [G["liquid_ru" + c] for c in "xyz"], [G["liquid_v" + c] for c in "xyz"]
    else:
        ru, v = G["walk_ru"], G["walk_v"]
    msd = an.compute_msd_tau_timeorig(*ru, **kw)
    vacf = an.compute_vacf_tau_timeorig(*v, **kw)
    want_m, want_v = G[f"{traj}_msd_{tag}"], G[f"{traj}_vacf_{tag}"]
    assert msd.shape == want_m.shape and vacf.shape == want_v.shape
    assert np.allclose(msd, want_m, rtol=RTOL, atol=0.0)
    # VACF crosses zero: absolute floor of 1e-12 x VACF(0)
    assert np.allclose(vacf, want_v, rtol=RTOL, atol=RTOL * abs(float(G[f"{traj}_vacf_full"][0])))


def test_msd_vacf_kernels_short_inputs_equal_the_reference_output(an, G):
    for ns in (1, 2):
        ru = [G["liquid_ru" + c][:ns] for c in "xyz"]
        v = [G["liquid_v" + c][:ns] for c in "xyz"]
        assert np.allclose(an.compute_msd_tau_timeorig(*ru), G[f"liquid_msd_first{ns}"], rtol=RTOL, atol=0.0)
        assert np.allclose(an.compute_vacf_tau_timeorig(*v), G[f"liquid_vacf_first{ns}"], rtol=RTOL, atol=0.0)
