"""CPU tests of the trajectory-consumer oracle (oracle/analysis_oracle.py)
against known answers, of the rva.dat writer against an independent reader, and of the host-side
pieces of the product mirror (analysis.py).

The comparison with OUTPUT OF THE REFERENCE'S OWN md_one_run_analysis.py is tests/test_analysis_golden.py
(fixture tests/golden/analysis_ref.npz); the tests here are the independent known answers."""
import importlib
import struct

import numpy as np
import pytest

from conftest import make_config


@pytest.fixture(scope="module")
def A():
    from oracle import analysis_oracle

    return analysis_oracle


@pytest.fixture(scope="module")
def an(lj):
    return importlib.import_module(lj.__name__ + ".analysis")


def _fcc(k, a):
    g = np.arange(k) * a
    x0, y0, z0 = [c.reshape(-1, 1) for c in np.meshgrid(g, g, g, indexing="ij")]
    h = 0.5 * a
    x = np.concatenate([x0, x0, x0 + h, x0 + h], axis=1).reshape(1, -1)
    y = np.concatenate([y0, y0 + h, y0, y0 + h], axis=1).reshape(1, -1)
    z = np.concatenate([z0, z0 + h, z0 + h, z0], axis=1).reshape(1, -1)
    return x, y, z


def test_rdf_fcc_shell_counts(A):
    # FCC, a = 2, L = 8 (k = 4, N = 256): neighbour shells at a*sqrt(m/2), m = 1, 2, 3, 4, 5, 6 hold
    # 12, 6, 24, 12, 24, 8 atoms; every shell below rmax = L/2 is resolved by bins of 0.01
    x, y, z = _fcc(4, 2.0)
    n, L, nbins, rmax = 256, 8.0, 390, 3.9
    r, g, counts = A.compute_rdf(x, y, z, L, nbins=nbins, rmax=rmax, return_counts=True)
    shells = {1: 12, 2: 6, 3: 24, 4: 12, 5: 24, 6: 8, 7: 48}
    expect = np.zeros(nbins, dtype=np.int64)
    for m, z_m in shells.items():
        d = 2.0 * np.sqrt(m / 2.0)
        if d < rmax:
            expect[int(d / (rmax / nbins))] += n * z_m // 2
    assert counts.sum() == expect.sum()
    # a shell that sits numerically on a bin edge may land one bin lower: compare with a 1-bin blur
    nz = np.flatnonzero(counts)
    for b in nz:
        assert expect[max(b - 1, 0):b + 2].sum() >= counts[b]
    assert len(nz) == len(np.flatnonzero(expect))
    # normalisation: g = 2 c / (n_samples n rho V_shell)
    b = nz[0]
    edges = np.linspace(0.0, rmax, nbins + 1)
    shell = 4.0 / 3.0 * np.pi * (edges[b + 1] ** 3 - edges[b] ** 3)
    assert abs(g[b] - 2.0 * counts[b] / (n * (n / L ** 3) * shell)) < 1e-12 * g[b]
    assert np.allclose(r, 0.5 * (edges[:-1] + edges[1:]))


def test_rdf_ideal_gas_is_flat_and_subsampling(A):
    rng = np.random.default_rng(7)
    L = 9.0
    pos = rng.random((3, 3, 1000)) * L           # 1000 > 800 particles: subsampled to 800
    r, g, counts = A.compute_rdf(pos[:, 0], pos[:, 1], pos[:, 2], L, nbins=30, return_counts=True)
    snaps, parts = A.rdf_subsample(3, 1000)
    assert len(snaps) == 3 and len(parts) == 800 and parts[0] == 0 and parts[-1] == 999
    assert len(np.unique(parts)) == 800
    assert counts.sum() <= 3 * 800 * 799 // 2
    assert np.all(np.abs(g[5:] - 1.0) < 0.1)     # ~1e4 pairs per bin beyond r = 0.75
    snaps, parts = A.rdf_subsample(1001, 10)
    assert len(snaps) == 200 and snaps[0] == 0 and snaps[-1] == 1000 and len(parts) == 10
    with pytest.raises(ValueError):
        A.compute_rdf(pos[:, 0, :1], pos[:, 1, :1], pos[:, 2, :1], L)


def test_rdf_minimum_image_and_pair_order(A):
    # two atoms across the periodic boundary: distance 1.0 through the wall, not 9.0
    x = np.array([[0.5, 9.5]]); y = np.array([[5.0, 5.0]]); z = np.array([[5.0, 5.0]])
    c = A.rdf_pair_counts(x[0], y[0], z[0], 10.0, 50, 5.0)
    assert c.sum() == 1 and c[10] == 1
    c2 = A.rdf_pair_counts(x[0, ::-1].copy(), y[0], z[0], 10.0, 50, 5.0)
    assert np.array_equal(c, c2)


def test_msd_vacf_known_answers(A):
    rng = np.random.default_rng(11)
    n_snap, n = 12, 37
    v = rng.normal(size=(3, n))
    r0 = rng.random((3, n))
    t = np.arange(n_snap, dtype=np.float64)[:, None]
    ru = [r0[q][None, :] + v[q][None, :] * t for q in range(3)]        # ballistic: MSD = <v^2> tau^2
    vv = [np.repeat(v[q][None, :], n_snap, axis=0) for q in range(3)]  # constant v: VACF = <v^2>
    v2 = np.mean(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    msd = A.compute_msd_tau_timeorig(*ru)
    vacf = A.compute_vacf_tau_timeorig(*vv)
    assert msd.shape == (n_snap,) and vacf.shape == (n_snap,)
    assert np.allclose(msd, v2 * np.arange(n_snap) ** 2, rtol=1e-12, atol=1e-13)
    assert np.allclose(vacf, v2, rtol=1e-13)
    # max_lag clamps, stride thins the origins, the last origin n_snap-1 is never used (:428)
    assert A.compute_msd_tau_timeorig(*ru, max_lag=100).shape == (n_snap,)
    m3 = A.compute_msd_tau_timeorig(*ru, max_lag=3, origin_stride=4)
    assert m3.shape == (4,) and np.allclose(m3, v2 * np.arange(4) ** 2, rtol=1e-12, atol=1e-13)
    # single snapshot (:416-417, :456-458)
    assert np.array_equal(A.compute_msd_tau_timeorig(*[a[:1] for a in ru]), [0.0])
    assert np.allclose(A.compute_vacf_tau_timeorig(*[a[:1] for a in vv]), [v2], rtol=1e-14)
    # max_lag = 0: every origin is skipped (L <= 0, :430), the result stays [0]
    assert np.array_equal(A.compute_msd_tau_timeorig(*ru, max_lag=0), [0.0])


def test_msd_origin_bookkeeping_against_direct_double_loop(A):
    rng = np.random.default_rng(3)
    n_snap, n, max_lag, stride = 9, 5, 6, 2
    a, b, c = rng.normal(size=(3, n_snap, n))
    want = np.zeros(max_lag + 1); cnt = np.zeros(max_lag + 1)
    for t0 in range(0, n_snap - 1, stride):
        for lag in range(0, min(max_lag, n_snap - 1 - t0) + 1):
            want[lag] += np.mean((a[t0 + lag] - a[t0]) ** 2 + (b[t0 + lag] - b[t0]) ** 2 + (c[t0 + lag] - c[t0]) ** 2)
            cnt[lag] += 1
    got = A.compute_msd_tau_timeorig(a, b, c, max_lag=max_lag, origin_stride=stride)
    assert np.allclose(got, want / cnt, rtol=1e-14)
    assert list(cnt) == [4, 4, 4, 3, 3, 2, 2]


def test_diffusion_estimates(A, an):
    t = np.linspace(0.0, 5.0, 101)
    msd = 6.0 * 0.37 * t + 0.2                     # D = 0.37 exactly
    assert abs(A.diffusion_from_msd_linear_fit(t, msd) - 0.37) < 1e-12
    assert abs(an.diffusion_from_msd_linear_fit(t, msd) - 0.37) < 1e-12
    vacf = 3.0 * np.exp(-2.0 * t)                  # (1/3) int_0^5 3 e^(-2t) dt = (1 - e^-10)/2
    for f in (A.diffusion_from_vacf_integral, an.diffusion_from_vacf_integral):
        assert abs(f(t, vacf) - 0.5 * (1.0 - np.exp(-10.0))) < 5e-4       # trapezoid error h^2/12 * |f'(0)| / 3 = 4.2e-4
    rng = np.random.default_rng(5)
    noisy = msd + rng.normal(scale=0.05, size=t.size)
    wig = vacf + rng.normal(scale=0.01, size=t.size)
    for frac in (0.5, 0.2, 1.0, 0.0, 7.0):
        assert abs(A.diffusion_from_msd_linear_fit(t, noisy, frac) - an.diffusion_from_msd_linear_fit(t, noisy, frac)) < 1e-11
        assert abs(A.diffusion_from_vacf_integral(t, wig, frac) - an.diffusion_from_vacf_integral(t, wig, frac)) < 1e-14
    assert np.isnan(an.diffusion_from_msd_linear_fit(t[:9], msd[:9])) and np.isnan(A.diffusion_from_msd_linear_fit(t[:9], msd[:9]))
    assert np.isnan(an.diffusion_from_vacf_integral(t[:1], vacf[:1])) and np.isnan(A.diffusion_from_vacf_integral(t[:1], vacf[:1]))


def test_rva_writer_is_read_by_the_independent_reader(A, lj, O, tmp_path):
    ic = importlib.import_module(lj.__name__ + ".initial_config")
    p, st = make_config(O, "cfg1")
    n = 500
    s = [a.copy() for a in st]
    ru = [a.copy() for a in s[:3]]
    snaps = []
    path = tmp_path / "rva.dat"
    with ic.RvaWriter(str(path), n, 10.0, 1e-4, 5, 3) as w:
        for _ in range(3):
            O.run_steps(p, 5, s, ru=ru)
            state = lj.SimState.from_arrays([a.copy() for a in s])
            w.write_sample(state, [a.copy() for a in ru])
            snaps.append(([a.copy() for a in s], [a.copy() for a in ru]))
    d = A.read_rva(path)
    assert (d["n"], d["L"], d["dt"], d["output_interval"], d["n_snapshots"]) == (500, 10.0, 1e-4, 5, 3)
    for k, (state, unwrapped) in enumerate(snaps):
        for q, name in enumerate(("rx", "ry", "rz")):
            assert np.array_equal(d[name][k], state[q])
            assert np.array_equal(d["ru" + name[1]][k], unwrapped[q])
            assert np.array_equal(d["v" + name[1]][k], state[3 + q])
    # header record is 28 bytes between two 4-byte markers (md_simulation_program.f90:257)
    raw = path.read_bytes()
    assert struct.unpack("<i", raw[:4])[0] == 28 and struct.unpack("<i", raw[32:36])[0] == 28
    assert len(raw) == 36 + 3 * 4 * (8 + 3 * n * 8)
    # the same content with 8-byte record markers (gfortran -frecord-marker=8) is read too (:279-285)
    out = bytearray()
    pos = 0
    while pos < len(raw):
        nb = struct.unpack("<i", raw[pos:pos + 4])[0]
        out += struct.pack("<q", nb) + raw[pos + 4:pos + 4 + nb] + struct.pack("<q", nb)
        pos += 8 + nb
    p8 = tmp_path / "rva8.dat"
    p8.write_bytes(bytes(out))
    d8 = A.read_rva(p8)
    assert np.array_equal(d8["vx"], d["vx"]) and d8["n_snapshots"] == 3
    # package reader agrees with the oracle reader, for both marker sizes
    for f in (path, p8):
        hdr, samples = ic.read_rva(str(f))
        assert hdr["n"] == 500 and hdr["n_snapshots"] == 3 and len(samples) == 3
        assert np.array_equal(samples[2][0][0], d["rx"][2]) and np.array_equal(samples[1][2][1], d["vy"][1])


def test_product_analysis_fails_loudly_without_gpu(lj, an):
    if lj.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu tests")
    x = np.zeros((2, 4))
    with pytest.raises(lj.LjmdError) as e:
        an.compute_rdf(x, x, x, 10.0)
    assert e.value.code == 100 and "no CPU fallback" in e.value.message
    with pytest.raises(lj.LjmdError) as e:
        an.compute_msd_tau_timeorig(x, x, x)
    assert e.value.code == 100


def test_product_analysis_argument_guards(lj, an):
    x = np.zeros((2, 4))
    with pytest.raises(ValueError):
        an.compute_rdf(x[:, :1], x[:, :1], x[:, :1], 10.0)       # the reference's ValueError (:563-564)
    with pytest.raises(ValueError):
        an.rdf_pair_counts(x, x[:1], x, 10.0, 10, 5.0)
    with pytest.raises(ValueError):
        an.compute_vacf_tau_timeorig(x, x, x[:, :2])
    lib = lj.load_library()
    assert lib.ljmd_rdf_histogram(None, None, None, 1, 2, 10.0, 10, 5.0, None, -1, None) == 6
    assert b"NULL" in lib.ljmd_last_error()


def test_product_rdf_host_part_equals_oracle_given_the_same_counts(A, an, monkeypatch):
    # subsampling, normalisation and the factor 2 of analysis.compute_rdf, with the GPU histogram
    # replaced by the oracle's: r and g must be bitwise those of the oracle
    def counts_from_oracle(rx, ry, rz, L, nbins, rmax, device=-1):
        c = np.zeros(nbins, dtype=np.int64)
        for s in range(rx.shape[0]):
            c += A.rdf_pair_counts(rx[s], ry[s], rz[s], L, nbins, rmax)
        return c

    monkeypatch.setattr(an, "rdf_pair_counts", counts_from_oracle)
    rng = np.random.default_rng(2)
    for n_snap, n, L, nbins, rmax in ((2, 900, 7.7, 40, None), (203, 6, 3.3, 10, 1.5), (1, 50, 6.0, 200, None)):
        x, y, z = rng.random((3, n_snap, n)) * L
        r, g = an.compute_rdf(x, y, z, L, nbins=nbins, rmax=rmax)
        r0, g0 = A.compute_rdf(x, y, z, L, nbins=nbins, rmax=rmax)
        assert np.array_equal(r, r0) and np.array_equal(g, g0)
