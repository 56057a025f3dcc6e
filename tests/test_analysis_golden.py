"""The trajectory-consumer oracle and the product's host pieces against OUTPUT OF THE REFERENCE ITSELF
(SURVEY.md section 8 row f4): tests/golden/analysis_ref.npz holds what the reference's own
md_one_run_analysis.py functions (read_rva :345-397, compute_msd_tau_timeorig :404-441,
compute_vacf_tau_timeorig :444-489, diffusion_from_msd_linear_fit :492-513, diffusion_from_vacf_integral
:516-534, compute_rdf :537-598) returned in the build container for the inputs stored next to them
(tests/golden/make_golden_analysis.py is the generating script).  This is what pins
oracle/analysis_oracle.py; the CUDA kernels are compared with the same fixture in
tests/test_zz_gpu_analysis.py.

Bars: the oracle keeps the reference's numpy expressions, so on the numpy the fixture was made with it is
expected to be BIT-IDENTICAL; the bound asserted is 4 ulp (1e-15 relative) so that a numpy whose pairwise
summation blocks differently still passes.  g(r) is integer counting followed by one division: exact."""
import hashlib
import importlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TIGHT = 1e-15


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "analysis_ref.npz"))


@pytest.fixture(scope="module")
def A():
    from oracle import analysis_oracle

    return analysis_oracle


@pytest.fixture(scope="module")
def an(lj):
    return importlib.import_module(lj.__name__ + ".analysis")


def _close(got, want, rtol=TIGHT, atol=0.0):
    got, want = np.asarray(got), np.asarray(want)
    return got.shape == want.shape and np.allclose(got, want, rtol=rtol, atol=atol)


def _write_liquid_file(lj, G, path):
    """The rva.dat file of the fixture, rebuilt by the product's RvaWriter from the stored arrays."""
    ic = importlib.import_module(lj.__name__ + ".initial_config")
    n, L, dt, interval, n_snap = G["liquid_header"]
    n, interval, n_snap = int(n), int(interval), int(n_snap)
    with ic.RvaWriter(str(path), n, float(L), float(dt), interval, n_snap) as w:
        for s in range(n_snap):
            arrays = [G["liquid_" + k][s] for k in ("rx", "ry", "rz", "vx", "vy", "vz")] + list(G["liquid_acc"][s])
            w.write_sample(lj.SimState.from_arrays([np.ascontiguousarray(a) for a in arrays]),
                           [G["liquid_ru" + c][s] for c in "xyz"])
    return n, float(L), float(dt), interval, n_snap


def test_rva_file_the_reference_read_is_what_the_writer_produces(lj, G, A, tmp_path):
    # the reference's read_rva accepted a file with this sha256 and returned the stored arrays; RvaWriter
    # must produce those very bytes from them, and both readers of this repo must return the same arrays
    path = tmp_path / "rva.dat"
    n, L, dt, interval, n_snap = _write_liquid_file(lj, G, path)
    assert hashlib.sha256(path.read_bytes()).digest() == G["liquid_rva_sha256"].tobytes()
    d = A.read_rva(path)
    assert (d["n"], d["L"], d["dt"], d["output_interval"], d["n_snapshots"]) == (n, L, dt, interval, n_snap)
    for name in ("rx", "ry", "rz", "rux", "ruy", "ruz", "vx", "vy", "vz"):
        assert np.array_equal(d[name], G["liquid_" + name]), name
    ic = importlib.import_module(lj.__name__ + ".initial_config")
    hdr, samples = ic.read_rva(str(path))
    assert hdr == {"n": n, "box_length": L, "dt": dt, "output_interval": interval, "n_snapshots": n_snap}
    assert len(samples) == n_snap
    for s in (0, n_snap - 1):
        r, ru, v, a = samples[s]
        assert np.array_equal(r[1], G["liquid_ry"][s]) and np.array_equal(ru[2], G["liquid_ruz"][s])
        assert np.array_equal(v[0], G["liquid_vx"][s]) and np.array_equal(a, G["liquid_acc"][s])


@pytest.mark.parametrize("tag,kw", [("full", {}), ("lag7_stride3", dict(max_lag=7, origin_stride=3)),
                                    ("lag100_stride2", dict(max_lag=100, origin_stride=2))])
def test_oracle_msd_vacf_equal_the_reference_on_the_liquid(A, G, tag, kw):
    ru = [G["liquid_ru" + c] for c in "xyz"]
    v = [G["liquid_v" + c] for c in "xyz"]
    assert _close(A.compute_msd_tau_timeorig(*ru, **kw), G[f"liquid_msd_{tag}"])
    assert _close(A.compute_vacf_tau_timeorig(*v, **kw), G[f"liquid_vacf_{tag}"], atol=1e-16)


def test_oracle_msd_vacf_short_inputs_equal_the_reference(A, G):
    for ns in (1, 2):
        ru = [G["liquid_ru" + c][:ns] for c in "xyz"]
        v = [G["liquid_v" + c][:ns] for c in "xyz"]
        assert _close(A.compute_msd_tau_timeorig(*ru), G[f"liquid_msd_first{ns}"])
        assert _close(A.compute_vacf_tau_timeorig(*v), G[f"liquid_vacf_first{ns}"])


@pytest.mark.parametrize("tag,kw", [("full", {}), ("lag20_stride4", dict(max_lag=20, origin_stride=4)),
                                    ("lag0", dict(max_lag=0))])
def test_oracle_msd_vacf_equal_the_reference_on_the_walk(A, G, tag, kw):
    assert _close(A.compute_msd_tau_timeorig(*G["walk_ru"], **kw), G[f"walk_msd_{tag}"])
    assert _close(A.compute_vacf_tau_timeorig(*G["walk_v"], **kw), G[f"walk_vacf_{tag}"], atol=1e-16)


def test_oracle_rdf_equals_the_reference(A, G):
    pos = [G["liquid_r" + c] for c in "xyz"]
    L = float(G["liquid_header"][1])
    r, g = A.compute_rdf(*pos, L)
    assert np.array_equal(r, G["liquid_rdf_r"]) and np.array_equal(g, G["liquid_rdf_g"])
    r, g = A.compute_rdf(*pos, L, nbins=64, rmax=2.0)
    assert np.array_equal(r, G["liquid_rdf64_r"]) and np.array_equal(g, G["liquid_rdf64_g"])
    assert g.max() > 2.0 and g[:20].max() == 0.0          # first peak of the liquid, excluded core
    # 864 > 800 particles: evenly spaced subsample (:559-561); 205 > 200 snapshots: thinning (:555-557)
    r, g = A.compute_rdf(G["gas864_x"], G["gas864_y"], G["gas864_z"], float(G["gas864_L"]), nbins=50)
    assert np.array_equal(r, G["gas864_rdf_r"]) and np.array_equal(g, G["gas864_rdf_g"])
    r, g = A.compute_rdf(G["many_x"], G["many_y"], G["many_z"], float(G["many_L"]), nbins=25)
    assert np.array_equal(r, G["many_rdf_r"]) and np.array_equal(g, G["many_rdf_g"])


def test_diffusion_estimates_equal_the_reference(A, an, G):
    t, msd, vacf = G["walk_t"], G["walk_msd_full"], G["walk_vacf_full"]
    for f, want in zip((0.5, 0.2, 1.0, 0.0), G["walk_D_msd"]):
        assert abs(A.diffusion_from_msd_linear_fit(t, msd, f) - want) <= 1e-13 * abs(want)
        # the product uses the closed form of the straight-line fit instead of lstsq
        assert abs(an.diffusion_from_msd_linear_fit(t, msd, f) - want) <= 1e-10 * abs(want)
    for f, want in zip((1.0, 0.8, 0.0), G["walk_D_vacf"]):
        assert abs(A.diffusion_from_vacf_integral(t, vacf, f) - want) <= 1e-15 * abs(want)
        assert abs(an.diffusion_from_vacf_integral(t, vacf, f) - want) <= 1e-14 * abs(want)


def test_product_rdf_host_part_equals_the_reference_given_oracle_counts(A, an, G, monkeypatch):
    # analysis.compute_rdf with the GPU histogram replaced by the (now pinned) oracle's pair counts: the
    # subsampling, factor 2 and normalisation on the host must reproduce the reference's g(r) to the bit
    def counts_from_oracle(rx, ry, rz, L, nbins, rmax, device=-1):
        c = np.zeros(nbins, dtype=np.int64)
        for s in range(rx.shape[0]):
            c += A.rdf_pair_counts(rx[s], ry[s], rz[s], L, nbins, rmax)
        return c

    monkeypatch.setattr(an, "rdf_pair_counts", counts_from_oracle)
    r, g = an.compute_rdf(G["gas864_x"], G["gas864_y"], G["gas864_z"], float(G["gas864_L"]), nbins=50)
    assert np.array_equal(r, G["gas864_rdf_r"]) and np.array_equal(g, G["gas864_rdf_g"])
    r, g = an.compute_rdf(G["many_x"], G["many_y"], G["many_z"], float(G["many_L"]), nbins=25)
    assert np.array_equal(r, G["many_rdf_r"]) and np.array_equal(g, G["many_rdf_g"])
    L = float(G["liquid_header"][1])
    r, g = an.compute_rdf(*[G["liquid_r" + c] for c in "xyz"], L, nbins=64, rmax=2.0)
    assert np.array_equal(r, G["liquid_rdf64_r"]) and np.array_equal(g, G["liquid_rdf64_g"])
