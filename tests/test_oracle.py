"""CPU tests of the oracle (oracle/ljmd_oracle.c) against analytic identities, the exactly
reproducible RNG stream, the SURVEY.md Appendix-B anchors and the committed golden fixtures.

PARITY UNPINNED: the reference has no tests/golden vectors and no Fortran compiler exists here,
so none of this compares with output of the reference's own build.
"""
import os

import numpy as np
import pytest

from conftest import make_config, melt, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_params_layout_and_derived(O):
    import ctypes

    assert ctypes.sizeof(O.OrcParams) == 96  # md_types.f90:27-50 as gfortran lays it out
    p = O.params_init(500, 10.0, 1e-4, 4.9, 5)
    assert p.inv_box_length == 1.0 / 10.0
    assert p.volume == 1000.0 and p.density == 0.5
    assert p.rc_square == 4.9 * 4.9
    assert p.dt_half == 0.5e-4 and p.dt_square_half == 0.5e-4 * 1e-4


@pytest.mark.parametrize("args,code", [
    ((500, 0.0, 1e-4, 4.9), 2), ((500, 10.0, 1e-4, 0.0), 4), ((500, 10.0, 1e-4, 5.0), 7),
    ((500, 10.0, 0.0, 4.9), 8),
])
def test_params_guards(O, args, code):
    with pytest.raises(O.OracleError) as e:
        O.params_init(*args)
    assert e.value.code == code


def test_rng_stream_exact(O):
    # all values are integers < 4e6 scaled by 2.5e-7 => exact
    g = O.Rng(-12345)
    d = [g.uniform() for _ in range(1500)]
    expect = [0.6283575, 0.81823825, 0.169642, 0.19663425, 0.9045145, 0.60583975]
    for a, b in zip(d[:6], expect):
        assert abs(a - b) < 1e-15
        assert abs(a / 2.5e-7 - round(a / 2.5e-7)) < 1e-6
    assert abs(sum(d) - 741.57097625) < 1e-9
    assert all(0.0 <= x < 1.0 for x in d)


def test_minimum_image_examples(O):
    # geometry_pbc.f90:70-74 worked example and ties (dnint rounds half away from zero)
    assert abs(O.minimum_image(9.5, 10.0, 0.1) - (-0.5)) < 1e-15
    assert O.minimum_image(5.0, 10.0, 0.1) == -5.0   # dnint(0.5) = 1
    assert O.minimum_image(-5.0, 10.0, 0.1) == 5.0   # dnint(-0.5) = -1
    assert O.minimum_image(0.3, 10.0, 0.1) == 0.3


def test_wrap_examples(O):
    # geometry_pbc.f90:33-37
    rx = np.array([9.8, 10.2, -0.4, 20.7, -1e-17])
    ry, rz = rx.copy(), rx.copy()
    O.wrap_positions(rx, ry, rz, 10.0)
    assert np.allclose(rx[:4], [9.8, 0.2, 9.6, 0.7], atol=1e-14)
    assert rx[4] == 10.0  # tiny negative wraps to exactly L
    assert np.all(ry == rx) and np.all(rz == rx)


def test_tail_constants_analytic(O):
    p = O.params_init(500, 10.0, 1e-4, 4.9, 5)
    tu, td, tdd = O.tail_corrections(p)
    rho, rc, n = 0.5, 4.9, 500
    # standard LJ tails: U = (8/3) pi N rho [1/(3 rc^9) - 1/rc^3]
    u_an = (8.0 / 3.0) * np.pi * n * rho * (1.0 / (3 * rc ** 9) - 1.0 / rc ** 3)
    assert abs(tu - u_an) < 1e-12 * abs(u_an)
    assert abs(tu - (-17.80163591736684)) < 1e-10
    assert abs(td - 106.80724319074815) < 1e-9
    assert abs(tdd - (-747.6198345738022)) < 1e-9


def test_cfg1_anchors(O):
    p, st = make_config(O, "cfg1")
    ax, ay, az, e, d, dd = O.compute_lj(p, *st[:3])
    assert O.count_pairs_in_cutoff(p, *st[:3]) == 62000
    assert abs(e - (-1617.29012097504)) < 1e-8
    assert abs(d - 8566.376918596876) < 1e-7
    assert abs(dd - (-46316.27274313783)) < 1e-6
    assert max(np.abs(ax).max(), np.abs(ay).max(), np.abs(az).max()) < 1e-13  # perfect FCC
    k = O.ekin_driver(*st[3:6])
    assert abs(k - 1062.29012097504) < 1e-8
    assert abs((e + k) - (-555.0)) < 1e-10
    s = [a.copy() for a in st]
    o1 = O.run_steps(p, 1, s)
    assert abs(o1[0] - (-1617.2902880702816)) < 1e-8 and abs(o1[1] - 1062.290288070596) < 1e-8
    o10 = O.run_steps(p, 9, s)
    assert abs(o10[0] - (-1616.8239881221884)) < 1e-7
    assert abs(o10[0] + o10[1] - (-554.516553827825)) < 1e-7


def _fcc_shell_sums(a, rc):
    """Independent known answer for a perfect FCC crystal: sums over the lattice vectors
    R = (a/2)(i,j,k), i+j+k even, 0 < |R| < rc, of R^-12 and R^-6, plus the neighbour count.  No pair
    loop and no minimum image: valid while rc < L/2 so every neighbour is a distinct image."""
    m = int(np.ceil(2.0 * rc / a)) + 1
    g = np.arange(-m, m + 1)
    i, j, k = np.meshgrid(g, g, g, indexing="ij")
    keep = ((i + j + k) % 2 == 0) & ((i != 0) | (j != 0) | (k != 0))
    r2 = (0.5 * a) ** 2 * (i[keep] ** 2 + j[keep] ** 2 + k[keep] ** 2).astype(np.float64)
    r2 = np.sort(r2[r2 < rc * rc])[::-1]  # small terms first
    return float(np.sum(r2 ** -6.0)), float(np.sum(r2 ** -3.0)), int(r2.size)


def test_fcc_lattice_sum_known_answer(O):
    # the enumerator itself against the literature FCC lattice sums (in nearest-neighbour units):
    # A12 = 12.131880196..., A6 = 14.453920885... (Kittel; converged to 1e-9 / 3e-3 at rc = 12 d)
    s12, s6, _ = _fcc_shell_sums(np.sqrt(2.0), 12.0)
    assert abs(s12 - 12.131880196) < 1e-8
    assert 0.0 < 14.453920885 - s6 < 4.0 * np.pi * np.sqrt(2.0) / (3.0 * 12.0 ** 3) * 1.01  # missing tail: integral bound
    # cfg1 (k = 5, L = 10, rc = 4.9): U, r dU/dr and r^2 d2U/dr2 sums of lj_potential_energy.f90:132-193
    p, st = make_config(O, "cfg1")
    ax, ay, az, e, d, dd = O.compute_lj(p, *st[:3])
    tu, td, tdd = O.tail_corrections(p)
    s12, s6, nnb = _fcc_shell_sums(10.0 / 5.0, 4.9)
    n = 500
    assert nnb == 248 and O.count_pairs_in_cutoff(p, *st[:3]) == n * nnb // 2
    half = 0.5 * n
    assert rel_err(e, half * 4.0 * (s12 - s6) + tu) < 1e-12
    assert rel_err(d, half * 24.0 * (s6 - 2.0 * s12) + td) < 1e-12
    assert rel_err(dd, half * 24.0 * (26.0 * s12 - 7.0 * s6) + tdd) < 1e-12


def test_virial_sums_are_the_scaling_derivatives_of_the_energy(O):
    # Independent consistency of the three reductions (lj_potential_energy.f90:140,143,178 + tails
    # :205-223): scaling positions, box and cutoff by lambda keeps every r/rc, so no pair crosses the
    # cutoff and U(lambda) is smooth;  dU/dlambda = sum r dU/dr = d_epot and d2U/dlambda2 =
    # sum r^2 d2U/dr2 = dd_epot at lambda = 1 (the tail terms scale like rho, rc and obey the same identity)
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 300)
    n, L, rc = 500, 10.0, 4.9
    _, _, _, e0, d0, dd0 = O.compute_lj(p, *s[:3])

    def energy(lam):
        q = O.params_init(n, L * lam, 1e-4, rc * lam, 5)
        return O.compute_lj(q, *[a * lam for a in s[:3]])[3]

    h = 1e-3
    em, ep, em2, ep2 = energy(1 - h), energy(1 + h), energy(1 - 2 * h), energy(1 + 2 * h)
    d_fd = (8.0 * (ep - em) - (ep2 - em2)) / (12.0 * h)                          # 4th-order central differences
    dd_fd = (-(ep2 + em2) + 16.0 * (ep + em) - 30.0 * e0) / (12.0 * h * h)
    assert abs(d_fd - d0) < 1e-8 * abs(d0)
    assert abs(dd_fd - dd0) < 1e-6 * abs(dd0)


def test_accelerations_are_minus_the_energy_gradient(O):
    # a_i = -dU/dr_i (unit mass; lj_potential_energy.f90:148-159,189-191), checked by central differences
    # on atoms none of whose pairs lies within 1e-3 of the cutoff (the potential is truncated, not shifted)
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 300)
    x, y, z = [a.copy() for a in s[:3]]
    ax, ay, az, *_ = O.compute_lj(p, x, y, z)
    L, rc, h = 10.0, 4.9, 1e-5
    checked = 0
    for i in range(0, 500, 7):
        d = np.stack([x - x[i], y - y[i], z - z[i]])
        d -= L * np.rint(d / L)
        r = np.sqrt((d * d).sum(axis=0))
        if np.any(np.abs(r - rc) < 1e-3):
            continue
        for arr, acc in ((x, ax), (y, ay), (z, az)):
            keep = arr[i]
            arr[i] = keep + h
            ep = O.compute_lj(p, x, y, z)[3]
            arr[i] = keep - h
            em = O.compute_lj(p, x, y, z)[3]
            arr[i] = keep
            assert abs(-(ep - em) / (2.0 * h) - acc[i]) < 1e-5 * max(1.0, abs(acc[i]))
        checked += 1
    assert checked >= 10


def _extended_lj(p, pos):
    """lj_potential_energy.f90:109-223 restated in numpy longdouble over all i<j pairs (no code shared with the
    oracle).  Returns (acc[3][n], U, d_epot, dd_epot, in-cutoff mask, pair-force magnitude scale per atom, S12, S6)."""
    ld = np.longdouble
    n, L, rc = p.n, ld(p.box_length), ld(p.rc)
    x, y, z = (np.asarray(a, dtype=ld) for a in pos)
    i, j = np.triu_indices(n, 1)
    d = [c[i] - c[j] for c in (x, y, z)]
    d = [c - L * np.rint(c / L) for c in d]
    r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    inside = r2 < rc * rc
    ir2 = 1 / r2[inside]
    ir6 = ir2 * ir2 * ir2
    ir12 = ir6 * ir6
    du = -2 * ir12 + ir6
    pi = ld("3.1415926535897932384626433832795")
    tail = 8 * pi * ld(n) * ld(n) / (ld(p.volume) * rc ** 3)
    rc6 = rc ** 6
    u = 4 * np.sum(ir12 - ir6) + tail * (1 / (3 * rc6) - 1) / 3
    d_u = 24 * np.sum(du) + 2 * tail * (-2 / (3 * rc6) + 1)
    dd_u = 24 * np.sum(26 * ir12 - 7 * ir6) + 2 * tail * (26 / (3 * rc6) - 7)
    acc = np.zeros((3, n), dtype=ld)
    mag = np.zeros(n, dtype=ld)
    ii, jj = i[inside], j[inside]
    for c in range(3):
        f = -du * d[c][inside] * ir2
        np.add.at(acc[c], ii, f)
        np.add.at(acc[c], jj, -f)
    w = 24 * np.abs(du * ir2) * np.sqrt(r2[inside])
    np.add.at(mag, ii, w)
    np.add.at(mag, jj, w)
    return 24 * acc, u, d_u, dd_u, (i, j, inside), mag, float(np.sum(ir12)), float(np.sum(ir6))


def _need_extended():
    if np.finfo(np.longdouble).nmant < 63:
        pytest.skip("numpy longdouble is not the 80-bit extended type on this platform")


def test_forces_and_sums_against_extended_precision_on_a_disordered_state(O):
    """An independent restatement of lj_potential_energy.f90:109-223 in numpy `longdouble` (x87 80-bit: 64-bit
    significand, 11 more bits than the oracle's arithmetic), vectorised over all i<j pairs, sharing no code with
    the oracle: the formulas of the reference file, minimum image as d - L*rint(d/L), strict cutoff, tails.  On a
    melted config-1 state the oracle's U, d_epot, dd_epot and every acceleration component must agree with it to
    a few ulp of what was summed -- which bounds how far the checker itself is from exact arithmetic, and so how
    much of the 1e-12 budget of the GPU parity tests the checker may use up (none worth mentioning)."""
    _need_extended()
    ld = np.longdouble
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 400)
    acc, want_u, want_du, want_ddu, (i, j, inside), mag, s12, s6 = _extended_lj(p, s[:3])
    # the cutoff decision is the reference's: taken on r2 in DOUBLE, in its operation order (:129,:132)
    dd = [np.asarray(s[c])[i] - np.asarray(s[c])[j] for c in range(3)]
    dd = [c - p.box_length * np.round(c * p.inv_box_length) for c in dd]
    assert np.array_equal(inside, ((dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2]) < p.rc_square)
    ax, ay, az, u, du_o, ddu_o = O.compute_lj(p, *s[:3])
    # scales: the sums of the MAGNITUDES that were added up (cancellation-aware, like conftest.scalar_scales)
    eps = 2.0 ** -52
    assert abs(ld(u) - want_u) <= 32 * eps * (4 * (s12 + s6) + 20)
    assert abs(ld(du_o) - want_du) <= 32 * eps * (24 * (2 * s12 + s6) + 110)
    assert abs(ld(ddu_o) - want_ddu) <= 32 * eps * (24 * (26 * s12 + 7 * s6) + 750)
    for got, want in zip((ax, ay, az), acc):
        assert np.max(np.abs(np.asarray(got, dtype=ld) - want) / mag) <= 8 * eps


def test_verlet_step_against_extended_precision(O):
    """verlet.f90:58-95 restated in longdouble around the extended-precision force routine: one step from a melted
    state.  Positions agree to an ulp of the box length (one atom may land on the other side of a wrap, which the
    comparison takes modulo L), velocities to 1e-15 of the thermal velocity, E_kin to 1e-14."""
    _need_extended()
    ld = np.longdouble
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 400)
    L, dt = ld(p.box_length), ld(p.dt)
    r = [np.asarray(a, dtype=ld) for a in s[:3]]
    v = [np.asarray(a, dtype=ld) for a in s[3:6]]
    a0 = [np.asarray(a, dtype=ld) for a in s[6:9]]
    rn = [(r[c] + v[c] * dt) + a0[c] * (dt * dt / 2) for c in range(3)]
    rn = [c - L * np.floor(c / L) for c in rn]
    vh = [v[c] + a0[c] * (dt / 2) for c in range(3)]
    a1, want_u, _, _, _, mag, _, _ = _extended_lj(p, rn)
    vn = [vh[c] + a1[c] * (dt / 2) for c in range(3)]
    want_k = (np.sum(vn[0] ** 2) + np.sum(vn[1] ** 2) + np.sum(vn[2] ** 2)) / 2
    got = [a.copy() for a in s]
    obs = O.verlet_step(p, got)
    eps = 2.0 ** -52
    for c in range(3):
        dr = np.asarray(got[c], dtype=ld) - rn[c]
        dr = dr - L * np.rint(dr / L)
        assert np.max(np.abs(dr)) <= 4 * eps * float(L)
        assert np.max(np.abs(np.asarray(got[3 + c], dtype=ld) - vn[c])) <= 16 * eps * float(np.max(np.abs(vn[c])))
        assert np.max(np.abs(np.asarray(got[6 + c], dtype=ld) - a1[c]) / mag) <= 8 * eps
    assert abs(ld(obs[1]) - want_k) <= 64 * eps * float(want_k)
    assert abs(ld(obs[0]) - want_u) <= 1e-13 * abs(float(want_u))


def test_newton3_and_momentum(O):
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 200)
    ax, ay, az, *_ = O.compute_lj(p, *s[:3])
    scale = np.abs(ax).max()
    assert abs(ax.sum()) < 1e-10 * scale and abs(ay.sum()) < 1e-10 * scale
    assert abs(s[3].sum()) < 1e-9 and abs(s[4].sum()) < 1e-9  # COM velocity stays ~0


def test_energy_conservation_smooth_case(O):
    # NVE: the potential is truncated, not shifted (lj_potential_energy.f90:132-140), so every
    # cutoff crossing jumps E by |U(rc)| = 0.016; the total stays within 1 % over 300 steps
    p, st = make_config(O, "small_cells")
    s = [a.copy() for a in st]
    _, ser = O.run_steps(p, 300, s, want_series=True)
    etot = ser[:, 0] + ser[:, 1]
    assert np.ptp(etot) < 1e-2 * abs(etot[0])


def test_velocity_verlet_is_time_reversible(O):
    # verlet.f90:58-95 is the symmetric kick-drift-kick scheme: n steps, v -> -v, n steps returns to
    # the start up to rounding (the wrap of geometry_pbc.f90 only shifts by multiples of L)
    p, st = make_config(O, "small_cells")
    s = melt(O, p, st, 100)
    start = [a.copy() for a in s]
    O.run_steps(p, 200, s)
    assert max(np.abs(s[q] - start[q]).max() for q in range(3)) > 1e-2  # it did move
    for q in (3, 4, 5):
        s[q] *= -1.0
    O.run_steps(p, 200, s)
    L = 10.259855680060184
    for q in range(3):
        d = s[q] - start[q]
        d -= L * np.rint(d / L)
        assert np.abs(d).max() < 1e-9
        assert np.abs(s[q + 3] + start[q + 3]).max() < 1e-9


def test_golden_cfg1(O):
    g = np.load(os.path.join(GOLD, "cfg1_n500.npz"))
    p, st = make_config(O, "cfg1")
    assert np.array_equal(np.array(st), g["state0"])  # bit-exact regression
    s = [a.copy() for a in st]
    o1 = O.run_steps(p, 1, s)
    assert np.array_equal(np.array(s), g["state1"]) and np.array_equal(o1, g["obs1"])
    o10 = O.run_steps(p, 9, s)
    assert np.array_equal(o10, g["obs10"])
    assert int(g["pairs_in_cutoff_t0"]) == 62000


def test_cell_oracle_matches_allpairs(O):
    g = np.load(os.path.join(GOLD, "cells_n864.npz"))
    p = O.params_init(864, float(g["L"]), 1e-3, 2.5, 6)
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    a = O.compute_lj(p, *s[:3])
    b = O.compute_lj_cells(p, *s[:3])
    for k in range(3):
        assert rel_err(b[k], a[k]) < 1e-13
    for k in range(3, 6):
        assert abs(a[k] - b[k]) <= 1e-13 * abs(a[k])
    assert b[7] == O.count_pairs_in_cutoff(p, *s[:3]) == int(g["n_cut"])
    cell = O.cell_assign(p, *s[:3])
    start, order = O.cell_sort(p, cell)
    assert np.array_equal(cell, g["cell"]) and np.array_equal(order, g["order"])
    assert start[-1] == 864 and np.all(np.diff(start) >= 0)
    assert np.all(np.diff(cell[order]) >= 0)  # sorted by cell
    for c in range(len(start) - 1):           # stable: ascending original index within a cell
        seg = order[start[c]:start[c + 1]]
        assert np.all(np.diff(seg) > 0)


def test_cell_oracle_32k_anchor(O):
    p, st = make_config(O, "cfg3")
    r = O.compute_lj_cells(p, *st[:3])
    tu = O.tail_corrections(p)[0]
    assert r[7] == 864000                       # 54 neighbours/atom on FCC
    assert abs((r[3] - tu) - (-203671.888067)) < 1e-5
    assert O.cells_per_edge(p) == 13


def test_unwrap_matches_displacement(O):
    p, st = make_config(O, "cfg1")
    s = [a.copy() for a in st]
    ru = [a.copy() for a in s[:3]]
    O.run_steps(p, 50, s, ru=ru)
    # unwrapped and wrapped differ by integer multiples of L
    for w, u in zip(s[:3], ru):
        k = (u - w) / p.box_length
        assert np.abs(k - np.round(k)).max() < 1e-9


def test_neighbor_list_oracle_matches_bruteforce(O):
    """orc_neighbor_list (cells, range rc+skin) == brute-force numpy minimum-image search."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cells_n864.npz"))
    L = float(g["L"])
    p = O.params_init(864, L, 1e-3, 2.5, 6)
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    rng = 2.8
    assert O.cells_per_edge_range(p, rng) == 3 and O.cells_per_edge_range(p, 2.5) == O.cells_per_edge(p) == 4
    counts, nbr = O.neighbor_list(p, rng, *s[:3], cap=160)
    pos = np.stack(s[:3], -1)
    d = pos[:, None, :] - pos[None, :, :]
    d = d - L * np.sign(d * (1.0 / L)) * np.floor(np.abs(d * (1.0 / L)) + 0.5)  # dnint
    r2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
    np.fill_diagonal(r2, np.inf)
    want = r2 < rng * rng
    assert np.array_equal(counts, want.sum(1).astype(np.int32))
    for i in (0, 17, 400, 863):
        assert np.array_equal(nbr[i, :counts[i]], np.nonzero(want[i])[0].astype(np.int32))
    # symmetric: j in list(i) <=> i in list(j)
    pairs = {(i, int(j)) for i in range(864) for j in nbr[i, :counts[i]]}
    assert all((j, i) in pairs for (i, j) in pairs)
    # range cells: same definition as the rc cells with another edge
    cell = O.cell_assign_range(p, rng, *s[:3])
    start, order = O.cell_sort_range(p, rng, cell)
    assert start[-1] == 864 and np.array_equal(np.sort(order), np.arange(864))
    assert np.all(np.diff(cell[order]) >= 0)


def test_cell_binning_wraps_out_of_box_input(O):
    p = O.params_init(864, 12.0, 1e-3, 2.5, 6)
    x = np.array([0.5, 12.5, -11.5, 24.5] + [1.0] * 860)
    y = np.full(864, 1.0)
    cell = O.cell_assign(p, x, y, y.copy())
    assert cell[0] == cell[1] == cell[2] == cell[3]


def test_golden_nlist_n864(O):
    """The committed neighbour-list fixture (tests/golden/make_golden.py) against the oracle."""
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    g, gn = np.load(os.path.join(here, "cells_n864.npz")), np.load(os.path.join(here, "nlist_n864.npz"))
    p = O.params_init(864, float(g["L"]), 1e-3, 2.5, 6)
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    rng = float(gn["range"])
    cell = O.cell_assign_range(p, rng, *s[:3])
    start, order = O.cell_sort_range(p, rng, cell)
    counts, nbr = O.neighbor_list(p, rng, *s[:3], cap=144)
    assert np.array_equal(cell, gn["cell"]) and np.array_equal(start, gn["start"]) and np.array_equal(order, gn["order"])
    assert np.array_equal(counts, gn["counts"]) and np.array_equal(nbr, gn["nbr"].astype(np.int32))


@pytest.mark.parametrize("seed,k,L,rc", [(1, 5, 7.6, 2.5), (2, 6, 10.3, 2.5), (3, 6, 12.9, 2.5), (4, 7, 9.1, 2.9),
                                          (5, 5, 9.0, 3.0)])
def test_cell_paths_on_random_boxes_with_boundary_atoms(O, seed, k, L, rc):
    """Randomised cross-check of the cell machinery the GPU large-N path is compared with: jittered
    lattices in boxes with 3, 4 and 5 cells per edge, some atoms exactly at 0 and exactly at L (wrap can
    return L).  Cell-list forces / sums == all-pairs oracle (1e-12), neighbour sets == brute force."""
    rng = np.random.default_rng(seed)
    n = 4 * k ** 3
    a = L / k
    g = (np.arange(k) + 0.25) * a
    base = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    pos = (base[:, None, :] + basis[None, :, :]).reshape(-1, 3) + rng.uniform(-0.12, 0.12, (n, 3)) * a
    pos = pos - L * np.floor(pos / L)
    edge = rng.choice(n, 12, replace=False)
    pos[edge[:6], rng.integers(0, 3, 6)] = 0.0
    pos[edge[6:], rng.integers(0, 3, 6)] = L
    x, y, z = [np.ascontiguousarray(pos[:, q]) for q in range(3)]
    p = O.params_init(n, L, 1e-3, rc, k)
    nc = O.cells_per_edge(p)
    if L == 3.0 * rc:
        # documented back-off: a cell edge within 1e-12 of the range gives up one cell per edge, and two
        # cells per edge are refused (a 27-cell stencil would visit a neighbour cell twice)
        assert nc == 2
        with pytest.raises(O.OracleError):
            O.compute_lj_cells(p, x, y, z)
        return
    assert nc == int(np.floor(L / rc)) and nc >= 3
    ref = O.compute_lj(p, x, y, z)
    got = O.compute_lj_cells(p, x, y, z)
    scale = max(np.abs(ref[q]).max() for q in range(3))
    for q in range(3):
        assert np.abs(got[q] - ref[q]).max() < 1e-12 * scale
    for q in (3, 4, 5):
        assert abs(got[q] - ref[q]) < 1e-12 * abs(ref[q])
    assert got[7] == O.count_pairs_in_cutoff(p, x, y, z)
    # neighbour list with a skin against a brute-force minimum-image search
    rl = rc + 0.3
    if O.cells_per_edge_range(p, rl) >= 3:
        counts, nbr = O.neighbor_list(p, rl, x, y, z, cap=400)
        d = pos[:, None, :] - pos[None, :, :]
        d = d - L * np.sign(d * (1.0 / L)) * np.floor(np.abs(d * (1.0 / L)) + 0.5)  # dnint
        r2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
        np.fill_diagonal(r2, np.inf)
        want = r2 < rl * rl
        assert np.array_equal(counts, want.sum(1).astype(np.int32))
        for i in list(edge) + [0, n - 1]:
            assert np.array_equal(nbr[i, :counts[i]], np.nonzero(want[i])[0].astype(np.int32))
