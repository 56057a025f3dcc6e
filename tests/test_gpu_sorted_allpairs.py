"""GPU parity tests of the spatially sorted all-pairs kernel (lj_pairs_sorted_kernel: large N, sparse
cutoff -- BASELINE.json configs[2], N = 32,000 at rho = 0.8, rc = 2.5).

Every unordered pair is still distance-tested exactly once like the reference loop
(/root/reference/scripts/physics/lj_potential_energy.f90:109-132); what is checked here is that the
per-tile image shift and the two-pass tile evaluation reproduce the reference's pair set (in-cutoff
pair count bit-exact) and its forces / U / virial terms to 1e-12 (forces: per atom against the sum of
pair-force magnitudes AND array-normalised), through upload / download in the caller's atom order.
The CPU checker is the cell oracle, itself pinned to the all-pairs oracle in tests/test_oracle.py."""
import os

import numpy as np
import pytest

from conftest import assert_forces_close, force_scale, make_config, melt, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _state(lj, arrs):
    return lj.SimState.from_arrays([a.copy() for a in arrs])


@pytest.fixture(scope="module")
def liquid(O):
    p_o, st = make_config(O, "cfg3")
    s = melt(O, p_o, st, 30, cells=True)
    ax, ay, az, e, d_, dd, ncand, ncut = O.compute_lj_cells(p_o, *s[:3])
    s[6][:], s[7][:], s[8][:] = ax, ay, az
    return p_o, s, (ax, ay, az), (e, d_, dd), ncut


def test_sorted_forces_and_pair_count_32k(lj, O, liquid):
    p_o, s, acc, obs_o, ncut = liquid
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    scale = force_scale(p_o.box_length, p_o.rc, *s[:3])
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, s))
        obs = ctx.compute_forces()
        got = ctx.download()
        cand, cut = ctx.pair_counts()
    n = p_o.n
    assert cand == n * (n - 1) // 2 and cut == ncut          # the reference's pair set, bit-exact count
    for k in range(6):                                       # positions / velocities come back in the caller's order, untouched
        assert np.array_equal(got.arrays()[k], s[k])
    assert_forces_close((got.ax, got.ay, got.az), acc, scale, TOL)
    for c in range(3):
        assert rel_err(got.arrays()[6 + c], acc[c]) <= TOL
    for g_, w_ in zip(obs, obs_o):
        assert abs(g_ - w_) <= TOL * abs(w_)


def test_sorted_equals_unsorted_kernel_and_is_deterministic(lj, O, liquid):
    """LJMD_FLAG_NO_SORT keeps the caller's order (every tile then takes the general minimum-image path):
    same pair set, forces equal to summation order; two sorted runs are bitwise equal."""
    p_o, s, acc, obs_o, ncut = liquid
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    outs = []
    for flags in (0, 0, lj.FLAG_NO_SORT):
        with lj.Context(p, method=lj.METHOD_ALLPAIRS, flags=flags) as ctx:
            ctx.upload(_state(lj, s))
            obs = ctx.step(3)
            d = ctx.download()
            outs.append((np.array(obs), [a.copy() for a in d.arrays()]))
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)
    assert np.all(np.abs(outs[0][0] - outs[2][0]) <= 1e-11 * np.abs(outs[2][0]))
    for a, b in zip(outs[0][1], outs[2][1]):
        assert rel_err(a, b) <= 1e-11


def test_sorted_multistep_vs_oracle_32k(lj, O, liquid):
    p_o, s, _, _, _ = liquid
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ref = [a.copy() for a in s]
    ref_obs = O.verlet_step(p_o, ref, cells=True)
    scale = force_scale(p_o.box_length, p_o.rc, *ref[:3])
    with lj.Context(p, method=lj.METHOD_ALLPAIRS, flags=lj.FLAG_TRACK_UNWRAPPED) as ctx:
        ctx.upload(_state(lj, s))
        obs = ctx.step(1)
        d = ctx.download()
        ru = ctx.download_unwrapped()
    for g_, w_ in zip(obs, ref_obs):
        assert abs(g_ - w_) <= TOL * abs(w_)
    for k in range(3):
        assert np.array_equal(d.arrays()[k], ref[k])          # drift + wrap: bitwise
        assert rel_err(d.arrays()[3 + k], ref[3 + k]) <= TOL
    assert_forces_close((d.ax, d.ay, d.az), ref[6:9], scale, TOL)
    # unwrapped coordinates: ru(t=0) = r(t=0), then += MIC(r_new - r_old), md_simulation_program.f90:339-353
    for k in range(3):
        dx = ref[k] - s[k]
        dx -= p_o.box_length * np.round(dx / p_o.box_length)
        assert np.allclose(ru[k], s[k] + dx, rtol=0, atol=1e-13)


def test_sorted_accepts_unwrapped_positions(lj, O, liquid):
    """The force routine is specified for positions outside the box too (geometry_pbc.f90:18-19)."""
    p_o, s, acc, obs_o, ncut = liquid
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    rng = np.random.default_rng(3)
    L = p_o.box_length
    u = [a.copy() for a in s]
    for k in range(3):
        u[k] = u[k] + L * rng.integers(-2, 3, size=p_o.n)
    ax, ay, az, e, d_, dd, _, ncut_u = O.compute_lj_cells(p_o, *u[:3])
    scale = force_scale(L, p_o.rc, *s[:3])
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, u))
        obs = ctx.compute_forces()
        got = ctx.download(("ax", "ay", "az"))
        _, cut = ctx.pair_counts()
    assert cut == ncut_u
    assert_forces_close((got.ax, got.ay, got.az), (ax, ay, az), scale, 4 * TOL)  # dx of shifted images carries the rounding of x +- 2L
    for g_, w_ in zip(obs, (e, d_, dd)):
        assert abs(g_ - w_) <= TOL * abs(w_)


def test_sorted_ragged_size(lj, O):
    """N not a multiple of 128 / 32: ragged last row and chunk."""
    n, L, rc = 16384 + 77, 27.5, 2.5
    rng = np.random.default_rng(11)
    # jittered simple-cubic positions (no close contacts), random order
    m = int(np.ceil(n ** (1 / 3)))
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n].astype(np.float64)
    pos = (g + 0.5 + 0.2 * (rng.random((n, 3)) - 0.5)) * (L / m)
    pos = pos[rng.permutation(n)]
    p_o = O.params_init(n, L, 1e-3, rc, 0)
    x, y, z = [np.ascontiguousarray(pos[:, c]) for c in range(3)]
    ax, ay, az, e, d_, dd, _, ncut = O.compute_lj_cells(p_o, x, y, z)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    st = [x, y, z] + [np.zeros(n) for _ in range(6)]
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, st))
        obs = ctx.compute_forces()
        got = ctx.download(("ax", "ay", "az"))
        _, cut = ctx.pair_counts()
    assert cut == ncut
    scale = force_scale(L, rc, x, y, z)
    assert_forces_close((got.ax, got.ay, got.az), (ax, ay, az), scale, TOL)
    for g_, w_ in zip(obs, (e, d_, dd)):
        assert abs(g_ - w_) <= TOL * abs(w_)


def test_sorted_resort_points_depend_on_the_step_count_only(lj, O, liquid):
    """The periodic re-sort (LJMD_SORT_INTERVAL steps after the last complete upload) changes the resident order,
    hence the summation order, hence the last bits -- but WHEN it happens depends on the step count alone:
    10 steps in one call and in calls of 3 + 1 + 6 are bitwise equal, and both stay within 1e-11 of a run that
    never re-sorts (and of the oracle)."""
    p_o, s, _, _, _ = liquid
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    outs = []
    for interval, calls in ((4, (10,)), (4, (3, 1, 6)), (0, (10,))):
        os.environ["LJMD_SORT_INTERVAL"] = str(interval)
        try:
            with lj.Context(p, method=lj.METHOD_ALLPAIRS, flags=lj.FLAG_TRACK_UNWRAPPED) as ctx:
                ctx.upload(_state(lj, s))
                for m in calls:
                    obs = ctx.step(m)
                d = ctx.download()
                ru = ctx.download_unwrapped()
                outs.append((np.array(obs), [a.copy() for a in d.arrays()] + [a.copy() for a in ru]))
        finally:
            del os.environ["LJMD_SORT_INTERVAL"]
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)
    assert np.all(np.abs(outs[0][0] - outs[2][0]) <= 1e-11 * np.abs(outs[2][0]))
    for a, b in zip(outs[0][1], outs[2][1]):
        assert rel_err(a, b) <= 1e-11
    ref = [a.copy() for a in s]
    for _ in range(10):
        ref_obs = O.verlet_step(p_o, ref, cells=True)
    assert np.all(np.abs(outs[0][0] - np.array(ref_obs)) <= 1e-10 * np.abs(np.array(ref_obs)))
    assert rel_err(outs[0][1][0], ref[0]) <= 1e-10
