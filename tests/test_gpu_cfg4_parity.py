"""Parity at BASELINE.json configs[3] itself: N = 1,048,576 atoms, rho = 0.8, rc = 2.5, neighbour list.
One velocity-Verlet step (physics/verlet.f90:41-97) of the GPU path against the CPU cell oracle (about
10 s of CPU work), plus the bit-exact integer tables (cell of every atom, cell starts, sorted order) at
full size.  Forces: per atom against the sum of the pair-force magnitudes AND array-normalised, 1e-12."""
import numpy as np
import pytest

from conftest import assert_forces_close, force_scale, rel_err, scalar_scales

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
TOL = 1e-12


def test_cfg4_single_step_vs_cell_oracle(lj, O):
    import bench

    w = bench.WORKLOADS["cfg4_nlist"]
    n, arrs = bench.synthetic_state(w, seed=3)
    assert n == 1048576
    p_o = O.params_init(n, w["L"], w["dt"], w["rc"], w["k"])
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ax, ay, az, e0, d0, dd0, ncand, ncut = O.compute_lj_cells(p_o, *arrs[:3])   # a(t), what the drivers do at t = 0
    s0 = [a.copy() for a in arrs]
    s0[6][:], s0[7][:], s0[8][:] = ax, ay, az
    ref = [a.copy() for a in s0]
    ref_obs = O.verlet_step(p_o, ref, cells=True)
    with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
        ctx.upload(lj.SimState.from_arrays([a.copy() for a in s0]))
        f_obs = ctx.compute_forces()
        f = ctx.download(("ax", "ay", "az"))
        info = ctx.nlist_info()
        cell, start, order = ctx.download_cells()
        _, cut = ctx.pair_counts()
        # the step starts from the oracle's a(t) (compute_forces above left the GPU's own, equal to 1e-12 but not
        # bitwise, and the drift r + v dt + a dt^2/2 of the step would inherit its last bits)
        ctx.upload(lj.SimState.from_arrays([a.copy() for a in s0]))
        obs = ctx.step(1)
        got = ctx.download()
    # t = 0 forces and the reference's pair set
    assert cut == ncut
    scale0, s12, s6 = force_scale(w["L"], w["rc"], *s0[:3], return_sums=True)
    assert_forces_close((f.ax, f.ay, f.az), (ax, ay, az), scale0, TOL)
    # U, d_epot, dd_epot: 1e-12 of the summed magnitudes (on this jittered lattice dd_epot = 24 (26 S12 - 7 S6)
    # + tail = 7e5 is the difference of 1e8-size sums: a bound relative to the value itself would ask for
    # more digits than a 2.7e7-term FP64 sum has in any order); U also 1e-12 relative to itself
    sc = scalar_scales(s12, s6, lj.tail_corrections(p))
    for g_, w_, s_ in zip(f_obs, (e0, d0, dd0), sc):
        assert abs(g_ - w_) <= TOL * s_
    assert abs(f_obs[0] - e0) <= TOL * abs(e0)
    # integer tables, bit-exact at full size
    rng = p_o.rc + info["skin"]
    cell_o = O.cell_assign_range(p_o, rng, *s0[:3])
    start_o, order_o = O.cell_sort_range(p_o, rng, cell_o)
    assert np.array_equal(cell, cell_o) and np.array_equal(start, start_o) and np.array_equal(order, order_o)
    # one step
    scale1, s12, s6 = force_scale(w["L"], w["rc"], *ref[:3], return_sums=True)
    sc = scalar_scales(s12, s6, lj.tail_corrections(p))
    assert abs(obs[0] - ref_obs[0]) <= TOL * abs(ref_obs[0]) and abs(obs[1] - ref_obs[1]) <= TOL * abs(ref_obs[1])
    assert abs(obs[2] - ref_obs[2]) <= TOL * sc[1] and abs(obs[3] - ref_obs[3]) <= TOL * sc[2]
    for k in range(3):
        assert np.array_equal(got.arrays()[k], ref[k])                 # drift + wrap: bitwise
        assert rel_err(got.arrays()[3 + k], ref[3 + k]) <= TOL
        assert rel_err(got.arrays()[6 + k], ref[6 + k]) <= TOL
    assert_forces_close((got.ax, got.ay, got.az), ref[6:9], scale1, TOL)
