"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
inputs.  Tolerance (north star): forces, U, virial terms within 1e-12 relative per step.
Forces are compared relative to max|a| of the configuration; scalars relative to their value."""
import os

import numpy as np
import pytest

from conftest import assert_forces_close, force_scale, make_config, melt, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _state(lj, arrs):
    return lj.SimState.from_arrays([a.copy() for a in arrs])


def _check_forces(lj, O, p_o, arrs, tol=TOL):
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = _state(lj, arrs)
    e, d, dd = lj.compute_lj_potential_energy(p, s)
    ax, ay, az, e0, d0, dd0 = O.compute_lj(p_o, *arrs[:3])
    assert rel_err(s.ax, ax) <= tol and rel_err(s.ay, ay) <= tol and rel_err(s.az, az) <= tol
    # ... and per atom, relative to the sum of the magnitudes of the pair forces on that atom
    assert_forces_close((s.ax, s.ay, s.az), (ax, ay, az), force_scale(p_o.box_length, p_o.rc, *arrs[:3]), tol)
    assert abs(e - e0) <= tol * abs(e0)
    assert abs(d - d0) <= tol * abs(d0)
    assert abs(dd - dd0) <= tol * abs(dd0)
    return s


def test_device_present(lj):
    assert lj.device_count() >= 1


def test_compute_lj_cfg1_melted(lj, O):
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 200)
    _check_forces(lj, O, p, s)


def test_compute_lj_cfg1_fcc_t0(lj, O):
    # perfect lattice: every a_i ~ 0 (cancellation); compare absolutely against the pair scale
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = _state(lj, st)
    e, d, dd = lj.compute_lj_potential_energy(p, s)
    _, _, _, e0, d0, dd0 = O.compute_lj(p_o, *st[:3])
    assert max(np.abs(s.ax).max(), np.abs(s.ay).max(), np.abs(s.az).max()) < 1e-12
    assert abs(e - e0) <= TOL * abs(e0) and abs(d - d0) <= TOL * abs(d0) and abs(dd - dd0) <= TOL * abs(dd0)


def test_golden_cfg1(lj):
    g = np.load(os.path.join(GOLD, "cfg1_n500.npz"))
    p = lj.init_params(500, 10.0, 1e-4, 4.9, 5)
    s = _state(lj, g["state0"])
    obs = lj.verlet_step(p, s)
    for got, want in zip(obs, g["obs1"]):
        assert abs(got - want) <= TOL * abs(want)
    for k, name in enumerate(lj.SimState.NAMES):
        assert rel_err(getattr(s, name), g["state1"][k]) <= TOL, name
    s = _state(lj, g["state200"])
    e, d, dd = lj.compute_lj_potential_energy(p, s)
    for got, want in zip((e, d, dd), g["obs_f200"]):
        assert abs(got - want) <= TOL * abs(want)


def test_verlet_step_cfg2(lj, O):
    p_o, st = make_config(O, "cfg2")
    s0 = melt(O, p_o, st, 3)
    ref = [a.copy() for a in s0]
    ref_obs = O.verlet_step(p_o, ref)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = _state(lj, s0)
    obs = lj.verlet_step(p, s)
    for got, want in zip(obs, ref_obs):
        assert abs(got - want) <= TOL * abs(want)
    for name, want in zip(lj.SimState.NAMES, ref):
        assert rel_err(getattr(s, name), want) <= TOL, name
    # positions and half-kicked velocities use the reference's exact operation order
    assert np.array_equal(s.rx, ref[0]) and np.array_equal(s.ry, ref[1]) and np.array_equal(s.rz, ref[2])
    # accelerations per atom, relative to the sum of the magnitudes of the pair forces on that atom
    assert_forces_close((s.ax, s.ay, s.az), ref[6:9], force_scale(p_o.box_length, p_o.rc, *ref[:3]), TOL)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 127, 128, 129, 500, 1000, 1537])
def test_ragged_sizes_random_positions(lj, O, n):
    rng = np.random.default_rng(1234 + n)
    L = 12.0
    p = O.params_init(n, L, 1e-3, 3.0, 0)
    # random gas, but keep pairs apart so r^-12 stays finite: jittered grid
    m = int(np.ceil(n ** (1 / 3)))
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
    pos = (g + 0.5 + 0.3 * (rng.random((n, 3)) - 0.5)) * (L / m)
    arrs = [np.ascontiguousarray(pos[:, k]) for k in range(3)] + [np.zeros(n) for _ in range(6)]
    _check_forces(lj, O, p, arrs)
    assert True


def test_unwrapped_inputs_match(lj, O):
    """The force routine is specified for unwrapped inputs too (geometry_pbc.f90:18-19)."""
    p, st = make_config(O, "cfg1")
    s = melt(O, p, st, 100)
    rng = np.random.default_rng(7)
    for k in range(3):
        s[k] = s[k] + p.box_length * rng.integers(-2, 3, size=p.n)
    _check_forces(lj, O, p, s, tol=1e-11)


def test_pair_counts_match(lj, O):
    p_o, st = make_config(O, "cfg1")
    s = melt(O, p_o, st, 200)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, s))
        cand, cut = ctx.pair_counts()
    assert cand == 500 * 499 // 2
    assert cut == O.count_pairs_in_cutoff(p_o, *s[:3])


def test_newton3_exact_cancellation(lj, O):
    p_o, st = make_config(O, "cfg2")
    s0 = melt(O, p_o, st, 2)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = _state(lj, s0)
    lj.compute_lj_potential_energy(p, s)
    scale = np.abs(s.ax).max()
    assert abs(s.ax.sum()) < 1e-10 * scale * np.sqrt(p.n)


def test_run_to_run_determinism(lj, O):
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    outs = []
    for _ in range(3):
        with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
            ctx.upload(_state(lj, st))
            ctx.compute_forces()
            obs = ctx.step(200)
            d = ctx.download()
            outs.append((obs.copy(), [a.copy() for a in d.arrays()]))
    for o, arrs in outs[1:]:
        assert np.array_equal(o, outs[0][0])
        for a, b in zip(arrs, outs[0][1]):
            assert np.array_equal(a, b)  # bitwise


def test_resident_multistep_vs_oracle(lj, O):
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    nsteps = 100
    ref = [a.copy() for a in st]
    _, ref_ser = O.run_steps(p_o, nsteps, ref, want_series=True)
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, st))
        ser = ctx.step_series(nsteps)
        d = ctx.download()
    # first step within 1e-12; afterwards round-off differences grow (chaotic dynamics)
    assert np.all(np.abs(ser[0] - ref_ser[0]) <= TOL * np.abs(ref_ser[0]))
    assert np.all(np.abs(ser - ref_ser) <= 1e-9 * np.abs(ref_ser))
    assert rel_err(d.rx, ref[0]) < 1e-9


def test_graph_path_equals_single_steps(lj, O):
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as a, lj.Context(p, method=lj.METHOD_ALLPAIRS) as b:
        a.upload(_state(lj, st)); b.upload(_state(lj, st))
        oa = a.step(70)                      # graph chunks + remainder
        for _ in range(70):
            ob = b.step(1)                   # direct launches
        assert np.array_equal(oa, ob)
        da, db = a.download(), b.download()
        for x, y in zip(da.arrays(), db.arrays()):
            assert np.array_equal(x, y)
        # per step: force + fused finalize; plus one stand-alone drift per step() call
        assert a.launch_count == 1 + 70 * 2 and b.launch_count == 70 * 3


def test_unwrapped_tracking(lj, O):
    # FCC corner atoms sit at x=0: those moving in -x wrap to ~L in the first step
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ref = [a.copy() for a in st]
    ru = [a.copy() for a in ref[:3]]
    O.run_steps(p_o, 50, ref, ru=ru)
    with lj.Context(p, method=lj.METHOD_ALLPAIRS, flags=lj.FLAG_TRACK_UNWRAPPED) as ctx:
        ctx.upload(_state(lj, st))
        ctx.step(50)
        got = ctx.download_unwrapped()
        d = ctx.download()
    assert np.abs(np.array(ru) - np.array(ref[:3])).max() > 0.5 * p.box_length  # crossings happened
    for a, b in zip(got, ru):
        assert np.abs(a - b).max() < 1e-10
    assert rel_err(d.rx, ref[0]) < 1e-10


def test_replicas_are_independent(lj, O):
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    R = 3
    states = []
    for r in range(R):
        s_r, _, _ = O.initial_config(p_o, -(12345 + r), -555.0)
        states.append(s_r)
    batched = [np.ascontiguousarray(np.stack([states[r][k] for r in range(R)])) for k in range(9)]
    with lj.Context(p, method=lj.METHOD_ALLPAIRS, nreplicas=R) as ctx:
        ctx.upload(lj.SimState.from_arrays(batched))
        obs = ctx.step(20)
        d = ctx.download()
    for r in range(R):
        ref = [a.copy() for a in states[r]]
        ro = O.run_steps(p_o, 20, ref)
        assert np.all(np.abs(obs[r] - ro) <= 1e-10 * np.abs(ro))
        assert rel_err(d.rx[r], ref[0]) < 1e-10


def test_cfg5_64_replicas_per_context(lj, O):
    """BASELINE.json configs[4] as specified: 64 replicas of the shipped N = 500 system in ONE context, replica r
    started with the velocity seed -(12345 + r).  Every replica, single step from identical states: the four
    observables to 1e-12, positions bitwise, velocities and accelerations to 1e-12 (per atom for the latter)."""
    p_o, _ = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    R = 64
    states = []
    for r in range(R):
        s_r, _, _ = O.initial_config(p_o, -(12345 + r), -555.0)
        states.append(melt(O, p_o, s_r, 40))        # off the perfect lattice
    batched = [np.ascontiguousarray(np.stack([states[r][k] for r in range(R)])) for k in range(9)]
    with lj.Context(p, method=lj.METHOD_ALLPAIRS, nreplicas=R) as ctx:
        ctx.upload(lj.SimState.from_arrays(batched))
        obs = ctx.step(1)
        d = ctx.download()
    assert obs.shape == (R, 4)
    for r in range(R):
        ref = [a.copy() for a in states[r]]
        ro = O.verlet_step(p_o, ref)
        assert np.all(np.abs(obs[r] - ro) <= TOL * np.abs(ro)), r
        got = [a[r] for a in d.arrays()]
        for k in range(3):
            assert np.array_equal(got[k], ref[k])
            assert rel_err(got[3 + k], ref[3 + k]) <= TOL
        if r % 16 == 0:
            scale = force_scale(p_o.box_length, p_o.rc, *ref[:3])
            assert_forces_close(got[6:9], ref[6:9], scale, TOL)
        else:
            for k in range(3):
                assert rel_err(got[6 + k], ref[6 + k]) <= TOL


def test_kinetic_energy_driver_order(lj, O):
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, st))
        k = ctx.kinetic_energy()
    k0 = O.ekin_driver(*st[3:6])
    assert abs(k - k0) <= TOL * k0


def test_error_paths_on_gpu(lj):
    p = lj.init_params(500, 10.0, 1e-4, 4.9, 5)
    with pytest.raises(lj.LjmdError) as e:
        lj.Context(p, method=lj.METHOD_CELLS)   # floor(L/rc) = 2 < 3
    assert e.value.code == 10
    with pytest.raises(lj.LjmdError):
        lj.Context(p, nreplicas=0)


def test_fp64_peak_measurement(lj):
    tf = lj.measure_fp64_peak()
    assert 10.0 < tf < 60.0, tf


# ---------------------------------------------------------------------------------------------
# Cell-list path (no reference counterpart: integer assignments are pinned against the CPU
# cell oracle, forces against the all-pairs oracle)
# ---------------------------------------------------------------------------------------------
def test_cells_golden_n864(lj, O):
    g = np.load(os.path.join(GOLD, "cells_n864.npz"))
    p_o = O.params_init(864, float(g["L"]), 1e-3, 2.5, 6)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    with lj.Context(p, method=lj.METHOD_CELLS) as ctx:
        assert ctx.method == lj.METHOD_CELLS and ctx.cells_per_edge == O.cells_per_edge(p_o) == 4
        ctx.upload(_state(lj, s))
        obs = ctx.compute_forces()
        d = ctx.download(("ax", "ay", "az"))
        cell, start, order = ctx.download_cells()
        cand, cut = ctx.pair_counts()
    # integer assignments: bit-exact
    assert np.array_equal(cell, g["cell"]) and np.array_equal(start, g["start"]) and np.array_equal(order, g["order"])
    assert cut == int(g["n_cut"]) and cand >= int(g["n_cand"])
    for k, name in enumerate(("ax", "ay", "az")):
        assert rel_err(getattr(d, name), g["acc"][k]) <= TOL
    for got, want in zip(obs, g["obs"]):
        assert abs(got - want) <= TOL * abs(want)


def test_cells_vs_oracle_32k(lj, O):
    p_o, st = make_config(O, "cfg3")
    s = melt(O, p_o, st, 20, cells=True)
    ax, ay, az, e, d_, dd, ncand, ncut = O.compute_lj_cells(p_o, *s[:3])
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with lj.Context(p, method=lj.METHOD_CELLS) as ctx:   # 13 cells per edge at N=32000
        assert ctx.method == lj.METHOD_CELLS and ctx.cells_per_edge == 13
        ctx.upload(_state(lj, s))
        obs = ctx.compute_forces()
        got = ctx.download(("ax", "ay", "az"))
        cell, start, order = ctx.download_cells()
        _, cut = ctx.pair_counts()
    assert cut == ncut
    assert np.array_equal(cell, O.cell_assign(p_o, *s[:3]))
    st_o, order_o = O.cell_sort(p_o, cell)
    assert np.array_equal(start, st_o) and np.array_equal(order, order_o)
    assert rel_err(got.ax, ax) <= TOL and rel_err(got.ay, ay) <= TOL and rel_err(got.az, az) <= TOL
    for g_, w_ in zip(obs, (e, d_, dd)):
        assert abs(g_ - w_) <= TOL * abs(w_)
    # the all-pairs GPU kernel on the same state agrees too
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, s))
        obs2 = ctx.compute_forces()
        got2 = ctx.download(("ax", "ay", "az"))
        _, cut2 = ctx.pair_counts()
    assert cut2 == ncut
    assert rel_err(got2.ax, ax) <= TOL and np.all(np.abs(obs2 - obs) <= TOL * np.abs(obs))


def test_cells_multistep_and_determinism(lj, O):
    p_o, st = make_config(O, "cfg3")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ref = [a.copy() for a in st]
    for _ in range(5):
        ref_obs = O.verlet_step(p_o, ref, cells=True)
    outs = []
    for _ in range(2):
        with lj.Context(p, method=lj.METHOD_CELLS) as ctx:
            ctx.upload(_state(lj, st))
            obs = ctx.step(5)
            d = ctx.download()
            outs.append((obs.copy(), [a.copy() for a in d.arrays()]))
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)
    assert np.all(np.abs(outs[0][0] - np.array(ref_obs)) <= 1e-11 * np.abs(np.array(ref_obs)))
    assert rel_err(outs[0][1][0], ref[0]) <= 1e-11


def test_cells_edge_positions(lj, O):
    """Atoms exactly on cell faces / at x == L (wrapped tiny negatives) bin like the oracle."""
    g = np.load(os.path.join(GOLD, "cells_n864.npz"))
    L = float(g["L"])
    p_o = O.params_init(864, L, 1e-3, 2.5, 6)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = [np.ascontiguousarray(a).copy() for a in g["state300"]]
    edge = L / 4
    s[0][:8] = [0.0, L, edge, 2 * edge, 3 * edge, np.nextafter(edge, 0), np.nextafter(L, 0), np.nextafter(2 * edge, 9)]
    with lj.Context(p, method=lj.METHOD_CELLS) as ctx:
        ctx.upload(_state(lj, s))
        ctx.compute_forces()
        cell, start, order = ctx.download_cells()
    assert np.array_equal(cell, O.cell_assign(p_o, *s[:3]))


# ---------------------------------------------------------------------------------------------
# i-range sharding, two "ranks" emulated as two contexts on one GPU (the exchange is done by
# hand; the NCCL path itself is exercised by bench.py --gpus N and tests/test_multi_gpu_host.py)
# ---------------------------------------------------------------------------------------------
def test_sharded_contexts_match_single(lj, O):
    import importlib

    import torch
    mg = importlib.import_module(lj.__name__ + ".multi_gpu")
    g = np.load(os.path.join(GOLD, "cells_n864.npz"))
    n = 864
    p_o = O.params_init(n, float(g["L"]), 1e-3, 2.5, 6)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    nsteps, world = 6, 2
    dev = torch.device("cuda", 0)
    with lj.Context(p, method=lj.METHOD_CELLS) as ref:
        ref.upload(_state(lj, s))
        ref.compute_forces()
        ref_obs = ref.step(nsteps)
        ref_state = ref.download()
    ctxs = [lj.Context(p, method=lj.METHOD_CELLS) for _ in range(world)]
    try:
        views, bufs = [], []
        for r, c in enumerate(ctxs):
            c.upload(_state(lj, s))
            c.set_shard(*mg.shard_range(n, r, world))
            c.compute_forces()
            p0, p1, _, _, pobs = c.device_ptrs()
            views.append(([mg.torch_view(p0, 3 * n, dev), mg.torch_view(p1, 3 * n, dev)], mg.torch_view(pobs, 4, dev)))
            bufs.append(c.shard_begin())
        for _ in range(nsteps):
            torch.cuda.synchronize()
            b = bufs[0]
            assert all(x == b for x in bufs)
            # "all-gather": every context receives the owned ranges of the others
            for r in range(world):
                lo, hi = mg.shard_range(n, r, world)
                for q in range(world):
                    if q != r:
                        for comp in range(3):
                            views[q][0][b][comp * n + lo:comp * n + hi] = views[r][0][b][comp * n + lo:comp * n + hi]
            torch.cuda.synchronize()
            bufs = [c.shard_step() for c in ctxs]
        torch.cuda.synchronize()
        parts = [v[1].cpu().numpy().copy() for v in views]
        tails = lj.tail_corrections(p)
        tot = parts[0] + parts[1]
        tot[0] += tails[0]; tot[2] += tails[1]; tot[3] += tails[2]
        assert np.all(np.abs(tot - ref_obs) <= 1e-13 * np.abs(ref_obs))
        for r, c in enumerate(ctxs):
            c.shard_end()
            d = c.download()
            lo, hi = mg.shard_range(n, r, world)
            assert np.array_equal(d.rx, ref_state.rx)               # all positions, bitwise
            assert np.array_equal(d.vx[lo:hi], ref_state.vx[lo:hi])  # owned velocities / forces
            assert np.array_equal(d.az[lo:hi], ref_state.az[lo:hi])
    finally:
        for c in ctxs:
            c.close()


def test_nve_energy_drift_10k_steps(lj, O):
    """North star: energy drift over a 10^4-step NVE run statistically indistinguishable from the
    reference (the shipped configuration, N=500).  The potential is truncated, not shifted
    (lj_potential_energy.f90:132-140), so E_tot is a jump process (pairs crossing rc) on top of the
    O(dt^2) integrator fluctuation; the two trajectories decorrelate (chaos), hence statistics."""
    p_o, st = make_config(O, "cfg1")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    nsteps = 10000
    ref = [a.copy() for a in st]
    _, ref_ser = O.run_steps(p_o, nsteps, ref, want_series=True)
    with lj.Context(p, method=lj.METHOD_ALLPAIRS) as ctx:
        ctx.upload(_state(lj, st))
        ser = ctx.step_series(nsteps)
    e_gpu, e_ref = ser[:, 0] + ser[:, 1], ref_ser[:, 0] + ref_ser[:, 1]
    # identical start: the first 50 steps agree to round-off amplification
    assert np.all(np.abs(e_gpu[:50] - e_ref[:50]) <= 1e-9 * np.abs(e_ref[:50]))
    # drift = mean of the last 2000 steps minus mean of the first 2000 steps (after the lattice melts)
    def drift(e):
        return e[-2000:].mean() - e[1000:3000].mean()
    def block_std(e):  # fluctuation of 500-step block means: the noise level of `drift`
        b = e[1000:].reshape(-1, 500).mean(axis=1)
        return b.std()
    noise = max(block_std(e_ref), block_std(e_gpu), 1e-6)
    assert abs(drift(e_gpu) - drift(e_ref)) <= 6.0 * noise, (drift(e_gpu), drift(e_ref), noise)
    # same fluctuation level (within a factor of 2) and same mean energy (within the noise)
    r = e_gpu[1000:].std() / e_ref[1000:].std()
    assert 0.5 < r < 2.0, r
    assert abs(e_gpu[1000:].mean() - e_ref[1000:].mean()) <= 6.0 * noise
    # kinetic temperature statistics agree too
    t_gpu, t_ref = ser[2000:, 1].mean(), ref_ser[2000:, 1].mean()
    assert abs(t_gpu - t_ref) <= 0.02 * abs(t_ref)


def test_initial_config_program_sequence_on_gpu(lj, O):
    """md_initial_config_program.f90:73-116 through the host mirror + the GPU force routine: the
    shipped run's start (seed -12345, E_target -555) and 20 warm-up steps vs the oracle."""
    import importlib
    ic = importlib.import_module(lj.__name__ + ".initial_config")
    p_o = O.params_init(500, 10.0, 1e-4, 4.9, 5)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ref, e_ref, k_ref = O.initial_config(p_o, -12345, -555.0)
    st, epot, ekin = ic.initial_config(lj, p, seed=-12345, target_total_energy=-555.0)
    assert abs(epot - e_ref) <= TOL * abs(e_ref) and abs(ekin - k_ref) <= TOL * k_ref
    assert abs((epot + ekin) + 555.0) <= 1e-10
    assert np.array_equal(st.rx, ref[0]) and rel_err(st.vx, ref[3]) <= TOL and rel_err(st.vz, ref[5]) <= TOL
    st2, e2, k2 = ic.initial_config(lj, p, seed=-12345, target_total_energy=-555.0, warmup_steps=20)
    ref2 = [a.copy() for a in ref]
    obs = O.run_steps(p_o, 20, ref2)
    assert abs(e2 - obs[0]) <= 1e-10 * abs(obs[0]) and abs(k2 - obs[1]) <= 1e-10 * abs(obs[1])
    assert rel_err(st2.rx, ref2[0]) <= 1e-10


def test_replica_ensemble_end_to_end(lj, O):
    """Config 5 in miniature: 4 replicas (velocity seeds -(12345+r)), batched in one
    context, sampled every 100 steps after 100 warm-up steps; per-replica means -> thermodynamic
    coefficients -> ensemble mean / population std, against the same pipeline driven by the oracle."""
    import importlib
    ob = importlib.import_module(lj.__name__ + ".observables")
    p_o = O.params_init(500, 10.0, 1e-4, 4.9, 5)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    R, total, interval, warm = 4, 400, 100, 100
    states, coefs_ref = [], []
    for r in range(R):
        s_r, _, _ = O.initial_config(p_o, -(12345 + r), -555.0)
        states.append(lj.SimState.from_arrays([a.copy() for a in s_r]))
        ref = [a.copy() for a in s_r]
        acc = O.Means(500, p_o.volume)
        for step in range(interval, total + 1, interval):
            obs = O.run_steps(p_o, interval, ref)
            if step > warm:
                acc.add(obs[0], obs[1], -obs[2], obs[2], obs[3])
        m, _ = acc.get()
        mm = dict(zip(ob.QUANTITIES, m))
        coefs_ref.append(O.thermo_coefs(p_o, mm["K"], mm["P"], mm["Kinv"], mm["dU"], mm["ddU"], mm["dU_Kinv"], mm["dU2_Kinv"]))
    coefs_ref = np.array(coefs_ref)                      # [R][16]
    out = ob.run_ensemble(lj, p, states, total, interval, warm)
    assert out["n_replicas"] == R and out["n_samples"] == 3
    mean_ref, std_ref = coefs_ref.mean(axis=0), coefs_ref.std(axis=0)
    for k, name in enumerate(ob.COEFS):
        assert abs(out["mean"][name] - mean_ref[k]) <= 1e-7 * abs(mean_ref[k]), name
        assert abs(out["std"][name] - std_ref[k]) <= 1e-5 * abs(mean_ref[k]) + 1e-12, name
