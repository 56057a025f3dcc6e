"""GPU parity tests of the neighbour-list path (LJMD_METHOD_NLIST): counting-sort cells + Verlet
list with skin + one fused kernel per step, state kept in cell-sorted order on the device.

No reference counterpart (the reference is all-pairs only, lj_potential_energy.f90:109-183):
forces / U / virial are checked against the all-pairs (or cell) oracle to 1e-12 relative per
step, the integer cell assignment and the neighbour sets bit-exactly against the oracle's
definitions (orc_cell_assign_range / orc_cell_sort_range / orc_neighbor_list)."""
import os

import numpy as np
import pytest

from conftest import assert_forces_close, force_scale, make_config, melt, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _state(lj, arrs):
    return lj.SimState.from_arrays([a.copy() for a in arrs])


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            os.environ[k] = str(v)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _small(O):
    g = np.load(os.path.join(GOLD, "cells_n864.npz"))
    p_o = O.params_init(864, float(g["L"]), 1e-3, 2.5, 6)
    s = [np.ascontiguousarray(a).copy() for a in g["state300"]]
    return p_o, s


def _check_integers(lj, O, ctx, p_o, s):
    info = ctx.nlist_info()
    rng = p_o.rc + info["skin"]
    assert info["cells_per_edge"] == O.cells_per_edge_range(p_o, rng)
    cell, start, order = ctx.download_cells()
    cell_o = O.cell_assign_range(p_o, rng, *s[:3])
    start_o, order_o = O.cell_sort_range(p_o, rng, cell_o)
    assert np.array_equal(cell, cell_o) and np.array_equal(start, start_o) and np.array_equal(order, order_o)
    counts, nbr = ctx.download_nlist()
    counts_o, nbr_o = O.neighbor_list(p_o, rng, *s[:3], cap=info["capacity"])
    assert np.array_equal(counts, counts_o)
    assert np.array_equal(nbr, nbr_o)
    assert info["entries"] == int(counts_o.sum())
    return info


def test_nlist_forces_and_integers_n864(lj, O):
    p_o, s = _small(O)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ax, ay, az, e, d_, dd = O.compute_lj(p_o, *s[:3])
    with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
        assert ctx.method == lj.METHOD_NLIST
        ctx.upload(_state(lj, s))
        obs = ctx.compute_forces()
        got = ctx.download()
        info = _check_integers(lj, O, ctx, p_o, s)
        cand, cut = ctx.pair_counts()
    assert info["rebuilds"] == 1
    assert cut == O.count_pairs_in_cutoff(p_o, *s[:3]) and 2 * cand == info["entries"]
    assert rel_err(got.ax, ax) <= TOL and rel_err(got.ay, ay) <= TOL and rel_err(got.az, az) <= TOL
    for g_, w_ in zip(obs, (e, d_, dd)):
        assert abs(g_ - w_) <= TOL * abs(w_)
    # the state comes back in ORIGINAL order, untouched
    for k, name in enumerate(("rx", "ry", "rz", "vx", "vy", "vz")):
        assert np.array_equal(getattr(got, name), s[k]), name


def test_nlist_golden_n864(lj):
    """Cells (edge >= rc+skin) and neighbour sets against the committed fixture -- no oracle at test time."""
    g, gn = np.load(os.path.join(GOLD, "cells_n864.npz")), np.load(os.path.join(GOLD, "nlist_n864.npz"))
    p = lj.init_params(864, float(g["L"]), 1e-3, 2.5, 6)
    s = [np.ascontiguousarray(a) for a in g["state300"]]
    with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
        ctx.upload(_state(lj, s))
        obs = ctx.compute_forces()
        got = ctx.download(("ax", "ay", "az"))
        cell, start, order = ctx.download_cells()
        counts, nbr = ctx.download_nlist()
    assert np.array_equal(cell, gn["cell"]) and np.array_equal(start, gn["start"]) and np.array_equal(order, gn["order"])
    assert np.array_equal(counts, gn["counts"])
    assert np.array_equal(nbr[:, :144], gn["nbr"].astype(np.int32)[:, :nbr.shape[1]])
    for k, name in enumerate(("ax", "ay", "az")):
        assert rel_err(getattr(got, name), g["acc"][k]) <= TOL
    for g_, w_ in zip(obs, g["obs"]):
        assert abs(g_ - w_) <= TOL * abs(w_)


def test_nlist_vs_oracle_32k(lj, O):
    p_o, st = make_config(O, "cfg3")
    s = melt(O, p_o, st, 20, cells=True)
    ax, ay, az, e, d_, dd, _, ncut = O.compute_lj_cells(p_o, *s[:3])
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with lj.Context(p) as ctx:  # AUTO -> neighbour list at N=32000
        assert ctx.method == lj.METHOD_NLIST and ctx.cells_per_edge == 12
        ctx.upload(_state(lj, s))
        obs = ctx.compute_forces()
        got = ctx.download(("ax", "ay", "az"))
        _check_integers(lj, O, ctx, p_o, s)
        _, cut = ctx.pair_counts()
    assert cut == ncut
    assert rel_err(got.ax, ax) <= TOL and rel_err(got.ay, ay) <= TOL and rel_err(got.az, az) <= TOL
    assert_forces_close((got.ax, got.ay, got.az), (ax, ay, az), force_scale(p_o.box_length, p_o.rc, *s[:3]), TOL)
    for g_, w_ in zip(obs, (e, d_, dd)):
        assert abs(g_ - w_) <= TOL * abs(w_)


def test_nlist_single_step_parity_32k(lj, O):
    """One verlet_step through the drop-in C-ABI call (AUTO -> neighbour list) from a melted state."""
    p_o, st = make_config(O, "cfg3")
    s0 = melt(O, p_o, st, 10, cells=True)
    ref = [a.copy() for a in s0]
    ref_obs = O.verlet_step(p_o, ref, cells=True)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = _state(lj, s0)
    obs = lj.verlet_step(p, s)
    for got, want in zip(obs, ref_obs):
        assert abs(got - want) <= TOL * abs(want)
    for name, want in zip(lj.SimState.NAMES, ref):
        assert rel_err(getattr(s, name), want) <= TOL, name
    assert np.array_equal(s.rx, ref[0]) and np.array_equal(s.ry, ref[1]) and np.array_equal(s.rz, ref[2])
    assert_forces_close((s.ax, s.ay, s.az), ref[6:9], force_scale(p_o.box_length, p_o.rc, *ref[:3]), TOL)
    lj.compat_release()


def test_nlist_multistep_rebuilds_graph_and_determinism(lj, O):
    """70 steps with a small skin (several list rebuilds): CUDA-graph path (IF-node rebuilds) ==
    eager path bitwise, run-to-run bitwise, and within 1e-9 of the oracle trajectory."""
    p_o, st = make_config(O, "cfg3")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    nsteps = 70
    ref = [a.copy() for a in st]
    for _ in range(nsteps):
        ref_obs = O.verlet_step(p_o, ref, cells=True)

    def run(**env):
        with _env(LJMD_NL_SKIN=0.12, **env):
            with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
                ctx.upload(_state(lj, st))
                ctx.compute_forces()
                obs = ctx.step(nsteps)
                mom = float(np.abs(ctx.nlist_momentum()).max())
                d = ctx.download()
                info = ctx.nlist_info()
        return obs.copy(), [a.copy() for a in d.arrays()], info, mom

    a = run()
    b = run()
    c = run(LJMD_NL_NOGRAPH=1)
    assert a[2]["rebuilds"] >= 3, a[2]
    assert a[2]["rebuilds"] == c[2]["rebuilds"]
    assert a[3] < 1e-8  # total momentum stays zero (Newton 3 + symmetric ordered pairs): built-in self-check
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[0], c[0])
    for x, y, z in zip(a[1], b[1], c[1]):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    assert np.all(np.abs(a[0] - np.array(ref_obs)) <= 1e-9 * np.abs(np.array(ref_obs)))
    assert rel_err(a[1][0], ref[0]) <= 1e-9 and rel_err(a[1][3], ref[3]) <= 1e-9


def test_nlist_skin_independence(lj, O):
    """The forces do not depend on the skin (the list always contains every in-cutoff pair)."""
    p_o, s = _small(O)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    outs = []
    for skin in (0.0, 0.3, 0.6):
        with _env(LJMD_NL_SKIN=skin):
            with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
                ctx.upload(_state(lj, s))
                ctx.compute_forces()
                obs = ctx.step(12)
                outs.append((obs.copy(), ctx.download().ax.copy(), ctx.nlist_info()["rebuilds"]))
    assert outs[0][2] > outs[2][2]  # skin 0 rebuilds every step
    for o in outs[1:]:
        assert np.all(np.abs(o[0] - outs[0][0]) <= 1e-11 * np.abs(outs[0][0]))
        assert rel_err(o[1], outs[0][1]) <= 1e-10


def test_nlist_unwrapped_tracking(lj, O):
    p_o, s = _small(O)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    nsteps = 25
    a0 = O.compute_lj(p_o, *s[:3])
    ref = [x.copy() for x in s[:6]] + [a0[0].copy(), a0[1].copy(), a0[2].copy()]
    ru = [x.copy() for x in s[:3]]
    O.run_steps(p_o, nsteps, ref, ru=ru)
    with _env(LJMD_NL_SKIN=0.1):
        with lj.Context(p, method=lj.METHOD_NLIST, flags=lj.FLAG_TRACK_UNWRAPPED) as ctx:
            ctx.upload(_state(lj, s))
            ctx.compute_forces()
            ctx.step(nsteps)
            got_ru = ctx.download_unwrapped()
            got = ctx.download()
            assert ctx.nlist_info()["rebuilds"] >= 2
    for k in range(3):
        assert rel_err(got_ru[k], ru[k]) <= 1e-11
    assert rel_err(got.rx, ref[0]) <= 1e-11


def test_nlist_unwrapped_inputs(lj, O):
    """The force routine is specified for unwrapped inputs too (geometry_pbc.f90:18-19)."""
    p_o, s = _small(O)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    ax, ay, az, e, d_, dd = O.compute_lj(p_o, *s[:3])
    rng = np.random.default_rng(11)
    su = [a.copy() for a in s]
    for k in range(3):
        su[k] = su[k] + p_o.box_length * rng.integers(-2, 3, size=p_o.n)
    with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
        ctx.upload(_state(lj, su))
        obs = ctx.compute_forces()
        got = ctx.download(("ax", "ay", "az"))
        obs5 = ctx.step(5)
    assert rel_err(got.ax, ax) <= 1e-11 and rel_err(got.az, az) <= 1e-11
    assert abs(obs[0] - e) <= 1e-11 * abs(e)
    ref = [a.copy() for a in su[:6]] + [ax.copy(), ay.copy(), az.copy()]
    ref_obs = O.run_steps(p_o, 5, ref)
    assert np.all(np.abs(obs5 - np.array(ref_obs)) <= 1e-10 * np.abs(np.array(ref_obs)))


def test_nlist_partial_upload_and_restart(lj, O):
    """Uploading only velocities into a sorted resident state keeps positions / order intact."""
    p_o, s = _small(O)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
        ctx.upload(_state(lj, s))
        ctx.compute_forces()
        ctx.step(7)
        mid = ctx.download()
        newv = [0.5 * mid.vx, 0.5 * mid.vy, 0.5 * mid.vz]
        ctx.upload(vx=newv[0], vy=newv[1], vz=newv[2])
        after = ctx.download()
        assert np.array_equal(after.rx, mid.rx) and np.array_equal(after.ax, mid.ax)
        assert np.array_equal(after.vx, newv[0])
        obs = ctx.step(3)
    ref = [mid.rx.copy(), mid.ry.copy(), mid.rz.copy()] + [v.copy() for v in newv] + [mid.ax.copy(), mid.ay.copy(), mid.az.copy()]
    ref_obs = O.run_steps(p_o, 3, ref)
    assert np.all(np.abs(obs - np.array(ref_obs)) <= 1e-11 * np.abs(np.array(ref_obs)))


def test_nlist_capacity_overflow_is_an_error(lj, O):
    p_o, s = _small(O)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    with _env(LJMD_NL_CAP=16):
        with lj.Context(p, method=lj.METHOD_NLIST) as ctx:
            ctx.upload(_state(lj, s))
            with pytest.raises(lj.LjmdError) as ei:
                ctx.compute_forces()
            assert ei.value.code == 10 and "capacity" in ei.value.message


def test_nlist_needs_three_cells(lj):
    p = lj.init_params(500, 10.0, 1e-4, 4.9, 5)
    with pytest.raises(lj.LjmdError) as ei:
        lj.Context(p, method=lj.METHOD_NLIST)
    assert ei.value.code == 10


# a spin-wait kernel that never gets its flag would otherwise hang the GPU box: fail hard instead
@pytest.mark.timeout(300, method="thread")
def test_nlist_slab_two_ranks_emulated(lj, O):
    """Slab sharding with two 'ranks' emulated as two contexts on one GPU (peer pointers are the
    other context's device pointers; the all-gather is done by hand with device copies).  One step
    per launch, alternating, so no kernel ever has to wait for a concurrently running peer.
    The trajectory must be BITWISE equal to the single-context neighbour-list run."""
    import importlib

    import torch
    mg = importlib.import_module(lj.__name__ + ".multi_gpu")
    p_o, st = make_config(O, "cfg3")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    nsteps, world = 60, 2
    dev = torch.device("cuda", 0)
    with _env(LJMD_NL_SKIN=0.12):
        with lj.Context(p, method=lj.METHOD_NLIST) as ref:
            ref.upload(_state(lj, st))
            ref.compute_forces()
            ref_obs = ref.step(nsteps)
            ref_state = ref.download()
            ref_rebuilds = ref.nlist_info()["rebuilds"]
        ctxs = [lj.Context(p, method=lj.METHOD_NLIST) for _ in range(world)]
        try:
            for c in ctxs:
                c.upload(_state(lj, st))
                c.compute_forces()
            raw = [c.slab_export()[1] for c in ctxs]
            for r, c in enumerate(ctxs):
                c.slab_attach(r, world, raw_ptrs=raw)
                c.slab_drift()
                c.slab_rebuild()
            narr = ctxs[0].slab_narr()

            def gather():
                bounds = ctxs[0].slab_bounds()
                assert all(np.array_equal(c.slab_bounds(), bounds) for c in ctxs)
                cmax = int(np.max(np.diff(bounds)))
                recv = torch.zeros(world * narr * cmax, dtype=torch.float64, device=dev)
                for r, c in enumerate(ctxs):
                    c.slab_pack(recv[r * narr * cmax:].data_ptr(), cmax)
                torch.cuda.synchronize()
                for c in ctxs:
                    c.slab_unpack(recv.data_ptr(), cmax)
                torch.cuda.synchronize()

            done_total, rebuilds, obs = 0, 1, None
            while done_total < nsteps:
                res = []
                for c in ctxs:
                    res.append(c.slab_run(1))
                    c.slab_publish()  # emulated ranks: the peer's next launch cannot do it in time
                    torch.cuda.synchronize()
                assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
                if res[0][0] == 1:
                    done_total += 1
                    obs = res[0][2] + res[1][2]
                    # the ranks' momentum shares add up to zero at every step
                    assert np.abs(ctxs[0].nlist_momentum() + ctxs[1].nlist_momentum()).max() < 1e-8
                if res[0][1]:
                    gather()
                    for c in ctxs:
                        c.slab_rebuild()
                    rebuilds += 1
            gather()
            tails = lj.tail_corrections(p)
            obs[0] += tails[0]; obs[2] += tails[1]; obs[3] += tails[2]
            assert rebuilds >= 2  # the initial build + at least one rebuild after a halt
            assert np.all(np.abs(obs - ref_obs) <= 1e-13 * np.abs(ref_obs))
            for c in ctxs:
                d = c.download()
                for x, y in zip(d.arrays(), ref_state.arrays()):
                    assert np.array_equal(x, y)
        finally:
            for c in ctxs:
                c.close()


@pytest.mark.timeout(120, method="thread")
def test_nlist_slab_flag_wait_is_bounded_and_poison_halts(lj, O):
    """A peer that never publishes its per-step flag must not leave the GPU spinning: the wait ends after
    LJMD_SLAB_SPIN_TIMEOUT_S with LJMD_ERR_CUDA and the state frozen at the last completed step; a peer that
    publishes the poison flag makes the waiting rank halt at once (no error, halted = 1)."""
    p_o, st = make_config(O, "cfg3")
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    import time
    with _env(LJMD_SLAB_SPIN_TIMEOUT_S=1.5):
        ctxs = [lj.Context(p, method=lj.METHOD_NLIST) for _ in range(2)]
        try:
            for c in ctxs:
                c.upload(_state(lj, st))
                c.compute_forces()
            raw = [c.slab_export()[1] for c in ctxs]
            for r, c in enumerate(ctxs):
                c.slab_attach(r, 2, raw_ptrs=raw)
                c.slab_drift()
                c.slab_rebuild()
            # step 1 needs no flag; step 2 waits for rank 1's flag of step 1, which never comes
            t0 = time.perf_counter()
            with pytest.raises(lj.LjmdError) as ei:
                ctxs[0].slab_run(2)
            waited = time.perf_counter() - t0
            assert ei.value.code == 100 and "flag" in ei.value.message
            assert 1.0 < waited < 30.0
        finally:
            for c in ctxs:
                c.close()
        ctxs = [lj.Context(p, method=lj.METHOD_NLIST) for _ in range(2)]
        try:
            for c in ctxs:
                c.upload(_state(lj, st))
                c.compute_forces()
            raw = [c.slab_export()[1] for c in ctxs]
            for r, c in enumerate(ctxs):
                c.slab_attach(r, 2, raw_ptrs=raw)
                c.slab_drift()
                c.slab_rebuild()
            ctxs[1].slab_poison()           # rank 1 "failed"
            done, halted, _ = ctxs[0].slab_run(3)
            assert done == 1 and halted     # step 1 ran (it waits for nothing), step 2 saw the poison and halted
        finally:
            for c in ctxs:
                c.close()


def test_drop_in_call_keeps_the_list_between_calls(lj, O):
    """Consecutive ljmd_verlet_step calls (host arrays in / out, AUTO -> neighbour list) reuse the cell-sorted order
    and the list of the previous call; the skin test of the drift decides when to rebuild.  Every call is one
    reference verlet_step from the state it was given: checked at 1e-12 on the first calls, late in a list's life
    and right after a rebuild, and for a caller that suddenly passes a different configuration."""
    p_o, st = make_config(O, "cfg3")
    s0 = melt(O, p_o, st, 10, cells=True)
    p = lj.SimParams.from_buffer_copy(bytes(p_o))
    s = _state(lj, s0)
    lj.compat_release()
    check_at = {0, 1, 2, 30, 44, 45, 46, 47, 59}
    for call in range(60):
        before = [a.copy() for a in s.arrays()] if call in check_at else None
        obs = lj.verlet_step(p, s)
        if before is not None:
            ref_obs = O.verlet_step(p_o, before, cells=True)
            for got, want in zip(obs, ref_obs):
                assert abs(got - want) <= TOL * abs(want), call
            for k in range(3):
                assert np.array_equal(s.arrays()[k], before[k]), call
                assert rel_err(s.arrays()[3 + k], before[3 + k]) <= TOL, call
                assert rel_err(s.arrays()[6 + k], before[6 + k]) <= TOL, call
    # another configuration with the same n: the kept list is useless, the skin test must notice
    other = melt(O, p_o, st, 3, cells=True)
    perm = np.random.default_rng(0).permutation(p_o.n)
    other = [np.ascontiguousarray(a[perm]) for a in other]
    s2 = _state(lj, other)
    ref = [a.copy() for a in other]
    ref_obs = O.verlet_step(p_o, ref, cells=True)
    obs = lj.verlet_step(p, s2)
    for got, want in zip(obs, ref_obs):
        assert abs(got - want) <= TOL * abs(want)
    for k in range(3):
        assert np.array_equal(s2.arrays()[k], ref[k])
        assert rel_err(s2.arrays()[6 + k], ref[6 + k]) <= TOL
    lj.compat_release()
