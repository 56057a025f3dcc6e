"""CPU tests of the host-side initial-configuration mirror and the binary hand-off files
 against the oracle's restatement of the same reference routines
(md_initial_config_program.f90:132-266, base/random_numbers.f90:48-116).  Bit-exact."""
import importlib
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ic(lj):
    return importlib.import_module(lj.__name__ + ".initial_config")


def test_rng_stream_matches_oracle_and_anchors(ic, O):
    g, o = ic.RandomUniform(-12345), O.Rng(-12345)
    mine = g.draw(3000)
    ref = np.array([o.uniform() for _ in range(3000)])
    assert np.array_equal(mine, ref)
    # first six draws and the sum of the first 1500 (exact multiples of 2.5e-7)
    assert np.allclose(mine[:6], [0.6283575, 0.81823825, 0.169642, 0.19663425, 0.9045145, 0.60583975], rtol=0, atol=1e-15)
    assert abs(mine[:1500].sum() - 741.57097625) < 1e-9
    # a positive seed after the first call does not re-initialise (random_numbers.f90:72)
    g2 = ic.RandomUniform(-7)
    a = g2()
    assert g2.seed == 1 and g2() != a


@pytest.mark.parametrize("k,L", [(5, 10.0), (6, 10.259855680060184), (3, 7.3)])
def test_fcc_lattice_bit_exact(ic, O, k, L):
    p = O.params_init(4 * k ** 3, L, 1e-3, 0.3 * L, k)
    rx, ry, rz = ic.build_fcc_lattice(k, L)
    ox, oy, oz = O.fcc_lattice(p)
    assert np.array_equal(rx, ox) and np.array_equal(ry, oy) and np.array_equal(rz, oz)


def test_full_initial_configuration_bit_exact(ic, O):
    """Lattice, velocities(seed -12345), COM removal, rescale to E = -555 -- the shipped run -- with the
    oracle supplying the (CPU) force evaluation the rescale needs."""
    n, L, k = 500, 10.0, 5
    p = O.params_init(n, L, 1e-4, 4.9, k)
    ref, e_ref, k_ref = O.initial_config(p, -12345, -555.0)
    rx, ry, rz = ic.build_fcc_lattice(k, L)
    vx, vy, vz = ic.assign_random_velocities(n, ic.RandomUniform(-12345))
    ic.remove_center_of_mass_velocity(vx, vy, vz)
    _, _, _, epot, _, _ = O.compute_lj(p, rx, ry, rz)
    scale = ic.rescale_velocities_to_target_energy(vx, vy, vz, -555.0, epot)
    assert abs(scale - 4.1061483183471585) < 1e-9
    for got, want in zip((rx, ry, rz, vx, vy, vz), ref[:6]):
        assert np.array_equal(got, want)
    assert ic.kinetic_energy(vx, vy, vz) == k_ref
    with pytest.raises(ValueError):
        ic.rescale_velocities_to_target_energy(vx.copy(), vy.copy(), vz.copy(), -1e9, epot)


def test_rv_init_and_rva_files(ic, tmp_path):
    class S:
        pass
    rng = np.random.default_rng(3)
    s = S()
    n = 37
    for name in ("rx", "ry", "rz", "vx", "vy", "vz", "ax", "ay", "az"):
        setattr(s, name, rng.random(n))
    f = os.path.join(tmp_path, "rv_init.dat")
    ic.write_rv_init(f, s)
    raw = open(f, "rb").read()
    # two gfortran sequential records of 3N real64 with 4-byte little-endian length markers
    assert len(raw) == 2 * (4 + 24 * n + 4)
    assert struct.unpack("<i", raw[:4])[0] == 24 * n == struct.unpack("<i", raw[4 + 24 * n:8 + 24 * n])[0]
    back = ic.read_rv_init(f, n)
    for got, name in zip(back, ("rx", "ry", "rz", "vx", "vy", "vz")):
        assert np.array_equal(got, getattr(s, name))
    with pytest.raises(ValueError):
        ic.read_rv_init(f, n + 1)
    g = os.path.join(tmp_path, "rva.dat")
    ru = [rng.random(n) for _ in range(3)]
    with ic.RvaWriter(g, n, 10.0, 1e-4, 100, 2) as w:
        w.write_sample(s, ru)
        w.write_sample(s, ru)
    raw = open(g, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 28               # header record: int32, 2 x real64, 2 x int32
    assert len(raw) == (4 + 28 + 4) + 2 * 4 * (4 + 24 * n + 4)
    hdr, samples = ic.read_rva(g)
    assert hdr == {"n": n, "box_length": 10.0, "dt": 1e-4, "output_interval": 100, "n_snapshots": 2}
    assert len(samples) == 2 and np.array_equal(samples[1][1][2], ru[2]) and np.array_equal(samples[0][3][0], s.ax)


def test_rng_block_draw_is_bit_identical_to_scalar_calls(ic):
    # draw() evaluates the lagged recurrence 24 values at a time; values AND the state left behind
    # must equal `count` scalar calls (base/random_numbers.f90:93-116)
    for seed in (-12345, -7):
        a, b = ic.RandomUniform(seed), ic.RandomUniform(seed)
        want = np.array([a() for _ in range(3000)])
        got = np.concatenate([b.draw(k) for k in (10, 300, 1, 1500, 255, 256, 678)])
        assert np.array_equal(want, got)
        assert [a() for _ in range(120)] == [b() for _ in range(120)]
