"""CPU tests of the host-side consumers of the per-step reductions:
running means, T / P per sample and the thermodynamic coefficients, against the oracle's
restatement of stats/md_means.f90:192-366 and physics/thermodynamic_coefs.f90:87-205.  Bit-exact
(same operations in the same order)."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ob(lj):
    return importlib.import_module(lj.__name__ + ".observables")


def _series(seed, nsamp, n):
    rng = np.random.default_rng(seed)
    U = -3.2 * n + 5.0 * rng.standard_normal(nsamp)
    K = 2.1 * n + 5.0 * rng.standard_normal(nsamp)
    dU = 17.0 * n + 40.0 * rng.standard_normal(nsamp)
    ddU = -90.0 * n + 300.0 * rng.standard_normal(nsamp)
    return np.stack([U, K, dU, ddU], axis=1)


def test_means_and_coefficients_match_oracle_bitwise(ob, O):
    n, L = 500, 10.0
    p = O.params_init(n, L, 1e-4, 4.9, 5)
    R, nsamp = 3, 200
    series = [_series(10 + r, nsamp, n) for r in range(R)]
    acc = ob.MeansAccumulator(n, p.volume, R)
    orc = [O.Means(n, p.volume) for _ in range(R)]
    for s in range(nsamp):
        T, P = acc.add_sample(np.stack([series[r][s] for r in range(R)]))
        for r in range(R):
            e, k, d, dd = series[r][s]
            t_o, p_o = orc[r].add(e, k, -d, d, dd)   # virial = -d_epot, md_simulation_program.f90:366
            assert T[r] == t_o and P[r] == p_o
    mean, std = acc.get()
    for r in range(R):
        m_o, s_o = orc[r].get()
        assert np.array_equal(mean[:, r], m_o) and np.array_equal(std[:, r], s_o)
    coefs = ob.coefs_from_means(n, p.volume, mean)
    for r in range(R):
        m = dict(zip(ob.QUANTITIES, mean[:, r]))
        want = O.thermo_coefs(p, m["K"], m["P"], m["Kinv"], m["dU"], m["ddU"], m["dU_Kinv"], m["dU2_Kinv"])
        got = np.array([coefs[k][r] for k in ob.COEFS])
        assert np.array_equal(got, want)
    # the two temperature definitions coexist: 2<K>/(3N) vs 2<K>/(3N-3)
    assert np.allclose(coefs["temperature"] * (3 * n - 3), mean[3] * 3 * n, rtol=1e-14)
    # alpha_E1 and alpha_E2 are two estimators of one quantity and K_S_inv is 1/K_S
    assert np.array_equal(coefs["K_S_inv"], 1.0 / coefs["K_S"])


def test_means_guards_mirror_reference_stops(ob):
    with pytest.raises(ValueError):
        ob.MeansAccumulator(0, 1000.0)
    with pytest.raises(ValueError):
        ob.MeansAccumulator(500, 0.0)
    acc = ob.MeansAccumulator(500, 1000.0)
    with pytest.raises(ValueError):
        acc.get()
    with pytest.raises(ValueError):
        acc.add_sample(np.array([[-1600.0, 0.0, 8000.0, -46000.0]]))   # ekin must be > 0
    with pytest.raises(ValueError):
        ob.thermodynamic_coefs_compute(1, 1000.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)  # 3N-3 = 0
