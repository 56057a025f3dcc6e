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


def test_weak_constant_virial_known_answers(ob, O):
    # Constant kinetic energy K and a constant, non-fluctuating dU = c, ddU = 0: every fluctuation term of
    # thermodynamic_coefs.f90:118-203 cancels and the microcanonical formulas reduce to closed forms with
    # f = 3N - 3:  Ca_v = f/2, gamma = 2N/f, K_S = N T (1 + gamma)/V - 2c/(9V), K_T = N T/V - 2c/(9V),
    # and the two estimators alpha_E1, alpha_E2 coincide at -3f/(2c).  (c = 0, the ideal gas, makes both
    # alpha_E denominators vanish and the reference stops there -- the guard is mirrored.)
    n, V, K, c = 500, 1000.0, 750.0, 12.5
    f = 3.0 * n - 3.0
    T = 2.0 * K / f
    P = (n / V) * T - c / (3.0 * V)          # md_means.f90:227 with W = -dU
    args = (K, P, 1.0 / K, c, 0.0, c / K, c * c / K)
    got = ob.thermodynamic_coefs_compute(n, V, *args)
    p = O.params_init(n, 10.0, 1e-4, 4.9, 5)
    assert np.array_equal(np.array([float(got[k]) for k in ob.COEFS]), O.thermo_coefs(p, *args))
    gamma = 2.0 * n / f
    k_s = n * T * (1.0 + gamma) / V - 2.0 * c / (9.0 * V)
    k_t = n * T / V - 2.0 * c / (9.0 * V)
    want = {"temperature": T, "pressure": P, "Ca_v": f / 2.0, "Ce_v": f / (2.0 * n), "gamma": gamma,
            "K_S": k_s, "K_T": k_t, "Ca_p": (f / 2.0) * k_s / k_t, "alpha_P": n / (V * k_t),
            "alpha_S": -1.0 / (gamma * T), "alpha_E1": -3.0 * f / (2.0 * c), "alpha_E2": -3.0 * f / (2.0 * c)}
    for k, w in want.items():
        assert abs(float(got[k]) - w) < 1e-8 * abs(w), k   # Ca_v = 1/(1 - (1 - 2/f)) loses ~4 digits
    with pytest.raises(ValueError):
        ob.thermodynamic_coefs_compute(n, V, K, (n / V) * T, 1.0 / K, 0.0, 0.0, 0.0, 0.0)
