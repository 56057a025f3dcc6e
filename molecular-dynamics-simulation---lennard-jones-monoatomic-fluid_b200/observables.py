"""Host-side consumers of the hot path's four per-step scalars:
running means, instantaneous T and P, the thermodynamic coefficients, and the replica-ensemble
driver of config 5.  O(1) work per sample on the host, vectorised over replicas with numpy in the
reference's operation order (every element sees exactly the scalar arithmetic of the Fortran).

Reference (file:line relative to /root/reference/scripts/):
  stats/md_means.f90:192-275   md_means_add_sample: T = 2K/(3N) (:221), P = rho*T + W/(3V) (:227),
                               22 running sums
  stats/md_means.f90:284-366   md_means_get; stats/stats_math.f90:56-61 std from moments
  physics/thermodynamic_coefs.f90:87-205  thermodynamic_coefs_compute (T = 2<K>/(3N-3), :118)
  md_simulation_program.f90:361-372       sampling cadence, virial = -d_epot (:366)
  run_many_md_simuations/simulation_results.f90:330-350,391-398  ensemble mean / population std (intent;
                               that program does not build as shipped)
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

QUANTITIES = ("U", "K", "E", "T", "P", "Kinv", "dU", "ddU", "dU_Kinv", "dU2_Kinv", "ddU_Kinv")
COEFS = ("temperature", "pressure", "Ca_v", "Ce_v", "gamma", "K_S_aux", "K_S", "K_S_inv", "K_T", "K_T_inv",
         "Ca_p", "Ce_p", "alpha_E1", "alpha_E2", "alpha_S", "alpha_P")


class MeansAccumulator:
    """type(md_means_accum) for `nreplicas` independent systems at once (virial always present)."""

    def __init__(self, n: int, volume: float, nreplicas: int = 1):
        if n <= 0:
            raise ValueError("md_means_add_sample(): accumulator not initialized (n <= 0).")
        if not volume > 0.0:
            raise ValueError("md_means_add_sample(): invalid volume (accumulator not initialized?).")
        self.n, self.volume, self.nreplicas = int(n), float(volume), int(nreplicas)
        self.n_samples = 0
        self.sum = np.zeros((11, nreplicas))
        self.sum_sq = np.zeros((11, nreplicas))

    def add_sample(self, obs4: np.ndarray):
        """obs4[..., 4] = (epot, ekin, d_epot, dd_epot) as the hot path returns them; the virial is
        W = -d_epot (md_simulation_program.f90:366).  Returns (T_inst, P_inst) per replica."""
        o = np.asarray(obs4, dtype=np.float64).reshape(self.nreplicas, 4)
        U, K, dU, ddU = o[:, 0], o[:, 1], o[:, 2], o[:, 3]
        virial = -dU
        npd = float(self.n)
        rho = npd / self.volume
        E = U + K
        T = 2.0 * K / (3.0 * npd)
        P = rho * T + virial / (3.0 * self.volume)
        if np.any(K <= 0.0):
            raise ValueError("md_means_add_sample(): ekin must be > 0 to accumulate 1/ekin terms.")
        Kinv = 1.0 / K
        x = (U, K, E, T, P, Kinv, dU, ddU, dU * Kinv, (dU * dU) * Kinv, ddU * Kinv)
        self.n_samples += 1
        for q, v in enumerate(x):
            self.sum[q] = self.sum[q] + v
            self.sum_sq[q] = self.sum_sq[q] + v * v
        return T, P

    def get(self):
        """(mean[11][R], std[11][R]) in the order of QUANTITIES (md_means_get)."""
        if self.n_samples <= 0:
            raise ValueError("md_means_get(): no samples accumulated.")
        inv_ns = 1.0 / float(self.n_samples)
        mean = self.sum * inv_ns
        m2 = self.sum_sq * inv_ns
        std = np.sqrt(np.maximum(0.0, m2 - mean * mean))
        return mean, std


def thermodynamic_coefs_compute(n: int, volume: float, ekin_mean, press_mean, ekinInv_mean, d_epot_mean,
                                dd_epot_mean, d_epot_ekinInv_mean, d_epot2_ekinInv_mean) -> Dict[str, np.ndarray]:
    """physics/thermodynamic_coefs.f90:87-205, element-wise over arrays of time averages."""
    a = [np.asarray(v, dtype=np.float64) for v in (ekin_mean, press_mean, ekinInv_mean, d_epot_mean, dd_epot_mean,
                                                   d_epot_ekinInv_mean, d_epot2_ekinInv_mean)]
    K, P, Kinv, d, dd, dKinv, d2Kinv = a
    npd = float(n)
    f = 3.0 * npd - 3.0
    if f <= 0.0:
        raise ValueError("thermodynamic_compute(): degrees_of_freedom <= 0 (check N).")
    aux1, aux2 = 1.0 - 2.0 / f, f / 2.0 - 1.0
    V = float(volume)

    def guard(x, msg):
        if np.any(np.abs(x) < 1.0e-14):
            raise ValueError(msg)

    T = 2.0 * K / f
    denom = 1.0 - aux1 * K * Kinv
    guard(denom, "thermodynamic_compute(): Ca_v denominator ~ 0 (numerical instability).")
    Ca_v = 1.0 / denom
    Ce_v = Ca_v / npd
    guard(Ce_v, "thermodynamic_compute(): Ce_v ~ 0 (check inputs).")
    gamma = 1.0 / Ce_v + (aux2 / 3.0) * (d * Kinv - dKinv)
    K_S_aux = ((npd * T * (1.0 + 2.0 * gamma - 1.0 / Ce_v)) / V) + (dd - 2.0 * d) / (9.0 * V)
    K_S = K_S_aux - (aux2 * (d2Kinv - 2.0 * d * dKinv + (d * d) * Kinv)) / (9.0 * V * V)
    guard(K_S, "thermodynamic_compute(): K_S ~ 0 (cannot invert).")
    K_S_inv = 1.0 / K_S
    K_T = K_S - (T * Ca_v * (gamma * gamma)) / V
    guard(K_T, "thermodynamic_compute(): K_T ~ 0 (cannot invert / compute Cp, alpha_P).")
    K_T_inv = 1.0 / K_T
    Ca_p = Ca_v * (K_S / K_T)
    Ce_p = Ca_p / npd
    denom = (P * V / Ca_v) - (gamma * T)
    guard(denom, "thermodynamic_compute(): alpha_E1 denominator ~ 0.")
    alpha_E1 = 1.0 / denom
    denom = (1.0 / 3.0) * (aux1 * K * dKinv - d)
    guard(denom, "thermodynamic_compute(): alpha_E2 denominator ~ 0.")
    alpha_E2 = 1.0 / denom
    denom = gamma * T
    guard(denom, "thermodynamic_compute(): gamma*T ~ 0 (alpha_S undefined).")
    alpha_S = -1.0 / denom
    alpha_P = (Ca_v * gamma) / V * K_T_inv
    vals = (T, P, Ca_v, Ce_v, gamma, K_S_aux, K_S, K_S_inv, K_T, K_T_inv, Ca_p, Ce_p, alpha_E1, alpha_E2, alpha_S, alpha_P)
    return dict(zip(COEFS, vals))


def coefs_from_means(n: int, volume: float, mean: np.ndarray) -> Dict[str, np.ndarray]:
    """The call of md_simulation_program.f90:411 with the outputs of md_means_get."""
    m = dict(zip(QUANTITIES, mean))
    return thermodynamic_coefs_compute(n, volume, m["K"], m["P"], m["Kinv"], m["dU"], m["ddU"], m["dU_Kinv"],
                                       m["dU2_Kinv"])


def run_ensemble(lj, params, states: Sequence, total_steps: int, output_interval: int, warmup_steps: int = 0,
                 device: int = -1, group=None) -> Dict[str, Dict[str, float]]:
    """Config 5: `len(states)` independent replicas of one parameter set, batched in ONE context on
    this rank's GPU (replica = grid dimension of the all-pairs kernels), sampled every
    `output_interval` steps after `warmup_steps` exactly like md_simulation_program.f90:361, with no
    host traffic in between.  Per replica: running means -> thermodynamic coefficients; then the
    ensemble mean and population standard deviation of every coefficient.  With torch.distributed
    initialised (one process per GPU, replicas dealt out by the caller) the per-rank sums are
    combined in rank order by multi_gpu.ensemble_reduce -- the only collective of the run."""
    R = len(states)
    batched = [np.ascontiguousarray(np.stack([getattr(s, k) for s in states])) for k in lj.SimState.NAMES]
    acc = MeansAccumulator(params.n, params.volume, R)
    with lj.Context(params, method=lj.METHOD_ALLPAIRS, device=device, nreplicas=R) as ctx:
        ctx.upload(lj.SimState.from_arrays(batched))
        ctx.compute_forces()
        step = 0
        while step + output_interval <= total_steps:
            obs = np.asarray(ctx.step(output_interval)).reshape(R, 4)
            step += output_interval
            if step > warmup_steps:  # md_simulation_program.f90:361 (step is a multiple of the interval here)
                acc.add_sample(obs)
    mean, _ = acc.get()
    coefs = coefs_from_means(params.n, params.volume, mean)
    names = list(COEFS)
    vals = np.stack([np.asarray(coefs[k], dtype=np.float64).reshape(R) for k in names])  # [16][R]
    sums, sumsq, count = vals.sum(axis=1), (vals * vals).sum(axis=1), R
    try:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    except Exception:
        distributed = False
    if distributed:
        import importlib
        mg = importlib.import_module(lj.__name__ + ".multi_gpu")
        emean, estd, total = mg.ensemble_reduce(sums, sumsq, count, group=group)
    else:
        total = count
        emean = sums / total
        estd = np.sqrt(np.maximum(sumsq / total - emean * emean, 0.0))
    return {"n_replicas": int(total), "n_samples": acc.n_samples,
            "mean": {k: float(v) for k, v in zip(names, emean)},
            "std": {k: float(v) for k, v in zip(names, estd)}}
