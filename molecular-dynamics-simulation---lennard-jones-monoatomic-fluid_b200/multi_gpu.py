"""Multi-GPU host logic: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU
tests) as plumbing only.  No reference counterpart: the reference is a serial Fortran program whose
"ensemble" is a serial do-loop (/root/reference/scripts/run_many_md_simuations/run_many_md.f90:57).

* i-particle-range sharding of one large system: every rank keeps all positions,
  integrates its own contiguous atom range, and the positions are all-gathered every step.
* replica ensembles: replicas are dealt round-robin to the ranks, zero communication during the
  run, one final reduction of the per-replica observable sums.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous i-particle range [begin, end) of `rank`; requires n % world == 0 so that an
    in-place all-gather of the SoA position arrays needs no padding."""
    if n % world != 0:
        raise ValueError(f"n={n} must be divisible by world size {world} for i-range sharding")
    per = n // world
    return rank * per, (rank + 1) * per


def replicas_of_rank(nreplicas: int, rank: int, world: int) -> List[int]:
    """Replica r runs on rank r mod world."""
    return [r for r in range(nreplicas) if r % world == rank]


def reduce_observables(partial, tails: Sequence[float], group=None):
    """Sum the per-rank shares of (epot, ekin, d_epot, dd_epot) IN RANK ORDER on every rank
    (bit-reproducible, independent of the collective's algorithm) and add the tail corrections
    (lj_potential_energy.f90:221-223) once.  `partial` is a tensor of 4 doubles on the rank's
    device (CPU tensor with gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    gathered = [torch.empty_like(partial) for _ in range(world)]
    dist.all_gather(gathered, partial, group=group)
    total = torch.zeros_like(partial)
    for g in gathered:  # fixed order
        total = total + g
    out = total.detach().cpu().numpy().astype(np.float64).copy()
    out[0] += tails[0]
    out[2] += tails[1]
    out[3] += tails[2]
    return out


def allgather_positions(buf, n: int, rank: int, world: int, group=None) -> None:
    """In-place all-gather of one SoA position buffer laid out [x(n) | y(n) | z(n)]: every rank
    contributes its own contiguous atom range of each component."""
    import torch.distributed as dist

    b, e = shard_range(n, rank, world)
    for c in range(3):
        comp = buf[c * n:(c + 1) * n]
        dist.all_gather_into_tensor(comp, comp[b:e], group=group)


def ensemble_reduce(sums: np.ndarray, sumsq: np.ndarray, count: int, group=None):
    """Final reduction of a replica ensemble: `sums[k]`, `sumsq[k]` are this rank's sums over its
    replicas of observable k and its square, `count` its number of replicas.  Returns
    (mean[k], population_std[k], total_count) on every rank
    (intent of run_many_md_simuations/simulation_results.f90:330-350,391-398)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    pack = torch.tensor(np.concatenate([sums, sumsq, [float(count)]]), dtype=torch.float64, device=dev)
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(gathered, pack, group=group)
    tot = np.zeros(len(pack))
    for g in gathered:  # rank order => reproducible
        tot = tot + g.cpu().numpy()
    k = len(sums)
    ntot = tot[-1]
    mean = tot[:k] / ntot
    var = np.maximum(tot[k:2 * k] / ntot - mean * mean, 0.0)
    return mean, np.sqrt(var), int(round(ntot))


class DeviceArray:
    """Wraps a raw device address so that torch.as_tensor(...) views it without a copy."""

    def __init__(self, ptr: int, nelem: int):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3}


def torch_view(ptr: int, nelem: int, device):
    import torch

    return torch.as_tensor(DeviceArray(ptr, nelem), device=device)
