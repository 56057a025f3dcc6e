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


def slab_layers(nc: int, rank: int, world: int) -> Tuple[int, int]:
    """Cell layers cx in [lo, hi) owned by `rank` (same rule as the library)."""
    return rank * nc // world, (rank + 1) * nc // world


class SlabRunner:
    """Drives one rank of a slab-sharded neighbour-list run (include/ljmd.h, ljmd_ctx_slab_*).

    Per step nothing goes through the host or NCCL: the step kernels push their boundary layers
    into the neighbours' position buffers over NVLink and synchronise through per-step flags.
    torch.distributed is used (i) once, to exchange the CUDA IPC handles, and (ii) whenever a rank
    reports a skin violation.  Then, by default (`rebuild_local`): two cell layers are exchanged with
    each of the two neighbours (NCCL send / recv), the cell histogram is completed by ONE all-reduce,
    and every rank re-sorts and rebuilds its own slab and halo layers only -- nothing scales with the
    whole system.  LJMD_SLAB_LOCAL=0 (or fewer than two cell layers per rank) selects the first
    version: one all-gather of the packed slabs, after which every rank re-sorts globally.  Both give
    the same slots, lists and trajectory, bit for bit.  After local rebuilds a rank holds current data
    for its slab and halo only; finish() gathers the complete state.

    `ctx` must hold the COMPLETE state (uploaded on every rank) with forces computed
    (ctx.compute_forces()).  Reference semantics of a step: physics/verlet.f90:41-97.
    """

    def __init__(self, ctx, rank: int, world: int, group=None, chunk: int = 32):
        import torch
        import torch.distributed as dist

        self.ctx, self.rank, self.world, self.group, self.chunk = ctx, rank, world, group, chunk
        self.torch, self.dist = torch, dist
        self.dev = torch.device("cuda", torch.cuda.current_device())
        # one explicit stream for the library's kernels AND the NCCL calls (torch's default stream
        # has handle 0, which the library reads as "use the context's own stream")
        self.stream = torch.cuda.Stream(device=self.dev)
        ctx.set_stream(self.stream.cuda_stream)
        blob, _ = ctx.slab_export()
        blobs = [None] * world
        dist.all_gather_object(blobs, blob, group=group)
        ctx.slab_attach(rank, world, handles=blobs)
        self.narr = ctx.slab_narr()
        self.n = ctx.params.n
        self.send = self.recv = None
        self.rebuilds = 0
        self.steps = 0
        import os
        self.debug = bool(os.environ.get("LJMD_SLAB_DEBUG_PY"))
        self.local = False
        if world >= 2 and os.environ.get("LJMD_SLAB_LOCAL", "1") != "0":
            try:
                ptr, ncells, cap, na = ctx.slab_local_info()
                self.local = True
            except Exception:
                self.local = False  # fewer than two cell layers per rank: global re-sort
        if self.local:
            self.hist = torch.as_tensor(DeviceArray(ptr, ncells, "<i4"), device=self.dev)
            self.halo_cap = cap
            self.halo = [torch.empty(na * cap, dtype=torch.float64, device=self.dev) for _ in range(4)]  # sendL sendR recvL recvR
        self.t_run_dev_ms = self.t_run = self.t_gather = self.t_rebuild = 0.0  # profile of step()
        ctx.slab_drift()     # r(t+dt) of every atom
        ctx.slab_rebuild()   # only if the drift already violated the skin of the t=0 list
        self.bounds = ctx.slab_bounds()
        # one gather now (idempotent: every rank already holds the complete state): NCCL sets up its channels on
        # the first collective of a size class, which would otherwise cost the first real rebuild a millisecond
        self.gather()
        if self.local:  # the same for the point-to-point channels and the histogram all-reduce of a local rebuild
            with torch.cuda.stream(self.stream):
                self._exchange_halos()
                dist.all_reduce(torch.zeros_like(self.hist), group=group)
        self.stream.synchronize()
        dist.barrier(group=group)  # every rank is attached before anybody writes into a peer

    def _buffers(self, cmax: int):
        torch = self.torch
        if self.send is None or self.send.numel() < self.narr * cmax:
            cap = int(cmax * 1.2) + 1024
            self.send = torch.empty(self.narr * cap, dtype=torch.float64, device=self.dev)
            self.recv = torch.empty(self.world * self.narr * cap, dtype=torch.float64, device=self.dev)
        return self.send[: self.narr * cmax], self.recv[: self.world * self.narr * cmax]

    def gather(self) -> None:
        """Complete state on every rank: pack own slab, one all-gather, unpack."""
        cmax = int(np.max(np.diff(self.bounds)))
        send, recv = self._buffers(cmax)
        with self.torch.cuda.stream(self.stream):
            self.ctx.slab_pack(send.data_ptr(), cmax)
            self.dist.all_gather_into_tensor(recv, send, group=self.group)
            self.ctx.slab_unpack(recv.data_ptr(), cmax)

    def _exchange_halos(self) -> None:
        """send buffers -> the two neighbours, their buffers -> recv (NCCL send / recv on the current stream)."""
        dist = self.dist
        send_l, send_r, recv_l, recv_r = self.halo
        left, right = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        # send right first, receive from the left first: with two ranks both neighbours are the SAME peer and
        # NCCL matches the sends and receives of a pair in order
        ops = [dist.P2POp(dist.isend, send_r, right, group=self.group), dist.P2POp(dist.isend, send_l, left, group=self.group),
               dist.P2POp(dist.irecv, recv_l, left, group=self.group), dist.P2POp(dist.irecv, recv_r, right, group=self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def rebuild_local(self) -> None:
        """Slab-local rebuild after a halt (include/ljmd.h, ljmd_ctx_slab_rebuild_local)."""
        dist, ctx, cap = self.dist, self.ctx, self.halo_cap
        send_l, send_r, recv_l, recv_r = self.halo
        with self.torch.cuda.stream(self.stream):
            ctx.slab_halo(send_l.data_ptr(), cap, 0, False)   # my first two cell layers, for the left neighbour
            ctx.slab_halo(send_r.data_ptr(), cap, 1, False)   # my last two, for the right neighbour
            self._exchange_halos()
            ctx.slab_halo(recv_l.data_ptr(), cap, 0, True)
            ctx.slab_halo(recv_r.data_ptr(), cap, 1, True)
            ctx.slab_rebuild_local(0)
            dist.all_reduce(self.hist, group=self.group)      # every rank counted the cells of its own layers
            ctx.slab_rebuild_local(1)

    def prepare(self, nsteps: int) -> None:
        """Captures the CUDA graphs step(nsteps) will replay (chunks of `chunk` steps and the remainder) ahead of
        time.  A halt in the middle of a chunk leaves an odd-sized remainder whose graph is captured when it occurs."""
        for m in {min(self.chunk, nsteps), nsteps % self.chunk}:
            if m >= 2:
                self.ctx.slab_prepare(m)

    def step(self, nsteps: int) -> np.ndarray:
        """nsteps velocity-Verlet steps; returns this rank's observable share of the last step."""
        left = nsteps
        obs = np.zeros(4)
        import os
        import time
        while left > 0:
            t0 = time.perf_counter()
            try:
                done, halted, obs = self.ctx.slab_run(min(self.chunk, left))
            except Exception:
                try:  # tell the peers: their step kernels halt instead of spinning on this rank's flag
                    self.ctx.slab_poison()
                except Exception:
                    pass
                raise
            t1 = time.perf_counter()
            self.t_run += t1 - t0
            self.t_run_dev_ms += self.ctx.last_step_ms
            if self.debug:  # every rank must report the same progress
                rep = [None] * self.world
                self.dist.all_gather_object(rep, (done, halted), group=self.group)
                if any(r != rep[0] for r in rep):
                    print(f"[slab debug] rank {self.rank}: ranks disagree on progress at step {self.steps}: {rep}", flush=True)
            left -= done
            self.steps += done
            if halted:
                shadow = os.environ.get("LJMD_SLAB_SHADOW")
                if shadow:
                    d0 = self.ctx.download(("vx", "ax", "ay"))
                if self.local:
                    t2 = time.perf_counter()
                    self.rebuild_local()
                else:
                    self.gather()
                    self.stream.synchronize()
                    t2 = time.perf_counter()
                    self.ctx.slab_rebuild()
                    self.bounds = self.ctx.slab_bounds()
                self.rebuilds += 1
                if shadow:
                    d1 = self.ctx.download(("vx", "ax", "ay"))
                    for nm in ("vx", "ax", "ay"):
                        bad = np.nonzero(getattr(d0, nm) != getattr(d1, nm))[0]
                        if len(bad):
                            print(f"[slab shadow] rank {self.rank} rebuild {self.rebuilds} step {self.steps}: {nm} changed for "
                                  f"{len(bad)} atoms, orig idx {bad[:4]}..{bad[-1]}", flush=True)
                if self.debug:  # the complete state must be identical on every rank here
                    cs = self.ctx.debug_checksum()
                    allcs = [None] * self.world
                    self.dist.all_gather_object(allcs, cs, group=self.group)
                    if any(c != allcs[0] for c in allcs):
                        print(f"[slab debug] rank {self.rank}: state differs after rebuild {self.rebuilds} "
                              f"(step {self.steps}): {allcs}", flush=True)
                t3 = time.perf_counter()
                self.t_gather += t2 - t1
                self.t_rebuild += t3 - t2
            elif done == 0 and left > 0:
                raise RuntimeError("slab run made no progress")
        return obs

    def finish(self) -> None:
        """Leave the complete, current state on every rank (for download)."""
        self.bounds = self.ctx.slab_bounds()
        self.gather()


class DeviceArray:
    """Wraps a raw device address so that torch.as_tensor(...) views it without a copy."""

    def __init__(self, ptr: int, nelem: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3}


def torch_view(ptr: int, nelem: int, device):
    import torch

    return torch.as_tensor(DeviceArray(ptr, nelem), device=device)
