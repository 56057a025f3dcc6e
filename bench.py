#!/usr/bin/env python
"""bench.py -- headline benchmark of the LJ force / velocity-Verlet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one velocity-Verlet step (one pass of the hot path: drift+wrap+half-kick, LJ
force/energy/virial, half-kick+E_kin+reductions) over the whole system.  Metric: atom-steps/s
(N x steps / s), BASELINE.json.  Default workload = BASELINE.json configs[1]: N=4,000 FCC,
all-pairs O(N^2) with minimum image, FP64, L=20, rc=0.49 L, dt=1e-4.
Other workloads (--workload): cfg1 (N=500 shipped input), cfg3_allpairs / cfg3_cells (N=32,000,
rho=0.8, rc=2.5).

With N>1 GPUs (torchrun, one rank per GPU) the default workload does not shard (a 4,000-atom
system is one GPU's work): every rank runs an independent replica ("replicas only", weak
scaling, no data-path collective); torch.distributed (NCCL) is used only for the barrier and
the max-over-ranks time.

The JSON line carries `roofline` (FP64 pipe, algorithmic flops of SURVEY.md 8d / measured DFMA
peak), `cpu_baseline` (the C oracle timed on this box's host, 1 core: the reference is serial),
`e2e` (the drop-in ljmd_verlet_step C-ABI call with HOST arrays, H2D+D2H inside every step),
`clocks` and `gpu_launches`.

--impl reference times the reference's CPU algorithm (the oracle port compiled with the shipped
-O2 flags; the Fortran itself cannot be compiled in this image) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: k, L, rc, dt, method, temperature for the synthetic velocities
    "cfg1": dict(k=5, L=10.0, rc=4.9, dt=1e-4, method="allpairs", T=1.416),
    "cfg2": dict(k=10, L=20.0, rc=9.8, dt=1e-4, method="allpairs", T=1.416),
    "cfg3_allpairs": dict(k=20, L=34.19951893353394, rc=2.5, dt=1e-3, method="allpairs", T=1.0),
    "cfg3_cells": dict(k=20, L=34.19951893353394, rc=2.5, dt=1e-3, method="cells", T=1.0),
    "dense32k": dict(k=20, L=40.0, rc=19.6, dt=1e-4, method="allpairs", T=1.416),  # tuning only
    "cfg4_cells": dict(k=64, L=109.4384605873086, rc=2.5, dt=1e-3, method="cells", T=1.0),
    "cfg5_ensemble": dict(k=5, L=10.0, rc=4.9, dt=1e-4, method="allpairs", T=1.416, replicas=64),  # per GPU
    "cfg3_nlist": dict(k=20, L=34.19951893353394, rc=2.5, dt=1e-3, method="nlist", T=1.0),
    "cfg4_nlist": dict(k=64, L=109.4384605873086, rc=2.5, dt=1e-3, method="nlist", T=1.0),
}


def synthetic_state(w, seed=0):
    """FCC lattice (md_initial_config_program.f90:132-187 ordering) with a 2 % jitter and uniform
    random velocities (COM removed, scaled to temperature T).  Pure numpy: the benchmark input
    does not come from the oracle."""
    k, L = w["k"], w["L"]
    n = 4 * k ** 3
    a = L / k
    rng = np.random.default_rng(seed)
    idx = np.stack(np.meshgrid(np.arange(k), np.arange(k), np.arange(k), indexing="ij"), -1)
    idx = idx.reshape(-1, 1, 3).astype(np.float64)
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]], dtype=np.float64)
    pos = ((idx + basis[None]) * a).reshape(-1, 3)
    pos = pos + 0.02 * a * (rng.random((n, 3)) - 0.5)
    pos = pos - L * np.floor(pos / L)
    vel = rng.random((n, 3)) - 0.5
    vel -= vel.mean(axis=0)
    vel *= np.sqrt(3.0 * n * w["T"] / (vel ** 2).sum())
    arrs = [np.ascontiguousarray(pos[:, c]) for c in range(3)]
    arrs += [np.ascontiguousarray(vel[:, c]) for c in range(3)]
    arrs += [np.zeros(n) for _ in range(3)]
    return n, arrs


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
             0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
             0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
             0x100: "display_clock_setting"}

    def run(self):
        if self.nv is None:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def time_oracle(w, arrs, budget_s, variant="O2", max_steps=None):
    """Times the CPU oracle's verlet_step (1 core; the reference is serial) on the same workload
    for about budget_s seconds.  Returns (atom_steps_per_s, steps_run, seconds)."""
    from oracle import oracle_py as O  # the one place bench.py executes oracle/: the CPU baseline

    n = 4 * w["k"] ** 3
    p = O.params_init(n, w["L"], w["dt"], w["rc"], w["k"])
    st = [a.copy() for a in arrs]
    cells = w["method"] in ("cells", "nlist")
    if cells:
        O.compute_lj_cells(p, *st[:3])
    O.verlet_step(p, st, variant=variant, cells=cells)  # warm-up (also gives a(t))
    t0 = time.perf_counter()
    steps = 0
    while True:
        O.verlet_step(p, st, variant=variant, cells=cells)
        steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or (max_steps is not None and steps >= max_steps):
            break
    return n * steps / el, steps, el


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, arrs = synthetic_state(w)
    # every "step" is one CPU verlet_step of the same workload; bounded so the run ends in minutes
    budget = 90.0
    val, steps, el = time_oracle(w, arrs, budget, variant="O2", max_steps=max(1, args.steps))
    line = {
        "impl": "reference",
        "metric": "atom_steps_per_s", "value": val, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wname, "n_atoms": n, "box_length": w["L"], "rc": w["rc"],
                   "dt": w["dt"], "method": "cells(cpu)" if w["method"] in ("cells", "nlist") else "allpairs"},
        "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} verlet_step calls of the C restatement of the reference "
                                   f"(gcc -O2 -ffp-contract=off = shipped gfortran flags), 1 thread "
                                   f"(the reference is serial Fortran; no Fortran compiler in this image), "
                                   f"{el:.1f} s on {cpu_model()} ({physical_cores()} cores present)"},
        "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_slabs(args, w, lj, mg, dist, rank, world, local, n, arrs, p):
    """Strong scaling of ONE large system with the neighbour-list path: slab sharding in the
    cell-sorted order, halo positions pushed into peer memory by the step kernel itself, per-step
    flags instead of collectives; NCCL only for the all-gather at list rebuilds."""
    import torch
    dev = torch.device("cuda", local)
    ctx = lj.Context(p, method=lj.METHOD_NLIST, device=local)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()
    runner = mg.SlabRunner(ctx, rank, world, chunk=32)
    runner.step(args.warmup)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0, r0 = ctx.launch_count, runner.rebuilds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(runner.stream)
    share = runner.step(args.steps)
    e1.record(runner.stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    clocks = sampler.finish()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    obs = mg.reduce_observables(torch.tensor(share, dtype=torch.float64, device=dev), lj.tail_corrections(p))
    if not np.all(np.isfinite(obs)):
        raise SystemExit(f"bench.py: non-finite observables {obs}")
    info = ctx.nlist_info()
    if rank == 0:
        line = {
            "metric": "atom_steps_per_s", "value": n * args.steps / elapsed, "unit": "atom-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_atoms": n, "box_length": w["L"], "rc": w["rc"], "dt": w["dt"],
                       "method": "nlist", "skin": info["skin"],
                       "multi_gpu": "slabs of cell layers in the cell-sorted order; boundary-layer positions written "
                                    "into the neighbour ranks' buffers by the step kernel (peer memory over NVLink), "
                                    "per-step flags, no collective per step; NCCL all-gather of the state only at list "
                                    "rebuilds (%d in the timed region)" % (runner.rebuilds - r0),
                       "l2": "state %.0f MB per GPU (> L2 with the neighbour list: %.0f MB)" %
                             (96 * n / 1e6, 4 * info["entries"] / 1e6 / world)},
            "roofline": None, "cpu_baseline": None, "e2e": None, "clocks": clocks,
            "gpu_launches": int(launches * world), "wall_s": wall,
            "obs_last": [float(x) for x in obs],
            "rank0_profile_ms": {"step_kernels_device": runner.t_run_dev_ms, "slab_run_wall": 1e3 * runner.t_run,
                                 "gather_wall": 1e3 * runner.t_gather, "rebuild_wall": 1e3 * runner.t_rebuild,
                                 "note": "whole run incl. warm-up"},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def run_sharded(args, w, lj, dist, rank, world, local, n, arrs, p):
    """Strong scaling of ONE large system (cell list) over the ranks: i-particle-range sharding,
    NCCL all-gather of the positions every step, observables reduced in rank order."""
    import importlib

    import torch
    mg = importlib.import_module(lj.__name__ + ".multi_gpu")
    dev = torch.device("cuda", local)
    if w["method"] == "nlist":
        return run_slabs(args, w, lj, mg, dist, rank, world, local, n, arrs, p)
    ctx = lj.Context(p, method=lj.METHOD_CELLS, device=local)
    stream = torch.cuda.Stream(device=dev)  # library kernels and NCCL calls on ONE explicit stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ib, ie = mg.shard_range(n, rank, world)
    ctx.set_shard(ib, ie)
    ctx.compute_forces()
    p0, p1, _, _, pobs = ctx.device_ptrs()
    pos = [mg.torch_view(p0, 3 * n, dev), mg.torch_view(p1, 3 * n, dev)]
    obs_t = mg.torch_view(pobs, 4, dev)
    state = {"buf": ctx.shard_begin()}

    def one_step():
        mg.allgather_positions(pos[state["buf"]], n, rank, world)
        state["buf"] = ctx.shard_step()

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    clocks = sampler.finish()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    obs = mg.reduce_observables(obs_t.clone(), lj.tail_corrections(p))
    if not np.all(np.isfinite(obs)):
        raise SystemExit(f"bench.py: non-finite observables {obs}")
    if rank == 0:
        line = {
            "metric": "atom_steps_per_s", "value": n * args.steps / elapsed, "unit": "atom-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_atoms": n, "box_length": w["L"], "rc": w["rc"], "dt": w["dt"],
                       "method": "cells", "multi_gpu": "i-particle ranges, NCCL all-gather of positions every step "
                       "(3 x %d doubles per rank), observables summed in rank order" % (n // world),
                       "l2": "state %.0f MB per GPU; positions re-gathered every step" % (96 * n / 1e6)},
            "roofline": None, "cpu_baseline": None,
            "e2e": None, "clocks": clocks, "gpu_launches": int(launches * world), "wall_s": wall,
            "allgather_bytes_per_rank_per_step": 24 * n * (world - 1) // world,
            "obs_last": [float(x) for x in obs],
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=2000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU oracle timing")
    ap.add_argument("--e2e-steps", type=int, default=500)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, w, args.workload)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    import ljmd_b200 as lj

    if lj.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")

    sharded = (world > 1 or os.environ.get("LJMD_BENCH_FORCE_SLABS")) and w["method"] in ("cells", "nlist")
    if sharded and dist is None:  # diagnostic: the sharded code path with a single rank
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist_mod.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
        dist = dist_mod
    n, arrs = synthetic_state(w, seed=0 if sharded else rank)
    p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
    method = {"cells": lj.METHOD_CELLS, "nlist": lj.METHOD_NLIST}.get(w["method"], lj.METHOD_ALLPAIRS)
    if sharded:
        return run_sharded(args, w, lj, dist, rank, world, local, n, arrs, p)
    R = int(w.get("replicas", 1))
    if R > 1:  # ensemble of independent replicas: replica r of rank q gets its own seed
        reps = [synthetic_state(w, seed=1000 * rank + r)[1] for r in range(R)]
        arrs = [np.ascontiguousarray(np.stack([reps[r][k] for r in range(R)])) for k in range(9)]
    ctx = lj.Context(p, method=method, device=local, nreplicas=R)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()

    def barrier():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    # ---- device-resident timing: W warm-up steps, then exactly K steps ----------------------
    ctx.step(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    obs = ctx.step(args.steps)  # returns after a stream synchronize
    t1 = time.perf_counter()
    launches = ctx.launch_count - l0
    clocks = sampler.finish()
    wall = t1 - t0
    # device time: CUDA events recorded on the context's stream around the K steps
    elapsed = ctx.last_step_ms * 1e-3
    if dist is not None:
        import torch
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = n * R * args.steps * world / elapsed
    if not np.all(np.isfinite(obs)):
        raise SystemExit(f"bench.py: non-finite observables {obs}")

    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- dominant kernel, CUDA events per launch (rank 0) ------------------------------------
    ctx.force_timing(True)
    nk = 200 if n <= 8000 else 20
    ctx.step(nk)
    ms_total, nev = ctx.force_timing(False)
    force_ms = ms_total / max(nev, 1)
    cand, cut = ctx.pair_counts()
    flops = 20.0 * cand + 25.0 * cut  # unordered pairs, reference operation count
    peak = lj.measure_fp64_peak(local)
    achieved = flops / (force_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": None,
        "kernel": "lj force kernel (%s)" % w["method"], "kernel_ms": force_ms,
        "flops_per_launch": flops, "pairs_candidate": cand, "pairs_in_cutoff": cut,
        "peak_source": "measured in this run: register-resident DFMA chains (ljmd_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 entry (nominal 37 TFLOP/s)",
        "pair_interactions_per_s": cut / (force_ms * 1e-3),
        "candidate_pairs_per_s": cand / (force_ms * 1e-3),
    }
    if w["method"] == "nlist":
        info = ctx.nlist_info()
        hbm_bytes = 144.0 * n + 4.0 * info["entries"] + 32.0 * n  # SURVEY.md 8d + 4 B per stored list entry + packed r
        roofline["hbm"] = {"bound": "hbm", "achieved": hbm_bytes / (force_ms * 1e-3) / 1e9, "peak": 6545.3, "unit": "GB/s",
                           "frac": hbm_bytes / (force_ms * 1e-3) / 1e9 / 6545.3, "algorithmic_bytes": hbm_bytes,
                           "traffic": 515.8e6 if n == 1048576 else None,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                           "note": "fused step kernel: 144 B/atom state + 4 B per list entry (%.1f entries/atom) + 32 B/atom "
                                   "packed positions; traffic = dram read+write of profiles/r1_ncu_nlist_step_v2.txt" % (info["entries"] / n)}
        roofline["note"] = ("the fused neighbour-list step is co-limited: L1 tag stage for the position gathers (one 128-byte "
                            "line per clock per SM, profiles/), HBM for the list stream, FP64 pipe for the pair arithmetic")
    elif w["method"] == "allpairs":
        roofline["hbm_note"] = "state is %.1f KB, L2-resident; HBM is not the bound for all-pairs" % (72 * n * R / 1e3)
        if args.workload == "cfg2":
            roofline["traffic"] = 203.5e3  # dram read+write per launch, profiles/r1_ncu_d_cfg2_summary.txt (ncu --set full)

    # ---- end to end through the drop-in C ABI call with host arrays ---------------------------
    d = ctx.download()
    ke = max(3, min(args.e2e_steps, args.steps))
    if R == 1:
        host = lj.SimState.from_arrays([a.copy() for a in d.arrays()])
        for _ in range(5):
            lj.verlet_step(p, host)
        t0 = time.perf_counter()
        for _ in range(ke):
            lj.verlet_step(p, host)
        te = time.perf_counter() - t0
        api = "ljmd_verlet_step (drop-in for verlet_step, physics/verlet.f90:41): host arrays in/out every step"
    else:  # the reference has no batched call: upload + one resident step + download of all replicas
        host = lj.SimState.from_arrays([a.copy() for a in d.arrays()])
        t0 = time.perf_counter()
        for _ in range(ke):
            ctx.upload(host)
            ctx.step(1)
            host = ctx.download()
        te = time.perf_counter() - t0
        api = "ljmd_ctx_upload + ljmd_ctx_step(1) + ljmd_ctx_download of all replicas every step"
    e2e = {"value": n * R * ke * world / te, "unit": "atom-steps/s",
           "h2d_bytes_per_step": 9 * n * R * 8, "d2h_bytes_per_step": 9 * n * R * 8 + 32 * R,
           "steps": ke, "ms_per_step": 1e3 * te / ke, "api": api}

    # ---- CPU baseline (bounded sample) --------------------------------------------------------
    cpu = None
    if not args.no_cpu:
        host_arrs = [a.copy() for a in d.arrays()]
        v2, s2, e2 = time_oracle(w, host_arrs, args.cpu_budget, "O2")
        v3, s3, e3 = time_oracle(w, host_arrs, args.cpu_budget / 2, "O3")
        cpu = {"value": v2, "unit": "atom-steps/s", "cores": 1, "kind": "port",
               "sample": f"{s2} verlet_step calls ({e2:.1f} s) of the C restatement of the reference, "
                         f"gcc -O2 -ffp-contract=off (shipped gfortran flags), 1 thread (reference is serial); "
                         f"-O3: {v3:.4g} atom-steps/s over {s3} steps; host {cpu_model()}, "
                         f"{physical_cores()} physical cores present",
               "value_O3": v3}

    line = {
        "metric": "atom_steps_per_s", "value": value, "unit": "atom-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "n_atoms": n, "replicas_per_gpu": R, "box_length": w["L"], "rc": w["rc"],
                   "dt": w["dt"], "method": w["method"],
                   "multi_gpu": "replicas only (independent systems per rank, no data-path collective)" if world > 1 else
                                ("single system" if R == 1 else "%d independent replicas on one GPU" % R),
                   "l2": ("not flushed: the MD state (%.0f KB) is L2/shared-memory resident by construction "
                          "and re-read every step; the kernel is FP64-pipe bound" % (72 * n * R / 1e3)) if w["method"] == "allpairs"
                         else "inputs larger than L2: state + neighbour list = %.0f MB, streamed from HBM every step" %
                              ((96 * n + 300 * n) / 1e6)},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": int(launches), "wall_s": wall,
        "obs_last": [float(x) for x in np.ravel(obs)[:4]],
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
