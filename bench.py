#!/usr/bin/env python
"""bench.py -- benchmark of the LJ force / velocity-Verlet hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one velocity-Verlet step (one pass of the hot path: drift+wrap+half-kick, LJ
force/energy/virial, half-kick+E_kin+reductions) over the whole system.  Metric: atom-steps/s
(N x steps / s), BASELINE.json.

Default workload, for EVERY --gpus N: cfg4_nlist = BASELINE.json configs[3], N = 1,048,576 atoms at
rho = 0.8, rc = 2.5, neighbour list -- the largest configuration that fits one GPU and the one the
north star's multi-GPU target is quoted on.
  * N = 1: the single device-resident context; the JSON line carries `roofline` (HBM and FP64),
    `cpu_baseline`, `e2e` (the drop-in ljmd_verlet_step call with HOST arrays), `long_run`
    (list rebuilds amortised) and a `secondary` block with the other single-GPU configs:
    cfg3_allpairs (N = 32,000 all-pairs, the north star's FP64 target) and cfg2 (N = 4,000).
  * N > 1 (torchrun, one rank per GPU): STRONG scaling of the same system, slab sharding with the
    halo exchange fused into the step kernel over NVLink peer memory (multi_gpu.SlabRunner).
Other workloads (--workload): cfg1, cfg2, cfg3_allpairs, cfg3_nlist, cfg3_cells, cfg4_cells,
cfg5_ensemble.  Workloads that do not shard (all-pairs) run as independent replicas per rank.

--impl reference times the reference's CPU algorithm (the C restatement compiled with the shipped
-O2 flags; the Fortran itself cannot be compiled in this image) on the same workload and config.
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
DEFAULT_WORKLOAD = "cfg4_nlist"
HBM_FALLBACK_GBS = 6545.3  # this pool's measured copy bandwidth, used only if MEASURED_PEAKS.json is absent


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


def shards(w):
    return w["method"] in ("cells", "nlist")


def config_of(wname, world):
    """The `config` object of the JSON line.  Both arms (--impl ours / reference) print exactly this
    dict: it names the workload, nothing measured goes in here."""
    w = WORKLOADS[wname]
    n = 4 * w["k"] ** 3
    R = int(w.get("replicas", 1))
    if world > 1 and shards(w):
        mg = ("strong scaling of ONE system over %d GPUs: slabs of cell layers in the cell-sorted order; the step kernel "
              "writes its boundary-layer positions into the neighbour ranks' buffers (peer memory over NVLink) and the ranks "
              "synchronise through per-step flags -- no collective per step; NCCL only at list rebuilds" % world
              if w["method"] == "nlist" else
              "strong scaling over %d GPUs: i-particle ranges, NCCL all-gather of the positions every step" % world)
    elif world > 1:
        mg = "replicas only: an independent system per rank, no data-path collective (the workload does not shard)"
    else:
        mg = "single system" if R == 1 else "%d independent replicas on one GPU" % R
    if w["method"] == "allpairs":
        l2 = ("not flushed: the MD state (%.0f KB) is L2/shared-memory resident by construction and re-read every "
              "step; the kernel is FP64-pipe bound" % (72 * n * R / 1e3))
    else:
        l2 = ("inputs larger than L2: state + neighbour list = about %.0f MB, streamed from HBM every step"
              % ((96 * n + 300 * n) / 1e6))
    return {"workload": wname, "n_atoms": n, "replicas_per_gpu": R, "box_length": w["L"], "rc": w["rc"],
            "dt": w["dt"], "method": w["method"], "multi_gpu": mg, "l2": l2}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (MEASURED_PEAKS.json absent): this pool's earlier copy-bandwidth measurement"


def profiled_traffic(kernel_key):
    """DRAM bytes per launch of a kernel from a committed `ncu --set full` capture, or None.
    profiles/traffic.json: {key: {"bytes": dram read+write per launch, "source": file under profiles/}}."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(kernel_key)
        return (float(t["bytes"]), t["source"]) if t else (None, None)
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.004):
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


def time_oracle(w, arrs, budget_s, variant="O2", max_steps=None, warm=1):
    """Times the CPU oracle's verlet_step (1 core; the reference is serial) on the same workload
    for about budget_s seconds.  Returns (atom_steps_per_s, steps_run, seconds)."""
    from oracle import oracle_py as O  # the one place bench.py executes oracle/: the CPU baseline

    n = 4 * w["k"] ** 3
    p = O.params_init(n, w["L"], w["dt"], w["rc"], w["k"])
    st = [a.copy() for a in arrs]
    cells = shards(w)
    if cells:
        O.compute_lj_cells(p, *st[:3])
    for _ in range(warm):
        O.verlet_step(p, st, variant=variant, cells=cells)  # warm-up (the first one also gives a(t))
    t0 = time.perf_counter()
    steps = 0
    while True:
        O.verlet_step(p, st, variant=variant, cells=cells)
        steps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or (max_steps is not None and steps >= max_steps):
            break
    return n * steps / el, steps, el


def cpu_note(w):
    if shards(w):
        return ("CPU *cell-list* oracle (same pair arithmetic as the reference, builder-defined binning): NOT a reference "
                "capability -- the reference is all-pairs O(N^2) only")
    return "C restatement of the reference's all-pairs loop"


def allpairs_extrapolation(w, n):
    """The reference's own O(N^2) algorithm at this N, extrapolated from a measured N=4,000 all-pairs step."""
    w2 = WORKLOADS["cfg2"]
    n2, arrs2 = synthetic_state(w2)
    v, s, el = time_oracle(w2, arrs2, 1.0, "O2", max_steps=4)
    pairs_per_s = (n2 * (n2 - 1) / 2) * s / el
    sec_per_step = (n * (n - 1) / 2) / pairs_per_s
    return {"atom_steps_per_s": n / sec_per_step, "s_per_step": sec_per_step,
            "basis": "%.3g candidate pairs/s measured on %d all-pairs steps at N=%d" % (pairs_per_s, s, n2)}


def run_reference(args, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    w = WORKLOADS[wname]
    n, arrs = synthetic_state(w)
    # every "step" is one CPU verlet_step of the same workload; bounded so the run ends in minutes
    val, steps, el = time_oracle(w, arrs, 150.0, variant="O2", max_steps=max(1, args.steps), warm=max(1, args.warmup))
    sample = (f"{steps} verlet_step calls ({el:.1f} s, after {max(1, args.warmup)} warm-up calls) of the {cpu_note(w)}; "
              f"gcc -O2 -ffp-contract=off (= the shipped gfortran flags), 1 thread (the reference is serial Fortran; "
              f"no Fortran compiler in this image); host {cpu_model()}, {physical_cores()} physical cores present")
    line = {
        "impl": "reference",
        "metric": "atom_steps_per_s", "value": val, "unit": "atom-steps/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps,
        "higher_is_better": True, "scaling": "strong" if (world > 1 and shards(w)) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wname, world),
        "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if shards(w) and n > 100000:
        line["cpu_baseline"]["allpairs_extrapolated"] = allpairs_extrapolation(w, n)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# Single GPU: one workload through the device-resident context + the drop-in host-array call
# ------------------------------------------------------------------------------------------------
def measure_single(lj, wname, steps, warmup, local, seed=0, e2e_steps=20, cpu_budget=12.0, long_steps=0,
                   barrier=None, allreduce_max=None, world=1):
    """Returns the measured fields of one workload on one GPU (the caller adds the contract keys)."""
    w = WORKLOADS[wname]
    n, arrs = synthetic_state(w, seed=seed)
    p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
    method = {"cells": lj.METHOD_CELLS, "nlist": lj.METHOD_NLIST}.get(w["method"], lj.METHOD_ALLPAIRS)
    R = int(w.get("replicas", 1))
    if R > 1:  # ensemble of independent replicas: replica r of rank q gets its own seed
        reps = [synthetic_state(w, seed=1000 * seed + r)[1] for r in range(R)]
        arrs = [np.ascontiguousarray(np.stack([reps[r][k] for r in range(R)])) for k in range(9)]
    ctx = lj.Context(p, method=method, device=local, nreplicas=R)
    ctx.upload(lj.SimState.from_arrays(arrs))
    ctx.compute_forces()

    # ---- device-resident timing: W warm-up steps, then exactly K steps ----------------------
    ctx.step(warmup)
    reb0 = ctx.nlist_info()["rebuilds"] if w["method"] == "nlist" else 0
    sampler = ClockSampler(local)
    sampler.start()
    if barrier:
        barrier()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    obs = ctx.step(steps)  # graphs are prepared before the first event; returns after a stream synchronize
    wall = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    clocks = sampler.finish()
    elapsed = ctx.last_step_ms * 1e-3  # CUDA events recorded on the context's stream around the K steps
    if allreduce_max:
        elapsed = allreduce_max(elapsed)
    if not np.all(np.isfinite(obs)):
        raise SystemExit(f"bench.py: non-finite observables {obs}")
    out = {"value": n * R * steps * world / elapsed, "ms_per_step": 1e3 * elapsed / steps, "clocks": clocks,
           "gpu_launches": int(launches), "wall_s": wall, "obs_last": [float(x) for x in np.ravel(obs)[:4]],
           "n": n, "R": R, "ctx": ctx, "p": p, "arrs": arrs}
    if w["method"] == "nlist":
        out["rebuilds_in_timed_region"] = int(ctx.nlist_info()["rebuilds"] - reb0)
    return out


def finish_single(lj, wname, m, local, steps, e2e_steps, cpu_budget, long_steps, no_cpu):
    """Rank-0 extras of measure_single: dominant-kernel roofline, long run, e2e, CPU baseline."""
    w = WORKLOADS[wname]
    ctx, p, n, R = m.pop("ctx"), m.pop("p"), m["n"], m["R"]
    arrs = m.pop("arrs")

    # ---- list rebuilds amortised: a run long enough to contain several -------------------------
    if long_steps and w["method"] == "nlist":
        r0 = ctx.nlist_info()["rebuilds"]
        ctx.step(long_steps)
        ms = ctx.last_step_ms
        nreb = int(ctx.nlist_info()["rebuilds"] - r0)
        m["long_run"] = {"steps": long_steps, "ms_per_step": ms / long_steps, "value": n * long_steps / (ms * 1e-3),
                         "rebuilds": nreb, "steps_per_rebuild": (long_steps / nreb) if nreb else None,
                         "note": "list rebuilds (device-triggered, inside the CUDA graph) amortised over the run"}

    # ---- dominant kernel, CUDA events per launch on the context's stream -------------------------
    ctx.force_timing(True)
    nk = 200 if n <= 8000 else 20
    ctx.step(nk)
    ms_total, nev = ctx.force_timing(False)
    force_ms = ms_total / max(nev, 1)
    cand, cut = ctx.pair_counts()
    flops = 20.0 * cand + 25.0 * cut  # unordered pairs, reference operation count
    peak = lj.measure_fp64_peak(local)
    peak_sus = lj.measure_fp64_peak_sustained(local, 1.0)
    fp64 = {"bound": "fp64", "achieved": flops / (force_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
            "frac": flops / (force_ms * 1e-3) / 1e12 / peak, "flops_per_launch": flops,
            "pairs_candidate": cand, "pairs_in_cutoff": cut,
            "peak_sustained": peak_sus, "frac_of_sustained": flops / (force_ms * 1e-3) / 1e12 / peak_sus,
            "peak_source": "measured in this run: register-resident DFMA chains (ljmd_measure_fp64_peak: best short burst "
                           "after idle = `peak`, the conservative denominator of `frac`; the same chains under 1 s of "
                           "back-to-back load, second half = `peak_sustained`); "
                           "MEASURED_PEAKS.json has no FP64 entry (nominal 37 TFLOP/s)",
            "pair_interactions_per_s": cut / (force_ms * 1e-3), "candidate_pairs_per_s": cand / (force_ms * 1e-3)}
    if w["method"] == "nlist":
        info = ctx.nlist_info()
        hbm_bytes = 144.0 * n + 4.0 * info["entries"] + 32.0 * n
        pk, src = hbm_peak()
        traffic, tsrc = profiled_traffic("nl_step_kernel:%d" % n)
        roofline = {"bound": "hbm", "achieved": hbm_bytes / (force_ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                    "frac": hbm_bytes / (force_ms * 1e-3) / 1e9 / pk, "traffic": traffic, "traffic_source": tsrc,
                    "kernel": "nl_step_kernel (fused neighbour-list force + velocity-Verlet step)", "kernel_ms": force_ms,
                    "algorithmic_bytes": hbm_bytes, "peak_source": src,
                    "bytes_note": "144 B/atom state (SURVEY.md 8d) + 4 B per stored list entry (%.1f entries/atom) + 32 B/atom "
                                  "packed positions" % (info["entries"] / n),
                    "fp64": fp64,
                    "note": "co-limited kernel: L1 tag stage for the position gathers, HBM for the list stream, FP64 pipe "
                            "for the pair arithmetic; the larger fraction is the bound"}
    else:
        traffic, tsrc = profiled_traffic("lj_pairs:%s" % wname)
        roofline = dict(fp64)
        roofline.update({"traffic": traffic, "traffic_source": tsrc, "kernel": "lj force kernel (%s)" % w["method"],
                         "kernel_ms": force_ms})
        if w["method"] == "allpairs":
            roofline["hbm_note"] = "state is %.1f KB, L2-resident; HBM is not the bound for all-pairs" % (72 * n * R / 1e3)
    m["roofline"] = roofline

    # ---- end to end through the drop-in C ABI call with host arrays ---------------------------
    d = ctx.download()
    ke = max(3, min(e2e_steps, steps))
    host = lj.SimState.from_arrays([a.copy() for a in d.arrays()])
    if R == 1:
        for _ in range(3):
            lj.verlet_step(p, host)
        t0 = time.perf_counter()
        for _ in range(ke):
            lj.verlet_step(p, host)
        te = time.perf_counter() - t0
        api = "ljmd_verlet_step (drop-in for verlet_step, physics/verlet.f90:41): host arrays in/out every step"
    else:  # the reference has no batched call: upload + one resident step + download of all replicas
        t0 = time.perf_counter()
        for _ in range(ke):
            ctx.upload(host)
            ctx.step(1)
            host = ctx.download()
        te = time.perf_counter() - t0
        api = "ljmd_ctx_upload + ljmd_ctx_step(1) + ljmd_ctx_download of all replicas every step"
    m["e2e"] = {"value": n * R * ke / te, "unit": "atom-steps/s", "h2d_bytes_per_step": 9 * n * R * 8,
                "d2h_bytes_per_step": 9 * n * R * 8 + 32 * R, "steps": ke, "ms_per_step": 1e3 * te / ke, "api": api}
    lj.compat_release()
    ctx.close()

    # ---- CPU baseline (bounded sample) --------------------------------------------------------
    m["cpu_baseline"] = None
    if not no_cpu:
        host_arrs = [a.copy() for a in d.arrays()]
        # the downloaded state carries a(t): no warm-up call needed where one step takes seconds
        v2, s2, e2 = time_oracle(w, host_arrs, cpu_budget, "O2", warm=0 if n >= 32000 else 1)
        m["cpu_baseline"] = {
            "value": v2, "unit": "atom-steps/s", "cores": 1, "kind": "port",
            "sample": f"{s2} verlet_step calls ({e2:.1f} s) of the {cpu_note(w)}; gcc -O2 -ffp-contract=off (the shipped "
                      f"gfortran flags), 1 thread (the reference is serial); host {cpu_model()}, "
                      f"{physical_cores()} physical cores present"}
        if shards(w) and n > 100000:
            m["cpu_baseline"]["allpairs_extrapolated"] = allpairs_extrapolation(w, n)
    del m["n"], m["R"]
    return m


# ------------------------------------------------------------------------------------------------
# Multi GPU
# ------------------------------------------------------------------------------------------------
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
    runner.prepare(args.steps)  # graphs of the timed region, like ljmd_ctx_step prepares its own before the first event
    # the sampler thread (NVML initialisation: milliseconds, different on every rank) starts BEFORE the barrier:
    # whatever a rank does between the barrier and its first launch is waited for by all the others
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    dist.barrier()
    l0, r0, d0 = ctx.launch_count, runner.rebuilds, runner.t_run_dev_ms
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(runner.stream)
    share = runner.step(args.steps)
    e1.record(runner.stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    clocks = sampler.finish()
    rebuilds = runner.rebuilds - r0
    step_dev_ms = runner.t_run_dev_ms - d0
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    obs = mg.reduce_observables(torch.tensor(share, dtype=torch.float64, device=dev), lj.tail_corrections(p))
    if not np.all(np.isfinite(obs)):
        raise SystemExit(f"bench.py: non-finite observables {obs}")

    # ---- a longer run with the rebuilds amortised (all ranks) -----------------------------------
    long_run = None
    if args.long_steps:
        torch.cuda.synchronize()
        dist.barrier()
        r1 = runner.rebuilds
        e0.record(runner.stream)
        runner.step(args.long_steps)
        e1.record(runner.stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nreb = runner.rebuilds - r1
        long_run = {"steps": args.long_steps, "ms_per_step": 1e3 * float(t.item()) / args.long_steps,
                    "value": n * args.long_steps / float(t.item()), "rebuilds": nreb,
                    "steps_per_rebuild": (args.long_steps / nreb) if nreb else None,
                    "note": "list rebuilds (NCCL all-gather of the slabs + re-sort + list build) amortised over the run"}
    info = ctx.nlist_info()  # sharded: entries of the owned slab
    bounds = runner.bounds
    own = int(bounds[rank + 1] - bounds[rank])
    nc = info["cells_per_edge"]
    lo, hi = mg.slab_layers(nc, rank, world)
    halo_atoms = 2.0 * own / max(hi - lo, 1)  # first + last owned cell layer
    if rank == 0:
        kernel_ms = step_dev_ms / max(args.steps, 1)
        hbm_bytes = (144.0 + 32.0) * own + 4.0 * info["entries"]
        pk, src = hbm_peak()
        roofline = {"bound": "hbm", "achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                    "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / pk, "traffic": None,
                    "kernel": "nl_step_kernel, sharded instantiation (rank 0's slab: %d of %d atoms)" % (own, n),
                    "kernel_ms": kernel_ms, "algorithmic_bytes": hbm_bytes, "peak_source": src,
                    "note": "rank 0; kernel_ms = device time of the step launches of the timed region / steps (CUDA events on the "
                            "runner's stream), i.e. it includes the wait for the neighbours' per-step flags"}
        line = {
            "metric": "atom_steps_per_s", "value": n * args.steps / elapsed, "unit": "atom-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, world),
            "roofline": roofline, "cpu_baseline": None,
            "e2e": None, "e2e_note": "not measured at N > 1: the host-array drop-in call is a single-GPU API; the N = 1 line has it",
            "clocks": clocks, "gpu_launches": int(launches * world), "wall_s": wall,
            "rebuilds_in_timed_region": int(rebuilds), "long_run": long_run,
            "halo_bytes_per_step_per_rank": 32.0 * halo_atoms,
            "nvlink_note": "each rank stores the packed positions (32 B) of its first and last owned cell layer into the two "
                           "neighbour ranks' buffers every step, plus %d 4-byte flags" % world,
            "skin": info["skin"], "obs_last": [float(x) for x in obs],
            "rank0_profile_ms": {"step_kernels_device": runner.t_run_dev_ms, "slab_run_wall": 1e3 * runner.t_run,
                                 "gather_wall": 1e3 * runner.t_gather, "rebuild_wall": 1e3 * runner.t_rebuild,
                                 "note": "whole run incl. warm-up and long run"},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def run_ensemble_bench(args, w, lj, dist, rank, world, local):
    """BASELINE.json configs[4]: an ensemble of independent N = 500 replicas (run_many_md_simuations intent,
    run_many_md.f90:57-82), 64 per GPU, replica r (global index) started like md_initial_config_program with the
    velocity seed -(12345 + r), batched in ONE context per GPU, sampled every 100 steps, no
    communication during the run, ONE final NCCL reduction of the per-replica thermodynamic coefficients
    (thermodynamic_coefs.f90:87-205) to the ensemble mean / population std.  A "step" is one velocity-Verlet
    step of every replica; the timed region is the whole run_ensemble call (sampling + the final reduction
    included) after a warm-up call of the same shape."""
    import importlib
    ic = importlib.import_module(lj.__name__ + ".initial_config")
    ob = importlib.import_module(lj.__name__ + ".observables")
    R = int(w.get("replicas", 64))
    n = 4 * w["k"] ** 3
    p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
    states = [ic.initial_config(lj, p, seed=-(12345 + rank * R + r), target_total_energy=-555.0 * n / 500.0)[0] for r in range(R)]
    lj.compat_release()
    interval = 100
    steps = max(interval, (args.steps // interval) * interval)
    warm = max(interval, (args.warmup // interval) * interval)
    ob.run_ensemble(lj, p, states, warm, interval, 0, device=local)  # warm-up: kernels, graphs, NCCL channels
    if dist is not None:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    out = ob.run_ensemble(lj, p, states, steps, interval, 0, device=local)
    el = time.perf_counter() - t0
    clocks = sampler.finish()
    if dist is not None:
        import torch
        t = torch.tensor([el], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    if rank == 0:
        line = {
            "metric": "atom_steps_per_s", "value": n * R * world * steps / el, "unit": "atom-steps/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args.workload, world),
            "roofline": None, "cpu_baseline": None, "e2e": None, "clocks": clocks,
            "gpu_launches": int(2 * steps * world + world),
            "timing": "host wall clock around observables.run_ensemble (max over ranks): H2D of the replicas, the steps, "
                      "the per-sample D2H of 4 observables per replica and the final reduction are all inside",
            "ensemble": {"n_replicas": out["n_replicas"], "n_samples": out["n_samples"],
                         "reduction": "multi_gpu.ensemble_reduce: one NCCL all-gather of (sum, sum of squares, count) per "
                                      "coefficient, added in rank order" if dist is not None else "single rank",
                         "mean": {k: out["mean"][k] for k in ("temperature", "pressure", "Ca_v", "Ce_v", "gamma", "K_S", "alpha_E1", "alpha_E2")
                                  if k in out["mean"]},
                         "std": {k: out["std"][k] for k in ("temperature", "pressure", "Ca_v", "Ce_v", "gamma", "K_S", "alpha_E1", "alpha_E2")
                                 if k in out["std"]}},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
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
            "config": config_of(args.workload, world),
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
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU oracle timing")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--long-steps", type=int, default=None,
                    help="extra run with the list rebuilds amortised (neighbour-list workloads; 0 = off)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    n_atoms = 4 * w["k"] ** 3
    if args.steps is None:  # defaults that finish within minutes: about a second of GPU time
        args.steps = 20 if args.impl == "reference" else (640 if n_atoms >= 100000 else 20000 if n_atoms <= 8000 else 1000)
    if args.warmup is None:
        args.warmup = 3 if args.impl == "reference" else max(3, args.steps // 10)
    if args.warmup < 3:
        args.warmup = 3
    if args.long_steps is None:
        args.long_steps = 256 if w["method"] == "nlist" else 0

    if args.impl == "reference":
        run_reference(args, args.workload)
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

    if args.workload == "cfg5_ensemble" and not os.environ.get("LJMD_BENCH_SYNTHETIC_REPLICAS"):
        return run_ensemble_bench(args, w, lj, dist, rank, world, local)
    sharded = (world > 1 or os.environ.get("LJMD_BENCH_FORCE_SLABS")) and shards(w)
    if sharded and dist is None:  # diagnostic: the sharded code path with a single rank
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist_mod.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local))
        dist = dist_mod
    if sharded:
        n, arrs = synthetic_state(w, seed=0)
        p = lj.init_params(n, w["L"], w["dt"], w["rc"], w["k"])
        return run_sharded(args, w, lj, dist, rank, world, local, n, arrs, p)

    def barrier():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def allreduce_max(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m = measure_single(lj, args.workload, args.steps, args.warmup, local, seed=rank, barrier=barrier,
                       allreduce_max=allreduce_max, world=world)
    if rank != 0:
        m["ctx"].close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    # rank 0: roofline of the dominant kernel always; e2e and the CPU leg only at N = 1 (at N > 1 the other
    # ranks would idle in the barrier while rank 0 runs them)
    m = finish_single(lj, args.workload, m, local, args.steps, args.e2e_steps, args.cpu_budget,
                      args.long_steps if world == 1 else 0, args.no_cpu or world > 1)
    if world > 1:
        m["e2e"] = None
        m["e2e_note"] = "not measured at N > 1 (replicas): the N = 1 line has the host-array call"
    line = {
        "metric": "atom_steps_per_s", "value": m.pop("value"), "unit": "atom-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m.pop("ms_per_step"), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload, world),
    }
    line.update(m)

    # ---- the other single-GPU configurations of BASELINE.json, same process, same box ----------
    if world == 1 and not args.no_secondary and args.workload == DEFAULT_WORKLOAD:
        sec = {}
        for name, (k_steps, k_warm, budget) in {"cfg3_allpairs": (40, 5, 5.0), "cfg2": (2000, 200, 2.0)}.items():
            s = measure_single(lj, name, k_steps, k_warm, local)
            s = finish_single(lj, name, s, local, k_steps, 20, budget, 0, args.no_cpu)
            s.update({"steps": k_steps, "warmup": k_warm, "config": config_of(name, 1)})
            # the top-level keys a reader looks for first
            s["frac"] = s["roofline"]["frac"]
            s["kernel_ms"] = s["roofline"]["kernel_ms"]
            sec[name] = s
        line["secondary"] = sec
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
