"""B200-native Lennard-Jones MD hot path: host-side mirror of the reference interface.

This package is a thin ctypes layer over ``lib/libljmd.so`` (hand-written sm_100a CUDA behind
the C ABI declared in ``include/ljmd.h``).  It mirrors the reference's module procedures

* ``compute_lj_potential_energy(params, state, epot, d_epot, dd_epot)``
  (/root/reference/scripts/physics/lj_potential_energy.f90:46)
* ``verlet_step(params, state, epot, ekin, d_epot, dd_epot)``
  (/root/reference/scripts/physics/verlet.f90:41)
* ``init_params`` / ``init_state`` (/root/reference/scripts/base/md_types.f90:105,175)

with the same argument meaning and error behaviour (a reference ``stop '<msg>'`` becomes an
:class:`LjmdError` carrying the same message), plus the device-resident :class:`Context`.

There is NO CPU fallback here: if the CUDA library is missing or no GPU is present, compute
calls raise.  The directory name contains '-' so import it with
``importlib.import_module("molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200")``
or through the ``ljmd_b200`` alias module at the repository root.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LJMD_LIB_PATH") or os.path.join(_HERE, "lib", "libljmd.so")  # override: A/B builds

METHOD_AUTO, METHOD_ALLPAIRS, METHOD_CELLS, METHOD_NLIST = 0, 1, 2, 3
FLAG_TRACK_UNWRAPPED = 1
FLAG_NO_SORT = 2

ERR_NAMES = {
    1: "LJMD_ERR_N", 2: "LJMD_ERR_BOX", 3: "LJMD_ERR_VOLUME", 4: "LJMD_ERR_RC",
    5: "LJMD_ERR_RC_SQUARE", 6: "LJMD_ERR_NULL", 7: "LJMD_ERR_RC_HALF_BOX", 8: "LJMD_ERR_DT",
    9: "LJMD_ERR_ARG", 10: "LJMD_ERR_UNSUPPORTED", 100: "LJMD_ERR_CUDA", 101: "LJMD_ERR_NOMEM",
}


class LjmdError(RuntimeError):
    """Raised where the reference would ``stop``; ``code`` is the C-ABI status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{ERR_NAMES.get(code, code)}] {message}")
        self.code = code
        self.message = message


class SimParams(C.Structure):
    """type(sim_params), base/md_types.f90:27-50 (same field names, 96 bytes)."""

    _fields_ = [
        ("n", C.c_int32),
        ("num_cells", C.c_int32),
        ("box_length", C.c_double),
        ("inv_box_length", C.c_double),
        ("volume", C.c_double),
        ("density", C.c_double),
        ("dt", C.c_double),
        ("dt_half", C.c_double),
        ("dt_square_half", C.c_double),
        ("rc", C.c_double),
        ("rc_square", C.c_double),
        ("length_scale", C.c_double),
        ("inv_length_scale", C.c_double),
    ]


class _Options(C.Structure):
    _fields_ = [("method", C.c_int32), ("device", C.c_int32), ("nreplicas", C.c_int32),
                ("flags", C.c_int32)]


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)
_lib: Optional[C.CDLL] = None

# Every symbol include/ljmd.h declares (tests check that the library exports all of them).
ABI_SYMBOLS = [
    "ljmd_version", "ljmd_last_error", "ljmd_device_count", "ljmd_params_init",
    "ljmd_compute_lj", "ljmd_verlet_step", "ljmd_compat_release", "ljmd_ctx_create",
    "ljmd_ctx_destroy", "ljmd_ctx_upload", "ljmd_ctx_download", "ljmd_ctx_download_unwrapped",
    "ljmd_ctx_compute_forces", "ljmd_ctx_step", "ljmd_ctx_step_series",
    "ljmd_ctx_kinetic_energy", "ljmd_ctx_method", "ljmd_ctx_cells_per_edge",
    "ljmd_ctx_launch_count", "ljmd_ctx_last_step_ms", "ljmd_ctx_force_timing", "ljmd_ctx_download_cells",
    "ljmd_ctx_pair_counts", "ljmd_measure_fp64_peak", "ljmd_measure_fp64_peak_sustained",
    "ljmd_ctx_set_stream", "ljmd_ctx_device_ptrs", "ljmd_ctx_set_shard", "ljmd_ctx_shard_begin",
    "ljmd_ctx_shard_step", "ljmd_ctx_shard_end", "ljmd_tail_corrections",
    "ljmd_ctx_nlist_info", "ljmd_ctx_download_nlist",
    "ljmd_ctx_slab_export", "ljmd_ctx_slab_attach", "ljmd_ctx_slab_bounds", "ljmd_ctx_slab_pack",
    "ljmd_ctx_slab_unpack", "ljmd_ctx_slab_drift", "ljmd_ctx_slab_rebuild", "ljmd_ctx_slab_run",
    "ljmd_ctx_slab_poison", "ljmd_ctx_slab_local_info", "ljmd_ctx_slab_halo", "ljmd_ctx_slab_rebuild_local",
    "ljmd_ctx_slab_prepare",
    "ljmd_ctx_series_arm", "ljmd_ctx_series_fetch",
    "ljmd_ctx_nlist_momentum", "ljmd_rdf_histogram", "ljmd_time_correlation",
]
# Tuning / debug exports declared in csrc/ljmd_debug.h, not in the product header.
DEBUG_SYMBOLS = ["ljmd_bench_chunk_cycles", "ljmd_ctx_debug_warp_times", "ljmd_ctx_debug_checksum",
                 "ljmd_ctx_slab_publish", "ljmd_ctx_debug_slab_timeline"]
MAX_RANKS = 8


def load_library() -> C.CDLL:
    """Load libljmd.so; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  This package has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.ljmd_last_error.restype = C.c_char_p
        L.ljmd_ctx_launch_count.restype = C.c_int64
        L.ljmd_ctx_launch_count.argtypes = [C.c_void_p]
        L.ljmd_ctx_method.argtypes = [C.c_void_p]
        L.ljmd_ctx_cells_per_edge.argtypes = [C.c_void_p]
        L.ljmd_ctx_destroy.argtypes = [C.c_void_p]
        L.ljmd_params_init.argtypes = [C.POINTER(SimParams), C.c_int32, C.c_double, C.c_double,
                                       C.c_double, C.c_int32]
        # the drop-in pair takes raw addresses (cached per SimState): the ctypes conversions of nine
        # numpy arrays used to cost ~30 us per call, a third of the whole host-array call at N=4000
        L.ljmd_compute_lj.argtypes = [C.POINTER(SimParams)] + [C.c_void_p] * 6 + [_DP] * 3
        L.ljmd_verlet_step.argtypes = [C.POINTER(SimParams)] + [C.c_void_p] * 9 + [_DP] * 4
        _lib = L
    return _lib


def _chk(code: int) -> None:
    if code != 0:
        raise LjmdError(code, load_library().ljmd_last_error().decode())


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    if a.dtype != np.float64 or not a.flags.c_contiguous:
        raise TypeError("state arrays must be C-contiguous float64")
    return a.ctypes.data_as(_DP)


def device_count() -> int:
    return int(load_library().ljmd_device_count())


# ---------------------------------------------------------------------------------------------
# md_types mirror
# ---------------------------------------------------------------------------------------------
def init_params(n: int, box_length: float, dt: float, rc: float, num_cells: int = 0) -> SimParams:
    """init_params + compute_derived_params, base/md_types.f90:105-169."""
    p = SimParams()
    _chk(load_library().ljmd_params_init(C.byref(p), n, box_length, dt, rc, num_cells))
    return p


class SimState:
    """type(sim_state), base/md_types.f90:56-60: nine SoA float64 arrays, zero-initialised
    (init_state, :175-182)."""

    NAMES = ("rx", "ry", "rz", "vx", "vy", "vz", "ax", "ay", "az")

    def __init__(self, n: int, nreplicas: int = 1):
        shape = (n,) if nreplicas == 1 else (nreplicas, n)
        for name in self.NAMES:
            setattr(self, name, np.zeros(shape, dtype=np.float64))

    def arrays(self):
        return [getattr(self, k) for k in self.NAMES]

    def _addresses(self):
        """Device-call addresses of the nine arrays, validated once and cached until an array
        object is replaced (the drop-in calls update the arrays in place, like the reference)."""
        arrs = (self.rx, self.ry, self.rz, self.vx, self.vy, self.vz, self.ax, self.ay, self.az)
        cache = self.__dict__.get("_addr_cache")
        if cache is not None and all(a is b for a, b in zip(arrs, cache[0])):
            return cache[1]
        for a in arrs:
            if a.dtype != np.float64 or not a.flags.c_contiguous:
                raise TypeError("state arrays must be C-contiguous float64")
        addrs = tuple(a.ctypes.data for a in arrs)
        self.__dict__["_addr_cache"] = (arrs, addrs)
        return addrs

    def _addresses_for(self, n: int):
        """_addresses() after checking that every array holds exactly n values: the C side copies
        params%n doubles from / to each address (the reference's arrays are allocated with exactly
        params%n elements, md_types.f90:198-200)."""
        for k in self.NAMES if n > 0 else ():  # n <= 0 is the C side's error (the reference's own first guard)
            a = getattr(self, k)
            if a is None or a.size != n:
                raise ValueError(f"state.{k} holds {0 if a is None else a.size} values, params%n is {n}")
        return self._addresses()

    def copy(self) -> "SimState":
        s = SimState.__new__(SimState)
        for k in self.NAMES:
            setattr(s, k, getattr(self, k).copy())
        return s

    @classmethod
    def from_arrays(cls, arrs: Sequence[np.ndarray]) -> "SimState":
        s = cls.__new__(cls)
        for k, a in zip(cls.NAMES, arrs):
            setattr(s, k, np.ascontiguousarray(a, dtype=np.float64))
        return s


def init_state(params: SimParams, nreplicas: int = 1) -> SimState:
    if params.n <= 0:
        raise LjmdError(1, "allocate_state(): params%n must be > 0.")
    return SimState(params.n, nreplicas)


# ---------------------------------------------------------------------------------------------
# Stateless drop-in pair (host arrays in, host arrays out; H2D/D2H inside the call)
# ---------------------------------------------------------------------------------------------
def compute_lj_potential_energy(params: SimParams, state: SimState):
    """physics/lj_potential_energy.f90:46.  Overwrites state.ax/ay/az; returns
    (epot, d_epot, dd_epot)."""
    e, d, dd = C.c_double(), C.c_double(), C.c_double()
    ad = state._addresses_for(params.n)
    _chk(load_library().ljmd_compute_lj(params, ad[0], ad[1], ad[2], ad[6], ad[7], ad[8], e, d, dd))
    return e.value, d.value, dd.value


def verlet_step(params: SimParams, state: SimState):
    """physics/verlet.f90:41.  Updates all nine arrays in place; returns
    (epot, ekin, d_epot, dd_epot)."""
    e, k, d, dd = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    _chk(load_library().ljmd_verlet_step(params, *state._addresses_for(params.n), e, k, d, dd))
    return e.value, k.value, d.value, dd.value


def compat_release() -> None:
    _chk(load_library().ljmd_compat_release())


def bench_chunk_cycles(ri: int, rj: int, dense: bool, warps_per_cta: int) -> float:
    c = C.c_double()
    _chk(load_library().ljmd_bench_chunk_cycles(ri, rj, 1 if dense else 0, warps_per_cta, C.byref(c)))
    return c.value


def tail_corrections(params: SimParams):
    """lj_potential_energy.f90:205-223: (tail_epot, tail_d_epot, tail_dd_epot)."""
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    _chk(load_library().ljmd_tail_corrections(C.byref(params), C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def measure_fp64_peak(device: int = -1) -> float:
    t = C.c_double()
    _chk(load_library().ljmd_measure_fp64_peak(device, C.byref(t)))
    return t.value


def measure_fp64_peak_sustained(device: int = -1, seconds: float = 1.0) -> float:
    """DFMA throughput sustained under `seconds` of back-to-back load (clocks settled), TFLOP/s."""
    t = C.c_double()
    _chk(load_library().ljmd_measure_fp64_peak_sustained(device, C.c_double(seconds), C.byref(t)))
    return t.value


# ---------------------------------------------------------------------------------------------
# Device-resident context
# ---------------------------------------------------------------------------------------------
class Context:
    """State resident in HBM; ``step(n)`` runs n velocity-Verlet steps with no host round trip."""

    def __init__(self, params: SimParams, method: int = METHOD_AUTO, device: int = -1,
                 nreplicas: int = 1, flags: int = 0):
        self._L = load_library()
        self.params = params
        self.nreplicas = nreplicas
        self._h = C.c_void_p()
        o = _Options(method, device, nreplicas, flags)
        _chk(self._L.ljmd_ctx_create(C.byref(params), C.byref(o), C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._L.ljmd_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shape(self):
        n = self.params.n
        return (n,) if self.nreplicas == 1 else (self.nreplicas, n)

    def upload(self, state: Optional[SimState] = None, **arrays) -> None:
        if state is not None:
            arrays = {k: getattr(state, k) for k in SimState.NAMES}
        ptrs = []
        for k in SimState.NAMES:
            a = arrays.get(k)
            if a is not None and a.size != self.params.n * self.nreplicas:
                raise ValueError(f"{k}: expected {self.params.n * self.nreplicas} values")
            ptrs.append(_p(a))
        _chk(self._L.ljmd_ctx_upload(self._h, *ptrs))

    def download(self, names: Sequence[str] = SimState.NAMES) -> SimState:
        s = SimState.__new__(SimState)
        ptrs = []
        for k in SimState.NAMES:
            if k in names:
                a = np.empty(self.shape, dtype=np.float64)
                setattr(s, k, a)
                ptrs.append(_p(a))
            else:
                setattr(s, k, None)
                ptrs.append(None)
        _chk(self._L.ljmd_ctx_download(self._h, *ptrs))
        return s

    def download_unwrapped(self):
        ru = [np.empty(self.shape, dtype=np.float64) for _ in range(3)]
        _chk(self._L.ljmd_ctx_download_unwrapped(self._h, *[_p(a) for a in ru]))
        return ru

    def compute_forces(self) -> np.ndarray:
        """compute_lj_potential_energy on the resident state: (epot, d_epot, dd_epot)[replica]."""
        obs = np.zeros((self.nreplicas, 3))
        _chk(self._L.ljmd_ctx_compute_forces(self._h, _p(obs)))
        return obs[0] if self.nreplicas == 1 else obs

    def step(self, nsteps: int = 1) -> np.ndarray:
        """nsteps x verlet_step; returns (epot, ekin, d_epot, dd_epot)[replica] of the last."""
        obs = np.zeros((self.nreplicas, 4))
        _chk(self._L.ljmd_ctx_step(self._h, int(nsteps), _p(obs)))
        return obs[0] if self.nreplicas == 1 else obs

    def step_series(self, nsteps: int) -> np.ndarray:
        ser = np.zeros((nsteps, self.nreplicas, 4))
        _chk(self._L.ljmd_ctx_step_series(self._h, int(nsteps), _p(ser)))
        return ser[:, 0, :] if self.nreplicas == 1 else ser

    def series_arm(self, cap: int) -> None:
        _chk(self._L.ljmd_ctx_series_arm(self._h, int(cap)))

    def series_fetch(self, nsteps: int) -> np.ndarray:
        ser = np.zeros((nsteps, self.nreplicas, 4))
        _chk(self._L.ljmd_ctx_series_fetch(self._h, int(nsteps), _p(ser)))
        return ser[:, 0, :] if self.nreplicas == 1 else ser

    def nlist_momentum(self) -> np.ndarray:
        """Sum of v over the atoms the last step launch handled (self-check: conserved)."""
        p3 = np.zeros(3)
        _chk(self._L.ljmd_ctx_nlist_momentum(self._h, _p(p3)))
        return p3

    def slab_publish(self) -> None:
        _chk(self._L.ljmd_ctx_slab_publish(self._h))

    def slab_poison(self) -> None:
        """Tell the peer ranks that this rank cannot go on (their step kernels halt instead of waiting)."""
        _chk(self._L.ljmd_ctx_slab_poison(self._h))

    def kinetic_energy(self) -> np.ndarray:
        k = np.zeros(self.nreplicas)
        _chk(self._L.ljmd_ctx_kinetic_energy(self._h, _p(k)))
        return k[0] if self.nreplicas == 1 else k

    @property
    def method(self) -> int:
        return int(self._L.ljmd_ctx_method(self._h))

    @property
    def cells_per_edge(self) -> int:
        return int(self._L.ljmd_ctx_cells_per_edge(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.ljmd_ctx_launch_count(self._h))

    @property
    def last_step_ms(self) -> float:
        """Device time (CUDA events on the context stream) of the last step() call."""
        ms = C.c_double()
        _chk(self._L.ljmd_ctx_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    def force_timing(self, enable: bool):
        """Enable/disable per-launch CUDA-event timing of the force kernel; returns the
        (total_ms, n_evals) accumulated since the previous call."""
        ms, ne = C.c_double(), C.c_int64()
        _chk(self._L.ljmd_ctx_force_timing(self._h, 1 if enable else 0, C.byref(ms), C.byref(ne)))
        return ms.value, ne.value

    def pair_counts(self):
        a, b = C.c_int64(), C.c_int64()
        _chk(self._L.ljmd_ctx_pair_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- multi-GPU plumbing ------------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        """Run this context on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        _chk(self._L.ljmd_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def device_ptrs(self):
        """(pos0, pos1, vel, acc, obs) device addresses as ints."""
        ptrs = [C.c_void_p() for _ in range(5)]
        _chk(self._L.ljmd_ctx_device_ptrs(self._h, *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)

    def set_shard(self, i_begin: int, i_end: int) -> None:
        _chk(self._L.ljmd_ctx_set_shard(self._h, int(i_begin), int(i_end)))

    def shard_begin(self) -> int:
        b = C.c_int32()
        _chk(self._L.ljmd_ctx_shard_begin(self._h, C.byref(b)))
        return b.value

    def shard_step(self) -> int:
        b = C.c_int32()
        _chk(self._L.ljmd_ctx_shard_step(self._h, C.byref(b)))
        return b.value

    def shard_end(self) -> None:
        _chk(self._L.ljmd_ctx_shard_end(self._h))

    def nlist_info(self):
        """(skin, cells_per_edge, capacity, rebuilds, entries) of the neighbour-list path."""
        skin, nc, cap = C.c_double(), C.c_int32(), C.c_int32()
        reb, ent = C.c_int64(), C.c_int64()
        _chk(self._L.ljmd_ctx_nlist_info(self._h, C.byref(skin), C.byref(nc), C.byref(cap),
                                         C.byref(reb), C.byref(ent)))
        return {"skin": skin.value, "cells_per_edge": nc.value, "capacity": cap.value,
                "rebuilds": reb.value, "entries": ent.value}

    def download_nlist(self):
        """Neighbour list of the last build in original numbering: (counts[n], nbr[n, cap])."""
        cap = self.nlist_info()["capacity"]
        n = self.params.n
        counts = np.zeros(n, dtype=np.int32)
        nbr = np.full((n, cap), -1, dtype=np.int32)
        _chk(self._L.ljmd_ctx_download_nlist(self._h, counts.ctypes.data_as(_IP),
                                             nbr.ctypes.data_as(_IP), cap))
        return counts, nbr

    # ---- slab sharding of the neighbour-list path (see include/ljmd.h) ------------------------
    def slab_export(self):
        """(128-byte IPC handle blob, (p4_ptr, flags_ptr))"""
        h = (C.c_ubyte * 128)()
        raw = (C.c_uint64 * 2)()
        _chk(self._L.ljmd_ctx_slab_export(self._h, h, raw))
        return bytes(h), (int(raw[0]), int(raw[1]))

    def slab_attach(self, rank: int, world: int, handles=None, raw_ptrs=None) -> None:
        hb = None
        if handles is not None:
            blob = b"".join(handles)
            assert len(blob) == 128 * world
            hb = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        rp = None
        if raw_ptrs is not None:
            flat = [int(x) for pair in raw_ptrs for x in pair]
            rp = (C.c_uint64 * len(flat))(*flat)
        _chk(self._L.ljmd_ctx_slab_attach(self._h, int(rank), int(world), hb, rp))
        self._slab_world = world

    def slab_bounds(self) -> np.ndarray:
        b = np.zeros(self._slab_world + 1, dtype=np.int32)
        _chk(self._L.ljmd_ctx_slab_bounds(self._h, b.ctypes.data_as(_IP)))
        return b

    def slab_narr(self) -> int:
        k = C.c_int32()
        _chk(self._L.ljmd_ctx_slab_pack(self._h, None, 0, C.byref(k)))
        return k.value

    def slab_pack(self, dev_ptr: int, cmax: int) -> None:
        _chk(self._L.ljmd_ctx_slab_pack(self._h, C.c_void_p(dev_ptr), int(cmax), None))

    def slab_unpack(self, dev_ptr: int, cmax: int) -> None:
        _chk(self._L.ljmd_ctx_slab_unpack(self._h, C.c_void_p(dev_ptr), int(cmax)))

    def slab_drift(self) -> None:
        _chk(self._L.ljmd_ctx_slab_drift(self._h))

    def slab_rebuild(self, force: bool = False) -> None:
        _chk(self._L.ljmd_ctx_slab_rebuild(self._h, 1 if force else 0))

    def slab_local_info(self):
        """(device address of the int32 cell histogram, ncells, halo_cap, halo_narr) of a slab-local rebuild."""
        ptr, nc, cap, na = C.c_uint64(), C.c_int32(), C.c_int32(), C.c_int32()
        _chk(self._L.ljmd_ctx_slab_local_info(self._h, C.byref(ptr), C.byref(nc), C.byref(cap), C.byref(na)))
        return int(ptr.value), nc.value, cap.value, na.value

    def slab_halo(self, dev_ptr: int, cap: int, side: int, unpack: bool) -> None:
        _chk(self._L.ljmd_ctx_slab_halo(self._h, C.c_void_p(dev_ptr), int(cap), int(side), 1 if unpack else 0))

    def slab_prepare(self, max_steps: int) -> None:
        _chk(self._L.ljmd_ctx_slab_prepare(self._h, int(max_steps)))

    def slab_rebuild_local(self, phase: int) -> None:
        _chk(self._L.ljmd_ctx_slab_rebuild_local(self._h, int(phase)))

    def debug_checksum(self):
        out = (C.c_uint64 * 4)()
        _chk(self._L.ljmd_ctx_debug_checksum(self._h, out))
        return [int(x) for x in out]

    def slab_run(self, max_steps: int):
        """(steps_done, halted, obs_share[4])"""
        done, halted = C.c_int32(), C.c_int32()
        obs = np.zeros(4)
        _chk(self._L.ljmd_ctx_slab_run(self._h, int(max_steps), C.byref(done), C.byref(halted), _p(obs)))
        return done.value, bool(halted.value), obs

    def download_cells(self):
        nc = self.cells_per_edge
        n = self.params.n
        cell = np.zeros(n, dtype=np.int32)
        start = np.zeros(nc ** 3 + 1, dtype=np.int32)
        order = np.zeros(n, dtype=np.int32)
        _chk(self._L.ljmd_ctx_download_cells(self._h, cell.ctypes.data_as(_IP),
                                             start.ctypes.data_as(_IP),
                                             order.ctypes.data_as(_IP)))
        return cell, start, order
