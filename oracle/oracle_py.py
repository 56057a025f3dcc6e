"""ctypes binding of the CPU oracle (oracle/ljmd_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED (see oracle/ljmd_oracle.h): the reference has no golden vectors and cannot be
compiled in this image.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}


class OrcParams(C.Structure):
    """Mirror of type(sim_params), base/md_types.f90:27-50 (96 bytes)."""

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


_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)


def build(force: bool = False) -> None:
    out = os.path.join(_HERE, "_build", "libljmd_oracle.so")
    src = os.path.join(_HERE, "ljmd_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib(variant: str = "O2") -> C.CDLL:
    if variant not in _LIBS:
        build()
        name = "libljmd_oracle.so" if variant == "O2" else "libljmd_oracle_O3.so"
        L = C.CDLL(os.path.join(_HERE, "_build", name))
        L.orc_minimum_image.restype = C.c_double
        L.orc_minimum_image.argtypes = [C.c_double] * 3
        L.orc_ekin_driver.restype = C.c_double
        L.orc_random_uniform.restype = C.c_double
        L.orc_count_pairs_in_cutoff.restype = C.c_int64
        L.orc_cells_per_edge.restype = C.c_int32
        L.orc_cells_per_edge_range.restype = C.c_int32
        _LIBS[variant] = L
    return _LIBS[variant]


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_DP)


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_IP)


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error code {code}")
        self.code = code


def _chk(code: int) -> None:
    if code != 0:
        raise OracleError(code)


def params_init(n: int, box_length: float, dt: float, rc: float, num_cells: int = 0) -> OrcParams:
    p = OrcParams()
    _chk(lib().orc_params_init(C.byref(p), n, C.c_double(box_length), C.c_double(dt),
                               C.c_double(rc), num_cells))
    return p


def minimum_image(dx: float, L: float, invL: float) -> float:
    return lib().orc_minimum_image(dx, L, invL)


def wrap_positions(rx, ry, rz, L: float) -> None:
    _chk(lib().orc_wrap_positions(len(rx), _p(rx), _p(ry), _p(rz), C.c_double(L)))


def compute_lj(p: OrcParams, rx, ry, rz, variant: str = "O2"):
    n = p.n
    ax, ay, az = (np.zeros(n) for _ in range(3))
    e, d, dd = C.c_double(), C.c_double(), C.c_double()
    _chk(lib(variant).orc_compute_lj(C.byref(p), _p(rx), _p(ry), _p(rz), _p(ax), _p(ay), _p(az),
                                     C.byref(e), C.byref(d), C.byref(dd)))
    return ax, ay, az, e.value, d.value, dd.value


def compute_lj_cells(p: OrcParams, rx, ry, rz):
    n = p.n
    ax, ay, az = (np.zeros(n) for _ in range(3))
    e, d, dd = C.c_double(), C.c_double(), C.c_double()
    nc, ni = C.c_int64(), C.c_int64()
    _chk(lib().orc_compute_lj_cells(C.byref(p), _p(rx), _p(ry), _p(rz), _p(ax), _p(ay), _p(az),
                                    C.byref(e), C.byref(d), C.byref(dd), C.byref(nc),
                                    C.byref(ni)))
    return ax, ay, az, e.value, d.value, dd.value, nc.value, ni.value


def tail_corrections(p: OrcParams):
    e, d, dd = C.c_double(), C.c_double(), C.c_double()
    lib().orc_tail_corrections(C.byref(p), C.byref(e), C.byref(d), C.byref(dd))
    return e.value, d.value, dd.value


def count_pairs_in_cutoff(p: OrcParams, rx, ry, rz) -> int:
    return lib().orc_count_pairs_in_cutoff(C.byref(p), _p(rx), _p(ry), _p(rz))


def verlet_step(p: OrcParams, state, variant: str = "O2", cells: bool = False):
    """state: list/tuple of 9 float64 arrays (rx..az), updated in place. Returns 4 observables."""
    e, k, d, dd = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    fn = lib(variant).orc_verlet_step_cells if cells else lib(variant).orc_verlet_step
    _chk(fn(C.byref(p), *[_p(a) for a in state], C.byref(e), C.byref(k), C.byref(d),
            C.byref(dd)))
    return e.value, k.value, d.value, dd.value


def run_steps(p: OrcParams, nsteps: int, state, ru=None, want_series: bool = False,
              variant: str = "O2"):
    obs = np.zeros(4)
    series = np.zeros((nsteps, 4)) if want_series else None
    ru_ptrs = [_p(a) for a in ru] if ru is not None else [None, None, None]
    _chk(lib(variant).orc_run_steps(C.byref(p), nsteps, *[_p(a) for a in state], *ru_ptrs,
                                    _p(obs), _p(series) if want_series else None))
    return (obs, series) if want_series else obs


def ekin_driver(vx, vy, vz) -> float:
    return lib().orc_ekin_driver(len(vx), _p(vx), _p(vy), _p(vz))


class Rng:
    """base/random_numbers.f90:48-116 with the SAVEd state made explicit."""

    class _S(C.Structure):
        _fields_ = [("ma", C.c_double * 56), ("inext", C.c_int32), ("inextp", C.c_int32),
                    ("iff", C.c_int32)]

    def __init__(self, seed: int = -12345):
        self.s = Rng._S()
        lib().orc_rng_reset(C.byref(self.s))
        self.seed = C.c_int32(seed)

    def uniform(self) -> float:
        return lib().orc_random_uniform(C.byref(self.s), C.byref(self.seed))


def fcc_lattice(p: OrcParams):
    n = p.n
    rx, ry, rz = (np.zeros(n) for _ in range(3))
    _chk(lib().orc_build_fcc_lattice(C.byref(p), _p(rx), _p(ry), _p(rz)))
    return rx, ry, rz


def initial_config(p: OrcParams, seed: int, target_total_energy: float = 0.0,
                   temp_target: float = 0.0, cells: bool = False):
    """md_initial_config_program.f90:73-105 (no warm-up). Returns (state[9], epot, ekin).
    temp_target > 0 replaces the energy target by U(0) + 1.5*N*T; cells=True evaluates the
    forces with the cell-list oracle (large N)."""
    n = p.n
    st = [np.zeros(n) for _ in range(9)]
    e, k = C.c_double(), C.c_double()
    _chk(lib().orc_initial_config_ex(C.byref(p), seed, C.c_double(target_total_energy),
                                     C.c_double(temp_target), 1 if cells else 0,
                                     *[_p(a) for a in st], C.byref(e), C.byref(k)))
    return st, e.value, k.value


def cells_per_edge(p: OrcParams) -> int:
    return lib().orc_cells_per_edge(C.byref(p))


def cell_assign(p: OrcParams, rx, ry, rz) -> np.ndarray:
    cell = np.zeros(p.n, dtype=np.int32)
    _chk(lib().orc_cell_assign(C.byref(p), _p(rx), _p(ry), _p(rz), _ip(cell)))
    return cell


def cell_sort(p: OrcParams, cell: np.ndarray):
    nc = cells_per_edge(p)
    start = np.zeros(nc ** 3 + 1, dtype=np.int32)
    order = np.zeros(p.n, dtype=np.int32)
    _chk(lib().orc_cell_sort(C.byref(p), _ip(cell), _ip(start), _ip(order)))
    return start, order


def cells_per_edge_range(p: OrcParams, rng: float) -> int:
    return lib().orc_cells_per_edge_range(C.byref(p), C.c_double(rng))


def cell_assign_range(p: OrcParams, rng: float, rx, ry, rz) -> np.ndarray:
    cell = np.zeros(p.n, dtype=np.int32)
    _chk(lib().orc_cell_assign_range(C.byref(p), C.c_double(rng), _p(rx), _p(ry), _p(rz), _ip(cell)))
    return cell


def cell_sort_range(p: OrcParams, rng: float, cell: np.ndarray):
    nc = cells_per_edge_range(p, rng)
    start = np.zeros(nc ** 3 + 1, dtype=np.int32)
    order = np.zeros(p.n, dtype=np.int32)
    _chk(lib().orc_cell_sort_range(C.byref(p), C.c_double(rng), _ip(cell), _ip(start), _ip(order)))
    return start, order


def neighbor_list(p: OrcParams, rng: float, rx, ry, rz, cap: int):
    """Full Verlet list within `rng` (original numbering, ascending): (counts[n], nbr[n, cap])."""
    counts = np.zeros(p.n, dtype=np.int32)
    nbr = np.full((p.n, cap), -1, dtype=np.int32)
    _chk(lib().orc_neighbor_list(C.byref(p), C.c_double(rng), _p(rx), _p(ry), _p(rz), cap,
                                 _ip(counts), _ip(nbr)))
    return counts, nbr


class OrcMeans(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_samples", C.c_int32), ("volume", C.c_double),
                ("sum", C.c_double * 11), ("sum_sq", C.c_double * 11)]


class Means:
    """stats/md_means.f90:192-366 (virial always present)."""

    def __init__(self, n: int, volume: float):
        self.m = OrcMeans()
        lib().orc_means_init(C.byref(self.m), n, C.c_double(volume))

    def add(self, epot, ekin, virial, d_epot, dd_epot):
        t, p = C.c_double(), C.c_double()
        _chk(lib().orc_means_add(C.byref(self.m), C.c_double(epot), C.c_double(ekin), C.c_double(virial),
                                 C.c_double(d_epot), C.c_double(dd_epot), C.byref(t), C.byref(p)))
        return t.value, p.value

    def get(self):
        mean, std = np.zeros(11), np.zeros(11)
        _chk(lib().orc_means_get(C.byref(self.m), _p(mean), _p(std)))
        return mean, std


def thermo_coefs(p: OrcParams, ekin_mean, press_mean, ekinInv_mean, d_mean, dd_mean, d_kinv_mean, d2_kinv_mean):
    out = np.zeros(16)
    _chk(lib().orc_thermo_coefs(C.byref(p), *[C.c_double(x) for x in (ekin_mean, press_mean, ekinInv_mean, d_mean,
                                                                      dd_mean, d_kinv_mean, d2_kinv_mean)], _p(out)))
    return out
