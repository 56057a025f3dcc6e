"""Host-side mirror of the reference's initial-configuration program and of its two binary
hand-off files: what a non-Fortran host needs to START the GPU hot
path from exactly the state the reference would start from, and to exchange states with the
reference's own programs.  Pure numpy / Python: O(N) set-up work, once per run, not on the hot path
(the force evaluation it needs for the energy rescale is the GPU library's).

Reference (file:line relative to /root/reference/scripts/):
  md_initial_config_program.f90:132-187  build_fcc_lattice
  md_initial_config_program.f90:196-209  assign_random_velocities (vx, vy, vz per particle)
  md_initial_config_program.f90:218-234  remove_center_of_mass_velocity
  md_initial_config_program.f90:244-266  rescale_velocities_to_target_energy
  md_initial_config_program.f90:73-116   the program's sequence (lattice, velocities, COM, force,
                                         rescale, force, warm-up steps)
  md_initial_config_program.f90:275-290  write_rv_init;  md_simulation_program.f90:573-588 read_rv_init
  md_simulation_program.f90:257,384-387  rva.dat header + 4 records per sample
  base/random_numbers.f90:48-116         random_uniform (Knuth subtractive generator in real64)

Arithmetic notes: Fortran `sum()` at the shipped -O2 is an in-order sum, so `_seqsum` accumulates
left to right (numpy's pairwise `sum` would differ in the last bits); the RNG state consists of
integers below 4e6 held in float64, so the stream is exactly reproducible.
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np

MBIG = 4.0e6
MSEED = 1618033.0
FAC = 1.0 / MBIG


class RandomUniform:
    """base/random_numbers.f90:48-116 with the SAVEd state made explicit.  `seed <= 0` (or the
    first call) initialises the table; afterwards the reference sets seed = 1."""

    def __init__(self, seed: int = -12345):
        self.seed = int(seed)
        self.iff = 0
        self.ma = [0.0] * 56  # 1-based like the Fortran array
        self.inext = 0
        self.inextp = 0

    def _init(self) -> None:
        self.iff = 1
        mj = abs(MSEED - abs(float(self.seed)))      # :73
        mj = mj % MBIG                               # :74
        ma = self.ma
        ma[55] = mj
        mk = 1.0
        for i in range(1, 55):                       # :78-84
            ii = (21 * i) % 55
            ma[ii] = mk
            mk = mj - mk
            if mk < 0.0:
                mk += MBIG
            mj = ma[ii]
        for _ in range(4):                           # :86-91
            for i in range(1, 56):
                ma[i] = ma[i] - ma[1 + (i + 30) % 55]
                if ma[i] < 0.0:
                    ma[i] += MBIG
        self.inext, self.inextp, self.seed = 0, 31, 1

    def __call__(self) -> float:
        if self.seed <= 0 or self.iff == 0:
            self._init()
        self.inext += 1
        if self.inext == 56:
            self.inext = 1
        self.inextp += 1
        if self.inextp == 56:
            self.inextp = 1
        mj = self.ma[self.inext] - self.ma[self.inextp]
        if mj < 0.0:
            mj += MBIG
        self.ma[self.inext] = mj
        return mj * FAC

    def draw(self, count: int) -> np.ndarray:
        return np.array([self() for _ in range(count)], dtype=np.float64)


def _seqsum(x: np.ndarray) -> float:
    """In-order sum, like gfortran's `sum` intrinsic without reassociation."""
    return float(np.cumsum(np.asarray(x, dtype=np.float64))[-1]) if len(x) else 0.0


def build_fcc_lattice(num_cells: int, box_length: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """md_initial_config_program.f90:132-187; a = L / num_cells (:73), loop order ix, iy, iz,
    basis (0,0,0), (0,1/2,1/2), (1/2,0,1/2), (1/2,1/2,0)."""
    k = int(num_cells)
    a = box_length / float(k)
    idx = np.arange(k, dtype=np.float64) * a          # dble(ix) * a
    x0, y0, z0 = np.meshgrid(idx, idx, idx, indexing="ij")
    x0, y0, z0 = x0.reshape(-1, 1), y0.reshape(-1, 1), z0.reshape(-1, 1)
    h = 0.5 * a                                        # 0.5d0*a, added to the cell origin as in :166-181
    rx = np.concatenate([x0, x0, x0 + h, x0 + h], axis=1)
    ry = np.concatenate([y0, y0 + h, y0, y0 + h], axis=1)
    rz = np.concatenate([z0, z0 + h, z0 + h, z0], axis=1)
    return (np.ascontiguousarray(rx.reshape(-1)), np.ascontiguousarray(ry.reshape(-1)),
            np.ascontiguousarray(rz.reshape(-1)))


def assign_random_velocities(n: int, rng: RandomUniform) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """md_initial_config_program.f90:196-209: draws in the order vx(i), vy(i), vz(i)."""
    d = rng.draw(3 * n).reshape(n, 3) - 0.5
    return (np.ascontiguousarray(d[:, 0]), np.ascontiguousarray(d[:, 1]), np.ascontiguousarray(d[:, 2]))


def remove_center_of_mass_velocity(vx, vy, vz) -> None:
    """md_initial_config_program.f90:218-234 (in place)."""
    n = float(len(vx))
    for v in (vx, vy, vz):
        v -= _seqsum(v) / n


def kinetic_energy(vx, vy, vz) -> float:
    """0.5*sum(vx*vx + vy*vy + vz*vz), md_initial_config_program.f90:254 / md_simulation_program.f90:238-240."""
    return 0.5 * _seqsum((vx * vx + vy * vy) + vz * vz)


def rescale_velocities_to_target_energy(vx, vy, vz, target_energy: float, epot: float) -> float:
    """md_initial_config_program.f90:244-266 (in place); returns the scale factor.  Raises where the
    reference stops."""
    ekin_old = kinetic_energy(vx, vy, vz)
    ekin_new = target_energy - epot
    if ekin_new <= 0.0:
        raise ValueError("rescale_velocities_to_target_energy(): target energy too low (zero or negative kinetic).")
    if ekin_old <= 0.0:
        raise ValueError("rescale_velocities_to_target_energy(): ekin_old <= 0 (cannot rescale).")
    scale = float(np.sqrt(ekin_new / ekin_old))
    vx *= scale
    vy *= scale
    vz *= scale
    return scale


def initial_config(lj, params, seed: int = -12345, target_total_energy: Optional[float] = None,
                   warmup_steps: int = 0):
    """The sequence of md_initial_config_program.f90:73-116 on the GPU library `lj` (the package):
    lattice, velocities (seed), COM removal, force, rescale to E_target, force, `warmup_steps`
    velocity-Verlet steps.  Returns (SimState, epot, ekin).  `params.num_cells` must be set
    (N = 4 * num_cells**3, base/read_input_files.f90:168)."""
    n = params.n
    if 4 * params.num_cells ** 3 != n:
        raise ValueError("build_fcc_lattice(): unexpected particle count.")
    st = lj.init_state(params)
    st.rx, st.ry, st.rz = build_fcc_lattice(params.num_cells, params.box_length)
    rng = RandomUniform(seed)
    st.vx, st.vy, st.vz = assign_random_velocities(n, rng)
    remove_center_of_mass_velocity(st.vx, st.vy, st.vz)
    epot, _, _ = lj.compute_lj_potential_energy(params, st)          # :91
    if target_total_energy is not None:
        rescale_velocities_to_target_energy(st.vx, st.vy, st.vz, target_total_energy, epot)  # :101
    epot, _, _ = lj.compute_lj_potential_energy(params, st)          # :104
    ekin = kinetic_energy(st.vx, st.vy, st.vz)
    if warmup_steps > 0:                                             # :113-116
        with lj.Context(params) as ctx:
            ctx.upload(st)
            obs = ctx.step(warmup_steps)
            st = ctx.download()
        epot, ekin = float(obs[0]), float(obs[1])
    return st, epot, ekin


# ---------------------------------------------------------------------------------------------
# gfortran unformatted sequential records: [int32 nbytes][payload][int32 nbytes], little endian
# ---------------------------------------------------------------------------------------------
def _write_record(f, payload: bytes) -> None:
    f.write(struct.pack("<i", len(payload)))
    f.write(payload)
    f.write(struct.pack("<i", len(payload)))


def _read_record(f) -> bytes:
    head = f.read(4)
    if len(head) != 4:
        raise EOFError("unformatted record: missing length marker")
    (nb,) = struct.unpack("<i", head)
    payload = f.read(nb)
    (nb2,) = struct.unpack("<i", f.read(4))
    if nb2 != nb or len(payload) != nb:
        raise ValueError("unformatted record: length markers do not match")
    return payload


def _pack3(a, b, c) -> bytes:
    return np.concatenate([np.asarray(a, "<f8"), np.asarray(b, "<f8"), np.asarray(c, "<f8")]).tobytes()


def write_rv_init(filename: str, state) -> None:
    """outputs/rv_init.dat, md_initial_config_program.f90:275-290: record 1 = rx,ry,rz; record 2 = vx,vy,vz."""
    with open(filename, "wb") as f:
        _write_record(f, _pack3(state.rx, state.ry, state.rz))
        _write_record(f, _pack3(state.vx, state.vy, state.vz))


def read_rv_init(filename: str, n: int):
    """md_simulation_program.f90:573-588.  Returns (rx, ry, rz, vx, vy, vz)."""
    with open(filename, "rb") as f:
        r = np.frombuffer(_read_record(f), dtype="<f8")
        v = np.frombuffer(_read_record(f), dtype="<f8")
    if len(r) != 3 * n or len(v) != 3 * n:
        raise ValueError("read_rv_init(): record length does not match params%n")
    return tuple(np.ascontiguousarray(x) for x in (r[:n], r[n:2 * n], r[2 * n:], v[:n], v[n:2 * n], v[2 * n:]))


class RvaWriter:
    """outputs/one_run/rva.dat: header record (n:int32, L:real64, dt:real64, output_interval:int32,
    n_snapshots:int32 = 28 bytes, md_simulation_program.f90:257), then per sample 4 records
    r (wrapped), ru (unwrapped), v, a (:384-387) -- the layout md_one_run_analysis.py:345-397 reads."""

    def __init__(self, filename: str, n: int, box_length: float, dt: float, output_interval: int, n_snapshots: int):
        self.n = n
        self.f = open(filename, "wb")
        _write_record(self.f, struct.pack("<iddii", n, box_length, dt, output_interval, n_snapshots))

    def write_sample(self, state, ru) -> None:
        _write_record(self.f, _pack3(state.rx, state.ry, state.rz))
        _write_record(self.f, _pack3(*ru))
        _write_record(self.f, _pack3(state.vx, state.vy, state.vz))
        _write_record(self.f, _pack3(state.ax, state.ay, state.az))

    def close(self) -> None:
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_rva(filename: str):
    """Reads what RvaWriter / the reference writes: (header dict, list of (r, ru, v, a) arrays [3][n])."""
    with open(filename, "rb") as f:
        n, L, dt, interval, nsnap = struct.unpack("<iddii", _read_record(f))
        samples = []
        while True:
            try:
                recs = [np.frombuffer(_read_record(f), dtype="<f8").reshape(3, n) for _ in range(4)]
            except EOFError:
                break
            samples.append(tuple(recs))
    return {"n": n, "box_length": L, "dt": dt, "output_interval": interval, "n_snapshots": nsnap}, samples
