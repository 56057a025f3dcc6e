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
        """`count` successive values.  The recurrence x[k] = (x[k-55] - x[k-24]) mod MBIG only looks
        24 or more values back, so it is evaluated 24 values at a time with numpy (3 N draws for
        N = 2^20 would otherwise be a 3-million-iteration Python loop); short requests and the
        table initialisation use the scalar path.  Bit-identical to calling the object `count` times."""
        count = int(count)
        if count < 256:
            return np.array([self() for _ in range(count)], dtype=np.float64)
        first = self()                                   # initialises the table if needed
        # history in generation order: value generated i calls ago sits at hist[-i]
        order = [(self.inext - 1 - i) % 55 + 1 for i in range(55)]   # newest ... oldest slot (1-based)
        hist = np.empty(55 + count - 1, dtype=np.float64)
        hist[:55] = [self.ma[j] for j in reversed(order)]
        for k in range(55, 55 + count - 1, 24):
            m = min(24, 55 + count - 1 - k)
            v = hist[k - 55:k - 55 + m] - hist[k - 24:k - 24 + m]
            v[v < 0.0] += MBIG
            hist[k:k + m] = v
        out = np.empty(count, dtype=np.float64)
        out[0] = first
        out[1:] = hist[55:] * FAC
        # write the last 55 values back into the circular table and advance the cursors
        adv = count - 1
        tail = hist[-55:]
        self.inext = (self.inext - 1 + adv) % 55 + 1
        self.inextp = (self.inextp - 1 + adv) % 55 + 1
        for i in range(55):                              # value generated i calls ago -> slot inext - i
            self.ma[(self.inext - 1 - i) % 55 + 1] = float(tail[54 - i])
        return out


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


def _marker_size(f) -> int:
    """4-byte record markers (gfortran default) or 8-byte ones (-frecord-marker=8), detected like
    the reference's reader does (md_one_run_analysis.py:262-291): the leading length must be
    plausible and repeated after the payload."""
    pos = f.tell()
    head = f.read(64)
    f.seek(pos)
    for size, code in ((4, "<i"), (8, "<q")):
        if len(head) >= 2 * size:
            (nb,) = struct.unpack(code, head[:size])
            if 0 < nb < 1_000_000_000:
                f.seek(pos + size + nb)
                tail = f.read(size)
                f.seek(pos)
                if len(tail) == size and struct.unpack(code, tail)[0] == nb:
                    return size
    return 4


def _read_record(f, marker: int = 4) -> bytes:
    code = "<i" if marker == 4 else "<q"
    head = f.read(marker)
    if len(head) != marker:
        raise EOFError("unformatted record: missing length marker")
    (nb,) = struct.unpack(code, head)
    if nb < 0:
        raise ValueError("unformatted record: negative length")
    payload = f.read(nb)
    tail = f.read(marker)
    if len(payload) != nb or len(tail) != marker or struct.unpack(code, tail)[0] != nb:
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
        ms = _marker_size(f)
        r = np.frombuffer(_read_record(f, ms), dtype="<f8")
        v = np.frombuffer(_read_record(f, ms), dtype="<f8")
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
        ms = _marker_size(f)
        n, L, dt, interval, nsnap = struct.unpack("<iddii", _read_record(f, ms))
        samples = []
        while True:
            try:
                recs = [np.frombuffer(_read_record(f, ms), dtype="<f8").reshape(3, n) for _ in range(4)]
            except EOFError:
                break
            samples.append(tuple(recs))
    return {"n": n, "box_length": L, "dt": dt, "output_interval": interval, "n_snapshots": nsnap}, samples
