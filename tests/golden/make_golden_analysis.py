"""Generates tests/golden/analysis_ref.npz by RUNNING THE REFERENCE'S OWN PYTHON.

What is executed: /root/reference/scripts/md_one_run_analysis.py is imported as a module (its top level
holds only imports, definitions and an `if __name__ == "__main__"` guard -- read before running) with
`matplotlib` stubbed (not installed here; the functions called below never touch it), and these functions
are called on the inputs built below:
    read_rva (:345-397), compute_msd_tau_timeorig (:404-441), compute_vacf_tau_timeorig (:444-489),
    diffusion_from_msd_linear_fit (:492-513), diffusion_from_vacf_integral (:516-534), compute_rdf (:537-598).
`np.trapz` (used at :534) no longer exists in numpy 2.x; it is aliased to `np.trapezoid`, the same
function under its new name, before the call.

The reference cannot travel to the GPU box, so BOTH the inputs and the reference's outputs are committed in
the fixture; tests/test_analysis_golden.py (CPU: oracle/analysis_oracle.py and the product's host pieces)
and tests/test_zz_gpu_analysis.py (GPU: csrc/md_analysis.cu through the C ABI) compare against them.

Inputs:
  liquid   a real trajectory: N = 108 LJ liquid (rho = 0.8, rc = 2.5, dt = 2e-3) advanced by the CPU oracle,
           24 snapshots 5 steps apart, written as rva.dat by the PRODUCT's RvaWriter and read back by the
           reference's read_rva -- pins the file format too (the file's sha256 is stored)
  gas864   2 snapshots x 864 uniform random particles: the 800-particle subsample of compute_rdf (:559-561)
  many     205 snapshots x 12 particles: the 200-snapshot thinning (:555-557)
  walk     60 snapshots x 40 particles random walk: MSD / VACF lag / stride variants and the diffusion fits

Re-run (only where /root/reference exists):  python tests/golden/make_golden_analysis.py
"""
import hashlib
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

sys.dont_write_bytecode = True   # /root/reference is read-only; do not try to drop a __pycache__ there
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/scripts/md_one_run_analysis.py"


def load_reference():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    spec = importlib.util.spec_from_file_location("ref_md_one_run_analysis", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod       # dataclasses look the module up while the class body executes
    spec.loader.exec_module(mod)
    return mod


def liquid_trajectory():
    """N = 108 liquid from the CPU oracle; the rva.dat file is written by the product's RvaWriter."""
    from oracle import oracle_py as O
    import ljmd_b200 as lj
    from ljmd_b200 import initial_config as ic

    O.build()
    n, k = 108, 3
    L = (n / 0.8) ** (1.0 / 3.0)
    dt, interval, n_snap = 2e-3, 5, 24
    p = O.params_init(n, L, dt, 2.5, k)
    st, _, _ = O.initial_config(p, -12345, 0.0, 1.0)
    s = [a.copy() for a in st]
    O.run_steps(p, 400, s)                       # melt
    ru = [a.copy() for a in s[:3]]
    path = os.path.join(tempfile.mkdtemp(prefix="ljmd_golden_"), "rva.dat")
    acc = np.empty((n_snap, 3, n))
    with ic.RvaWriter(path, n, L, dt, interval, n_snap) as w:
        for q in range(n_snap):
            O.run_steps(p, interval, s, ru=ru)
            w.write_sample(lj.SimState.from_arrays([a.copy() for a in s]), [a.copy() for a in ru])
            acc[q] = s[6:9]
    return path, acc


def main():
    ref = load_reference()
    out = {}

    # ---- liquid: file -> reference reader -> reference analysis -------------------------------------
    path, acc = liquid_trajectory()
    with open(path, "rb") as f:
        out["liquid_rva_sha256"] = np.frombuffer(hashlib.sha256(f.read()).digest(), dtype=np.uint8)
    from pathlib import Path
    d = ref.read_rva(Path(path))
    out["liquid_header"] = np.array([d.n, d.L, d.dt, d.output_interval, d.n_snapshots], dtype=np.float64)
    for name in ("rx", "ry", "rz", "rux", "ruy", "ruz", "vx", "vy", "vz"):
        out["liquid_" + name] = getattr(d, name)
    out["liquid_acc"] = acc                      # record 4 (ignored by the reader) -- needed to rebuild the file
    for tag, kw in (("full", {}), ("lag7_stride3", dict(max_lag=7, origin_stride=3)),
                    ("lag100_stride2", dict(max_lag=100, origin_stride=2))):
        out[f"liquid_msd_{tag}"] = ref.compute_msd_tau_timeorig(d.rux, d.ruy, d.ruz, **kw)
        out[f"liquid_vacf_{tag}"] = ref.compute_vacf_tau_timeorig(d.vx, d.vy, d.vz, **kw)
    r, g = ref.compute_rdf(d.rx, d.ry, d.rz, d.L)
    out["liquid_rdf_r"], out["liquid_rdf_g"] = r, g
    r, g = ref.compute_rdf(d.rx, d.ry, d.rz, d.L, nbins=64, rmax=2.0)
    out["liquid_rdf64_r"], out["liquid_rdf64_g"] = r, g
    # single snapshot / two snapshots (:416-417, :456-458)
    for ns in (1, 2):
        out[f"liquid_msd_first{ns}"] = ref.compute_msd_tau_timeorig(d.rux[:ns], d.ruy[:ns], d.ruz[:ns])
        out[f"liquid_vacf_first{ns}"] = ref.compute_vacf_tau_timeorig(d.vx[:ns], d.vy[:ns], d.vz[:ns])

    # ---- RDF subsampling ------------------------------------------------------------------------------
    rng = np.random.default_rng(20261018)
    L = 9.5
    x, y, z = rng.random((3, 2, 864)) * L
    out["gas864_L"] = np.float64(L)
    out["gas864_x"], out["gas864_y"], out["gas864_z"] = x, y, z
    out["gas864_rdf_r"], out["gas864_rdf_g"] = ref.compute_rdf(x, y, z, L, nbins=50)
    L = 4.0
    x, y, z = rng.random((3, 205, 12)) * L
    out["many_L"] = np.float64(L)
    out["many_x"], out["many_y"], out["many_z"] = x, y, z
    out["many_rdf_r"], out["many_rdf_g"] = ref.compute_rdf(x, y, z, L, nbins=25)

    # ---- random walk: lag / stride variants and the two diffusion estimates ----------------------------
    v = rng.normal(size=(3, 60, 40))
    ru = np.cumsum(v, axis=1) * 0.05 + rng.random((3, 1, 40)) * 5.0
    out["walk_v"], out["walk_ru"] = v, ru
    for tag, kw in (("full", {}), ("lag20_stride4", dict(max_lag=20, origin_stride=4)), ("lag0", dict(max_lag=0))):
        out[f"walk_msd_{tag}"] = ref.compute_msd_tau_timeorig(*ru, **kw)
        out[f"walk_vacf_{tag}"] = ref.compute_vacf_tau_timeorig(*v, **kw)
    t = np.arange(60, dtype=np.float64) * 0.01
    out["walk_t"] = t
    out["walk_D_msd"] = np.array([ref.diffusion_from_msd_linear_fit(t, out["walk_msd_full"], f)
                                  for f in (0.5, 0.2, 1.0, 0.0)])
    out["walk_D_vacf"] = np.array([ref.diffusion_from_vacf_integral(t, out["walk_vacf_full"], f)
                                   for f in (1.0, 0.8, 0.0)])
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(HERE, "analysis_ref.npz"), **out)
    print("analysis_ref.npz written:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
