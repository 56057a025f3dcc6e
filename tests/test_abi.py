"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/ljmd.h
declares, mirrors md_types, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ljmd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ljmd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lj):
    L = lj.load_library()
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/ljmd.h but not exported"
    assert sorted(lj.ABI_SYMBOLS) == names


def test_params_struct_layout(lj):
    assert ctypes.sizeof(lj.SimParams) == 96
    offs = {f[0]: getattr(lj.SimParams, f[0]).offset for f in lj.SimParams._fields_}
    assert offs["n"] == 0 and offs["num_cells"] == 4 and offs["box_length"] == 8
    assert offs["rc_square"] == 72 and offs["inv_length_scale"] == 88


def test_params_init_matches_oracle(lj, O):
    p = lj.init_params(4000, 20.0, 1e-4, 9.8, 10)
    q = O.params_init(4000, 20.0, 1e-4, 9.8, 10)
    assert bytes(p) == bytes(q)


@pytest.mark.parametrize("args,code,msg", [
    ((500, -1.0, 1e-4, 4.9), 2, "box_length must be > 0"),
    ((500, 10.0, 1e-4, -1.0), 4, "rc (cutoff_radius) must be > 0"),
    ((500, 10.0, 1e-4, 5.0), 7, "must be < L/2"),
    ((500, 10.0, 0.0, 4.9), 8, "dt must be > 0"),
])
def test_params_init_errors_mirror_reference_stops(lj, args, code, msg):
    # md_types.f90:143,150,153,161
    with pytest.raises(lj.LjmdError) as e:
        lj.init_params(*args)
    assert e.value.code == code and msg in e.value.message


def test_no_cpu_fallback(lj):
    """Without a CUDA device every compute entry point must fail loudly."""
    if lj.device_count() > 0:
        pytest.skip("a GPU is present")
    p = lj.init_params(500, 10.0, 1e-4, 4.9, 5)
    s = lj.init_state(p)
    with pytest.raises(lj.LjmdError) as e:
        lj.verlet_step(p, s)
    assert e.value.code == 100
    with pytest.raises(lj.LjmdError) as e:
        lj.compute_lj_potential_energy(p, s)
    assert e.value.code == 100
    with pytest.raises(lj.LjmdError):
        lj.Context(p)


def test_argument_guards_without_gpu(lj):
    import numpy as np

    L = lj.load_library()
    p = lj.init_params(500, 10.0, 1e-4, 4.9, 5)
    bad = lj.SimParams.from_buffer_copy(bytes(p))
    bad.n = 0
    s = lj.init_state(p)
    with pytest.raises(lj.LjmdError) as e:   # lj_potential_energy.f90:77
        lj.compute_lj_potential_energy(bad, s)
    assert e.value.code == 1 and "params%n must be > 0" in e.value.message
    with pytest.raises(lj.LjmdError) as e:   # verlet.f90:51
        lj.verlet_step(bad, s)
    assert e.value.code == 1 and "verlet_step()" in e.value.message
    bad = lj.SimParams.from_buffer_copy(bytes(p))
    bad.volume = 0.0
    with pytest.raises(lj.LjmdError) as e:   # lj_potential_energy.f90:79
        lj.compute_lj_potential_energy(bad, s)
    assert e.value.code == 3
    # NULL array == "state arrays are not allocated" (:82)
    e_, d_, dd_ = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    rc = L.ljmd_compute_lj(ctypes.byref(p), None, None, None, None, None, None,
                           ctypes.byref(e_), ctypes.byref(d_), ctypes.byref(dd_))
    assert rc == 6 and b"not allocated" in L.ljmd_last_error()
    assert np.all(s.ax == 0.0)


def test_product_does_not_import_oracle():
    """The shipped package must never reach into oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("no cpu fallback", ""), (dirpath, f)


def test_cpp_host_links_and_fails_loudly_without_gpu(lj, tmp_path):
    """examples/md_driver.cpp: a C++ host on the C ABI alone (no torch, no Python).  It must compile
    and link against include/ljmd.h + libljmd.so; on a machine without a CUDA device it must stop with
    the library's status instead of computing anything on the CPU; bad parameters stop with the
    reference's message (md_types.f90:152)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(lj.LIB_PATH)
    exe = os.path.join(tmp_path, "md_driver")
    subprocess.run(["g++", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "md_driver.cpp"),
                    "-L" + libdir, "-lljmd", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    bad = subprocess.run([exe, "5", "10", "5.0"], capture_output=True, text=True)
    assert bad.returncode == 7 and "must be < L/2" in bad.stderr
    run = subprocess.run([exe, "3", "6.0", "2.5", "1e-3", "40", "20", "0"], capture_output=True, text=True)
    if lj.device_count() == 0:
        assert run.returncode == 100 and "no CPU fallback" in run.stderr
    else:
        assert run.returncode == 0 and run.stdout.count("\n") >= 4


def test_debug_exports_are_not_in_the_product_header(lj):
    """Tuning / debug entry points live in csrc/ljmd_debug.h, not in include/ljmd.h."""
    names = header_functions()
    L = lj.load_library()
    for name in lj.DEBUG_SYMBOLS:
        assert name not in names and hasattr(L, name)


def test_drop_in_pair_rejects_arrays_of_the_wrong_length(lj):
    p = lj.init_params(500, 10.0, 1e-4, 4.9, 5)
    s = lj.SimState(400)
    with pytest.raises(ValueError):
        lj.verlet_step(p, s)
    with pytest.raises(ValueError):
        lj.compute_lj_potential_energy(p, s)
