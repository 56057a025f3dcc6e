"""A/B parity of code paths selected by environment switches (each leg in its own process).

* the drop-in host-array call ljmd_verlet_step is replayed as one CUDA graph (upload, drift, force,
  finalize, download) by default; LJMD_COMPAT_GRAPH=0 issues the separate launches.  Both must be
  bitwise equal (first run on a B200 in round 2: green).
* large systems: page-locked caller arrays vs the bounce buffer (LJMD_COMPAT_PIN), and the download of r(t+dt)
  overlapped with the step kernel vs after it (LJMD_COMPAT_OVERLAP), across list rebuilds.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import ljmd_b200 as lj
from oracle import oracle_py as O
O.build()
n, L, rc, dt = 500, 10.0, 4.9, 1e-4
op = O.params_init(n, L, dt, rc, 5)
st, _, _ = O.initial_config(op, -12345, -555.0)
p = lj.init_params(n, L, dt, rc, 5)
s = lj.SimState.from_arrays([a.copy() for a in st])
obs = []
for k in range(25):
    obs.append(lj.verlet_step(p, s))
    if k == 10:                      # a parameter change must drop the captured graph
        p = lj.init_params(n, L, 2e-4, rc, 5)
np.savez(sys.argv[1], state=np.array([getattr(s, nm) for nm in lj.SimState.NAMES]), obs=np.array(obs))
"""


def _run(tmp_path, name, env_extra):
    out = tmp_path / (name + ".npz")
    env = dict(os.environ)
    env.pop("LJMD_COMPAT_GRAPH", None)
    env.update(env_extra)
    subprocess.run([sys.executable, "-c", _SCRIPT.format(root=ROOT), str(out)], check=True, env=env, timeout=240)
    return np.load(out)


def test_compat_graph_is_bitwise_equal_to_the_default_drop_in_call(tmp_path):
    a = _run(tmp_path, "launches", {"LJMD_COMPAT_GRAPH": "0"})
    b = _run(tmp_path, "graph", {})
    assert np.array_equal(a["obs"], b["obs"])
    assert np.array_equal(a["state"], b["state"])


_SCRIPT_BIG = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import bench, ljmd_b200 as lj
k = 32                                             # N = 131,072: 1 MiB per array, the page-locking threshold
L = k * (4.0 / 0.8) ** (1.0 / 3.0)
w = dict(k=k, L=L, rc=2.5, dt=1e-3, method="nlist", T=1.0)
n, arrs = bench.synthetic_state(w, seed=9)
p = lj.init_params(n, L, 1e-3, 2.5, k)
s = lj.SimState.from_arrays(arrs)
e0 = lj.compute_lj_potential_energy(p, s)
obs = [lj.verlet_step(p, s) for _ in range(3)]
lj.compat_release()                                # before the arrays go away
np.savez(sys.argv[1], state=np.array([getattr(s, nm) for nm in lj.SimState.NAMES]), obs=np.array(obs), e0=np.array(e0))
"""


def test_page_locked_caller_arrays_are_bitwise_equal_to_the_bounce_buffer(tmp_path):
    """Large systems: the drop-in pair page-locks the caller's arrays and copies straight from / to them
    (LJMD_COMPAT_PIN=0: pinned bounce buffer + two memcpy passes).  Same results, bit for bit."""
    outs = []
    for name, env_extra in (("bounce", {"LJMD_COMPAT_PIN": "0"}), ("pinned", {})):
        out = tmp_path / (name + ".npz")
        env = dict(os.environ)
        env.pop("LJMD_COMPAT_PIN", None)
        env.update(env_extra)
        subprocess.run([sys.executable, "-c", _SCRIPT_BIG.format(root=ROOT), str(out)], check=True, env=env, timeout=240)
        outs.append(np.load(out))
    for key in ("obs", "state", "e0"):
        assert np.array_equal(outs[0][key], outs[1][key]), key


_SCRIPT_LONG = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import bench, ljmd_b200 as lj
k = 32                                             # N = 131,072: the page-locked path
L = k * (4.0 / 0.8) ** (1.0 / 3.0)
w = dict(k=k, L=L, rc=2.5, dt=1e-3, method="nlist", T=1.0)
n, arrs = bench.synthetic_state(w, seed=4)
p = lj.init_params(n, L, 1e-3, 2.5, k)
s = lj.SimState.from_arrays(arrs)
lj.compute_lj_potential_energy(p, s)
obs, snaps = [], []
for call in range(70):                             # the list is rebuilt about every 42 calls at this dt and T
    obs.append(lj.verlet_step(p, s))
    if call in (0, 1, 40, 45, 50, 69):
        snaps.append(np.array([getattr(s, nm) for nm in lj.SimState.NAMES]))
lj.compat_release()
np.savez(sys.argv[1], snaps=np.array(snaps), obs=np.array(obs))
"""


def test_overlapped_position_download_is_bitwise_equal_across_rebuilds(tmp_path):
    """Large systems: r(t+dt) is final after the drift, so the drop-in call copies it to the caller on a second stream
    while the force / kick kernel runs (LJMD_COMPAT_OVERLAP=0: after the step, with v and a).  70 consecutive calls
    (at least one list rebuild inside a call) must agree bit for bit in every array the caller sees."""
    outs = []
    for name, env_extra in (("serial", {"LJMD_COMPAT_OVERLAP": "0"}), ("overlap", {})):
        out = tmp_path / (name + ".npz")
        env = dict(os.environ)
        env.pop("LJMD_COMPAT_OVERLAP", None)
        env.update(env_extra)
        subprocess.run([sys.executable, "-c", _SCRIPT_LONG.format(root=ROOT), str(out)], check=True, env=env, timeout=240)
        outs.append(np.load(out))
    assert np.all(np.isfinite(outs[0]["obs"]))
    for key in ("obs", "snaps"):
        assert np.array_equal(outs[0][key], outs[1][key]), key
    # the positions really moved between the snapshots (the test is not comparing a frozen state)
    assert not np.array_equal(outs[0]["snaps"][0][:3], outs[0]["snaps"][-1][:3])
