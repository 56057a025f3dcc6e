"""REAL multi-process runs of the slab-sharded neighbour-list path: one
process per GPU under torchrun, NCCL for the plumbing, CUDA-IPC peer memory for the halo stores, step
kernels of different ranks running concurrently and waiting on each other's per-step flags (the
spin-wait path of nl_step_kernel and its halt logic, which the single-process emulation in
test_gpu_nlist.py cannot exercise).  Skipped when fewer GPUs are visible than ranks wanted.

Bar: the final state on EVERY rank is bitwise equal to the single-context run of the same system,
across at least two list rebuilds; the ranks' momentum shares add up to zero after every call."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(tmp_path, world, k, nsteps, skin, port):
    out = tmp_path / f"slab_w{world}.json"
    env = dict(os.environ)
    env.setdefault("LJMD_SLAB_SPIN_TIMEOUT_S", "20")  # a flag that never arrives fails the test instead of hanging the box
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "slab_worker.py"), str(out), str(k),
           str(nsteps), str(skin)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.load(open(out))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_processes_bitwise_equal_to_single_context(tmp_path, world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    # N = 131,072 at rho = 0.8: 19 cell layers; skin 0.12 => a rebuild every ~15 steps at T = 1, dt = 1e-3
    res = _run(tmp_path, world, 32, 150, 0.12, 29611 + world)
    assert len(res) == world
    for r in res:
        assert r["bitwise_equal_to_single_context"], r
        assert r["momentum_ok"], r
        assert r["rebuilds_sharded"] >= 2 and r["obs_rel_err"] <= 1e-13, r
