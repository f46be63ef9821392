"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches tests/dist_worker.py under
torchrun with one rank per GPU and requires every check in it to pass.  The q-vortex time-step case (64^3, hyperpow 8,
SVV on: Richardson bootstrap + 3 ABCN steps, 1e-12 per step) compares against oracle snapshots that
tests/oracle_vortex.py computes ONCE here, on the host cores, before the ranks start."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture(scope="module")
def qvortex_snapshots(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("qv64") / "qv64.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "oracle_vortex.py"), "--n", "64", "--steps", "3",
                        "--out", out], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    return out


def run_worker(world, snapshots=None, tmpdir="/tmp", timeout=1200):
    env = dict(os.environ)
    env["MLEGS_TEST_TMP"] = str(tmpdir)
    if snapshots:
        env["MLEGS_QVORTEX_SNAPSHOTS"] = snapshots
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=timeout)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_parity(world, tmp_path, request):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    snaps = request.getfixturevalue("qvortex_snapshots")
    r = run_worker(world, snaps, tmp_path)
    print(r.stdout[-8000:])
    logdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(logdir):          # the full log travels back from the GPU box
        with open(os.path.join(logdir, f"dist_worker_{world}gpu.log"), "w") as fh:
            fh.write(r.stdout)
    assert r.returncode == 0 and "DIST WORKER OK" in r.stdout
