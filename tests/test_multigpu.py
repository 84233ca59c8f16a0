"""Multi-GPU parity as pytest cases: the sharded runs -- one process per GPU with one NCCL all-reduce
per run, and one process driving several devices -- reproduce the single-GPU run of the same global walkers.  The
checks themselves live in tests/multigpu_check.py (run under torch.distributed.run); they need >= 2 GPUs and are
skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_per_gpu_reproduces_the_single_gpu_run(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29500 + 7 * world),
                        os.path.join(ROOT, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    for tag in ("blocking OK", "generic OK", "stored operator OK"):
        assert r.stdout.count(tag) == world, (tag, r.stdout[-4000:])


def test_one_process_driving_two_devices_reproduces_the_single_gpu_run():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multigpu_check.py"), "--single-process", "2"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
