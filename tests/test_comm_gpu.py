"""Multi-GPU parity for the NVLink collectives: spawns tests/multigpu_worker.py under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("nproc", [2])
def test_p2p_collectives_multigpu(nproc):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, PYTHONPATH=str(REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(REPO / "tests" / "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("WORKER OK") == nproc
