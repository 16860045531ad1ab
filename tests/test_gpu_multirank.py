"""Multi-GPU parity (``-m gpu``; needs >= 2 B200s, otherwise skipped): the public modules under torchrun/NCCL
against the per-rank outputs of the reference recorded in tests/golden/."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_torchrun_nccl_matches_reference(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    names = [n for n in golden_names() if load_golden(n)["world"] == world]
    assert names
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd + names, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
