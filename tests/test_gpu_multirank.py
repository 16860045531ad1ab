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


def _run_parity(world, cases, timeout=1500):
    script = os.path.join(ROOT, "tests", "dist_parity.py")
    if world == 1:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), script]
    res = subprocess.run(cmd + cases, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0 and "all green" in res.stdout, res.stdout[-6000:] + res.stderr[-3000:]


def test_production_size_parity_one_gpu():
    """BASELINE configs at full size against the reference's fp32 torch graph (tests/torch_ref.py), one GPU."""
    if not torch.cuda.is_available():
        pytest.fail("needs a B200 (sm_100a)")
    _run_parity(1, ["c3", "c2", "c2s100", "c2raw", "c4", "c4raw", "mpos", "ragged"])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_production_size_parity_multi_gpu(world):
    """The same over NVLink peer memory and over NCCL: gradients through the real collectives at the sizes the
    scaling numbers are quoted on (c3: N=32768, D=768; c2: N=4096, D=512; c4: SigLIP N=16384)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    _run_parity(world, ["c3", "c2", "c2s100", "c2raw", "c4", "c4raw", "mpos", "ragged", "exchange"])
