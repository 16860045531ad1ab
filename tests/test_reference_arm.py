"""bench.py's reference arm and the runner behind it (CPU): the unmodified reference module from oracle/_ref is timed
through its public API, in one process and as rank 0 of a gloo job, and the JSON line carries the contract's keys.
Skipped where oracle/_ref has not been made (python oracle/make_ref.py needs /root/reference)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_runner

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref has not been made on this machine")


def test_runner_times_the_reference_module():
    one = ref_runner.time_single(128, 64, steps=2, warmup=1)
    assert one["kind"] == "reference" and one["pairs_per_s"] > 0 and "ClipLoss(world_size=1)" in one["sample"]
    sig = ref_runner.time_single(128, 64, steps=2, warmup=1, kind="siglip", scale=10.0, bias=-10.0)
    assert sig["pairs_per_s"] > 0 and "SigLipLoss" in sig["sample"]
    # problem sizes whose full step would exceed the budget become one rank of a gloo job
    assert "gloo job" in ref_runner.time_reference(512, 64, 1, 1, budget_flop=12.0 * 256 * 512 * 64)["sample"]


def test_bench_reference_line_has_the_contract_keys():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "3",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ClipLoss fwd+bwd pairs/sec" and line["unit"] == "pairs/s"
    assert line["steps"] == 3 and line["warmup"] == 1 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[0]" in line["config"]["workload"]
    # the other ranks of a torchrun launch exit 0 without work
    env["RANK"], env["WORLD_SIZE"] = "3", "8"
    idle = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"],
                          capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert idle.returncode == 0 and idle.stdout.strip() == ""
