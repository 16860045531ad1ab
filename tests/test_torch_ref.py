"""Pins ``tests/torch_ref.py`` (the fp32 torch restatement the production-size GPU parity runs use as their reference)
against the unmodified reference's recorded outputs in tests/golden/: single process and gloo world_size 2."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import torch_ref
from conftest import golden_names, load_golden, rel_err


def _run(case, rank, world):
    m = case["meta"]
    n = case["image"].shape[0] // world
    rows = slice(rank * n, (rank + 1) * n)
    img, txt = torch.from_numpy(case["image"][rows]), torch.from_numpy(case["text"][rows])
    go = float(m["grad_output"])
    if m["kind"] == "clip":
        out = torch_ref.clip_reference(img, txt, float(m["scale"]), bool(m["local_loss"]), bool(m["gather_with_grad"]),
                                       rank, world, go)
    elif m["kind"] == "siglip":
        out = torch_ref.siglip_reference(img, txt, float(m["scale"]), float(m["bias"]), rank, world, go)
    else:
        out = torch_ref.mpos_reference(img, txt, float(m["scale"]), torch.from_numpy(case["ranks"][rank]["labels_in"]),
                                       float(m["delta"]), rank, world, go)
    return {k: v.numpy() for k, v in out.items()}


def _check(out, ref, kind):
    assert abs(float(out["loss"]) - float(ref["loss"])) <= 2e-6 * max(1.0, abs(float(ref["loss"])))
    assert rel_err(out["d_image"], ref["d_image"]) <= 2e-5
    assert rel_err(out["d_text"], ref["d_text"]) <= 2e-5
    assert abs(float(out["d_scale"]) - float(ref["d_scale"])) <= 2e-5 * max(abs(float(ref["d_scale"])), 1e-3)
    if kind == "clip":
        assert np.array_equal(out["labels"], ref["labels"])
    if kind == "siglip":
        assert abs(float(out["d_bias"]) - float(ref["d_bias"])) <= 2e-5 * max(abs(float(ref["d_bias"])), 1e-3)


@pytest.mark.parametrize("name", [n for n in golden_names() if "_w1" in n])
def test_single_process_matches_reference(name):
    case = load_golden(name)
    _check(_run(case, 0, 1), case["ranks"][0], case["meta"]["kind"])


def _worker(rank, world, init_file, names, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        ret[rank] = {name: _run(load_golden(name), rank, world) for name in names}
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_world2_matches_reference():
    names = [n for n in golden_names() if load_golden(n)["world"] == 2]
    assert names
    ret = mp.Manager().dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(2, os.path.join(td, "init"), names, ret), nprocs=2, join=True)
    for name in names:
        case = load_golden(name)
        for r in range(2):
            _check(ret[r][name], case["ranks"][r], case["meta"]["kind"])
