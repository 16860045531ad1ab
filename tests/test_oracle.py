"""Oracle vs the reference's own outputs (golden fixtures) -- CPU only.

Pins oracle/clip_oracle.py against every vector oracle/gen_golden.py recorded from the unmodified
reference (all four (local_loss, gather_with_grad) modes at W=2/4/8, W=1, GradScaler-style
grad_output, ragged sizes, SigLIP at W=1/3/4).
"""
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from oracle.clip_oracle import (bf16_round, clip_loss_oracle, cross_entropy_mean, ground_truth_labels,
                                multipositive_loss_oracle, siglip_loss_oracle)

LOSS_TOL = 2e-6      # fp32 reference vs float64 oracle
GRAD_TOL = 2e-5


def _parts(x, world):
    n = x.shape[0] // world
    return [x[r * n:(r + 1) * n] for r in range(world)]


@pytest.mark.parametrize("name", golden_names("clip_"))
def test_clip_oracle_matches_reference(name):
    g = load_golden(name)
    m, W = g["meta"], g["world"]
    out = clip_loss_oracle(_parts(g["image"], W), _parts(g["text"], W), m["scale"], bool(m["local_loss"]),
                           bool(m["gather_with_grad"]), float(m["grad_output"]))
    for r in range(W):
        ref = g["ranks"][r]
        assert abs(out[r]["loss"] - float(ref["loss"])) <= LOSS_TOL * max(1.0, abs(float(ref["loss"])))
        assert np.array_equal(out[r]["labels"], ref["labels"]) and out[r]["labels"].dtype == np.int64
        assert rel_err(out[r]["d_image"], ref["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_logit_scale"] - float(ref["d_scale"])) <= 5e-5 * max(abs(float(ref["d_scale"])), 1e-3)


@pytest.mark.parametrize("name", golden_names("siglip_"))
def test_siglip_oracle_matches_reference(name):
    g = load_golden(name)
    m, W = g["meta"], g["world"]
    out = siglip_loss_oracle(_parts(g["image"], W), _parts(g["text"], W), m["scale"], m["bias"], float(m["grad_output"]))
    for r in range(W):
        ref = g["ranks"][r]
        assert abs(out[r]["loss"] - float(ref["loss"])) <= 5e-6 * abs(float(ref["loss"]))
        assert rel_err(out[r]["d_image"], ref["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_logit_scale"] - float(ref["d_scale"])) <= 5e-5 * max(abs(float(ref["d_scale"])), 1e-3)
        assert abs(out[r]["d_logit_bias"] - float(ref["d_bias"])) <= 5e-5 * max(abs(float(ref["d_bias"])), 1e-3)


@pytest.mark.parametrize("num_logits,rank,world,local", [(8, 0, 1, False), (8, 3, 4, True), (8, 3, 4, False),
                                                         (4096, 7, 8, True), (1, 0, 1, True)])
def test_labels_bit_exact(num_logits, rank, world, local):
    lab = ground_truth_labels(num_logits, rank, world, local)
    assert lab.dtype == np.int64 and lab.shape == (num_logits,)
    off = num_logits * rank if (world > 1 and local) else 0
    assert np.array_equal(lab, np.arange(num_logits, dtype=np.int64) + off)


def test_mean_of_local_losses_is_global_loss():
    rng = np.random.default_rng(0)
    W, n, d = 4, 12, 16
    img = rng.standard_normal((W * n, d)); img /= np.linalg.norm(img, axis=1, keepdims=True)
    txt = rng.standard_normal((W * n, d)); txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    loc = clip_loss_oracle(_parts(img, W), _parts(txt, W), 20.0, True, True)
    glo = clip_loss_oracle(_parts(img, W), _parts(txt, W), 20.0, False, True)
    one = clip_loss_oracle([img], [txt], 20.0)
    assert abs(np.mean([o["loss"] for o in loc]) - glo[0]["loss"]) < 1e-12
    assert abs(glo[0]["loss"] - one[0]["loss"]) < 1e-12
    # gradient multiplicity table: (T,T) and (F,T) hand back W x the global gradient
    for r in range(W):
        rows = slice(r * n, (r + 1) * n)
        assert rel_err(loc[r]["d_image"], W * one[0]["d_image"][rows]) < 1e-12
        assert rel_err(glo[r]["d_text"], W * one[0]["d_text"][rows]) < 1e-12
    assert abs(sum(o["d_logit_scale"] for o in loc) - W * one[0]["d_logit_scale"]) < 1e-12
    ff = clip_loss_oracle(_parts(img, W), _parts(txt, W), 20.0, False, False)
    assert rel_err(ff[1]["d_image"], one[0]["d_image"][n:2 * n]) < 1e-12


def test_homogeneity_identity():
    """loss depends on scale*I*T only: scale * d_scale == sum(dI * I) == sum(dT * T) (W=1)."""
    rng = np.random.default_rng(1)
    img = rng.standard_normal((40, 24)); txt = rng.standard_normal((40, 24))
    o = clip_loss_oracle([img], [txt], 3.0)[0]
    assert abs(3.0 * o["d_logit_scale"] - (o["d_image"] * img).sum()) < 1e-10
    assert abs(3.0 * o["d_logit_scale"] - (o["d_text"] * txt).sum()) < 1e-10


def test_cross_entropy_and_bf16_round():
    z = np.array([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]])
    loss, dz = cross_entropy_mean(z, np.array([2, 0]))
    assert abs(loss - 0.5 * (np.log(np.exp(-2) + np.exp(-1) + 1) + np.log(3))) < 1e-12
    assert np.allclose(dz.sum(axis=1), 0)
    x = np.array([1.0, 1.00390625, 1.005859375, -3.1415927], dtype=np.float32)
    import torch
    assert np.array_equal(bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


@pytest.mark.parametrize("name", [n for n in golden_names() if "_w1" in n and not n.startswith("mpos")])
def test_timed_port_matches_reference(name):
    """oracle/clip_port.py (the CPU baseline bench.py times) reproduces the reference's W=1 outputs."""
    import torch
    from oracle.clip_port import clip_rank_step, siglip_rank_step
    g = load_golden(name)
    m, ref = g["meta"], g["ranks"][0]
    img, txt = torch.from_numpy(g["image"]), torch.from_numpy(g["text"])
    if m["kind"] == "clip":
        loss, di, dt, ds = clip_rank_step(img, txt, img, txt, float(m["scale"]), 0, float(m["grad_output"]))
    else:
        loss, di, dt, ds, db = siglip_rank_step(img, txt, float(m["scale"]), float(m["bias"]))
        assert abs(db.item() - float(ref["d_bias"])) <= 1e-4 * abs(float(ref["d_bias"]))
    assert abs(loss.item() - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    assert rel_err(di.numpy(), ref["d_image"]) <= 1e-4 and rel_err(dt.numpy(), ref["d_text"]) <= 1e-4
    assert abs(ds.item() - float(ref["d_scale"])) <= 1e-4 * abs(float(ref["d_scale"])) + 1e-6


def test_timed_port_rank_block_matches_oracle():
    import torch
    from oracle.clip_port import clip_rank_step
    rng = np.random.default_rng(3)
    W, n, d = 4, 8, 16
    img = rng.standard_normal((W * n, d)).astype(np.float32); img /= np.linalg.norm(img, axis=1, keepdims=True)
    txt = rng.standard_normal((W * n, d)).astype(np.float32); txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    ref = clip_loss_oracle(_parts(img, W), _parts(txt, W), 10.0, True, False)
    ti, tt = torch.from_numpy(img), torch.from_numpy(txt)
    for r in range(W):
        rows = slice(r * n, (r + 1) * n)
        loss, di, dt, ds = clip_rank_step(ti[rows], tt[rows], ti, tt, 10.0, r * n)
        assert abs(loss.item() - ref[r]["loss"]) < 1e-5
        assert rel_err(di.numpy(), ref[r]["d_image"]) < 1e-5 and rel_err(dt.numpy(), ref[r]["d_text"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("mpos_"))
def test_multipositive_oracle_matches_reference(name):
    """the reference's MultiPositiveClipLoss (loss.py:671-747), single process and gloo groups of 2 / 4"""
    g = load_golden(name)
    m, W = g["meta"], g["world"]
    labels = [g["ranks"][r]["labels_in"] for r in range(W)]
    out = multipositive_loss_oracle(_parts(g["image"], W), _parts(g["text"], W), labels, m["scale"], m["delta"],
                                    float(m["grad_output"]))
    for r in range(W):
        ref = g["ranks"][r]
        assert abs(out[r]["loss"] - float(ref["loss"])) <= LOSS_TOL * max(1.0, abs(float(ref["loss"])))
        assert rel_err(out[r]["d_image"], ref["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_logit_scale"] - float(ref["d_scale"])) <= 5e-5 * max(abs(float(ref["d_scale"])), 1e-3)


def test_ref_copy_is_unmodified():
    """oracle/_ref (what bench.py's reference arm and experiments/train_step.py run on the GPU box) is a byte-for-byte
    copy of the reference sources: every file matches the sha256 in its manifest and, where /root/reference exists,
    the original."""
    import hashlib
    import json
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    manifest_path = os.path.join(root, "MANIFEST.json")
    if not os.path.exists(manifest_path):
        pytest.skip("oracle/_ref has not been made on this machine (python oracle/make_ref.py)")
    manifest = json.load(open(manifest_path))
    assert "open_clip/loss.py" in manifest and "open_clip_train/train.py" in manifest
    for rel, meta in manifest.items():
        data = open(os.path.join(root, "src", rel), "rb").read()
        assert hashlib.sha256(data).hexdigest() == meta["sha256"], rel
        if os.path.exists(meta["source"]):
            assert open(meta["source"], "rb").read() == data, rel
