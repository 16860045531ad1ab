"""Drop-in check against the reference's own training loop (CPU, build container only).

The unmodified ``train_one_epoch`` of the reference (``src/open_clip_train/train.py:70-269``) drives a tiny two-tower
model twice from the same initial weights: once with the reference's loss module and once with ``mrclip_b200``'s
(kernels replaced by the float64 stand-in engine, so this exercises the call contract, not the CUDA code): the
keyword call ``loss(**model_out, output_dict=True)`` (``train.py:128``; ``tokenized_texts=labels, delta=...`` for the
multi-positive loss, ``:123``), ``sum(losses.values())``, ``backward(total_loss, scaler)``, the optimizer step and the
``logit_scale`` clamp.  The parameter updates must agree.

Needs /root/reference (present where the CPU suite runs, absent on the GPU box: skipped there; not a ``gpu`` test).
"""
import argparse
import copy
import math
import os
import sys
import types

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import mrclip_b200
from conftest import StandInEngine, rel_err

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_SRC, "open_clip_train")),
                                reason="the reference tree is not available on this machine")


@pytest.fixture(scope="module")
def ref():
    added = []
    if "ftfy" not in sys.modules:                      # the tokenizer's only missing import; never called here
        stub = types.ModuleType("ftfy")
        stub.fix_text = lambda s: s
        sys.modules["ftfy"] = stub
        added.append("ftfy")
    sys.path.insert(0, REF_SRC)
    try:
        import open_clip.loss as ref_loss
        import open_clip_train.train as ref_train
        yield types.SimpleNamespace(loss=ref_loss, train=ref_train)
    finally:
        sys.path.remove(REF_SRC)
        for name in added:
            sys.modules.pop(name, None)


class TinyTowers(nn.Module):
    """forward(images, texts) -> the dict open_clip's CLIP returns with output_dict=True (model.py:303-332)"""

    def __init__(self, with_bias):
        super().__init__()
        self.visual = nn.Linear(12, 16)
        self.embed = nn.Embedding(40, 16)
        self.proj = nn.Linear(16, 16)
        self.logit_scale = nn.Parameter(torch.tensor(math.log(1 / 0.07)))
        self.logit_bias = nn.Parameter(torch.tensor(-3.0)) if with_bias else None

    def forward(self, images, texts):
        out = {"image_features": F.normalize(self.visual(images), dim=-1),
               "text_features": F.normalize(self.proj(self.embed(texts).mean(1)), dim=-1),
               "logit_scale": self.logit_scale.exp()}
        if self.logit_bias is not None:
            out["logit_bias"] = self.logit_bias
        return out


class Loader(list):
    num_batches = 3
    num_samples = 3 * 24


class TrainData:
    def __init__(self, batches):
        self.dataloader = Loader(batches)

    def set_epoch(self, epoch):
        pass


def _args(multipositive):
    return argparse.Namespace(device="cpu", precision="fp32", distill=False, accum_freq=1, freeze=False,
                              skip_scheduler=True, distance=False, multipositiveloss=multipositive, delta=0.3,
                              horovod=False, grad_clip_norm=None, log_every_n_steps=1, world_size=1, batch_size=24,
                              wandb=False, rank=0, local_rank=0)


@pytest.mark.parametrize("kind", ["clip", "multipositive", "siglip"])
def test_reference_train_loop_accepts_the_drop_in(ref, kind, monkeypatch):
    monkeypatch.setenv("MRCLIP_BWD", "fused")          # the stand-in's exact backend (no bf16 rounding of G)
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(24, 12, generator=g), torch.randint(0, 40, (24, 5), generator=g),
                torch.randint(0, 6, (24,), generator=g)) for _ in range(3)]
    torch.manual_seed(1)
    init = TinyTowers(with_bias=(kind == "siglip"))
    make = {"clip": lambda m: m.ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1),
            "multipositive": lambda m: m.MultiPositiveClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True,
                                                               rank=0, world_size=1),
            "siglip": lambda m: m.SigLipLoss(rank=0, world_size=1)}[kind]
    finals, losses = [], []
    for module in (ref.loss, mrclip_b200):
        model = copy.deepcopy(init)
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        loss_mod = make(module)
        seen = []
        hook = loss_mod.register_forward_hook(lambda _m, _i, out: seen.append(float(sum(v.detach() for v in out.values()))))
        if module is mrclip_b200:
            mrclip_b200.set_engine(StandInEngine())
        try:
            ref.train.train_one_epoch(model, {"train": TrainData(batches)}, loss_mod, 0, opt, None, None, None,
                                      _args(kind == "multipositive"))
        finally:
            mrclip_b200.set_engine(None)
            hook.remove()
        finals.append({k: v.detach().clone() for k, v in model.named_parameters()})
        losses.append(seen)
    assert len(losses[0]) == len(losses[1]) == 3
    for a, b in zip(*losses):
        assert abs(a - b) <= 2e-3 * abs(a)               # the drop-in rounds the features to bf16
    start = dict(init.named_parameters())
    for name in finals[0]:
        upd_ref = (finals[0][name] - start[name].detach()).numpy()
        upd_new = (finals[1][name] - start[name].detach()).numpy()
        assert rel_err(upd_new, upd_ref) <= 1e-2, name      # measured 2e-4 .. 4e-3 (bf16 rounding of the features)
