"""Timed CPU port of the reference's ClipLoss / SigLipLoss step  --  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's implementation of this path *is* PyTorch CPU code (src/open_clip/loss.py builds the
logits with ``@``, calls ``F.cross_entropy`` / ``F.logsigmoid`` and lets autograd run the backward),
and /root/reference does not exist on the GPU box, so ``bench.py``'s ``cpu_baseline`` leg and its
``--impl reference`` arm time this port on the box's host cores instead (kind = "port").  It uses the
same torch operators in the same order as the reference (loss.py:116-124 operand order
``(scale * A) @ B.T``, :134-137 mean cross entropies halved, :349-362 sigmoid loss) on fp32 CPU
tensors with every host thread torch can use.  ``tests/test_oracle.py`` pins it against the
golden vectors recorded from the unmodified reference.

One "rank step" is what a single rank of the reference executes in local-loss mode with gathered
features already present: two logits blocks (image rows vs all texts, text rows vs all images),
their cross entropies, and the autograd backward through both.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def clip_rank_step(img_rows, txt_rows, img_all, txt_all, logit_scale, label_offset=0, grad_output=1.0):
    """ClipLoss forward+backward for one row block.  Returns (loss, d_img_rows, d_txt_rows, d_scale).

    With img_rows is img_all (world_size 1) this is exactly loss.py:123-137.
    """
    i = img_rows.detach().clone().requires_grad_(True)
    t = txt_rows.detach().clone().requires_grad_(True)
    s = torch.as_tensor(logit_scale, dtype=torch.float32).detach().clone().requires_grad_(True)
    same = img_rows.shape[0] == img_all.shape[0]
    ia = i if same else img_all
    ta = t if same else txt_all
    logits_per_image = s * i @ ta.T
    logits_per_text = s * t @ ia.T
    labels = torch.arange(i.shape[0], dtype=torch.long) + label_offset
    loss = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
    (loss * grad_output).backward()
    return loss.detach(), i.grad, t.grad, s.grad


def siglip_rank_step(img_rows, txt_all, logit_scale, logit_bias, label_offset=0):
    """SigLipLoss forward+backward of one rank's image rows against every text chunk (loss.py:349-362)."""
    i = img_rows.detach().clone().requires_grad_(True)
    t = txt_all.detach().clone().requires_grad_(True)
    s = torch.as_tensor(logit_scale, dtype=torch.float32).detach().clone().requires_grad_(True)
    b = torch.as_tensor(logit_bias, dtype=torch.float32).detach().clone().requires_grad_(True)
    n = i.shape[0]
    logits = s * i @ t.T + b
    labels = -torch.ones_like(logits)
    idx = torch.arange(n)
    labels[idx, idx + label_offset] = 1.0
    loss = -F.logsigmoid(labels * logits).sum() / n
    loss.backward()
    return loss.detach(), i.grad, t.grad, s.grad, b.grad


def time_clip_sample(n_total, d, sample_rows, steps=1, warmup=1, seed=1237, threads=None):
    """Time ``clip_rank_step`` on ``sample_rows`` pairs of an ``n_total``-pair problem.

    Returns dict(pairs_per_s, seconds_per_step, cores, sample).  Each step contrasts ``sample_rows``
    image/text pairs against all ``n_total`` candidates, forward and backward, like one rank of the
    reference's local-loss mode; pairs/s = sample_rows / seconds.
    """
    import os
    if threads is None:
        threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    img = F.normalize(torch.randn(n_total, d, generator=g), dim=-1).bfloat16().float()
    txt = F.normalize(0.5 * img + 0.5 * torch.randn(n_total, d, generator=g) / d ** 0.5, dim=-1).bfloat16().float()
    rows = slice(0, sample_rows)
    scale = torch.tensor(14.285714)
    for _ in range(warmup):
        clip_rank_step(img[rows], txt[rows], img, txt, scale)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        clip_rank_step(img[rows], txt[rows], img, txt, scale)
        times.append(time.perf_counter() - t0)
    sec = sorted(times)[len(times) // 2]
    return dict(pairs_per_s=sample_rows / sec, seconds_per_step=sec, cores=threads,
                sample=f"{sample_rows} of {n_total} pairs (rows 0..{sample_rows - 1}) vs all {n_total} candidates, "
                       f"D={d}, fp32, fwd+bwd, median of {steps} after {warmup} warm-up")
