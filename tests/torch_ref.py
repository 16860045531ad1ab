"""Plain PyTorch fp32 evaluation of the reference's losses on the GPU -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Every function restates, operator for operator, what one rank of the reference executes
(``src/open_clip/loss.py``): ``gather_features`` :21-65 (``torch.distributed.nn.all_gather`` when
``gather_with_grad``, plain ``all_gather`` + re-inserted local slot otherwise), ``ClipLoss.get_logits`` :104-126,
``ClipLoss.forward`` :128-139, ``SigLipLoss._loss`` :354-363 summed over every text chunk (its four ``dist_impl``
exchange schemes are the same loss), ``MultiPositiveClipLoss.forward`` :696-747 with
``multi_positive_cross_entropy_loss`` :626-644.  The collectives run on the process group that is already
initialised (NCCL on the GPU box); with ``world == 1`` nothing distributed is touched.

Used by ``tests/dist_parity.py`` (production-size multi-GPU parity) and by ``bench.py``'s pre-timing parity check.
The product (``mrclip_b200/``) never imports this module.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.distributed.nn
import torch.nn.functional as F


def _gather(x, world, with_grad, local_loss, rank):
    """loss.py:48-63 for one tensor."""
    if with_grad:
        return torch.cat(torch.distributed.nn.all_gather(x), dim=0)
    parts = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(parts, x.detach().contiguous())
    if not local_loss:
        parts[rank] = x
    return torch.cat(parts, dim=0)


def _finish(loss, leaves, grad_output):
    (loss * grad_output).backward()
    out = {"loss": loss.detach()}
    for k, v in leaves.items():
        out["d_" + k] = v.grad
    return out


def _leaves(image, text, scale, raw):
    """Leaf tensors and what the loss sees.  raw: the leaves are the towers' un-normalised outputs and the model's
    log-scale parameter; the loss sees F.normalize(.) and scale.exp() (model.py:282-301, :324), gradients flow back."""
    i0 = image.detach().float().clone().requires_grad_(True)
    t0 = text.detach().float().clone().requires_grad_(True)
    s0 = torch.as_tensor(scale, dtype=torch.float32, device=i0.device).detach().clone().requires_grad_(True)
    if raw:
        return (i0, t0, s0), (F.normalize(i0, dim=-1), F.normalize(t0, dim=-1), s0.exp())
    return (i0, t0, s0), (i0, t0, s0)


def clip_reference(image, text, scale, local_loss=False, gather_with_grad=False, rank=0, world=1, grad_output=1.0,
                   raw=False):
    """One rank's ClipLoss forward + backward in fp32 torch.  image/text: this rank's [n, D] rows (any float dtype,
    up-cast to fp32).  Returns dict(loss, d_image, d_text, d_scale) of fp32 tensors on the same device."""
    (i0, t0, s0), (i, t, s) = _leaves(image, text, scale, raw)
    n = i.shape[0]
    if world > 1:
        all_i = _gather(i, world, gather_with_grad, local_loss, rank)
        all_t = _gather(t, world, gather_with_grad, local_loss, rank)
        if local_loss:
            logits_per_image = s * i @ all_t.T
            logits_per_text = s * t @ all_i.T
        else:
            logits_per_image = s * all_i @ all_t.T
            logits_per_text = logits_per_image.T
    else:
        logits_per_image = s * i @ t.T
        logits_per_text = s * t @ i.T
    num = logits_per_image.shape[0]
    labels = torch.arange(num, device=i.device, dtype=torch.long)
    if world > 1 and local_loss:
        labels = labels + num * rank
    loss = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
    out = _finish(loss, {"image": i0, "text": t0, "scale": s0}, grad_output)
    out["labels"] = labels
    return out


def siglip_reference(image, text, scale, bias, rank=0, world=1, grad_output=1.0, raw=False):
    """One rank's SigLipLoss: its image rows against every rank's text chunk, positives in its own chunk only; the text
    features travel with gradient (the reference's neighbour exchanges are ``*_with_grad``, loss.py:279-311)."""
    (i0, t0, s0), (i, t, s) = _leaves(image, text, scale, raw)
    leaves = {"image": i0, "text": t0, "scale": s0}
    b = None
    if bias is not None:
        b = torch.as_tensor(bias, dtype=torch.float32, device=i.device).detach().clone().requires_grad_(True)
        leaves["bias"] = b
    n = i.shape[0]
    all_t = _gather(t, world, True, True, rank) if world > 1 else t
    logits = s * i @ all_t.T
    if b is not None:
        logits = logits + b
    labels = -torch.ones_like(logits)
    idx = torch.arange(n, device=i.device)
    labels[idx, idx + rank * n] = 1.0
    loss = -F.logsigmoid(labels * logits).sum() / n
    return _finish(loss, leaves, grad_output)


def mpos_reference(image, text, scale, labels, delta=0.5, rank=0, world=1, grad_output=1.0):
    """One rank's MultiPositiveClipLoss (local_loss=True, gather_with_grad=True when world > 1)."""
    i = image.detach().float().clone().requires_grad_(True)
    t = text.detach().float().clone().requires_grad_(True)
    s = torch.as_tensor(scale, dtype=torch.float32, device=i.device).detach().clone().requires_grad_(True)
    if world > 1:
        all_i = _gather(i, world, True, True, rank)
        all_t = _gather(t, world, True, True, rank)
        lab_all = torch.empty((world * labels.shape[0],), dtype=labels.dtype, device=labels.device)
        dist.all_gather_into_tensor(lab_all, labels.contiguous())
        logits_per_image = s * i @ all_t.T
        logits_per_text = s * t @ all_i.T
    else:
        lab_all = labels
        logits_per_image = s * i @ t.T
        logits_per_text = s * t @ i.T
    pos_mask = (labels.view(-1, 1) == lab_all.view(1, -1)).float()

    def mp_ce(logits):
        shifted = logits - logits.amax(dim=1, keepdim=True).detach()
        log_prob = shifted - torch.log(shifted.exp().sum(dim=1, keepdim=True) + 1e-12)
        return (-(pos_mask * log_prob).sum(dim=1) / pos_mask.sum(dim=1).clamp(min=1)).mean()

    loss = delta * mp_ce(logits_per_image) + (1.0 - delta) * mp_ce(logits_per_text)
    return _finish(loss, {"image": i, "text": t, "scale": s}, grad_output)


def rel(got, ref):
    """Norm-wise relative error of two tensors (float64 accumulation), as a Python float."""
    g, r = got.detach().double().flatten(), ref.detach().double().flatten()
    den = float(r.norm())
    return float((g - r).norm()) / (den if den > 0 else 1.0)


def compare(ours, ref, tol_loss=1e-3, tol_grad=1e-2):
    """ours / ref: dicts with loss, d_image, d_text, d_scale (, d_bias).  Returns (errors dict, list of keys out of
    tolerance).  Tolerances are BASELINE.json's: loss 1e-3 relative, gradients 1e-2 relative."""
    errs = {"loss": abs(float(ours["loss"]) - float(ref["loss"])) / max(abs(float(ref["loss"])), 1e-30)}
    for k in ("d_image", "d_text"):
        errs[k] = rel(ours[k].float(), ref[k])
    for k in ("d_scale", "d_bias"):
        if k in ref and ours.get(k) is not None:
            errs[k] = abs(float(ours[k]) - float(ref[k])) / max(abs(float(ref[k])), 1e-6)
    bad = [k for k, v in errs.items() if not (v <= (tol_loss if k == "loss" else tol_grad))]
    return errs, bad
