"""CPU oracle for MR-CLIP's distributed contrastive loss path  --  TEST INFRASTRUCTURE ONLY.

A float64 numpy restatement of what the reference computes in
``/root/reference/src/open_clip/loss.py`` (``gather_features`` :21-65, ``ClipLoss`` :68-139,
``SigLipLoss`` :314-448), including what its autograd graph hands back on every rank.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may import
this package; the product (``mrclip_b200``) never does.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the unmodified reference (single process and
multi-process gloo groups) in the build container and stores its inputs/outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against every stored vector.

The restatement is written rank by rank, the way the reference executes: each rank builds its own
graph from its local features and the gathered copies, and the backward of the collective
(``torch.distributed.nn.all_gather`` -> reduce-scatter-sum, torch/distributed/nn/functional.py:327-354;
``NeighbourExchange*`` -> reverse exchange, loss.py:279-311) routes gradients between ranks.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "ground_truth_labels",
    "cross_entropy_mean",
    "clip_loss_oracle",
    "siglip_loss_oracle",
    "multipositive_loss_oracle",
    "bf16_round",
]


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round float32 values to the nearest bfloat16 (ties to even), returned as float32."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def ground_truth_labels(num_logits: int, rank: int = 0, world_size: int = 1, local_loss: bool = False) -> np.ndarray:
    """loss.py:91-102 -- arange(num_logits) (+ num_logits*rank when world_size>1 and local_loss)."""
    labels = np.arange(num_logits, dtype=np.int64)
    if world_size > 1 and local_loss:
        labels = labels + num_logits * rank
    return labels


def _lse(z: np.ndarray, axis: int) -> np.ndarray:
    m = z.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def cross_entropy_mean(logits: np.ndarray, labels: np.ndarray):
    """F.cross_entropy(logits, labels) with mean reduction (loss.py:135-136) and d loss / d logits."""
    m = logits.shape[0]
    lse = _lse(logits, 1)
    picked = logits[np.arange(m), labels]
    loss = float((lse - picked).mean())
    dz = np.exp(logits - lse[:, None])
    dz[np.arange(m), labels] -= 1.0
    return loss, dz / m


def _as_parts(x, world_size):
    parts = [np.asarray(p, dtype=np.float64) for p in x]
    assert len(parts) == world_size
    return parts


def clip_loss_oracle(image_parts, text_parts, logit_scale: float, local_loss: bool = False,
                     gather_with_grad: bool = False, grad_output: float = 1.0):
    """Per-rank results of ``ClipLoss(local_loss, gather_with_grad, rank=r, world_size=W)`` fwd+bwd.

    image_parts / text_parts: length-W sequences of [n, D] arrays (rank r's local features).
    Returns a list (one dict per rank) with keys loss, d_image, d_text, d_logit_scale, labels.
    Follows get_logits (loss.py:104-126) / forward (:128-139) on every rank and then the
    collective's backward for the chosen gather mode (:50-63).
    """
    W = len(image_parts)
    I = _as_parts(image_parts, W)
    T = _as_parts(text_parts, W)
    n, _ = I[0].shape
    s = float(logit_scale)
    I_all = np.concatenate(I, 0)
    T_all = np.concatenate(T, 0)

    per_rank = []
    for r in range(W):
        rows = slice(r * n, (r + 1) * n)
        g_loc_i = np.zeros_like(I[r])
        g_loc_t = np.zeros_like(T[r])
        g_all_i = np.zeros_like(I_all)   # gradient w.r.t. the gathered image copy on rank r
        g_all_t = np.zeros_like(T_all)
        if W > 1 and local_loss:
            # logits_per_image = s * I_r @ T_all^T ; logits_per_text = s * T_r @ I_all^T   (:116-118)
            labels = ground_truth_labels(n, r, W, True)
            c_img = I[r] @ T_all.T
            c_txt = T[r] @ I_all.T
            l_img, dz_img = cross_entropy_mean(s * c_img, labels)
            l_txt, dz_txt = cross_entropy_mean(s * c_txt, labels)
            dz_img *= 0.5
            dz_txt *= 0.5
            g_loc_i += s * dz_img @ T_all
            g_all_t += s * dz_img.T @ I[r]
            g_loc_t += s * dz_txt @ I_all
            g_all_i += s * dz_txt.T @ T[r]
            g_s = float((dz_img * c_img).sum() + (dz_txt * c_txt).sum())
        else:
            # W>1 global: s * I_all @ T_all^T and its transpose (:119-121); W==1: both products (:123-124)
            labels = ground_truth_labels(I_all.shape[0], r, W, False)
            c = I_all @ T_all.T
            l_img, dz_img = cross_entropy_mean(s * c, labels)
            l_txt, dz_txt = cross_entropy_mean(s * c.T, labels)
            dz = 0.5 * (dz_img + dz_txt.T)
            g_all_i += s * dz @ T_all
            g_all_t += s * dz.T @ I_all
            g_s = float((dz * c).sum())
        per_rank.append(dict(loss=0.5 * (l_img + l_txt), labels=labels, g_loc_i=g_loc_i, g_loc_t=g_loc_t,
                             g_all_i=g_all_i, g_all_t=g_all_t, g_s=g_s))

    out = []
    for r in range(W):
        rows = slice(r * n, (r + 1) * n)
        pr = per_rank[r]
        if W == 1:
            d_i, d_t = pr["g_all_i"], pr["g_all_t"]
        elif gather_with_grad:
            # all_gather backward = reduce_scatter(SUM) of every rank's gradient for slot r  (:51-52)
            d_i = pr["g_loc_i"] + sum(q["g_all_i"][rows] for q in per_rank)
            d_t = pr["g_loc_t"] + sum(q["g_all_t"][rows] for q in per_rank)
        elif not local_loss:
            # no-grad gather, own slot re-inserted (:58-61): only this rank's graph reaches the leaf
            d_i, d_t = pr["g_all_i"][rows], pr["g_all_t"][rows]
        else:
            d_i, d_t = pr["g_loc_i"], pr["g_loc_t"]
        out.append(dict(loss=pr["loss"], labels=pr["labels"], d_image=grad_output * d_i,
                        d_text=grad_output * d_t, d_logit_scale=grad_output * pr["g_s"]))
    return out


def _softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def siglip_loss_oracle(image_parts, text_parts, logit_scale: float, logit_bias: float,
                       grad_output: float = 1.0):
    """Per-rank results of ``SigLipLoss(rank=r, world_size=W, dist_impl=*)`` fwd+bwd.

    Every dist_impl sums, on rank r, ``_loss(I_r, T_c)`` over all chunks c with positives only on the
    own chunk (loss.py:365-446); text gradients travel back to the chunk's owner.
    """
    W = len(image_parts)
    I = _as_parts(image_parts, W)
    T = _as_parts(text_parts, W)
    n = I[0].shape[0]
    s, b = float(logit_scale), float(logit_bias)
    d_text = [np.zeros_like(t) for t in T]
    out = []
    for r in range(W):
        loss = 0.0
        d_i = np.zeros_like(I[r])
        g_s = 0.0
        g_b = 0.0
        for c in range(W):
            cos = I[r] @ T[c].T
            y = -np.ones((n, n))
            if c == r:
                y = 2.0 * np.eye(n) + y                     # get_ground_truth (:338-342)
            z = s * cos + b                                  # get_logits (:344-348)
            loss += float(_softplus(-y * z).sum() / n)       # -logsigmoid(labels*logits).sum()/n (:362)
            g = (-y / (1.0 + np.exp(y * z))) / n             # d loss / d z
            d_i += s * g @ T[c]
            d_text[c] += s * g.T @ I[r]
            g_s += float((g * cos).sum())
            g_b += float(g.sum())
        out.append(dict(loss=loss, d_image=grad_output * d_i, d_logit_scale=grad_output * g_s,
                        d_logit_bias=grad_output * g_b))
    for r in range(W):
        out[r]["d_text"] = grad_output * d_text[r]
    return out


def multipositive_loss_oracle(image_parts, text_parts, label_parts, logit_scale: float, delta: float = 0.5,
                              grad_output: float = 1.0):
    """MR-CLIP's ``MultiPositiveClipLoss`` (loss.py:671-747 with ``multi_positive_cross_entropy_loss`` :626-644):
    samples that share a label are positives of each other.

    Per rank r (``local_loss=True, gather_with_grad=True`` when world_size > 1; plain when 1):
        pos_mask = labels_r[:, None] == labels_all[None, :]                                  (:707-712)
        loss_img = mean_i( -(pos_mask * log_softmax(s I_r T_all^T))_i.sum() / count_i )      (:626-644; the +1e-12 inside
        loss_txt = the same on s T_r I_all^T with the same mask                               the log is below fp32 eps)
        L_r = delta * loss_img + (1 - delta) * loss_txt                                      (:745)
    Gradients follow the reference's autograd graph: the gathered copies carry gradient, and
    ``torch.distributed.nn.all_gather``'s backward sums every rank's contribution (reduce-scatter).
    Returns one dict per rank: loss, d_image, d_text, d_logit_scale.
    """
    W = len(image_parts)
    I = [np.asarray(x, dtype=np.float64) for x in image_parts]
    T = [np.asarray(x, dtype=np.float64) for x in text_parts]
    L = [np.asarray(x).astype(np.int64) for x in label_parts]
    I_all, T_all, L_all = np.concatenate(I), np.concatenate(T), np.concatenate(L)
    n = I[0].shape[0]
    s = float(logit_scale)
    out = []
    dI_all = np.zeros_like(I_all)
    dT_all = np.zeros_like(T_all)
    for r in range(W):
        rows = slice(r * n, (r + 1) * n)
        mask = (L[r][:, None] == L_all[None, :]).astype(np.float64)
        cnt = np.maximum(mask.sum(axis=1), 1.0)
        c_img = I[r] @ T_all.T          # [n, N]
        c_txt = T[r] @ I_all.T
        terms = []
        loss = 0.0
        ds = 0.0
        for c, w in ((c_img, delta), (c_txt, 1.0 - delta)):
            z = s * c
            logp = z - _lse(z, 1)[:, None]
            loss += w * float((-(mask * logp).sum(axis=1) / cnt).mean())
            # d loss / d z = w/n * (softmax * (sum_j mask_ij / cnt_i) - mask / cnt) = w/n * (softmax - mask/cnt)
            g = w / n * (np.exp(logp) - mask / cnt[:, None])
            terms.append(g)
            ds += float((g * c).sum())
        g_img, g_txt = terms
        # image-direction graph: I_r (local) x T_all (gathered, with grad)
        dI_all[rows] += s * g_img @ T_all
        dT_all += s * g_img.T @ I[r]
        # text-direction graph: T_r (local) x I_all (gathered, with grad)
        dT_all[rows] += s * g_txt @ I_all
        dI_all += s * g_txt.T @ T[r]
        out.append(dict(loss=loss, d_logit_scale=grad_output * ds))
    for r in range(W):
        rows = slice(r * n, (r + 1) * n)
        out[r]["d_image"] = grad_output * dI_all[rows]
        out[r]["d_text"] = grad_output * dT_all[rows]
    return out
