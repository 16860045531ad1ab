"""Retrieval metrics on the GPU: drop-in for the reference's ``get_clip_metrics`` (open_clip_train/train.py:465-534).

The reference moves the features to the CPU, forms the full ``N_val x N_val`` logit matrix, argsorts every row and
walks Python loops to find where the samples of each row's class rank.  Here the S tiles of the loss kernels carry a
rank-of-label epilogue (``csrc/tile_kernel.cuh``, ``MODE_RANK``): nothing ``N x N`` is formed and nothing is sorted.
For row ``i`` with positives ``P(i)`` = the samples that share its label, ``m = |P(i)|``:

    best rank          = #{k not in P(i): S_ik > max_{j in P(i)} S_ij}
    mean positive rank = ( sum_{j in P(i)} #{k not in P(i): S_ik > S_ij}  +  m (m - 1) / 2 ) / m

(the positives rank among themselves at 0 .. m-1 whatever their order); comparisons are strict, i.e. a negative that ties
a positive is ranked after it -- ``torch.argsort`` leaves that case unspecified in the reference.  Same signature, same
dictionary keys; ``logit_scale`` must be positive (ranks do not depend on it).  Features are contracted as bf16 with
fp32 accumulation, like the loss.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi
from ._cabi import Shape

__all__ = ["get_clip_metrics", "rank_statistics", "positive_layout"]


def _dense_classes(labels):
    ids, table = [], {}
    for lab in labels:
        ids.append(table.setdefault(lab, len(table)))
    return np.asarray(ids, dtype=np.int64)


def positive_layout(class_ids):
    """Host-side layout of the positive lists for ``class_ids`` (ints, one per sample, shared by rows and columns).
    Returns (cls int64 [N], m int64 [N] = positives per row incl. the row's own column, ordinal int64 [N] = position of
    sample j inside its class, off int64 [N] = start of row i's list, total = sum(m)): column j of class c lands in slot
    off[i] + ordinal[j] of every row i of class c, so each list is filled exactly once without atomics."""
    cls = np.asarray(class_ids, dtype=np.int64)
    n = cls.shape[0]
    sizes = np.bincount(cls)
    m = sizes[cls]
    order = np.argsort(cls, kind="stable")
    ordinal = np.empty(n, dtype=np.int64)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    ordinal[order] = np.arange(n) - starts[cls[order]]
    off = np.concatenate(([0], np.cumsum(m)[:-1])).astype(np.int64)
    return cls, m, ordinal, off, int(m.sum())


def rank_statistics(row_features, col_features, class_ids):
    """For every row: (best, mean) = 0-based rank of its best positive and mean 0-based rank of all its positives among the
    columns, ``class_ids`` (int array [N], shared by rows and columns) defining the positives.  Returns two float64 numpy
    arrays of length N.  Everything N x N runs in ``tile_kernel<MODE_RANK>``."""
    from .engine import default_engine
    eng = default_engine()
    lib = eng.lib
    dev = row_features.device
    n, d = row_features.shape
    assert col_features.shape == (n, d) and len(class_ids) == n
    cls, m, ordinal, off, total = positive_layout(class_ids)
    ld, npad = eng.padded_dim(d), eng.padded_cols(n)
    a = torch.zeros((n, ld), dtype=torch.bfloat16, device=dev)
    b = torch.zeros((n, ld), dtype=torch.bfloat16, device=dev)
    eng.pack(row_features.detach() if row_features.stride(1) == 1 else row_features.detach().contiguous(), a)
    eng.pack(col_features.detach() if col_features.stride(1) == 1 else col_features.detach().contiguous(), b)
    col_cls = np.full(npad, -1, dtype=np.int32)
    col_cls[:n] = cls
    col_ord = np.zeros(npad, dtype=np.int32)
    col_ord[:n] = ordinal
    t = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x), dtype=dt).to(dev)
    row_cls_d, col_cls_d, col_ord_d = t(cls, torch.int32), t(col_cls, torch.int32), t(col_ord, torch.int32)
    off_d, m_d = t(off, torch.int64), t(m, torch.int32)
    pos = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    lmax = torch.empty(n, dtype=torch.float32, device=dev)
    pairs = torch.zeros(n, dtype=torch.int64, device=dev)
    best = torch.zeros(n, dtype=torch.int32, device=dev)
    shape = Shape(n, n, d, 0)
    st = torch.cuda.current_stream().cuda_stream
    _cabi.check(lib.mrclip_rank_collect(a.data_ptr(), b.data_ptr(), shape, ld, row_cls_d.data_ptr(), col_cls_d.data_ptr(),
                                        col_ord_d.data_ptr(), off_d.data_ptr(), pos.data_ptr(), st))
    _cabi.check(lib.mrclip_rank_lmax(pos.data_ptr(), off_d.data_ptr(), m_d.data_ptr(), n, lmax.data_ptr(), st))
    for chunk0 in range(0, int(m.max()), 32):                   # 32 positives per row are compared per pass
        _cabi.check(lib.mrclip_rank_count(a.data_ptr(), b.data_ptr(), shape, ld, row_cls_d.data_ptr(), col_cls_d.data_ptr(),
                                          off_d.data_ptr(), m_d.data_ptr(), pos.data_ptr(), lmax.data_ptr(), chunk0,
                                          pairs.data_ptr(), best.data_ptr(), st))
    best_h = best.cpu().numpy().astype(np.float64)
    pairs_h = pairs.cpu().numpy().astype(np.float64)
    mean_h = (pairs_h + m * (m - 1) / 2.0) / m
    return best_h, mean_h


def get_clip_metrics(image_features, text_features, logit_scale, ground_truth_general, ground_truth_unique=None,
                     trace=False):
    """Reference ``get_clip_metrics`` (train.py:465-534): mean / median / mean-of-mean rank and Recall@{1,5,10} of the
    class-matching samples, image->text and text->image, for the general (and, if given, the unique) label set."""
    if float(logit_scale) <= 0:
        raise ValueError("logit_scale must be positive")
    dev = torch.device("cuda", torch.cuda.current_device())
    img = torch.as_tensor(image_features).to(dev)
    txt = torch.as_tensor(text_features).to(dev)
    metrics, vocabulary = {}, {}
    for which, ground_truth in (("general", ground_truth_general), ("unique", ground_truth_unique)):
        if ground_truth is None:
            continue
        cls = _dense_classes(list(ground_truth))
        for name, (rows, cols) in (("image_to_text", (img, txt)), ("text_to_image", (txt, img))):
            key = f"{name}_{which}"
            preds, preds_mean = rank_statistics(rows, cols, cls)
            metrics[f"{key}_meanofmean_rank"] = preds_mean.mean() + 1
            metrics[f"{key}_mean_rank"] = preds.mean() + 1
            metrics[f"{key}_median_rank"] = np.floor(np.median(preds)) + 1
            for k in [1, 5, 10]:
                metrics[f"{key}_R@{k}"] = np.mean(preds < k)
            if trace and which == "general":
                # the reference lists the ten best columns of the first 201 rows (train.py:512-531): a 201 x N block
                head = min(len(ground_truth), 201)
                top = torch.topk(rows[:head].float() @ cols.float().t(), k=min(10, cols.shape[0]), dim=1).indices.cpu().tolist()
                gt = list(ground_truth)
                vocabulary[key] = {i: {"anchor": i, "gt": gt[i], "indices": top[i], "labels": [gt[j] for j in top[i]]}
                                   for i in range(head)}
    if trace:
        return metrics, vocabulary
    return metrics
