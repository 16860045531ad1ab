"""Float64 restatement of the reference's retrieval metrics  --  TEST INFRASTRUCTURE ONLY.

Follows ``get_clip_metrics`` (open_clip_train/train.py:465-534) step by step: logits ``scale * I @ T.T`` (:486) and their
transpose, ``argsort(descending)`` per row (:495), the positions of the samples that share the row's label (:501-504),
their minimum and mean (:507-508), then mean-of-mean / mean / median rank (+1) and Recall@{1,5,10} (:511-516).
``tests/test_oracle.py`` pins it against tests/golden/metrics_*.npz, which ``oracle/gen_golden.py`` recorded from the
unmodified reference function.  Ties are broken the way ``numpy.argsort(kind="stable")`` on the negated row does; the
fixtures have no ties between a positive and a negative.
"""
import numpy as np


def clip_metrics_oracle(image_features, text_features, logit_scale, ground_truth_general, ground_truth_unique=None):
    img = np.asarray(image_features, dtype=np.float64)
    txt = np.asarray(text_features, dtype=np.float64)
    logits_per_image = float(logit_scale) * img @ txt.T
    logits = {"image_to_text": logits_per_image, "text_to_image": logits_per_image.T}
    metrics = {}
    for which, ground_truth in (("general", ground_truth_general), ("unique", ground_truth_unique)):
        if ground_truth is None:
            continue
        table = {}
        gt = np.asarray([table.setdefault(x, len(table)) for x in ground_truth])      # labels may be strings
        for name, logit in logits.items():
            ranking = np.argsort(-logit, axis=1, kind="stable")
            preds, preds_mean = [], []
            for i in range(len(gt)):
                positions = np.nonzero(np.isin(ranking[i], np.nonzero(gt == gt[i])[0]))[0]
                preds.append(positions.min())
                preds_mean.append(positions.astype(np.float64).mean())
            preds, preds_mean = np.asarray(preds), np.asarray(preds_mean)
            key = f"{name}_{which}"
            metrics[f"{key}_meanofmean_rank"] = preds_mean.mean() + 1
            metrics[f"{key}_mean_rank"] = preds.mean() + 1
            metrics[f"{key}_median_rank"] = np.floor(np.median(preds)) + 1
            for k in [1, 5, 10]:
                metrics[f"{key}_R@{k}"] = np.mean(preds < k)
    return metrics
