"""Retrieval metrics (reference open_clip_train/train.py:465-534).

CPU: the float64 oracle (oracle/metrics_oracle.py) against what the unmodified reference function returned
(tests/golden/metrics/*.npz, recorded by oracle/gen_golden_metrics.py), and the sort-free counting identities the GPU
epilogue uses against the oracle.  GPU (``-m gpu``): ``mrclip_b200.metrics.get_clip_metrics`` through the C ABI against
the same fixtures and, at N = 4096, against the oracle's identities evaluated with torch."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, has_b200
from oracle.metrics_oracle import clip_metrics_oracle

FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "metrics", "*.npz")))


def _load(path):
    z = np.load(path)
    uniq = list(z["unique"]) if bool(z["has_unique"]) else None
    want = {k[2:]: float(z[k]) for k in z.files if k.startswith("m_")}
    return z["image"], z["text"], float(z["scale"]), list(z["general"]), uniq, want


def _counting_identities(s, cls):
    """best / mean 0-based rank of the positives of every row from strict pair counts (what tile_kernel<MODE_RANK> does)"""
    n = s.shape[0]
    best, mean = np.zeros(n), np.zeros(n)
    for i in range(n):
        pos = cls == cls[i]
        p, neg = s[i, pos], s[i, ~pos]
        m = p.size
        best[i] = (neg > p.max()).sum()
        mean[i] = ((neg[None, :] > p[:, None]).sum() + m * (m - 1) / 2.0) / m
    return best, mean


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_matches_reference_function(path):
    img, txt, scale, gen, uniq, want = _load(path)
    got = clip_metrics_oracle(img, txt, scale, gen, uniq)
    assert set(got) == set(want)
    for k in want:      # the reference contracts in fp32 and averages rank positions in fp32: 5e-6 on the mean ranks
        assert abs(got[k] - want[k]) <= 5e-6 * max(1.0, abs(want[k])), k


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_counting_identities_match_the_sorted_ranks(path):
    img, txt, scale, gen, uniq, want = _load(path)
    s = scale * img.astype(np.float64) @ txt.astype(np.float64).T
    for which, gt in (("general", gen), ("unique", uniq)):
        if gt is None:
            continue
        cls = np.asarray(gt)
        for name, mat in (("image_to_text", s), ("text_to_image", s.T)):
            best, mean = _counting_identities(mat, cls)
            assert abs(best.mean() + 1 - want[f"{name}_{which}_mean_rank"]) < 5e-6 * want[f"{name}_{which}_mean_rank"]
            assert abs(mean.mean() + 1 - want[f"{name}_{which}_meanofmean_rank"]) < 5e-6 * want[f"{name}_{which}_meanofmean_rank"]
            assert np.floor(np.median(best)) + 1 == want[f"{name}_{which}_median_rank"]
            for k in (1, 5, 10):
                assert abs(np.mean(best < k) - want[f"{name}_{which}_R@{k}"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_gpu_metrics_match_reference_function(path):
    if not has_b200():
        pytest.fail("needs a B200 (sm_100a)")
    from mrclip_b200.metrics import get_clip_metrics
    img, txt, scale, gen, uniq, want = _load(path)
    got = get_clip_metrics(torch.from_numpy(img), torch.from_numpy(txt), torch.tensor(scale), gen, uniq)
    assert set(got) == set(want)
    for k in want:           # integers underneath; the reference's own fp32 contraction / fp32 means allow 5e-6
        assert abs(float(got[k]) - want[k]) <= 5e-6 * max(1.0, abs(want[k])), (k, got[k], want[k])
    m2, voc = get_clip_metrics(torch.from_numpy(img), torch.from_numpy(txt), torch.tensor(scale), gen, uniq, trace=True)
    assert m2.keys() == got.keys() and set(voc) == {"image_to_text_general", "text_to_image_general"}
    first = voc["image_to_text_general"][0]
    assert first["anchor"] == 0 and first["gt"] == gen[0] and len(first["indices"]) == min(10, len(gen))
    s0 = img[0].astype(np.float64) @ txt.astype(np.float64).T
    assert abs(s0[first["indices"][0]] - s0.max()) <= 1e-6           # (duplicate captions tie exactly)


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,classes", [(4096, 512, 37), (5000, 200, 5000), (3000, 768, 3)])
def test_gpu_rank_statistics_large(n, d, classes):
    """class sizes from 1 (singletons) to ~1000 (several 32-positive passes), ragged N and D; against the counting
    identities evaluated on the same bf16 features in float64 (near-ties may flip under fp32 accumulation: 1e-4)."""
    if not has_b200():
        pytest.fail("needs a B200 (sm_100a)")
    from mrclip_b200.metrics import rank_statistics
    g = torch.Generator().manual_seed(n + d)
    cls = torch.randint(0, classes, (n,), generator=g).numpy() if classes < n else np.arange(n)
    img = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(0.2 * img + 0.8 * torch.randn(n, d, generator=g) / d ** 0.5, dim=-1)
    img, txt = img.bfloat16(), txt.bfloat16()
    best, mean = rank_statistics(img.cuda(), txt.cuda(), cls)
    s = (img.double().cuda() @ txt.double().cuda().t())
    c = torch.from_numpy(cls).cuda()
    pos = c[:, None] == c[None, :]
    pmax = torch.where(pos, s, torch.full_like(s, -1e30)).max(dim=1).values
    best_ref = ((s > pmax[:, None]) & ~pos).sum(dim=1).double().cpu().numpy()
    assert np.abs(best - best_ref).max() <= 1 and abs(best.mean() - best_ref.mean()) <= 1e-4 * max(best_ref.mean(), 1.0)
    m = pos.sum(dim=1).double()
    pairs_ref = torch.zeros(n, dtype=torch.float64, device="cuda")
    for i0 in range(0, n, 256):             # sum over positives t of #{negatives above t}
        blk = slice(i0, min(i0 + 256, n))
        sb, pb = s[blk], pos[blk]
        srt = torch.sort(torch.where(pb, torch.full_like(sb, 1e30), sb), dim=1).values      # negatives ascending, positives last
        nneg = (~pb).sum(dim=1)
        idx = torch.searchsorted(srt, torch.where(pb, sb, torch.full_like(sb, 1e30)), right=True)
        above = (nneg[:, None] - idx).clamp(min=0).double()
        pairs_ref[blk] = torch.where(pb, above, torch.zeros_like(above)).sum(dim=1)
    mean_ref = ((pairs_ref + m * (m - 1) / 2) / m).cpu().numpy()
    assert abs(mean.mean() - mean_ref.mean()) <= 1e-4 * mean_ref.mean()
    assert np.abs(mean - mean_ref).max() <= 1e-3 * max(mean_ref.max(), 1.0) + 1.0


def test_positive_layout_fills_every_list_exactly_once():
    """Host side of the rank epilogue: slot off[i] + ordinal[j] over the columns j of row i's class is a bijection onto
    row i's list [off[i], off[i] + m[i]), and the lists tile [0, total) without gaps or overlaps."""
    from mrclip_b200.metrics import positive_layout
    rng = np.random.default_rng(0)
    for n, classes in ((1, 1), (17, 1), (64, 64), (300, 7), (257, 40)):
        ids = rng.integers(0, classes, size=n) if classes < n else rng.permutation(n)
        cls, m, ordinal, off, total = positive_layout(ids)
        assert total == int(sum(np.bincount(cls) ** 2)) and off[0] == 0
        assert np.array_equal(off[1:], np.cumsum(m)[:-1])
        for i in range(n):
            cols = np.nonzero(cls == cls[i])[0]
            assert m[i] == cols.size
            assert sorted(off[i] + ordinal[cols]) == list(range(off[i], off[i] + m[i]))
