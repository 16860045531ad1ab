"""Executable specification of the emat backend's numerics (CPU, numpy): what the kernels compute, restated with the same
roundings, checked against the float64 oracle.  It pins the *design* -- bf16 exponentials relative to 32 x 64 sub-tile
references, exact positives, bf16 G, entropy-form d_scale, the flush guard -- independently of any GPU
(csrc/tile_kernel.cuh MODE_FWDE, aux_kernels.cuh emat_transform_kernel / emat_check_kernel)."""
import numpy as np
import pytest

from oracle.clip_oracle import bf16_round, clip_loss_oracle

LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


def _feats(n, d, seed, corr):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    y = corr * x + (1 - corr) * rng.standard_normal((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    return bf16_round(x.astype(np.float32)).astype(np.float64), bf16_round(y.astype(np.float32)).astype(np.float64)


def emat_pipeline(img, txt, scale):
    """one rank, ClipLoss: returns loss, dI, dT, d_scale and the guard flag the device check would raise"""
    n = img.shape[0]
    s2 = (scale * LOG2E) * (img @ txt.T).astype(np.float32).astype(np.float64)      # fp32 accumulators, log2 units
    # forward: sub-tile references, exponentials in fp32, statistics in fp32, E stored as bf16
    ref = np.zeros_like(s2)
    for b in range(0, n, 32):
        for c in range(0, n, 64):
            ref[b:b + 32, c:c + 64] = s2[b:b + 32, c:c + 64].max()
    e32 = np.exp2(s2 - ref)
    lse_r = np.log2((e32 * np.exp2(ref - ref.max(axis=1, keepdims=True))).sum(axis=1)) + ref.max(axis=1)
    lse_c = np.log2((e32 * np.exp2(ref - ref.max(axis=0, keepdims=True))).sum(axis=0)) + ref.max(axis=0)
    diag2 = np.diag(s2).copy()
    loss = LN2 / (2 * n) * (lse_r + lse_c - 2 * diag2).sum()
    e = bf16_round(e32.astype(np.float32)).astype(np.float64)
    # guard (emat_check_kernel): a flushed entry can matter only if an LSE lies > 80 orders below its reference
    flag = False
    for b in range(0, n, 32):
        for c in range(0, n, 64):
            r = ref[b, c]
            flag |= (r - lse_r[b:b + 32].min() > 80) or (r - lse_c[c:c + 64].min() > 80)
    # rescale pass: G = E * (2^(c - lse_r) + 2^(c - lse_c)), positives exactly, rounded to bf16 for the GEMMs
    g = e * (np.exp2(np.minimum(ref - lse_r[:, None], 120)) + np.exp2(np.minimum(ref - lse_c[None, :], 120)))
    idx = np.arange(n)
    g[idx, idx] = np.exp2(diag2 - lse_r) + np.exp2(diag2 - lse_c) - 2.0
    g = bf16_round(g.astype(np.float32)).astype(np.float64)
    coef = 0.5 / n
    d_img = coef * scale * g @ txt
    d_txt = coef * scale * g.T @ img
    # d_scale from the entropies of the stored exponentials (positives exact)
    with np.errstate(divide="ignore", invalid="ignore"):
        l2e = np.where(e > 0, np.log2(e), 0.0) + ref
    pr, pc = e * np.exp2(ref - lse_r[:, None]), e * np.exp2(ref - lse_c[None, :])
    tr, tc = pr * (l2e - lse_r[:, None]), pc * (l2e - lse_c[None, :])
    lpr, lpc = diag2 - lse_r, diag2 - lse_c
    tr[idx, idx], tc[idx, idx] = np.exp2(lpr) * lpr, np.exp2(lpc) * lpc
    d_scale = (loss + LN2 * coef * (tr.sum() + tc.sum())) / scale
    dot_scale = (d_img * img).sum() / scale          # the homogeneity shortcut used for one large rank
    return dict(loss=loss, d_image=d_img, d_text=d_txt, d_scale=d_scale, dot_scale=dot_scale, flag=flag)


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("n,d,scale,corr", [(256, 64, 14.285714, 0.3), (512, 96, 100.0, 0.1), (384, 48, 30.0, 0.9)])
def test_emat_numerics_match_the_oracle(n, d, scale, corr):
    img, txt = _feats(n, d, 100 + n, corr)
    ref = clip_loss_oracle([img], [txt], scale)[0]
    out = emat_pipeline(img, txt, scale)
    assert not out["flag"]
    assert abs(out["loss"] - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    assert _rel(out["d_image"], ref["d_image"]) <= 4e-3 and _rel(out["d_text"], ref["d_text"]) <= 4e-3
    assert abs(out["d_scale"] - ref["d_logit_scale"]) <= 2e-3 * abs(ref["d_logit_scale"]) + 1e-9
    # the dot shortcut carries the bf16 noise of G; it is only used for n*N >= 2^22
    assert abs(out["dot_scale"] - ref["d_logit_scale"]) <= 5e-2 * abs(ref["d_logit_scale"]) + 1e-6


def test_confident_model_needs_the_exact_positives():
    """a nearly converged model (P_ii ~ 0.99): G_ii = P_r + P_c - 2 from bf16 E would lose its leading digits"""
    n, d, scale = 256, 64, 100.0
    img, _ = _feats(n, d, 7, 0.0)
    txt = img.copy()
    ref = clip_loss_oracle([img], [txt], scale)[0]
    out = emat_pipeline(img, txt, scale)
    assert abs(out["loss"] - ref["loss"]) <= 1e-3 * abs(ref["loss"])
    assert _rel(out["d_image"], ref["d_image"]) <= 1e-2 and _rel(out["d_text"], ref["d_text"]) <= 1e-2


def test_guard_condition_fires_when_a_reference_dwarfs_its_band():
    n, d, scale = 256, 64, 100.0
    img, txt = _feats(n, d, 5, 0.0)
    img[0] = 0.0
    img[0, 0] = 1.0
    txt[0] = img[0]                         # S2_00 = 144, every other logit of the band is ~ +-15
    out = emat_pipeline(img, txt, scale)
    assert out["flag"]


def test_forward_side_row_sums_online_merge():
    """MODE_FWDEU's bookkeeping (tile_kernel.cuh) and row_ent_split_kernel, restated in fp32: per 64-column sub-tile
    u += sum e*(S2 - c) + c*sum e, merged across sub-tiles with the LSE's rescaling factors, one (m, l, u) triple per
    (row, column-chunk half); the split by column owner then gives R2(me, q) = sum_{i, j in q} Prow_ij * S2_ij."""
    n, N, d, scale, ranks = 128, 1024, 64, 100.0, 4
    img, txt = _feats(N, d, 3, 0.5)
    s2 = ((scale * LOG2E) * (img[:n] @ txt.T)).astype(np.float32)
    f32 = np.float32
    slot_cols = 128                                    # one chunk half (tiles_per_chunk = 1)
    slots = N // slot_cols
    m = np.full((slots, n), -np.inf, f32)
    l = np.zeros((slots, n), f32)
    u = np.zeros((slots, n), f32)
    for s in range(slots):
        for sub in range(slot_cols // 64):
            cols = slice(s * slot_cols + sub * 64, s * slot_cols + sub * 64 + 64)
            for b in range(0, n, 32):
                blk = s2[b:b + 32, cols]
                c = f32(blk.max())
                t = (blk - c).astype(f32)
                e = np.exp2(t).astype(f32)
                rowsum, row_t = e.sum(1, dtype=f32), (e * t).sum(1, dtype=f32)
                mnew = np.maximum(m[s, b:b + 32], c)
                f_old, f_new = np.exp2(m[s, b:b + 32] - mnew).astype(f32), np.exp2(c - mnew).astype(f32)
                l[s, b:b + 32] = l[s, b:b + 32] * f_old + rowsum * f_new
                u[s, b:b + 32] = u[s, b:b + 32] * f_old + (row_t + c * rowsum) * f_new
                m[s, b:b + 32] = mnew
    mx = m.max(0)
    lse = mx + np.log2((l * np.exp2(m - mx)).sum(0))
    per = slots // ranks
    got = np.array([(np.exp2(m[q * per:(q + 1) * per] - lse) * u[q * per:(q + 1) * per]).sum() for q in range(ranks)])
    s2d = s2.astype(np.float64)
    lse_ref = np.log2(np.exp2(s2d - s2d.max(1, keepdims=True)).sum(1)) + s2d.max(1)
    p = np.exp2(s2d - lse_ref[:, None])
    want = np.array([(p[:, q * (N // ranks):(q + 1) * (N // ranks)] * s2d[:, q * (N // ranks):(q + 1) * (N // ranks)]).sum()
                     for q in range(ranks)])
    assert np.abs(lse - lse_ref).max() < 1e-4
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
