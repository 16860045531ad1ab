"""Size-independent properties of the oracle (CPU only): the conventions of SURVEY.md section 3a that the GPU tests
rely on, checked on the float64 restatement itself with hypothesis-drawn shapes, plus a finite-difference check of
its gradients.  None of this touches the product; it pins the checker."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle.clip_oracle import clip_loss_oracle, multipositive_loss_oracle, siglip_loss_oracle


def _feats(rng, n, d):
    x = rng.standard_normal((n, d))
    y = 0.4 * x + 0.6 * rng.standard_normal((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    return x, y


def _parts(x, w):
    n = x.shape[0] // w
    return [x[r * n:(r + 1) * n] for r in range(w)]


@settings(max_examples=12, deadline=None)
@given(st.integers(1, 4), st.integers(2, 6), st.integers(2, 9), st.integers(0, 10_000))
def test_rank_partition_conventions(w, n, d, seed):
    """mean over ranks of the local losses = the single-rank loss; gather_with_grad hands every rank W x the
    global-mean gradient; the per-rank d_scale of the local-loss modes sums to W x the global one."""
    rng = np.random.default_rng(seed)
    img, txt = _feats(rng, w * n, d)
    one = clip_loss_oracle([img], [txt], 9.0)[0]
    tt = clip_loss_oracle(_parts(img, w), _parts(txt, w), 9.0, True, True)
    ft = clip_loss_oracle(_parts(img, w), _parts(txt, w), 9.0, False, True)
    ff = clip_loss_oracle(_parts(img, w), _parts(txt, w), 9.0, False, False)
    assert np.isclose(np.mean([o["loss"] for o in tt]), one["loss"], rtol=1e-10)
    for r in range(w):
        rows = slice(r * n, (r + 1) * n)
        assert np.isclose(ft[r]["loss"], one["loss"], rtol=1e-10) and np.isclose(ff[r]["loss"], one["loss"], rtol=1e-10)
        for o, mult in ((tt[r], w), (ft[r], w), (ff[r], 1)):
            assert np.allclose(o["d_image"], mult * one["d_image"][rows], rtol=1e-8, atol=1e-12)
            assert np.allclose(o["d_text"], mult * one["d_text"][rows], rtol=1e-8, atol=1e-12)
        assert np.isclose(ft[r]["d_logit_scale"], one["d_logit_scale"], rtol=1e-8, atol=1e-12)
    assert np.isclose(sum(o["d_logit_scale"] for o in tt), w * one["d_logit_scale"], rtol=1e-8, atol=1e-12)


@settings(max_examples=10, deadline=None)
@given(st.integers(3, 12), st.integers(2, 8), st.integers(0, 10_000))
def test_permutation_equivariance_and_homogeneity(n, d, seed):
    rng = np.random.default_rng(seed)
    img, txt = _feats(rng, n, d)
    s = 11.0
    base = clip_loss_oracle([img], [txt], s)[0]
    perm = rng.permutation(n)
    p = clip_loss_oracle([img[perm]], [txt[perm]], s)[0]
    assert np.isclose(p["loss"], base["loss"], rtol=1e-12)
    assert np.allclose(p["d_image"], base["d_image"][perm], rtol=1e-9, atol=1e-14)
    # S = s * I T^T is homogeneous in s, I and T:  s dL/ds = <dI, I> = <dT, T>
    assert np.isclose(s * base["d_logit_scale"], (base["d_image"] * img).sum(), rtol=1e-8, atol=1e-12)
    assert np.isclose(s * base["d_logit_scale"], (base["d_text"] * txt).sum(), rtol=1e-8, atol=1e-12)


def _fd(fn, x, eps=1e-6):
    g = np.zeros_like(x)
    for idx in np.ndindex(*x.shape):
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        g[idx] = (fn(xp) - fn(xm)) / (2 * eps)
    return g


def test_oracle_gradients_by_finite_differences():
    rng = np.random.default_rng(3)
    img, txt = _feats(rng, 5, 4)
    lab = np.array([0, 1, 0, 2, 1])
    s, b, delta = 7.0, -3.0, 0.3
    c = clip_loss_oracle([img], [txt], s)[0]
    assert np.allclose(c["d_image"], _fd(lambda x: clip_loss_oracle([x], [txt], s)[0]["loss"], img), rtol=1e-5, atol=1e-8)
    assert np.allclose(c["d_text"], _fd(lambda x: clip_loss_oracle([img], [x], s)[0]["loss"], txt), rtol=1e-5, atol=1e-8)
    g = siglip_loss_oracle([img], [txt], s, b)[0]
    assert np.allclose(g["d_image"], _fd(lambda x: siglip_loss_oracle([x], [txt], s, b)[0]["loss"], img), rtol=1e-5, atol=1e-8)
    eps = 1e-6
    fd_b = (siglip_loss_oracle([img], [txt], s, b + eps)[0]["loss"] - siglip_loss_oracle([img], [txt], s, b - eps)[0]["loss"]) / (2 * eps)
    assert np.isclose(g["d_logit_bias"], fd_b, rtol=1e-5)
    m = multipositive_loss_oracle([img], [txt], [lab], s, delta)[0]
    assert np.allclose(m["d_image"], _fd(lambda x: multipositive_loss_oracle([x], [txt], [lab], s, delta)[0]["loss"], img),
                       rtol=1e-5, atol=1e-8)
    assert np.allclose(m["d_text"], _fd(lambda x: multipositive_loss_oracle([img], [x], [lab], s, delta)[0]["loss"], txt),
                       rtol=1e-5, atol=1e-8)
    fd_s = (multipositive_loss_oracle([img], [txt], [lab], s + eps, delta)[0]["loss"] -
            multipositive_loss_oracle([img], [txt], [lab], s - eps, delta)[0]["loss"]) / (2 * eps)
    assert np.isclose(m["d_logit_scale"], fd_s, rtol=1e-5)


def test_multipositive_reduces_to_clip_when_labels_are_unique():
    """every sample its own class and delta = 1/2: the multi-positive loss is ClipLoss"""
    rng = np.random.default_rng(5)
    img, txt = _feats(rng, 12, 6)
    lab = np.arange(12)
    m = multipositive_loss_oracle([img], [txt], [lab], 10.0, 0.5)[0]
    c = clip_loss_oracle([img], [txt], 10.0)[0]
    assert np.isclose(m["loss"], c["loss"], rtol=1e-12)
    assert np.allclose(m["d_image"], c["d_image"], rtol=1e-9, atol=1e-14)
    assert np.isclose(m["d_logit_scale"], c["d_logit_scale"], rtol=1e-9)


@pytest.mark.parametrize("w", [2, 3])
def test_siglip_rank_mean_is_global(w):
    rng = np.random.default_rng(11)
    img, txt = _feats(rng, 6 * w, 5)
    one = siglip_loss_oracle([img], [txt], 10.0, -10.0)[0]
    many = siglip_loss_oracle(_parts(img, w), _parts(txt, w), 10.0, -10.0)
    assert np.isclose(np.mean([o["loss"] for o in many]), one["loss"], rtol=1e-10)
    n = 6
    for r in range(w):
        assert np.allclose(many[r]["d_image"], w * one["d_image"][r * n:(r + 1) * n], rtol=1e-8, atol=1e-14)
    assert np.isclose(np.mean([o["d_logit_bias"] for o in many]), one["d_logit_bias"], rtol=1e-8)


def test_dscale_identity_for_forward_side_sums():
    """with A_r = sum_{i in r, all j} Prow_ij C_ij and A'_r = sum_{all i, j in r} Prow_ij C_ij
    (C = I T^T), the local-loss d_scale of rank r is  <dT_r, T_r>/s + 1/(2n) (A_r - A'_r)  -- only row-softmax sums,
    which the forward can accumulate in fp32 per column chunk."""
    rng = np.random.default_rng(0)
    w, n, d, s = 4, 6, 5, 9.0
    img, txt = _feats(rng, w * n, d)
    ref = clip_loss_oracle(_parts(img, w), _parts(txt, w), s, True, True)
    c = img @ txt.T
    z = s * c
    p_row = np.exp(z - np.log(np.exp(z).sum(axis=1, keepdims=True)))
    for r in range(w):
        rows = slice(r * n, (r + 1) * n)
        a_r = (p_row[rows] * c[rows]).sum()
        ap_r = (p_row[:, rows] * c[:, rows]).sum()
        ds = (ref[r]["d_text"] * txt[rows]).sum() / s + 0.5 / n * (a_r - ap_r)
        assert np.isclose(ds, ref[r]["d_logit_scale"], rtol=1e-10)
