"""GPU parity tests (B200 only, ``-m gpu``): the sm_100a kernels, called through the C ABI, against

  * the reference's own outputs (tests/golden/, recorded from the unmodified reference),
  * the float64 oracle on fresh seeded inputs (sizes the oracle finishes in seconds),
  * a plain fp32 torch evaluation of the same formula on the GPU at BASELINE's full size, and
  * size-independent properties (homogeneity identity, rank-partition invariance).

Tolerances are BASELINE.json's: loss <= 1e-3 relative, gradients <= 1e-2 relative (norm-wise),
labels bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, has_b200, load_golden, rel_err

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-3
GRAD_TOL = 1e-2


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not has_b200():
        pytest.fail("these tests need a B200 (sm_100a); run them with -m gpu on the GPU box")
    import mrclip_b200
    mrclip_b200.set_engine(None)
    from mrclip_b200.engine import default_engine
    eng = default_engine()          # raises if libmrclip.so is missing: no silent fallback
    assert eng.name == "cuda-sm100a"
    yield


def _features(n, d, seed, corr=0.3):
    g = torch.Generator().manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(corr * img + (1 - corr) * torch.randn(n, d, generator=g) / d ** 0.5, dim=-1)
    return img.bfloat16().float(), txt.bfloat16().float()


# ------------------------------------------------------------------------------------------------
# emulate W ranks on one GPU: each rank's kernels run one after the other through the engine; the
# statistics exchange that NCCL performs in production is a tensor copy here.
# ------------------------------------------------------------------------------------------------
class EmulatedRanks:
    def __init__(self, img, txt, world):
        from mrclip_b200._cabi import Shape
        from mrclip_b200.engine import default_engine
        self.Shape = Shape
        self.eng = eng = default_engine()
        self.dev = dev = torch.device("cuda:0")
        self.W, self.N, self.d = world, img.shape[0], img.shape[1]
        self.n = self.N // world
        ld, npad = eng.padded_dim(self.d), eng.padded_cols(self.N)
        self.img_all = torch.zeros((self.N, ld), dtype=torch.bfloat16, device=dev)
        self.txt_all = torch.zeros((self.N, ld), dtype=torch.bfloat16, device=dev)
        eng.pack(img.to(dev), self.img_all)
        eng.pack(txt.to(dev), self.txt_all)
        self.img_t = torch.zeros((ld, npad), dtype=torch.bfloat16, device=dev)
        self.txt_t = torch.zeros((ld, npad), dtype=torch.bfloat16, device=dev)
        eng.transpose(self.img_all, self.img_t)
        eng.transpose(self.txt_all, self.txt_t)
        self.ws = torch.empty(int(eng.workspace_bytes(self.n, self.N, self.d)), dtype=torch.uint8, device=dev)
        self.npad = npad

    def shape(self, r):
        return self.Shape(self.n, self.N, self.d, r * self.n)

    def rows(self, r):
        return slice(r * self.n, (r + 1) * self.n)

    def clip(self, scale, local_loss, gather_with_grad, grad_output=1.0, backend="fused"):
        eng, dev, W, n, N = self.eng, self.dev, self.W, self.n, self.N
        gmat = None
        if backend in ("gmat", "emat"):
            gmat = torch.empty(int(eng.gmat_bytes(n, N)) // 2, dtype=torch.bfloat16, device=dev)
        s = torch.tensor([scale], dtype=torch.float32, device=dev)
        go = torch.tensor([grad_output], dtype=torch.float32, device=dev)
        stats = torch.zeros((W, 3, N), dtype=torch.float32, device=dev)
        diag = torch.zeros((W, n), dtype=torch.float32, device=dev)
        for r in range(W):
            eng.clip_fwd_tiles(self.img_all[self.rows(r)], self.txt_all, self.shape(r), s, 0, N, self.ws)
            eng.clip_fwd_reduce(self.shape(r), self.ws, stats[r, 2], stats[r, 0], stats[r, 1], diag[r])
        lse_col = torch.full((self.npad,), float("inf"), dtype=torch.float32, device=dev)
        lse_row = torch.full((self.npad,), float("inf"), dtype=torch.float32, device=dev)
        eng.lse2_merge(stats[0, 0], stats[0, 1], W, 3 * N, N, lse_col)
        lse_row[:N].view(W, n).copy_(stats[:, 2, :n])
        losses = torch.zeros((W,), dtype=torch.float32, device=dev)
        for r in range(W):
            eng.clip_loss(lse_row[self.rows(r)], lse_col, diag[r], n, r * n, losses[r:r + 1])
        global_mode = W > 1 and not local_loss
        coef = 0.5 / N if (global_mode and not gather_with_grad) else 0.5 / n
        w_oth = 0.0 if (W > 1 and local_loss and not gather_with_grad) else 1.0
        out = []
        ds_all = torch.zeros((W,), dtype=torch.float32, device=dev)
        if backend == "emat" and w_oth == 1.0:
            # production order per rank: forward keeps E, the rescale pass turns it into G, two plain GEMMs; the
            # text gradient is the sum over ranks of the [N, d] partials (NCCL reduce-scatter in production)
            dt_sum = torch.zeros((N, self.d), dtype=torch.float32, device=dev)
            msums = torch.zeros((W, 16, 2, W), dtype=torch.float32, device=dev)   # [rank][slot][direction][owner]
            for r in range(W):
                rows, sh = self.rows(r), self.shape(r)
                eng.clip_fwd_tiles_e(self.img_all[rows], self.txt_all, sh, s, 0, N, self.ws, gmat)
                eng.emat_to_gmat(self.img_all[rows], self.txt_all, sh, lse_row[rows], lse_col, diag[r], s, 1.0, 1.0,
                                 self.ws, gmat, msums[r], n, W)
                d_i = torch.empty((n, self.d), dtype=torch.float32, device=dev)
                part = torch.empty((N, self.d), dtype=torch.float32, device=dev)
                dot = torch.zeros((1,), dtype=torch.float32, device=dev)
                eng.gmat_gemm_dot(False, gmat, sh, self.txt_all, coef, s, go, self.ws, d_i, self.img_all[rows], dot)
                eng.gmat_gemm(True, gmat, sh, self.img_all[rows], coef, s, go, self.ws, part)
                dt_sum += part
                out.append(dict(d_image=d_i.cpu().numpy(), dot=float(dot)))
            # scale * dL_r/dscale = L_r + ln2/(2n) * (negative entropies of my rows + of my columns)
            msums = msums.sum(1)
            ds_r = [grad_output / scale * (float(losses[r]) + 0.6931471805599453 * 0.5 / n *
                                           float(msums[r, 0].sum() + msums[:, 1, r].sum())) for r in range(W)]
            for r in range(W):
                out[r]["d_text"] = dt_sum[self.rows(r)].cpu().numpy()
                out[r]["loss"] = float(losses.mean() if global_mode else losses[r])
                out[r]["d_scale"] = float(np.mean(ds_r)) if global_mode else ds_r[r]
                if W == 1:   # the GEMM's <dI, I> / scale must tell the same story (bf16 noise of G: large blocks only)
                    assert abs(out[r]["dot"] - ds_r[r]) <= 2e-2 * abs(ds_r[r]) + 1e-5 or n * N < (1 << 22)
            return out
        if backend == "emat":
            backend = "gmat"     # (local_loss, no gather_with_grad): the two gradients need different G blocks
        for r in range(W):
            d_i = torch.empty((n, self.d), dtype=torch.float32, device=dev)
            d_t = torch.empty((n, self.d), dtype=torch.float32, device=dev)
            ds = ds_all[r:r + 1]
            ld = self.img_all.shape[1]
            if backend == "gmat":
                eng.clip_gwrite(self.img_all[self.rows(r)], self.txt_all, self.shape(r), lse_row[self.rows(r)], lse_col,
                                s, 1.0, w_oth, coef, go, self.ws, gmat, ds, True, W == 1)
                eng.gmat_gemm(False, gmat, self.shape(r), self.txt_all, coef, s, go, self.ws, d_i)
                if W == 1:
                    eng.gmat_gemm(True, gmat, self.shape(r), self.img_all[self.rows(r)], coef, s, go, self.ws, d_t)
                else:
                    eng.clip_gwrite(self.txt_all[self.rows(r)], self.img_all, self.shape(r), lse_col[self.rows(r)],
                                    lse_row, s, 1.0, w_oth, coef, go, self.ws, gmat, ds, True, False)
                    eng.gmat_gemm(False, gmat, self.shape(r), self.img_all, coef, s, go, self.ws, d_t)
            else:
                eng.clip_bwd(self.img_all[self.rows(r)], self.txt_all, self.txt_t, self.shape(r),
                             lse_row[self.rows(r)], lse_col, s, 1.0, w_oth, coef, go, self.ws, d_i, ds, True)
                eng.clip_bwd(self.txt_all[self.rows(r)], self.img_all, self.img_t, self.shape(r),
                             lse_col[self.rows(r)], lse_row, s, 1.0, w_oth, coef, go, self.ws, d_t, ds, True)
            out.append(dict(d_image=d_i.cpu().numpy(), d_text=d_t.cpu().numpy()))
        ds_all = ds_all * ((0.5 / n) / coef)
        for r in range(W):
            out[r]["loss"] = float(losses.mean() if global_mode else losses[r])
            out[r]["d_scale"] = float(ds_all.mean() if global_mode else ds_all[r])
        return out

    def siglip(self, scale, bias, grad_output=1.0, backend="fused"):
        eng, dev, W, n, N = self.eng, self.dev, self.W, self.n, self.N
        gmat = None
        if backend in ("gmat", "emat"):
            gmat = torch.empty(int(eng.gmat_bytes(n, N)) // 2, dtype=torch.bfloat16, device=dev)
        s = torch.tensor([scale], dtype=torch.float32, device=dev)
        b = torch.tensor([bias], dtype=torch.float32, device=dev)
        go = torch.tensor([grad_output], dtype=torch.float32, device=dev)
        out = []
        if backend == "emat":
            dt_sum = torch.zeros((N, self.d), dtype=torch.float32, device=dev)
            for r in range(W):
                rows, sh = self.rows(r), self.shape(r)
                loss = torch.zeros((1,), dtype=torch.float32, device=dev)
                ds = torch.zeros((1,), dtype=torch.float32, device=dev)
                db = torch.zeros((1,), dtype=torch.float32, device=dev)
                d_i = torch.empty((n, self.d), dtype=torch.float32, device=dev)
                part = torch.empty((N, self.d), dtype=torch.float32, device=dev)
                eng.siglip_fwd_e(self.img_all[rows], self.txt_all, sh, s, b, self.ws, loss, gmat)
                eng.siglip_e_scalars(sh, self.ws, 1.0 / n, go, ds, db, False)
                eng.gmat_gemm(False, gmat, sh, self.txt_all, 1.0 / n, s, go, self.ws, d_i)
                eng.gmat_gemm(True, gmat, sh, self.img_all[rows], 1.0 / n, s, go, self.ws, part)
                dt_sum += part
                out.append(dict(loss=float(loss), d_image=d_i.cpu().numpy(), d_scale=float(ds), d_bias=float(db)))
            for r in range(W):
                out[r]["d_text"] = dt_sum[self.rows(r)].cpu().numpy()
            return out
        for r in range(W):
            loss = torch.zeros((1,), dtype=torch.float32, device=dev)
            ds = torch.zeros((1,), dtype=torch.float32, device=dev)
            db = torch.zeros((1,), dtype=torch.float32, device=dev)
            d_i = torch.empty((n, self.d), dtype=torch.float32, device=dev)
            d_t = torch.empty((n, self.d), dtype=torch.float32, device=dev)
            eng.siglip_fwd(self.img_all[self.rows(r)], self.txt_all, self.shape(r), s, b, self.ws, loss)
            ld = self.img_all.shape[1]
            if backend == "gmat":
                eng.siglip_gwrite(self.img_all[self.rows(r)], self.txt_all, self.shape(r), s, b, 1.0 / n, go, self.ws,
                                  gmat, ds, db, False)
                eng.gmat_gemm(False, gmat, self.shape(r), self.txt_all, 1.0 / n, s, go, self.ws, d_i)
                if W == 1:
                    eng.gmat_gemm(True, gmat, self.shape(r), self.img_all[self.rows(r)], 1.0 / n, s, go, self.ws, d_t)
                else:
                    eng.siglip_gwrite(self.txt_all[self.rows(r)], self.img_all, self.shape(r), s, b, 1.0 / n, go,
                                      self.ws, gmat, None, None, False)
                    eng.gmat_gemm(False, gmat, self.shape(r), self.img_all, 1.0 / n, s, go, self.ws, d_t)
            else:
                eng.siglip_bwd(self.img_all[self.rows(r)], self.txt_all, self.txt_t, self.shape(r), s, b, 1.0 / n, go,
                               self.ws, d_i, ds, db, False)
                eng.siglip_bwd(self.txt_all[self.rows(r)], self.img_all, self.img_t, self.shape(r), s, b, 1.0 / n, go,
                               self.ws, d_t, None, None, False)
            out.append(dict(loss=float(loss), d_image=d_i.cpu().numpy(), d_text=d_t.cpu().numpy(),
                            d_scale=float(ds), d_bias=float(db)))
        return out


def _check_rank(out, ref, kind):
    assert abs(out["loss"] - float(ref["loss"])) <= LOSS_TOL * abs(float(ref["loss"]))
    assert rel_err(out["d_image"], ref["d_image"]) <= GRAD_TOL
    assert rel_err(out["d_text"], ref["d_text"]) <= GRAD_TOL
    assert abs(out["d_scale"] - float(ref["d_scale"])) <= GRAD_TOL * abs(float(ref["d_scale"])) + 1e-7
    if kind == "siglip":
        assert abs(out["d_bias"] - float(ref["d_bias"])) <= GRAD_TOL * abs(float(ref["d_bias"])) + 1e-7


@pytest.mark.parametrize("backend", ["emat", "gmat", "fused"])
@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith("mpos")])
def test_kernels_match_reference_golden(name, backend):
    """Every fixture recorded from the reference, all ranks emulated on one GPU through the C ABI,
    for both backward backends (materialised-G GEMMs and the fused recompute row pass)."""
    g = load_golden(name)
    m, W = g["meta"], g["world"]
    em = EmulatedRanks(torch.from_numpy(g["image"]), torch.from_numpy(g["text"]), W)
    if m["kind"] == "clip":
        out = em.clip(float(m["scale"]), bool(m["local_loss"]), bool(m["gather_with_grad"]), float(m["grad_output"]),
                      backend=backend)
    else:
        out = em.siglip(float(m["scale"]), float(m["bias"]), float(m["grad_output"]), backend=backend)
    for r in range(W):
        _check_rank(out[r], g["ranks"][r], m["kind"])


@pytest.mark.parametrize("name", [n for n in golden_names() if "_w1" in n])
def test_module_api_matches_reference_golden(name):
    """The user-facing nn.Module (autograd path, workspace pool, dtype handling) at world_size 1."""
    from mrclip_b200 import ClipLoss, MultiPositiveClipLoss, SigLipLoss
    g = load_golden(name)
    m = g["meta"]
    dev = torch.device("cuda:0")
    img = torch.from_numpy(g["image"]).to(dev).requires_grad_(True)
    txt = torch.from_numpy(g["text"]).to(dev).requires_grad_(True)
    scale = torch.tensor(float(m["scale"]), device=dev, requires_grad=True)
    if m["kind"] == "mpos":
        lab = torch.from_numpy(g["ranks"][0]["labels_in"]).to(dev)
        loss = MultiPositiveClipLoss()(img, txt, scale, delta=float(m["delta"]), tokenized_texts=lab,
                                       output_dict=True)["multi contrastive_loss"]
    elif m["kind"] == "clip":
        mod = ClipLoss(cache_labels=True)
        loss = mod(img, txt, scale, output_dict=True)["contrastive_loss"]
        labels = mod.get_ground_truth(dev, img.shape[0])
        assert labels.dtype == torch.long and torch.equal(labels.cpu(), torch.from_numpy(g["ranks"][0]["labels"]))
        extra = {}
    else:
        bias = torch.tensor(float(m["bias"]), device=dev, requires_grad=True)
        loss = SigLipLoss()(img, txt, scale, bias)
    (loss * float(m["grad_output"])).backward()
    out = dict(loss=loss.item(), d_image=img.grad.cpu().numpy(), d_text=txt.grad.cpu().numpy(),
               d_scale=scale.grad.item())
    if m["kind"] == "siglip":
        out["d_bias"] = bias.grad.item()
    _check_rank(out, g["ranks"][0], m["kind"])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_module_accepts_amp_dtypes(dtype):
    from mrclip_b200 import ClipLoss
    from oracle.clip_oracle import clip_loss_oracle
    dev = torch.device("cuda:0")
    img, txt = _features(384, 512, 5)
    ref = clip_loss_oracle([img.numpy()], [txt.numpy()], 14.285714)[0]
    i = img.to(dev, dtype).requires_grad_(True)
    t = txt.to(dev, dtype).requires_grad_(True)
    s = torch.tensor(14.285714, device=dev, requires_grad=True)
    loss = ClipLoss()(i, t, s)
    loss.backward()
    assert i.grad.dtype == dtype and t.grad.dtype == dtype and s.grad.dtype == torch.float32
    assert abs(loss.item() - ref["loss"]) <= LOSS_TOL * abs(ref["loss"])
    assert rel_err(i.grad.float().cpu().numpy(), ref["d_image"]) <= (GRAD_TOL if dtype != torch.float32 else 5e-3)
    assert rel_err(t.grad.float().cpu().numpy(), ref["d_text"]) <= GRAD_TOL


@pytest.mark.parametrize("backend", ["emat", "gmat", "fused"])
@pytest.mark.parametrize("N,D,W,scale,mode", [
    (1024, 768, 1, 14.285714, (False, False)),
    (1536, 512, 4, 100.0, (True, True)),
    (1000, 200, 1, 30.0, (False, False)),      # ragged rows / cols / K
    (130, 40, 2, 14.285714, (True, False)),    # tiny, single partial tile
    (1, 8, 1, 5.0, (False, False)),            # degenerate batch of one
])
def test_clip_vs_oracle(N, D, W, scale, mode, backend):
    from oracle.clip_oracle import clip_loss_oracle
    img, txt = _features(N, D, 1000 + N + D, corr=0.15)
    n = N // W
    parts = lambda x: [x[r * n:(r + 1) * n].numpy() for r in range(W)]
    ref = clip_loss_oracle(parts(img), parts(txt), scale, mode[0], mode[1])
    out = EmulatedRanks(img, txt, W).clip(scale, mode[0], mode[1], backend=backend)
    for r in range(W):
        assert abs(out[r]["loss"] - ref[r]["loss"]) <= LOSS_TOL * abs(ref[r]["loss"]) + 1e-6
        assert rel_err(out[r]["d_image"], ref[r]["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref[r]["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_scale"] - ref[r]["d_logit_scale"]) <= GRAD_TOL * abs(ref[r]["d_logit_scale"]) + 1e-6


@pytest.mark.parametrize("backend", ["emat", "gmat", "fused"])
@pytest.mark.parametrize("N,D,W", [(1024, 768, 2), (520, 264, 1), (96, 24, 3)])
def test_siglip_vs_oracle(N, D, W, backend):
    from oracle.clip_oracle import siglip_loss_oracle
    img, txt = _features(N, D, 2000 + N)
    n = N // W
    parts = lambda x: [x[r * n:(r + 1) * n].numpy() for r in range(W)]
    ref = siglip_loss_oracle(parts(img), parts(txt), 10.0, -10.0)
    out = EmulatedRanks(img, txt, W).siglip(10.0, -10.0, backend=backend)
    for r in range(W):
        assert abs(out[r]["loss"] - ref[r]["loss"]) <= LOSS_TOL * abs(ref[r]["loss"])
        assert rel_err(out[r]["d_image"], ref[r]["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref[r]["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_scale"] - ref[r]["d_logit_scale"]) <= GRAD_TOL * abs(ref[r]["d_logit_scale"])
        assert abs(out[r]["d_bias"] - ref[r]["d_logit_bias"]) <= GRAD_TOL * abs(ref[r]["d_logit_bias"])


def _torch_fp32_clip(img, txt, scale):
    """Plain fp32 torch evaluation of the ClipLoss formula (W=1) on the GPU; TF32 off."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        i = img.clone().requires_grad_(True)
        t = txt.clone().requires_grad_(True)
        s = scale.clone().requires_grad_(True)
        logits = s * i @ t.T
        labels = torch.arange(i.shape[0], device=i.device)
        loss = 0.5 * (torch.nn.functional.cross_entropy(logits, labels) +
                      torch.nn.functional.cross_entropy(logits.T, labels))
        loss.backward()
        return loss.item(), i.grad, t.grad, s.grad.item()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("N,D,backend", [(8192, 512, "emat"), (8192, 512, "gmat"), (8192, 512, "fused"),
                                         (32768, 768, "emat")])
def test_full_size_vs_torch_fp32_and_properties(N, D, backend, monkeypatch):
    """BASELINE sizes (config 3 at world_size 1 is N=32768, D=768): fp32 torch on the same GPU as the
    checker, plus properties that need no checker at all."""
    from mrclip_b200 import ClipLoss
    monkeypatch.setenv("MRCLIP_BWD", backend)
    dev = torch.device("cuda:0")
    img, txt = _features(N, D, 1234 + 3)
    img, txt = img.to(dev), txt.to(dev)
    scale = torch.tensor(14.285714, device=dev)
    i = img.clone().requires_grad_(True)
    t = txt.clone().requires_grad_(True)
    s = scale.clone().requires_grad_(True)
    loss = ClipLoss()(i, t, s)
    loss.backward()
    # homogeneity: scale * dL/dscale == <dI, I> == <dT, T>
    lhs = (s * s.grad).item()
    assert abs((i.grad * img).sum().item() - lhs) <= 2e-2 * abs(lhs) + 1e-6
    assert abs((t.grad * txt).sum().item() - lhs) <= 2e-2 * abs(lhs) + 1e-6
    # permutation equivariance of pairs
    perm = torch.randperm(N, device=dev)
    ip = img[perm].clone().requires_grad_(True)
    tp = txt[perm].clone().requires_grad_(True)
    loss_p = ClipLoss()(ip, tp, scale)
    loss_p.backward()
    assert abs(loss_p.item() - loss.item()) <= 1e-5 * abs(loss.item())
    assert rel_err(ip.grad.cpu().numpy(), i.grad[perm].cpu().numpy()) <= 1e-3
    ref_loss, ref_di, ref_dt, ref_ds = _torch_fp32_clip(img, txt, scale)
    assert abs(loss.item() - ref_loss) <= LOSS_TOL * abs(ref_loss)
    assert rel_err(i.grad.cpu().numpy(), ref_di.cpu().numpy()) <= GRAD_TOL
    assert rel_err(t.grad.cpu().numpy(), ref_dt.cpu().numpy()) <= GRAD_TOL
    assert abs(s.grad.item() - ref_ds) <= GRAD_TOL * abs(ref_ds)


def test_rank_partition_invariance_large():
    """mean over ranks of the local losses equals the single-rank loss; gradients are W x."""
    N, D, W = 4096, 512, 8
    img, txt = _features(N, D, 77)
    one = EmulatedRanks(img, txt, 1).clip(14.285714, False, False)[0]
    many = EmulatedRanks(img, txt, W).clip(14.285714, True, True)
    n = N // W
    assert abs(np.mean([o["loss"] for o in many]) - one["loss"]) <= 1e-5 * abs(one["loss"])
    for r in range(W):
        assert rel_err(many[r]["d_image"], W * one["d_image"][r * n:(r + 1) * n]) <= 2e-3
        assert rel_err(many[r]["d_text"], W * one["d_text"][r * n:(r + 1) * n]) <= 2e-3
    assert abs(sum(o["d_scale"] for o in many) - W * one["d_scale"]) <= 1e-3 * abs(W * one["d_scale"])


def test_gradscaler_grad_output_and_workspace_reuse():
    """grad_output != 1 (train.py:63-67 GradScaler) and two forwards in flight before their backwards."""
    from mrclip_b200 import ClipLoss
    dev = torch.device("cuda:0")
    img, txt = _features(512, 512, 9)
    mod = ClipLoss(cache_labels=True)
    outs = []
    for k in range(2):
        i = img.to(dev).requires_grad_(True)
        t = txt.roll(k, 0).to(dev).requires_grad_(True)
        s = torch.tensor(14.285714, device=dev, requires_grad=True)
        outs.append((mod(i, t, s), i, t, s))
    for gscale, (loss, i, t, s) in zip((1.0, 65536.0), outs):
        (loss * gscale).backward()
    (l0, i0, t0, s0), (l1, i1, t1, s1) = outs
    fresh = ClipLoss()
    i = img.to(dev).requires_grad_(True)
    t = txt.roll(1, 0).to(dev).requires_grad_(True)
    s = torch.tensor(14.285714, device=dev, requires_grad=True)
    l = fresh(i, t, s)
    l.backward()
    assert abs(l.item() - l1.item()) <= 1e-6 * abs(l.item())
    assert rel_err((i1.grad / 65536.0).cpu().numpy(), i.grad.cpu().numpy()) <= 1e-5
    assert abs(s1.grad.item() / 65536.0 - s.grad.item()) <= 1e-5 * abs(s.grad.item())
    with torch.no_grad():
        assert abs(mod(img.to(dev), txt.to(dev), 14.285714).item() - l0.item()) <= 1e-6 * abs(l0.item())


def test_emat_guard_falls_back_to_exact_recompute(monkeypatch):
    """A row that dominates its 32 x 64 sub-tile by > 2^80 makes bf16 E of its neighbours flush to zero; the device
    guard must notice, the exact recompute must rewrite the block, and the result must still match the oracle."""
    from mrclip_b200 import ClipLoss
    from mrclip_b200.engine import default_engine
    monkeypatch.setenv("MRCLIP_BWD", "emat")
    dev = torch.device("cuda:0")
    N, D, scale = 256, 64, 100.0
    img, txt = _features(N, D, 5, corr=0.0)
    img[0] = 0.0
    img[0, 0] = 1.0
    txt[0] = img[0]                 # S_00 = 100 (144 in log2 units); every other logit of the band is ~ +-15
    mod = ClipLoss()
    i = img.to(dev).requires_grad_(True)
    t = txt.to(dev).requires_grad_(True)
    s = torch.tensor(scale, device=dev, requires_grad=True)
    loss = mod(i, t, s)
    loss.backward()
    ws = next(w for lst in mod._pool._free.values() for w in lst)
    eng = default_engine()
    from mrclip_b200._cabi import Shape
    off = eng.lib.mrclip_emat_flag(Shape(N, N, D, 0), ws.scratch.data_ptr()) - ws.scratch.data_ptr()
    assert int(ws.scratch[off:off + 4].view(torch.int32).item()) == 1, "guard flag not raised"
    from oracle.clip_oracle import clip_loss_oracle
    ref = clip_loss_oracle([img.numpy()], [txt.numpy()], scale)[0]
    assert abs(loss.item() - ref["loss"]) <= LOSS_TOL * abs(ref["loss"])
    assert rel_err(i.grad.cpu().numpy(), ref["d_image"]) <= GRAD_TOL
    assert rel_err(t.grad.cpu().numpy(), ref["d_text"]) <= GRAD_TOL
    assert abs(s.grad.item() - ref["d_logit_scale"]) <= GRAD_TOL * abs(ref["d_logit_scale"]) + 1e-7
    # two ranks, local loss: the per-rank d_scale then comes from the recompute's per-chunk partials
    N2 = 1024
    img2, txt2 = _features(N2, D, 6, corr=0.0)
    img2[0] = 0.0
    img2[0, 0] = 1.0
    txt2[0] = img2[0]
    out = EmulatedRanks(img2, txt2, 2).clip(scale, True, True, backend="emat")
    ref2 = clip_loss_oracle([img2[:512].numpy(), img2[512:].numpy()], [txt2[:512].numpy(), txt2[512:].numpy()], scale,
                            True, True)
    for r in range(2):
        assert abs(out[r]["loss"] - ref2[r]["loss"]) <= LOSS_TOL * abs(ref2[r]["loss"]) + 1e-6
        assert rel_err(out[r]["d_image"], ref2[r]["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref2[r]["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_scale"] - ref2[r]["d_logit_scale"]) <= GRAD_TOL * abs(ref2[r]["d_logit_scale"]) + 1e-6


@pytest.mark.parametrize("N,D,W", [(1000, 200, 1), (1536, 512, 4)])
def test_gemm_direct_epilogue(N, D, W, monkeypatch):
    """single-split GEMMs scale and store their result themselves (no fp32 partials, no reduce pass); forced here
    at small sizes, where the planner would normally split K"""
    from oracle.clip_oracle import clip_loss_oracle
    monkeypatch.setenv("MRCLIP_GEMM_MAX_KSPLIT", "1")
    img, txt = _features(N, D, 4000 + N, corr=0.2)
    n = N // W
    parts = lambda x: [x[r * n:(r + 1) * n].numpy() for r in range(W)]
    ref = clip_loss_oracle(parts(img), parts(txt), 14.285714, True, True)
    for backend in ("emat", "gmat"):
        out = EmulatedRanks(img, txt, W).clip(14.285714, True, True, backend=backend)
        for r in range(W):
            assert rel_err(out[r]["d_image"], ref[r]["d_image"]) <= GRAD_TOL
            assert rel_err(out[r]["d_text"], ref[r]["d_text"]) <= GRAD_TOL
            assert abs(out[r]["d_scale"] - ref[r]["d_logit_scale"]) <= GRAD_TOL * abs(ref[r]["d_logit_scale"]) + 1e-6


@pytest.mark.parametrize("mode", [(False, True), (True, True)])
def test_config2_gather_with_grad_w8(mode):
    """BASELINE config 2: global batch 4096, dim 512, eight ranks, gather_with_grad -- every rank's loss, gradients
    and d_scale against the float64 oracle (default backend, ranks emulated one after the other on this GPU)."""
    from oracle.clip_oracle import clip_loss_oracle
    N, D, W = 4096, 512, 8
    img, txt = _features(N, D, 1234 + 2, corr=0.5)
    n = N // W
    parts = lambda x: [x[r * n:(r + 1) * n].numpy() for r in range(W)]
    ref = clip_loss_oracle(parts(img), parts(txt), 14.285714, mode[0], mode[1])
    out = EmulatedRanks(img, txt, W).clip(14.285714, mode[0], mode[1], backend="emat")
    for r in range(W):
        assert abs(out[r]["loss"] - ref[r]["loss"]) <= LOSS_TOL * abs(ref[r]["loss"])
        assert rel_err(out[r]["d_image"], ref[r]["d_image"]) <= GRAD_TOL
        assert rel_err(out[r]["d_text"], ref[r]["d_text"]) <= GRAD_TOL
        assert abs(out[r]["d_scale"] - ref[r]["d_logit_scale"]) <= GRAD_TOL * abs(ref[r]["d_logit_scale"]) + 1e-7


def test_config4_siglip_with_bias_full_size():
    """BASELINE config 4 on one GPU: SigLipLoss, global batch 16384, dim 768, logit_bias, through the public module,
    against fp32 torch on the same GPU (the reference's formula, loss.py:342-363)."""
    from mrclip_b200 import SigLipLoss
    dev = torch.device("cuda:0")
    N, D = 16384, 768
    img, txt = _features(N, D, 1234 + 4)
    img, txt = img.to(dev), txt.to(dev)
    i = img.clone().requires_grad_(True)
    t = txt.clone().requires_grad_(True)
    s = torch.tensor(10.0, device=dev, requires_grad=True)
    b = torch.tensor(-10.0, device=dev, requires_grad=True)
    loss = SigLipLoss()(i, t, s, b)
    loss.backward()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        i2 = img.clone().requires_grad_(True)
        t2 = txt.clone().requires_grad_(True)
        s2 = torch.tensor(10.0, device=dev, requires_grad=True)
        b2 = torch.tensor(-10.0, device=dev, requires_grad=True)
        logits = s2 * i2 @ t2.T + b2
        labels = 2 * torch.eye(N, device=dev) - 1
        ref = -torch.nn.functional.logsigmoid(labels * logits).sum() / N
        ref.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert abs(loss.item() - ref.item()) <= LOSS_TOL * abs(ref.item())
    assert rel_err(i.grad.cpu().numpy(), i2.grad.cpu().numpy()) <= GRAD_TOL
    assert rel_err(t.grad.cpu().numpy(), t2.grad.cpu().numpy()) <= GRAD_TOL
    assert abs(s.grad.item() - s2.grad.item()) <= GRAD_TOL * abs(s2.grad.item())
    assert abs(b.grad.item() - b2.grad.item()) <= GRAD_TOL * abs(b2.grad.item())


def test_fp32_features_amp_contract():
    """Under AMP the reference hands fp32 features to the loss; gradients come back in fp32."""
    from mrclip_b200 import ClipLoss
    from oracle.clip_oracle import clip_loss_oracle
    dev = torch.device("cuda:0")
    img, txt = _features(2048, 512, 321)
    i = img.to(dev).requires_grad_(True)
    t = txt.to(dev).requires_grad_(True)
    s = torch.tensor(14.285714, device=dev, requires_grad=True)
    loss = ClipLoss(cache_labels=True)(i, t, s, output_dict=True)["contrastive_loss"]
    loss.backward()
    assert i.grad.dtype == torch.float32 and t.grad.dtype == torch.float32
    ref = clip_loss_oracle([img.numpy()], [txt.numpy()], 14.285714)[0]
    assert abs(loss.item() - ref["loss"]) <= LOSS_TOL * abs(ref["loss"])
    assert rel_err(i.grad.cpu().numpy(), ref["d_image"]) <= GRAD_TOL
    assert rel_err(t.grad.cpu().numpy(), ref["d_text"]) <= GRAD_TOL
    assert abs(s.grad.item() - ref["d_logit_scale"]) <= GRAD_TOL * abs(ref["d_logit_scale"])


def test_multipositive_full_size_vs_torch_fp32():
    """MR-CLIP's multi-positive loss at N=8192, D=512 with 600 label classes, against the reference's formula
    (loss.py:626-644, :745) evaluated in fp32 torch on the same GPU."""
    from mrclip_b200 import MultiPositiveClipLoss
    dev = torch.device("cuda:0")
    N, D, delta = 8192, 512, 0.4
    img, txt = _features(N, D, 4321)
    img, txt = img.to(dev), txt.to(dev)
    lab = torch.randint(0, 600, (N,), generator=torch.Generator().manual_seed(5), dtype=torch.long).to(dev)
    i = img.clone().requires_grad_(True)
    t = txt.clone().requires_grad_(True)
    s = torch.tensor(14.285714, device=dev, requires_grad=True)
    loss = MultiPositiveClipLoss()(i, t, s, delta=delta, tokenized_texts=lab)
    (loss * 2.0).backward()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        i2 = img.clone().requires_grad_(True)
        t2 = txt.clone().requires_grad_(True)
        s2 = torch.tensor(14.285714, device=dev, requires_grad=True)
        mask = (lab[:, None] == lab[None, :]).float()

        def mp(logits):
            logits = logits - logits.max(dim=1, keepdim=True).values.detach()
            logp = logits - torch.log(torch.exp(logits).sum(dim=1, keepdim=True) + 1e-12)
            return (-(mask * logp).sum(dim=1) / mask.sum(dim=1).clamp(min=1)).mean()
        ref = delta * mp(s2 * i2 @ t2.T) + (1 - delta) * mp(s2 * t2 @ i2.T)
        (ref * 2.0).backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert abs(loss.item() - ref.item()) <= LOSS_TOL * abs(ref.item())
    assert rel_err(i.grad.cpu().numpy(), i2.grad.cpu().numpy()) <= GRAD_TOL
    assert rel_err(t.grad.cpu().numpy(), t2.grad.cpu().numpy()) <= GRAD_TOL
    assert abs(s.grad.item() - s2.grad.item()) <= GRAD_TOL * abs(s2.grad.item())
