"""Shared test plumbing: markers, golden-vector loader and the CPU stand-in engine.

``-m "not gpu"`` tests run on the CPU-only build container; ``-m gpu`` tests need a B200 and go
through the C ABI.  Nothing here reads /root/reference (it does not exist on the GPU box); the
reference's outputs come from the committed fixtures in tests/golden/ (made by oracle/gen_golden.py).
"""
import glob
import math
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu on the GPU box")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    world = int(meta["world"])
    ranks = []
    for r in range(world):
        ranks.append({k[len(f"r{r}_"):]: z[k] for k in z.files if k.startswith(f"r{r}_")})
    return dict(meta=meta, image=z["image"], text=z["text"], ranks=ranks, world=world)


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.linalg.norm(ref.ravel())
    return float(np.linalg.norm((got - ref).ravel()) / (den if den > 0 else 1.0))


LOG2E = 1.4426950408889634


class StandInEngine:
    """CPU stand-in for ``mrclip_b200.engine.CudaEngine`` (tests only).

    Implements the engine interface with dense float64 torch math so that the multi-rank
    orchestration, the mode coefficients and the label logic in ``mrclip_b200/loss.py`` can be
    exercised under gloo without a GPU.  It keeps per-scratch state in a dict instead of device
    workspace.  The product never constructs this class.
    """

    name = "cpu-standin"

    exact_g = False     # True: the gradient GEMMs contract a float64 copy of G (sharp checks of scalar formulas)

    def __init__(self):
        self.state = {}
        self.calls = []

    def padded_dim(self, d):
        return (d + 7) // 8 * 8

    def padded_cols(self, n):
        return (n + 127) // 128 * 128

    def workspace_bytes(self, m, n, d):
        return 256

    def fwd_col_granule(self, m, n):
        return 128

    def launch_count(self):
        return len(self.calls)

    def pack(self, src, dst):
        self.calls.append("pack")
        dst.zero_()
        dst[:, :src.shape[1]] = src.to(torch.bfloat16)

    def transpose(self, src, dst):
        self.calls.append("transpose")
        dst.zero_()
        dst[:src.shape[1], :src.shape[0]] = src.t()

    @staticmethod
    def _cos(a_rows, b_all):
        return a_rows.double() @ b_all.double().t()

    def clip_fwd_tiles(self, a_rows, b_all, shape, scale, col_begin, col_end, ws):
        self.calls.append("clip_fwd_tiles")
        key = ws.data_ptr()
        s = float(scale.item())
        z = s * self._cos(a_rows, b_all)
        st = self.state.setdefault(key, {})
        st["z"] = z
        st.setdefault("cols", set()).update(range(col_begin, col_end))

    def clip_fwd_reduce(self, shape, ws, lse2_row, col_m, col_l, diag2):
        self.calls.append("clip_fwd_reduce")
        st = self.state.pop(ws.data_ptr())
        assert st["cols"] == set(range(shape.n_cols)), "forward tiles did not cover every column"
        z = st["z"]
        lse2_row[:shape.m_rows] = (torch.logsumexp(z, dim=1) * LOG2E).float()
        m = z.max(dim=0).values
        col_m[:shape.n_cols] = (m * LOG2E).float()
        col_l[:shape.n_cols] = torch.exp(z - m[None, :]).sum(dim=0).float()
        idx = torch.arange(shape.m_rows)
        diag2[:shape.m_rows] = (z[idx, idx + shape.label_offset] * LOG2E).float()

    def lse2_merge(self, part_m, part_l, parts, stride, n_cols, out):
        self.calls.append("lse2_merge")
        ms = torch.as_strided(part_m, (parts, n_cols), (stride, 1)).double()
        ls = torch.as_strided(part_l, (parts, n_cols), (stride, 1)).double()
        mx = ms.max(dim=0).values
        tot = (ls * torch.exp2(ms - mx[None, :])).sum(dim=0)
        out.fill_(float("inf"))
        out[:n_cols] = (mx + torch.log2(tot)).float()

    def clip_loss(self, lse2_row, lse2_col, diag2, m_rows, label_offset, loss):
        self.calls.append("clip_loss")
        cols = lse2_col[label_offset:label_offset + m_rows].double()
        tot = (lse2_row[:m_rows].double() + cols - 2 * diag2[:m_rows].double()).sum()
        loss.view(-1)[0] = float(tot / LOG2E / (2 * m_rows))

    def clip_bwd(self, a_rows, b_all, bt_all, shape, lse2_a, lse2_b, scale, w_own, w_oth, coef, grad_out, ws,
                 d_a, d_scale, accumulate):
        self.calls.append("clip_bwd")
        s = float(scale.item())
        go = 1.0 if grad_out is None else float(grad_out.item())
        cos = self._cos(a_rows, b_all)
        assert torch.equal(bt_all[:b_all.shape[1], :b_all.shape[0]], b_all.t()), "stale transposed operand"
        t2 = s * cos * LOG2E
        p_own = torch.exp2(t2 - lse2_a[:shape.m_rows].double()[:, None])
        p_oth = torch.exp2(t2 - lse2_b[:shape.n_cols].double()[None, :])
        g = w_own * p_own + w_oth * p_oth
        idx = torch.arange(shape.m_rows)
        g[idx, idx + shape.label_offset] -= (w_own + w_oth)
        d_a.copy_((coef * s * go * (g @ b_all.double()))[:, :d_a.shape[1]].to(d_a.dtype))
        if d_scale is not None:
            val = coef * go * w_own * ((p_own * cos).sum() - cos[idx, idx + shape.label_offset].sum())
            d_scale[0] = (float(d_scale[0]) if accumulate else 0.0) + float(val)

    # ---- gmat backend ------------------------------------------------------------------------
    def gmat_bytes(self, m, n):
        return m * n * 2

    @staticmethod
    def _gview(gmat, shape):
        return gmat[:shape.m_rows * shape.n_cols].view(shape.m_rows, shape.n_cols)

    def clip_gwrite(self, a_rows, b_all, shape, lse2_a, lse2_b, scale, w_own, w_oth, coef, grad_out, ws, gmat,
                    d_scale, accumulate, both_directions):
        self.calls.append("clip_gwrite")
        s = float(scale.item())
        go = 1.0 if grad_out is None else float(grad_out.item())
        cos = self._cos(a_rows, b_all)
        t2 = s * cos * LOG2E
        p_own = torch.exp2(t2 - lse2_a[:shape.m_rows].double()[:, None])
        p_oth = torch.exp2(t2 - lse2_b[:shape.n_cols].double()[None, :])
        g = w_own * p_own + w_oth * p_oth
        idx = torch.arange(shape.m_rows)
        g[idx, idx + shape.label_offset] -= (w_own + w_oth)
        self._gview(gmat, shape).copy_(g.to(torch.bfloat16))
        if d_scale is not None:
            dsum = cos[idx, idx + shape.label_offset].sum()
            val = w_own * ((p_own * cos).sum() - dsum)
            if both_directions:
                val = val + w_oth * ((p_oth * cos).sum() - dsum)
            d_scale[0] = (float(d_scale[0]) if accumulate else 0.0) + float(coef * go * val)

    def siglip_gwrite(self, a_rows, b_all, shape, scale, bias, coef, grad_out, ws, gmat, d_scale, d_bias, accumulate):
        self.calls.append("siglip_gwrite")
        s = float(scale.item())
        b = 0.0 if bias is None else float(bias.item())
        go = 1.0 if grad_out is None else float(grad_out.item())
        cos = self._cos(a_rows, b_all)
        g = torch.sigmoid(s * cos + b)
        idx = torch.arange(shape.m_rows)
        g[idx, idx + shape.label_offset] -= 1.0
        self._gview(gmat, shape).copy_(g.to(torch.bfloat16))
        if d_scale is not None:
            d_scale[0] = (float(d_scale[0]) if accumulate else 0.0) + float(coef * go * (g * cos).sum())
        if d_bias is not None:
            d_bias[0] = (float(d_bias[0]) if accumulate else 0.0) + float(coef * go * g.sum())

    def gmat_gemm(self, transposed, gmat, shape, feat, coef, scale, grad_out, ws, d_out):
        self.calls.append("gmat_gemm")
        s = float(scale.item())
        go = 1.0 if grad_out is None else float(grad_out.item())
        g = self._gview(gmat, shape).double()
        if self.exact_g and ("g64", gmat.data_ptr()) in self.state:
            g = self.state[("g64", gmat.data_ptr())]
        assert feat.shape[0] == (shape.m_rows if transposed else shape.n_cols)
        out = (g.t() if transposed else g) @ feat.double()
        d_out.copy_((coef * s * go * out)[:, :d_out.shape[1]].to(d_out.dtype))

    # ---- emat backend ------------------------------------------------------------------------
    def clip_fwd_tiles_e(self, a_rows, b_all, shape, scale, col_begin, col_end, ws, emat):
        self.clip_fwd_tiles(a_rows, b_all, shape, scale, col_begin, col_end, ws)
        self.calls[-1] = "clip_fwd_tiles_e"
        # what the kernel leaves behind is a function of the logits only; keep a marker so a stale block is caught
        self.state[("emat", emat.data_ptr())] = (a_rows.data_ptr(), shape.label_offset)

    def emat_to_gmat(self, a_rows, b_all, shape, lse2_row, lse2_col, diag2, scale, w_row, w_col, ws, emat,
                     msums=None, n_per_rank=0, ranks=1):
        self.calls.append("emat_to_gmat")
        assert self.state.pop(("emat", emat.data_ptr())) == (a_rows.data_ptr(), shape.label_offset), "stale E block"
        s = float(scale.item())
        t2 = s * self._cos(a_rows, b_all) * LOG2E
        p_row = torch.exp2(t2 - lse2_row[:shape.m_rows].double()[:, None])
        p_col = torch.exp2(t2 - lse2_col[:shape.n_cols].double()[None, :])
        idx = torch.arange(shape.m_rows)
        g = w_row * p_row + w_col * p_col
        g[idx, idx + shape.label_offset] -= (w_row + w_col)
        self._gview(emat, shape).copy_(g.to(torch.bfloat16))
        self.state[("g64", emat.data_ptr())] = g
        if msums is not None:
            assert n_per_rank * ranks == shape.n_cols and tuple(msums.shape[1:]) == (2, ranks)
            msums.zero_()
            msums = msums[0]
            lp_row = t2 - lse2_row[:shape.m_rows].double()[:, None]
            lp_col = t2 - lse2_col[:shape.n_cols].double()[None, :]
            for r in range(ranks):
                cols = slice(r * n_per_rank, (r + 1) * n_per_rank)
                msums[0, r] = float((w_row * p_row[:, cols] * lp_row[:, cols]).sum())
                msums[1, r] = float((w_col * p_col[:, cols] * lp_col[:, cols]).sum())

    # ---- MRCLIP_DS=fwd: forward-side row sums for d logit_scale ---------------------------------
    def fwd_row_ent_ok(self, m_rows, n_cols, n_per_rank):
        return n_cols % n_per_rank == 0

    def clip_fwd_tiles_eu(self, a_rows, b_all, shape, scale, col_begin, col_end, ws, emat):
        self.clip_fwd_tiles_e(a_rows, b_all, shape, scale, col_begin, col_end, ws, emat)
        self.calls[-1] = "clip_fwd_tiles_eu"
        self.state[("u", ws.data_ptr())] = self.state[ws.data_ptr()]["z"] * LOG2E

    def row_ent_split(self, shape, ws, lse2_row, n_per_rank, ranks, out_slots):
        self.calls.append("row_ent_split")
        t2 = self.state.pop(("u", ws.data_ptr()))
        p_row = torch.exp2(t2 - lse2_row[:shape.m_rows].double()[:, None])
        out_slots.zero_()
        for q in range(ranks):
            cols = slice(q * n_per_rank, (q + 1) * n_per_rank)
            out_slots[q % out_slots.shape[0], 0, q] = float((p_row[:, cols] * t2[:, cols]).sum())

    def sum_slots_dot(self, slots, d_out, feat, dot_slots):
        self.calls.append("sum_slots_dot")
        tot = slots.double().sum(0)
        d_out.copy_(tot.to(d_out.dtype))
        dot_slots.zero_()
        dot_slots[0] = float((tot * feat.double()[:, :tot.shape[1]]).sum())

    def gmat_gemm_dot(self, transposed, gmat, shape, feat, coef, scale, grad_out, ws, d_out, dot_feat, dot_out):
        self.gmat_gemm(transposed, gmat, shape, feat, coef, scale, grad_out, ws, d_out)
        if dot_feat is not None:
            s = float(scale.item())
            dot_out[0] = float(dot_out[0]) + float((d_out.double() * dot_feat.double()[:, :d_out.shape[1]]).sum() / s)

    # ---- MultiPositiveClipLoss ----------------------------------------------------------------
    def class_means(self, x_all, order, seg_start, seg_cnt, mean_out):
        self.calls.append("class_means")
        mean_out.zero_()
        xs = x_all.double()
        for c in range(seg_cnt.shape[0]):
            k = int(seg_cnt[c])
            if k > 0:
                members = order[int(seg_start[c]):int(seg_start[c]) + k].long()
                mean_out[c] = xs[members].mean(0).float()

    def mpos_forward(self, img_rows, txt_rows, d, cls, tmean, imean, lse2_row, lse2_col, scale, delta, loss):
        self.calls.append("mpos_forward")
        s = float(scale.item())
        c = cls.long()
        pos_img = s * (img_rows.double() * tmean[c].double()).sum(-1)
        pos_txt = s * (txt_rows.double() * imean[c].double()).sum(-1)
        li = delta * (lse2_row.double() / LOG2E - pos_img) + (1.0 - delta) * (lse2_col.double() / LOG2E - pos_txt)
        loss.view(-1)[0] = float(li.mean())

    def mpos_backward(self, d_img, d_txt, img_rows, txt_rows, d, cls, tmean, imean, coef, scale, grad_out):
        self.calls.append("mpos_backward")
        k = coef * float(scale.item()) * (1.0 if grad_out is None else float(grad_out.item()))
        c = cls.long()
        d_img += (k * (txt_rows.double() - tmean[c].double()))[:, :d].to(d_img.dtype)
        d_txt += (k * (img_rows.double() - imean[c].double()))[:, :d].to(d_txt.dtype)

    def siglip_fwd_e(self, a_rows, b_all, shape, scale, bias, ws, loss, gmat):
        self.siglip_fwd(a_rows, b_all, shape, scale, bias, ws, loss)
        self.calls[-1] = "siglip_fwd_e"
        s = float(scale.item())
        b = 0.0 if bias is None else float(bias.item())
        cos = self._cos(a_rows, b_all)
        g = torch.sigmoid(s * cos + b)
        idx = torch.arange(shape.m_rows)
        g[idx, idx + shape.label_offset] -= 1.0
        self._gview(gmat, shape).copy_(g.to(torch.bfloat16))
        self.state[("sig", ws.data_ptr())] = (float((g * cos).sum()), float(g.sum()))

    def siglip_e_scalars(self, shape, ws, coef, grad_out, d_scale, d_bias, accumulate):
        self.calls.append("siglip_e_scalars")
        gc, gs = self.state.pop(("sig", ws.data_ptr()))
        go = 1.0 if grad_out is None else float(grad_out.item())
        if d_scale is not None:
            d_scale[0] = (float(d_scale[0]) if accumulate else 0.0) + coef * go * gc
        if d_bias is not None:
            d_bias[0] = (float(d_bias[0]) if accumulate else 0.0) + coef * go * gs

    def siglip_fwd(self, a_rows, b_all, shape, scale, bias, ws, loss):
        self.calls.append("siglip_fwd")
        s = float(scale.item())
        b = 0.0 if bias is None else float(bias.item())
        z = s * self._cos(a_rows, b_all) + b
        y = -torch.ones_like(z)
        idx = torch.arange(shape.m_rows)
        y[idx, idx + shape.label_offset] = 1.0
        loss.view(-1)[0] = float(torch.nn.functional.softplus(-y * z).sum() / shape.m_rows)

    def siglip_bwd(self, a_rows, b_all, bt_all, shape, scale, bias, coef, grad_out, ws, d_a, d_scale, d_bias,
                   accumulate):
        self.calls.append("siglip_bwd")
        s = float(scale.item())
        b = 0.0 if bias is None else float(bias.item())
        go = 1.0 if grad_out is None else float(grad_out.item())
        cos = self._cos(a_rows, b_all)
        g = torch.sigmoid(s * cos + b)
        idx = torch.arange(shape.m_rows)
        g[idx, idx + shape.label_offset] -= 1.0
        d_a.copy_((coef * s * go * (g @ b_all.double()))[:, :d_a.shape[1]].to(d_a.dtype))
        if d_scale is not None:
            d_scale[0] = (float(d_scale[0]) if accumulate else 0.0) + float(coef * go * (g * cos).sum())
        if d_bias is not None:
            d_bias[0] = (float(d_bias[0]) if accumulate else 0.0) + float(coef * go * g.sum())


@pytest.fixture
def standin_engine():
    import mrclip_b200
    eng = StandInEngine()
    mrclip_b200.set_engine(eng)
    yield eng
    mrclip_b200.set_engine(None)


def has_b200():
    if not torch.cuda.is_available():
        return False
    return torch.cuda.get_device_capability(0)[0] == 10
