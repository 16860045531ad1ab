"""Kernel engine: the handful of device operations the loss modules are built from.

``CudaEngine`` forwards every call to the C ABI (``libmrclip.so``) on the current torch CUDA stream;
torch is used only to own device memory.  The interface is deliberately small so that the
multi-rank orchestration in ``loss.py`` can be exercised on CPU by tests with a stand-in engine
(tests provide it; the product never constructs anything but ``CudaEngine``).
"""
from __future__ import annotations

import torch

from . import _cabi

_DT = {torch.float32: _cabi.DT_F32, torch.bfloat16: _cabi.DT_BF16, torch.float16: _cabi.DT_F16}


def _ptr(t):
    return None if t is None else t.data_ptr()


class CudaEngine:
    """Runs the hot path through the hand-written sm_100a kernels.  No fallback."""

    name = "cuda-sm100a"

    def __init__(self):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("mrclip_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if not self.lib.mrclip_device_ok():
            raise RuntimeError("mrclip_b200 kernels are built for sm_100a (B200) only")

    # ---- geometry helpers -------------------------------------------------------------------
    def padded_dim(self, d):
        return self.lib.mrclip_padded_dim(d)

    def padded_cols(self, n):
        return self.lib.mrclip_padded_cols(n)

    def workspace_bytes(self, m, n, d):
        return self.lib.mrclip_workspace_bytes(m, n, d)

    def fwd_col_granule(self, m, n):
        return self.lib.mrclip_fwd_col_granule(m, n)

    def launch_count(self):
        return self.lib.mrclip_launch_count()

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    # ---- data movement ----------------------------------------------------------------------
    def pack(self, src, dst):
        """src [rows, d] (fp32/bf16/fp16, row stride arbitrary) -> dst bf16 [rows, ld] zero padded."""
        assert src.dim() == 2 and src.stride(1) == 1 and dst.dtype == torch.bfloat16 and dst.is_contiguous()
        _cabi.check(self.lib.mrclip_pack_bf16(src.data_ptr(), _DT[src.dtype], src.shape[0], src.shape[1],
                                              src.stride(0), dst.data_ptr(), dst.shape[1], self._stream()))

    def transpose(self, src, dst):
        """src bf16 [rows, ld] -> dst bf16 [ld, npad]."""
        _cabi.check(self.lib.mrclip_transpose_bf16(src.data_ptr(), src.shape[0], src.shape[1], src.stride(0),
                                                   dst.data_ptr(), dst.stride(0), self._stream()))

    # ---- ClipLoss ---------------------------------------------------------------------------
    def clip_fwd_tiles(self, a_rows, b_all, shape, scale, col_begin, col_end, ws):
        _cabi.check(self.lib.mrclip_clip_fwd_tiles(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                   scale.data_ptr(), col_begin, col_end, ws.data_ptr(),
                                                   self._stream()))

    def clip_fwd_reduce(self, shape, ws, lse2_row, col_m, col_l, diag2):
        _cabi.check(self.lib.mrclip_clip_fwd_reduce(shape, ws.data_ptr(), lse2_row.data_ptr(), col_m.data_ptr(),
                                                    col_l.data_ptr(), diag2.data_ptr(), self._stream()))

    def lse2_merge(self, part_m, part_l, parts, stride, n_cols, out):
        _cabi.check(self.lib.mrclip_lse2_merge(part_m.data_ptr(), part_l.data_ptr(), parts, stride, n_cols,
                                               out.data_ptr(), self._stream()))

    def clip_loss(self, lse2_row, lse2_col, diag2, m_rows, label_offset, loss):
        _cabi.check(self.lib.mrclip_clip_loss(lse2_row.data_ptr(), lse2_col.data_ptr(), diag2.data_ptr(), m_rows,
                                              label_offset, loss.data_ptr(), self._stream()))

    def clip_bwd(self, a_rows, b_all, bt_all, shape, lse2_a, lse2_b, scale, w_own, w_oth, coef, grad_out, ws,
                 d_a, d_scale, accumulate):
        _cabi.check(self.lib.mrclip_clip_bwd(a_rows.data_ptr(), b_all.data_ptr(), bt_all.data_ptr(),
                                             bt_all.stride(0), shape, b_all.shape[1], lse2_a.data_ptr(),
                                             lse2_b.data_ptr(), scale.data_ptr(), w_own, w_oth, coef,
                                             _ptr(grad_out), ws.data_ptr(), d_a.data_ptr(), _DT[d_a.dtype],
                                             d_a.stride(0), _ptr(d_scale), int(accumulate), self._stream()))

    # ---- SigLipLoss -------------------------------------------------------------------------
    def siglip_fwd(self, a_rows, b_all, shape, scale, bias, ws, loss):
        _cabi.check(self.lib.mrclip_siglip_fwd(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                               scale.data_ptr(), _ptr(bias), ws.data_ptr(), loss.data_ptr(),
                                               self._stream()))

    def siglip_bwd(self, a_rows, b_all, bt_all, shape, scale, bias, coef, grad_out, ws, d_a, d_scale, d_bias,
                   accumulate):
        _cabi.check(self.lib.mrclip_siglip_bwd(a_rows.data_ptr(), b_all.data_ptr(), bt_all.data_ptr(),
                                               bt_all.stride(0), shape, b_all.shape[1], scale.data_ptr(),
                                               _ptr(bias), coef, _ptr(grad_out), ws.data_ptr(), d_a.data_ptr(),
                                               _DT[d_a.dtype], d_a.stride(0), _ptr(d_scale), _ptr(d_bias),
                                               int(accumulate), self._stream()))


    # ---- "gmat" backward backend: materialised bf16 gradient block + plain GEMMs ----------------
    def gmat_bytes(self, m, n):
        return self.lib.mrclip_gmat_bytes(m, n)

    def clip_gwrite(self, a_rows, b_all, shape, lse2_a, lse2_b, scale, w_own, w_oth, coef, grad_out, ws, gmat,
                    d_scale, accumulate, both_directions):
        _cabi.check(self.lib.mrclip_clip_gwrite(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                lse2_a.data_ptr(), lse2_b.data_ptr(), scale.data_ptr(), w_own, w_oth,
                                                coef, _ptr(grad_out), ws.data_ptr(), gmat.data_ptr(), _ptr(d_scale),
                                                int(accumulate), int(both_directions), self._stream()))

    def siglip_gwrite(self, a_rows, b_all, shape, scale, bias, coef, grad_out, ws, gmat, d_scale, d_bias, accumulate):
        _cabi.check(self.lib.mrclip_siglip_gwrite(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                  scale.data_ptr(), _ptr(bias), coef, _ptr(grad_out), ws.data_ptr(),
                                                  gmat.data_ptr(), _ptr(d_scale), _ptr(d_bias), int(accumulate),
                                                  self._stream()))

    def gmat_gemm(self, transposed, gmat, shape, feat, coef, scale, grad_out, ws, d_out):
        """d_out = coef*scale*grad_out * (G . B) with feat = B [N, ld], or, transposed, (G^T . A) with feat = A [n, ld]."""
        _cabi.check(self.lib.mrclip_gmat_gemm(int(transposed), gmat.data_ptr(), shape, feat.data_ptr(), feat.shape[1],
                                              coef, scale.data_ptr(), _ptr(grad_out), ws.data_ptr(), d_out.data_ptr(),
                                              _DT[d_out.dtype], d_out.stride(0), self._stream()))


    # ---- "emat" backend: the forward keeps E (CLIP) or G (SigLIP); the backward never recomputes S ---------
    def clip_fwd_tiles_e(self, a_rows, b_all, shape, scale, col_begin, col_end, ws, emat):
        _cabi.check(self.lib.mrclip_clip_fwd_tiles_e(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                     scale.data_ptr(), col_begin, col_end, ws.data_ptr(),
                                                     emat.data_ptr(), self._stream()))

    def emat_to_gmat(self, a_rows, b_all, shape, lse2_row, lse2_col, diag2, scale, w_row, w_col, ws, emat,
                     msums=None, n_per_rank=0, ranks=1):
        """E -> G in place.  The guard decides on the device between the one-pass rewrite (normal) and the exact
        recompute of G (when a flushed E entry could carry gradient); exactly one of the two kernels does work.
        msums (float32 [slots, 2, ranks], optional) receives the d_logit_scale sums split by column owner, spread
        over `slots` copies that the caller adds up (msums.sum(0))."""
        st = self._stream()
        _cabi.check(self.lib.mrclip_emat_check(shape, ws.data_ptr(), lse2_row.data_ptr(), lse2_col.data_ptr(), st))
        flag = self.lib.mrclip_emat_flag(shape, ws.data_ptr())
        _cabi.check(self.lib.mrclip_clip_gwrite_if(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                   lse2_row.data_ptr(), lse2_col.data_ptr(), scale.data_ptr(),
                                                   w_row, w_col, ws.data_ptr(), emat.data_ptr(), flag, st))
        _cabi.check(self.lib.mrclip_emat_transform(shape, ws.data_ptr(), emat.data_ptr(), lse2_row.data_ptr(),
                                                   lse2_col.data_ptr(), diag2.data_ptr(), scale.data_ptr(), w_row,
                                                   w_col, flag, _ptr(msums), 1 if msums is None else msums.shape[0],
                                                   n_per_rank, ranks, st))

    def gmat_gemm_dot(self, transposed, gmat, shape, feat, coef, scale, grad_out, ws, d_out, dot_feat, dot_out):
        """gmat_gemm that also accumulates <d_out, dot_feat>/scale into dot_out (= d loss / d logit_scale)."""
        _cabi.check(self.lib.mrclip_gmat_gemm_dot(int(transposed), gmat.data_ptr(), shape, feat.data_ptr(),
                                                  feat.shape[1], coef, scale.data_ptr(), _ptr(grad_out), ws.data_ptr(),
                                                  d_out.data_ptr(), _DT[d_out.dtype], d_out.stride(0),
                                                  _ptr(dot_feat), _ptr(dot_out), self._stream()))

    def gmat_gemm_push(self, gmat, shape, feat, coef, scale, grad_out, ws, peer_ptrs, n_per_rank, my_rank, bf16=False):
        """G^T . A with the epilogue scattering row block q into rank q's receive buffer (peer_ptrs: int64 device
        tensor of NVLink-mapped addresses), slot my_rank.  bf16: the receive buffers are bf16 (MRCLIP_PUSH_DTYPE=bf16)."""
        fn = self.lib.mrclip_gmat_gemm_push_bf16 if bf16 else self.lib.mrclip_gmat_gemm_push
        _cabi.check(fn(gmat.data_ptr(), shape, feat.data_ptr(), feat.shape[1], coef, scale.data_ptr(), _ptr(grad_out),
                       ws.data_ptr(), peer_ptrs.data_ptr(), n_per_rank, my_rank, self._stream()))

    def sum_slots_bf16(self, slots, d_out, feat=None, dot_slots=None):
        """sum_slots / sum_slots_dot over bf16 slots."""
        assert slots.dtype == torch.bfloat16 and slots.is_contiguous() and slots.dim() == 3
        assert (feat is None) == (dot_slots is None)
        _cabi.check(self.lib.mrclip_sum_slots_bf16(slots.data_ptr(), slots.shape[0], slots.shape[1], slots.shape[2],
                                                   d_out.data_ptr(), _DT[d_out.dtype], d_out.stride(0), _ptr(feat),
                                                   0 if feat is None else feat.stride(0), _ptr(dot_slots),
                                                   self._stream()))

    def push_copy(self, src, peer_ptrs, dst_offset_bytes, skip_rank):
        """copy the contiguous tensor src into every peer's buffer at dst_offset_bytes (all-gather by NVLink stores)"""
        assert src.is_contiguous()
        _cabi.check(self.lib.mrclip_push_copy(src.data_ptr(), src.numel() * src.element_size(), peer_ptrs.data_ptr(),
                                              peer_ptrs.numel(), dst_offset_bytes, skip_rank, self._stream()))

    def sum_slots(self, slots, d_out):
        """d_out[rows, d] = sum_k slots[k] (fp32 [k, rows, d] contiguous)."""
        assert slots.dtype == torch.float32 and slots.is_contiguous() and slots.dim() == 3
        _cabi.check(self.lib.mrclip_sum_slots(slots.data_ptr(), slots.shape[0], slots.shape[1], slots.shape[2],
                                              d_out.data_ptr(), _DT[d_out.dtype], d_out.stride(0), self._stream()))

    # ---- MRCLIP_DS=fwd (opt-in, not validated on hardware yet): d logit_scale from forward-side row sums ----------
    def fwd_row_ent_ok(self, m_rows, n_cols, n_per_rank):
        """True when every column chunk of the forward plan has one owner (the per-owner split is then exact)."""
        return bool(self.lib.mrclip_fwd_row_ent_ok(m_rows, n_cols, n_per_rank))

    def clip_fwd_tiles_eu(self, a_rows, b_all, shape, scale, col_begin, col_end, ws, emat):
        """clip_fwd_tiles_e that also keeps u = sum_j 2^(S2 - m) S2 per row and column-chunk half in ws."""
        _cabi.check(self.lib.mrclip_clip_fwd_tiles_eu(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                      scale.data_ptr(), col_begin, col_end, ws.data_ptr(),
                                                      emat.data_ptr(), self._stream()))

    def row_ent_split(self, shape, ws, lse2_row, n_per_rank, ranks, out_slots):
        """out_slots[:, 0, q] (float32 [64, 2, ranks], summed over dim 0 by the caller) = R2(me, q): sum over my rows
        and the columns of rank q of Prow * S2 (log2 units).  After clip_fwd_reduce of a clip_fwd_tiles_eu forward."""
        assert out_slots.dtype == torch.float32 and out_slots.is_contiguous() and out_slots.shape == (64, 2, ranks)
        _cabi.check(self.lib.mrclip_row_ent_split(shape, ws.data_ptr(), lse2_row.data_ptr(), n_per_rank, ranks,
                                                  out_slots.data_ptr(), self._stream()))

    def sum_slots_dot(self, slots, d_out, feat, dot_slots):
        """sum_slots that also leaves <d_out, feat> (fp32, feat = packed bf16 rows) spread over dot_slots[64]."""
        assert slots.dtype == torch.float32 and slots.is_contiguous() and slots.dim() == 3
        assert feat.dtype == torch.bfloat16 and feat.stride(1) == 1 and dot_slots.numel() == 64
        _cabi.check(self.lib.mrclip_sum_slots_dot(slots.data_ptr(), slots.shape[0], slots.shape[1], slots.shape[2],
                                                  d_out.data_ptr(), _DT[d_out.dtype], d_out.stride(0), feat.data_ptr(),
                                                  feat.stride(0), dot_slots.data_ptr(), self._stream()))

    def siglip_fwd_e(self, a_rows, b_all, shape, scale, bias, ws, loss, gmat):
        _cabi.check(self.lib.mrclip_siglip_fwd_e(a_rows.data_ptr(), b_all.data_ptr(), shape, b_all.shape[1],
                                                 scale.data_ptr(), _ptr(bias), ws.data_ptr(), loss.data_ptr(),
                                                 gmat.data_ptr(), self._stream()))

    def siglip_e_scalars(self, shape, ws, coef, grad_out, d_scale, d_bias, accumulate):
        _cabi.check(self.lib.mrclip_siglip_e_scalars(shape, ws.data_ptr(), coef, _ptr(grad_out), _ptr(d_scale),
                                                     _ptr(d_bias), int(accumulate), self._stream()))


    # ---- MultiPositiveClipLoss: class means of the packed features and the terms built on them -------------
    def class_means(self, x_all, order, seg_start, seg_cnt, mean_out):
        """mean_out[c] (fp32 [N, ld]) = average of the rows of x_all (bf16 [N, ld]) of class id c."""
        assert x_all.dtype == torch.bfloat16 and x_all.is_contiguous() and mean_out.dtype == torch.float32
        assert order.dtype == seg_start.dtype == seg_cnt.dtype == torch.int32
        _cabi.check(self.lib.mrclip_class_means(x_all.data_ptr(), x_all.shape[1], order.data_ptr(), seg_start.data_ptr(),
                                                seg_cnt.data_ptr(), seg_cnt.shape[0], mean_out.data_ptr(), self._stream()))

    def mpos_forward(self, img_rows, txt_rows, d, cls, tmean, imean, lse2_row, lse2_col, scale, delta, loss):
        _cabi.check(self.lib.mrclip_mpos_forward(img_rows.data_ptr(), txt_rows.data_ptr(), img_rows.shape[1], img_rows.shape[0],
                                                 d, cls.data_ptr(), tmean.data_ptr(), imean.data_ptr(), lse2_row.data_ptr(),
                                                 lse2_col.data_ptr(), scale.data_ptr(), float(delta), loss.data_ptr(),
                                                 self._stream()))

    def mpos_backward(self, d_img, d_txt, img_rows, txt_rows, d, cls, tmean, imean, coef, scale, grad_out):
        _cabi.check(self.lib.mrclip_mpos_backward(d_img.data_ptr(), _DT[d_img.dtype], d_img.stride(0), d_txt.data_ptr(),
                                                  _DT[d_txt.dtype], d_txt.stride(0), img_rows.data_ptr(), txt_rows.data_ptr(),
                                                  img_rows.shape[1], img_rows.shape[0], d, cls.data_ptr(), tmean.data_ptr(),
                                                  imean.data_ptr(), coef, scale.data_ptr(), _ptr(grad_out), self._stream()))


_default_engine = None


def default_engine():
    """The process-wide CudaEngine (created on first use; raises when the GPU path is unavailable)."""
    global _default_engine
    if _default_engine is None:
        _default_engine = CudaEngine()
    return _default_engine
