"""Whole-step entries of the C ABI (``mrclip_step_forward`` / ``mrclip_step_backward``, include/mrclip.h).

One ctypes call per direction launches every kernel of ``ClipLoss`` / ``SigLipLoss`` (reference loss.py:128-139,
:365-448 and their autograd graph).  On several ranks nothing but the library's own kernels moves data: text rows,
LSE statistics, text-gradient tiles and scalars travel over NVLink peer memory (torch symmetric memory provides the
mapping only) and are ordered by device-side flags, so a step issues ~20 launches from C instead of ~60 from Python
and contains no NCCL call.  ``StepPlan`` owns the descriptor structs of one workspace and keeps the tensors they point
to alive.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import Shape

_DT = {torch.float32: _cabi.DT_F32, torch.bfloat16: _cabi.DT_BF16, torch.float16: _cabi.DT_F16}


class Peer(C.Structure):
    """``mrclip_peer`` (include/mrclip.h)."""
    _fields_ = [("ranks", C.c_int), ("rank", C.c_int), ("ctl_block_peers", C.c_void_p), ("ctl_block", C.c_void_p),
                ("ctl", C.c_void_p), ("txt_peers", C.c_void_p), ("stats_peers", C.c_void_p), ("recv_peers", C.c_void_p),
                ("recv", C.c_void_p), ("recv_bf16", C.c_int), ("ctl_block_peers_host", C.c_void_p),
                ("txt_peers_host", C.c_void_p), ("host_epoch", C.c_void_p)]


class Step(C.Structure):
    """``mrclip_step`` (include/mrclip.h)."""
    _fields_ = [("shape", Shape), ("ld", C.c_int), ("kind", C.c_int), ("local_loss", C.c_int), ("img_rows", C.c_void_p),
                ("txt_all", C.c_void_p), ("ws", C.c_void_p), ("emat", C.c_void_p), ("stats", C.c_void_p),
                ("lse2_row_all", C.c_void_p), ("lse2_col_all", C.c_void_p), ("msums", C.c_void_p), ("small", C.c_void_p),
                ("inv_norm", C.c_void_p), ("scale_buf", C.c_void_p), ("peer", Peer)]


KIND_CLIP, KIND_SIGLIP = 0, 1


def _ptr(t):
    return None if t is None else t.data_ptr()


class StepPlan:
    """Descriptors of one workspace for one loss kind / mode.  ``ws`` is a ``loss._Workspace``; on several ranks it must
    hold the symmetric buffers (``ws.sym``: two text buffers, statistics, control block) and, once a backward is
    needed, the receive slots of the fused reduce-scatter (``ws.push``)."""

    def __init__(self, eng, ws, kind, local_loss, rank):
        self.lib = eng.lib
        self.ws, self.kind, self.rank = ws, kind, rank
        dev = ws.img_all.device
        self.small = torch.zeros(int(self.lib.mrclip_step_small_floats()), dtype=torch.float32, device=dev)
        self.ctl = torch.zeros(128, dtype=torch.int32, device=dev)
        self.inv_norm = torch.zeros((2, ws.n), dtype=torch.float32, device=dev)      # raw forward: 1/||x|| of every row
        self.scale_buf = torch.zeros((1,), dtype=torch.float32, device=dev)          # raw forward: exp(log-scale)
        rows = slice(rank * ws.n, (rank + 1) * ws.n)
        self.host_epoch = C.c_int(0)          # steps issued on this plan's flags (host twin of the device epoch)
        self.host_ptrs = []                   # host copies of the peer address tables (kept alive here)
        self.steps = []
        for flip in ((0, 1) if ws.world > 1 else (0,)):
            st = Step()
            st.shape = Shape(ws.n, ws.N, ws.d, rank * ws.n)
            st.ld, st.kind, st.local_loss = ws.ld, kind, int(bool(local_loss))
            st.img_rows = ws.img_all[rows].data_ptr()
            st.ws = ws.scratch.data_ptr()
            st.emat = None
            st.lse2_row_all, st.lse2_col_all = ws.lse2_row_all.data_ptr(), ws.lse2_col_all.data_ptr()
            st.msums, st.small = ws.msums.data_ptr(), self.small.data_ptr()
            st.inv_norm, st.scale_buf = self.inv_norm.data_ptr(), self.scale_buf.data_ptr()
            if ws.world > 1:
                txt, _, txt_ptrs = ws.sym["txt"][flip]
                stats, _, stats_ptrs = ws.sym["stats"]
                blk, _, blk_ptrs = ws.sym["ctl"]
                st.txt_all, st.stats = txt.data_ptr(), stats.data_ptr()
                blk_h = (C.c_ulonglong * ws.world)(*[int(p) for p in blk_ptrs.tolist()])
                txt_h = (C.c_ulonglong * ws.world)(*[int(p) for p in txt_ptrs.tolist()])
                self.host_ptrs += [blk_h, txt_h]
                st.peer = Peer(ws.world, rank, blk_ptrs.data_ptr(), blk.data_ptr(), self.ctl.data_ptr(), txt_ptrs.data_ptr(),
                               stats_ptrs.data_ptr(), None, None, 0, C.addressof(blk_h), C.addressof(txt_h),
                               C.addressof(self.host_epoch))
            else:
                st.txt_all, st.stats = ws.txt_all.data_ptr(), ws.stats_local.data_ptr()
                st.peer = Peer(1, 0, None, None, self.ctl.data_ptr(), None, None, None, None, 0, None, None, None)
            self.steps.append(st)
        self.uses_fwd_ds = bool(self.lib.mrclip_step_uses_fwd_ds(C.byref(self.steps[0])))

    def _attach_grad_buffers(self, eng):
        """E / G block and, on several ranks, the receive slots (collective on first use)."""
        ws = self.ws
        emat = ws.gmat_buffer(eng)
        push = ws.push_buffers(self.rank) if ws.world > 1 else None
        if ws.world > 1 and push is None:
            return False
        for st in self.steps:
            st.emat = emat.data_ptr()
            if push is not None:
                recv, ptrs, _ = push
                st.peer.recv_peers, st.peer.recv = ptrs.data_ptr(), recv.data_ptr()
                st.peer.recv_bf16 = int(recv.dtype == torch.bfloat16)
        return True

    def forward(self, eng, flip, img, txt, scale, bias, need_grad, loss_out, raw=False):
        st = self.steps[flip]
        if need_grad and not st.emat and not self._attach_grad_buffers(eng):
            raise RuntimeError("mrclip_b200: peer receive buffers unavailable")
        _cabi.check(self.lib.mrclip_step_forward(C.byref(st), img.data_ptr(), _DT[img.dtype], img.stride(0), txt.data_ptr(),
                                                 _DT[txt.dtype], txt.stride(0), scale.data_ptr(), _ptr(bias),
                                                 int(bool(need_grad)), int(bool(raw)), loss_out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))

    def normalize_bwd(self, flip, d_img, d_txt):
        """Chain the gradients of the normalised features (as backward leaves them) to the raw tower outputs, in place."""
        ws, st = self.ws, self.steps[flip]
        stream = torch.cuda.current_stream().cuda_stream
        txt_rows = st.txt_all + self.rank * ws.n * ws.ld * 2
        for y, inv, g in ((st.img_rows, self.inv_norm[0], d_img), (txt_rows, self.inv_norm[1], d_txt)):
            _cabi.check(self.lib.mrclip_normalize_bwd(y, ws.ld, inv.data_ptr(), ws.n, ws.d, g.data_ptr(), _DT[g.dtype],
                                                      g.stride(0), stream))

    def backward(self, flip, scale, grad_out, coef, d_img, d_txt, d_scale, d_bias):
        st = self.steps[flip]
        _cabi.check(self.lib.mrclip_step_backward(C.byref(st), scale.data_ptr(), _ptr(grad_out), coef, d_img.data_ptr(),
                                                  _DT[d_img.dtype], d_img.stride(0), d_txt.data_ptr(), _DT[d_txt.dtype],
                                                  d_txt.stride(0), _ptr(d_scale), _ptr(d_bias),
                                                  torch.cuda.current_stream().cuda_stream))
