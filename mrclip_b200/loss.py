"""Drop-in ``ClipLoss`` / ``SigLipLoss`` / ``MultiPositiveClipLoss`` / ``gather_features`` for MR-CLIP on B200.

Same constructor and call signatures as the reference (``src/open_clip/loss.py``:
``gather_features`` :21-65, ``ClipLoss`` :68-139, ``SigLipLoss`` :314-448, ``MultiPositiveClipLoss`` :671-747), so
``open_clip.factory.create_loss`` (factory.py:432-503) and ``train_one_epoch`` (train.py:128) use it unchanged.  Behind
the signatures the N x N logit matrix never exists: the forward runs the tcgen05 tile kernel with a fused online
log-sum-exp (``csrc/tile_kernel.cuh``) and keeps only the bf16 exponentials E of this rank's row block; the backward
rescales E into the gradient of the logits in one HBM pass and contracts it twice with plain tcgen05 GEMMs
(``csrc/gemm2_kernel.cuh``): three N x N x D contractions per step, the algorithmic minimum ("emat" backend).  The older
backends (``MRCLIP_BWD=gmat``: recompute S and write G, 4-5 contractions; ``fused``: O(N*D) memory, 7) remain selectable
and tested.

Two orchestrations of the same kernels:

* default -- ``mrclip_step_forward`` / ``mrclip_step_backward`` (``mrclip_b200/step.py``, ``csrc/mrclip_cabi.cu``): one C call
  per direction.  On several ranks (one process per GPU) nothing but the library's kernels and the copy engines move
  data, over NVLink peer memory mapped by torch symmetric memory, ordered by device-side flags (``csrc/peer_sync.cuh``):
  no NCCL call and no barrier inside a step.
* ``MRCLIP_STEP=py``, NCCL transports, the gmat / fused backends, ``(local_loss, not gather_with_grad)``, the
  multi-positive loss and the CPU tests (stand-in engine, gloo): the per-kernel sequence written out below in Python.

Multi-rank decomposition: rank r owns rows [r*n, (r+1)*n) of both modalities.  Forward: all-gather of the bf16-packed
TEXT rows (the image rows of other ranks are never needed), row-block tiles -> exact row LSE + per-column partial
(max, sum) -> exchange of those statistics.  Backward: dI_r = G_r . T_all is complete on its rank; the text gradient is
the sum over ranks of G_q^T . I_q, i.e. one reduce-scatter of [N, D] partials fused into the GEMM that produces them -- it
replaces the reference's reduce-scatter of W full copies per modality (torch/distributed/nn/functional.py:343-347).
The per-rank values reproduce the reference's conventions exactly:

  (local_loss, gather_with_grad)   loss on rank r     d features                d logit_scale
  (F, F)                           L_global           1/(2N) * (Pr + Pc - 2d)   global
  (F, T)                           L_global           1/(2n) * (Pr + Pc - 2d)   global
  (T, F)                           L_r                1/(2n) * (P_own - d)      local
  (T, T)                           L_r                1/(2n) * (Pr + Pc - 2d)   local

There is no CPU path: tensors must live on one sm_100a device, otherwise the call raises.  ``logit_scale`` must be
positive (it is ``exp(.)`` in the reference, model.py:324): the forward takes its sub-tile references on the unscaled
accumulators.
"""
from __future__ import annotations

import functools
import os
from typing import Optional

import torch
import torch.nn as nn

try:
    import torch.distributed.nn  # noqa: F401  (kept for parity with the reference's import guard)
    from torch import distributed as dist
    has_distributed = True
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False

from ._cabi import Shape
from .engine import default_engine
from .step import KIND_CLIP, KIND_SIGLIP
from .exchange import (NeighbourExchange, NeighbourExchangeBidir, neighbour_exchange,  # noqa: F401  (reference names)
                       neighbour_exchange_bidir, neighbour_exchange_bidir_with_grad, neighbour_exchange_with_grad)

__all__ = ["ClipLoss", "SigLipLoss", "MultiPositiveClipLoss", "gather_features", "gather_features_with_tokens",
           "multi_positive_cross_entropy_loss", "set_engine"]

_engine_override = None
_FWD_DS_MIN_PAIRS = 1 << 22     # MRCLIP_DS=fwd: below this the bf16 noise of G in <dT_r, T_r> is not averaged out


def set_engine(engine):
    """Test hook: route the device operations through ``engine`` (None restores the CUDA engine)."""
    global _engine_override
    _engine_override = engine


def _engine():
    return _engine_override if _engine_override is not None else default_engine()


# --------------------------------------------------------------------------------------------
# Differentiable gather with the reference's semantics (API parity for callers such as the
# multi-positive subclasses; the fused losses below do not use it).
# --------------------------------------------------------------------------------------------
class _AllGatherWithGrad(torch.autograd.Function):
    """all_gather whose backward is reduce_scatter(SUM), as torch.distributed.nn.all_gather."""

    @staticmethod
    def forward(ctx, x, world_size):
        ctx.world_size = world_size
        out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous())
        return out

    @staticmethod
    def backward(ctx, grad):
        n = grad.shape[0] // ctx.world_size
        out = torch.empty((n,) + tuple(grad.shape[1:]), dtype=grad.dtype, device=grad.device)
        dist.reduce_scatter_tensor(out, grad.contiguous(), op=dist.ReduceOp.SUM)
        return out, None


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0,
                    world_size=1, use_horovod=False):
    """Reference ``gather_features`` (loss.py:21-65): rank-ordered concatenation of every rank's features."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if use_horovod:
        raise NotImplementedError("Horovod is not supported by mrclip_b200 (NCCL over NVLink only)")
    if gather_with_grad:
        all_image_features = _AllGatherWithGrad.apply(image_features, world_size)
        all_text_features = _AllGatherWithGrad.apply(text_features, world_size)
    else:
        with torch.no_grad():
            all_image_features = _AllGatherWithGrad.apply(image_features, world_size)
            all_text_features = _AllGatherWithGrad.apply(text_features, world_size)
        if not local_loss:
            # keep the graph for the local slot (loss.py:58-61)
            n = image_features.shape[0]
            parts_i = list(all_image_features.split(n, dim=0))
            parts_t = list(all_text_features.split(n, dim=0))
            parts_i[rank] = image_features
            parts_t[rank] = text_features
            all_image_features = torch.cat(parts_i, dim=0)
            all_text_features = torch.cat(parts_t, dim=0)
    return all_image_features, all_text_features


def gather_features_with_tokens(image_features, text_features, text_tokens=None, local_loss=False,
                                gather_with_grad=False, rank=0, world_size=1, use_horovod=False):
    """Reference ``gather_features_with_tokens`` (loss.py:450-509): ``gather_features`` plus the rank-ordered
    concatenation of every rank's integer labels (None stays None).  Labels carry no gradient."""
    all_image_features, all_text_features = gather_features(
        image_features, text_features, local_loss=local_loss, gather_with_grad=gather_with_grad, rank=rank,
        world_size=world_size, use_horovod=use_horovod)
    return all_image_features, all_text_features, _all_gather_tokens(text_tokens, world_size)


def _all_gather_tokens(text_tokens, world_size):
    if text_tokens is None:
        return None
    tok = text_tokens.contiguous()
    out = torch.empty((world_size * tok.shape[0],) + tuple(tok.shape[1:]), dtype=tok.dtype, device=tok.device)
    dist.all_gather_into_tensor(out, tok)
    return out


def multi_positive_cross_entropy_loss(logits, pos_mask):
    """Reference ``multi_positive_cross_entropy_loss`` (loss.py:626-644) on materialised logits:
    mean_i( -(1/|P(i)|) sum_{j in P(i)} log softmax(logits_i)_j ), with the reference's quirks -- the row maximum is
    detached, 1e-12 is added inside the log, and an empty positive set divides by 1.  Helper for callers that hold
    logits; ``MultiPositiveClipLoss.forward`` computes the same value without forming them."""
    shifted = logits - logits.amax(dim=1, keepdim=True).detach()
    log_prob = shifted - torch.log(shifted.exp().sum(dim=1, keepdim=True) + 1e-12)
    positives = pos_mask.sum(dim=1).clamp(min=1)
    return (-(pos_mask * log_prob).sum(dim=1) / positives).mean()


# --------------------------------------------------------------------------------------------
# Workspace: gathered bf16 operands, their transposes and the kernel scratch, reused across steps
# --------------------------------------------------------------------------------------------
class _Workspace:
    def __init__(self, eng, device, n, world, d):
        self.n, self.world, self.d = n, world, d
        self.N = n * world
        self.ld = eng.padded_dim(d)
        self.npad = eng.padded_cols(self.N)
        bf, f32 = torch.bfloat16, torch.float32
        self.img_all = torch.zeros((self.N, self.ld), dtype=bf, device=device)
        self.txt_all = torch.zeros((self.N, self.ld), dtype=bf, device=device)
        # symmetric (NVLink peer-mapped) twins for the all-gathers by peer stores: two text buffers, alternated per
        # forward (a peer may still read last step's buffer in its backward), and the statistics block
        self.sym = None
        if world > 1 and torch.device(device).type == "cuda" and os.environ.get("MRCLIP_AG", "push").lower() != "nccl":
            self.sym = _symmetric_buffers(self, device)
        self.flip = 0
        self.img_t = self.txt_t = None      # [ld, npad] transposes, fused backend only (allocated on first use)
        self.scratch = torch.empty(max(int(eng.workspace_bytes(n, self.N, d)), 256), dtype=torch.uint8, device=device)
        # statistics in log2 units
        self.stats_local = torch.zeros((3, self.N), dtype=f32, device=device)      # col_m, col_l, (row lse2 in [:n])
        self.stats_all = torch.zeros((world, 3, self.N), dtype=f32, device=device) if world > 1 else None
        if self.sym is not None:
            self.stats_all = self.sym["stats"][0]
        self.lse2_row_all = torch.full((self.npad,), float("inf"), dtype=f32, device=device)
        self.lse2_col_all = torch.full((self.npad,), float("inf"), dtype=f32, device=device)
        self.diag2 = torch.zeros((n,), dtype=f32, device=device)
        self.in_use = False
        self.generation = 0            # bumped every time the set is handed out (see _Lease)
        self.transposed = False
        self.gmat = None
        self.dt_partial = None
        self.push = None
        self.msums = torch.zeros((64, 2, world), dtype=f32, device=device)   # 64 slots against atomic contention
        self.dot_slots = torch.zeros((64,), dtype=f32, device=device)        # <dT_r, T_r> partials (MRCLIP_DS=fwd)
        self.has_emat = False
        self.plans = {}

    def step_plan(self, eng, kind, local_loss, rank):
        """Descriptors of the whole-step C entries for this workspace (mrclip_b200/step.py), made on first use."""
        key = (kind, bool(local_loss), rank)
        if key not in self.plans:
            from .step import StepPlan
            self.plans[key] = StepPlan(eng, self, kind, local_loss, rank)
        return self.plans[key]

    def gmat_buffer(self, eng):
        """bf16 scratch for the materialised gradient block G (n x N), allocated on first use."""
        if self.gmat is None:
            nbytes = int(eng.gmat_bytes(self.n, self.N))
            self.gmat = torch.empty(max(nbytes // 2, 8), dtype=torch.bfloat16, device=self.img_all.device)
        return self.gmat

    def push_buffers(self, rank):
        """Peer-mapped receive buffer of the fused GEMM -> reduce-scatter (fp32 [W, n, d] on every rank, symmetric
        memory over NVLink): returns (recv, int64 device tensor of every rank's mapped address, handle), or None when
        symmetric memory is unavailable (then NCCL's reduce_scatter is used).  Collective on first use."""
        if self.push is None:
            self.push = False
            if self.img_all.is_cuda and os.environ.get("MRCLIP_RS", "push").lower() != "nccl":
                # bf16 payload (default; MRCLIP_PUSH_DTYPE=fp32 for the wide one): half the NVLink bytes
                pdt = torch.float32
                if os.environ.get("MRCLIP_PUSH_DTYPE", "bf16").lower() == "bf16" and self.d % 8 == 0:
                    pdt = torch.bfloat16
                bufs = _symmetric_alloc([((self.world, self.n, self.d), pdt)], self.img_all.device, self.world,
                                        "the gradient reduce-scatter")
                if bufs is not None:
                    recv, hdl, ptrs = bufs[0]
                    self.push = (recv, ptrs, hdl)
        return self.push or None

    def dt_partial_buffer(self):
        """fp32 [N, d] partial text gradient of this rank's row block (reduce-scattered), world > 1 only."""
        if self.dt_partial is None:
            self.dt_partial = torch.empty((self.N, self.d), dtype=torch.float32, device=self.img_all.device)
        return self.dt_partial


def _all_ranks_ok(ok, device):
    """True only when every rank of the default group reports ok.  The peer-memory paths and their NCCL twins issue
    different collectives, so the choice between them must never be made by one rank alone."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def _symmetric_alloc(specs, device, world, what):
    """Collective: every rank allocates the symmetric (NVLink peer-mapped) tensors ``specs`` = [(shape, dtype), ...] in
    the same order and maps its peers' copies.  Returns [(tensor, handle, int64 device tensor of every rank's mapped
    address)] or None -- on EVERY rank -- when any rank could not allocate or map them, or when the module's
    world_size is not the default process group's (the handles are made on dist.group.WORLD)."""
    out, err = [], None
    if dist.get_world_size() != world:
        err = f"module world_size {world} != default process group size {dist.get_world_size()}"
    symm_mem = None
    if err is None:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            tensors = [symm_mem.empty(shape, dtype=dtype, device=device) for shape, dtype in specs]
        except Exception as exc:  # pragma: no cover  (depends on the driver / fabric setup of the box)
            err = f"{type(exc).__name__}: {exc}"
    if not _all_ranks_ok(err is None, device):
        import warnings
        warnings.warn(f"mrclip_b200: symmetric memory unavailable for {what} ({err or 'on another rank'}); using NCCL")
        return None
    try:
        for t in tensors:
            hdl = symm_mem.rendezvous(t, dist.group.WORLD)
            if len(hdl.buffer_ptrs) != world:
                raise RuntimeError(f"{len(hdl.buffer_ptrs)} mapped peers for world_size {world}")
            ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=device)
            t.zero_()
            out.append((t, hdl, ptrs))
    except Exception as exc:  # pragma: no cover
        err = f"{type(exc).__name__}: {exc}"
    if not _all_ranks_ok(err is None, device):
        import warnings
        warnings.warn(f"mrclip_b200: symmetric memory rendezvous failed for {what} ({err or 'on another rank'}); using NCCL")
        return None
    return out


def _symmetric_buffers(ws, device):
    """Collective: two text buffers (alternated per step) and the statistics block.  None when unavailable."""
    bufs = _symmetric_alloc([((ws.N, ws.ld), torch.bfloat16), ((ws.N, ws.ld), torch.bfloat16),
                             ((ws.world, 3, ws.N), torch.float32), ((1024,), torch.int32)], device, ws.world,
                            "the all-gathers")
    # "ctl": the flag / scalar block of csrc/peer_sync.cuh (mrclip_peer_block_bytes() = 4096 bytes, zeroed)
    return None if bufs is None else {"txt": bufs[:2], "stats": bufs[2], "ctl": bufs[3]}


def _backend(eng, ws):
    """MRCLIP_BWD = emat | gmat | fused | auto.  auto: emat (or gmat) while the n x N bf16 block stays under
    MRCLIP_GMAT_MAX_GIB (16), else fused."""
    mode = os.environ.get("MRCLIP_BWD", "auto").lower()
    if mode in ("emat", "gmat", "fused"):
        return mode
    limit = float(os.environ.get("MRCLIP_GMAT_MAX_GIB", "16")) * 2 ** 30
    return "emat" if int(eng.gmat_bytes(ws.n, ws.N)) <= limit else "fused"


def _use_gmat(eng, ws):
    """True for the backends that contract a materialised bf16 G block with plain GEMMs ('emat': the block comes
    from the forward's exponentials, 3 GEMM units per step; 'gmat': from a recompute pass, 4-5); 'fused' keeps
    O(N*D) memory but recomputes S per 384-wide slice of D (7 units)."""
    return _backend(eng, ws) in ("gmat", "emat")


class _Lease:
    """One hand-out of a workspace, owned by the autograd ctx of the forward that took it.  The set goes back to the
    pool when backward has consumed it or -- through __del__ -- when the graph is dropped without a backward (skipped
    step, exception, loss only logged), so a leaked ctx can never pin an n x N block for good.  A second backward
    through the same graph (retain_graph=True) finds the lease spent: backward rewrites the saved E block in place,
    so the state it needs no longer exists."""
    __slots__ = ("ws", "gen", "spent", "__weakref__")

    def __init__(self, ws):
        self.ws, self.gen, self.spent = ws, ws.generation, False

    def check(self):
        if self.spent or self.ws.generation != self.gen:
            raise RuntimeError("mrclip_b200: backward through this loss a second time -- the saved exponentials were "
                               "consumed (rewritten in place) by the first backward; call forward again "
                               "(retain_graph=True is not supported by the fused loss)")

    def release(self):
        if not self.spent:
            self.spent = True
            if self.ws.generation == self.gen:
                self.ws.in_use = False

    def __del__(self):
        self.release()


class _WorkspacePool:
    """Per-module cache keyed by (device, n, world, d); a set is handed out once until its lease ends.  At most
    4 sets per key and MRCLIP_POOL_KEYS (8) keys are kept: ragged batch sizes evict the least recently used idle key
    instead of growing without bound."""

    def __init__(self):
        self._free = {}          # key -> list of workspaces; dict order = least recently used first

    def take(self, eng, device, n, world, d):
        key = (str(device), n, world, d)
        lst = self._free.pop(key, [])
        self._free[key] = lst    # most recently used
        w = next((x for x in lst if not x.in_use), None)
        if w is None:
            w = _Workspace(eng, device, n, world, d)
            if len(lst) < 4:
                lst.append(w)
            max_keys = int(os.environ.get("MRCLIP_POOL_KEYS", "8"))
            for old in [k for k in self._free if k != key]:
                if len(self._free) <= max_keys:
                    break
                if not any(x.in_use for x in self._free[old]):
                    del self._free[old]
        w.in_use = True
        w.generation += 1
        w.transposed = False
        w.has_emat = False
        return w

    @staticmethod
    def lease(w):
        return _Lease(w)



def _step_path_ok(eng, ws, world, need_grad):
    """The whole-step C entries (csrc/mrclip_cabi.cu, mrclip_step_*) serve the emat pipeline on one rank and, on several,
    over NVLink peer memory.  MRCLIP_STEP=py keeps the per-kernel Python orchestration (also what the CPU tests with a
    stand-in engine and the NCCL fallbacks use)."""
    if _engine_override is not None or os.environ.get("MRCLIP_STEP", "c").lower() == "py":
        return False
    if need_grad and _backend(eng, ws) != "emat":
        return False
    if world > 1:
        if ws.sym is None or ws.n < 8 or world > 64 or ws.N % 4 != 0:
            return False
        if os.environ.get("MRCLIP_RS", "push").lower() == "nccl" or os.environ.get("MRCLIP_GEMM_CTA") == "1":
            return False
        if need_grad and ws.push_buffers(0) is None:
            return False
    return True


class _RawUnavailable(RuntimeError):
    """forward_raw cannot run fused for this call (no whole-step path); the caller composes F.normalize + forward."""


def _rows_major(x):
    x = x.detach()
    return x if x.stride(1) == 1 else x.contiguous()


def _scalar_f32(x, device, positive=False):
    if positive and not torch.is_tensor(x) and float(x) <= 0.0:
        raise ValueError(f"logit_scale must be positive, got {x}")
    if torch.is_tensor(x):
        return x.detach().reshape(-1)[:1].to(device=device, dtype=torch.float32).contiguous()
    return torch.tensor([float(x)], dtype=torch.float32, device=device)


@functools.lru_cache(maxsize=None)
def _capability(device):
    return torch.cuda.get_device_capability(device)


def _check_inputs(image_features, text_features):
    if not (torch.is_tensor(image_features) and torch.is_tensor(text_features)):
        raise TypeError("image_features and text_features must be tensors")
    if image_features.dim() != 2 or image_features.shape != text_features.shape:
        raise ValueError(f"expected two [n, D] feature matrices of equal shape, got {tuple(image_features.shape)} "
                         f"and {tuple(text_features.shape)}")
    if image_features.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise TypeError(f"unsupported feature dtype {image_features.dtype}")
    if image_features.shape[0] == 0:
        raise ValueError("empty batch")
    if _engine_override is None:
        # the kernels take raw device pointers: anything that is not on one sm_100a GPU would fault inside the launch
        if not (image_features.is_cuda and text_features.is_cuda):
            raise RuntimeError("mrclip_b200 has no CPU path: features must live on a B200 (sm_100a) device, got "
                               f"{image_features.device} / {text_features.device}")
        if image_features.device != text_features.device:
            raise RuntimeError(f"features on different devices: {image_features.device} vs {text_features.device}")
        if _capability(image_features.device)[0] != 10:
            raise RuntimeError(f"mrclip_b200 kernels are built for sm_100a only; {image_features.device} is "
                               f"sm_{''.join(map(str, _capability(image_features.device)))}")


def _scoped(fn):
    """forward(ctx, image_features, ...) / backward(ctx, grad) with the features' device current (see _on_device)."""
    @functools.wraps(fn)
    def inner(ctx, first, *rest):
        if fn.__name__ == "forward":
            _check_inputs(first, rest[0])
        device = first.device if fn.__name__ == "forward" else ctx.ws.img_all.device
        with _on_device(device):
            return fn(ctx, first, *rest)
    return inner


class _on_device:
    """Runs the block with the features' GPU current, so that the stream handed to the C ABI (the current stream of the
    current device) and the raw pointers belong to the same device.  No-op for the CPU stand-in engine of the tests."""

    def __init__(self, device):
        device = torch.device(device)
        need = device.type == "cuda" and device.index is not None and device.index != torch.cuda.current_device()
        self.guard = torch.cuda.device(device) if need else None

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _gather_packed(eng, ws, image_features, text_features, rank, world, gather_images=True):
    """Pack the local features to bf16 straight into their slot and all-gather in place.  The emat backend
    never touches another rank's image rows (dI_r = G_r . T_all, dT partial = G_r^T . I_r), so it gathers
    the text side only."""
    n = ws.n
    rows = slice(rank * n, (rank + 1) * n)
    img = image_features.detach()
    txt = text_features.detach()
    if img.stride(1) != 1:
        img = img.contiguous()
    if txt.stride(1) != 1:
        txt = txt.contiguous()
    push = world > 1 and ws.sym is not None and not gather_images
    if push:
        ws.flip ^= 1
        ws.txt_all, hdl, ptrs = ws.sym["txt"][ws.flip]
    eng.pack(img, ws.img_all[rows])
    eng.pack(txt, ws.txt_all[rows])
    if push:
        # all-gather by peer stores: my packed slice goes straight into every peer's buffer over NVLink
        eng.push_copy(ws.txt_all[rows], ptrs, rank * n * ws.ld * 2, rank)
        hdl.barrier()
    elif world > 1:
        if gather_images:
            _all_gather_rows(ws.img_all, rows)
        _all_gather_rows(ws.txt_all, rows)


def _all_gather_rows(buf, rows):
    """All-gather the [n, ld] slot ``buf[rows]`` of every rank into ``buf`` (in place on NCCL)."""
    src = buf[rows]
    if not buf.is_cuda:
        src = src.clone()  # gloo (CPU tests) does not promise in-place semantics
    dist.all_gather_into_tensor(buf, src)


def _ensure_transposed(eng, ws):
    if ws.img_t is None:
        ws.img_t = torch.zeros((ws.ld, ws.npad), dtype=torch.bfloat16, device=ws.img_all.device)
        ws.txt_t = torch.zeros((ws.ld, ws.npad), dtype=torch.bfloat16, device=ws.img_all.device)
    if not ws.transposed:
        eng.transpose(ws.img_all, ws.img_t)
        eng.transpose(ws.txt_all, ws.txt_t)
        ws.transposed = True


def _text_grad_scatter(eng, ws, gmat, shape, coef, scale, gout, rank, d_txt, dot_slots=None):
    """dT_r = sum over ranks q of (G_q^T . I_q)[rows of r].  Launches this rank's partial GEMM and returns the
    closure that completes d_txt; the caller runs the image-gradient GEMM in between.

    Default on GPUs: the GEMM's epilogue pushes each output tile into its owner's receive slot over NVLink peer
    memory (csrc/gemm2_kernel.cuh, mrclip_gmat_gemm_push: 128-byte bulk copies from a staging tile), then one
    device-side barrier and a slot sum on the owner.  MRCLIP_RS=nccl, CPU tests under gloo, or no symmetric memory:
    fp32 partial + reduce_scatter, asynchronous, overlapped with the image-gradient GEMM.
    dot_slots (float32 [64], optional): also receives <dT_r, T_r> with T_r the packed text rows (its sum)."""
    n, d, world = ws.n, ws.d, ws.world
    rows = slice(rank * n, (rank + 1) * n)
    push = ws.push_buffers(rank)
    if push is not None:
        recv, ptrs, hdl = push
        if recv.dtype == torch.bfloat16:
            eng.gmat_gemm_push(gmat, shape, ws.img_all[rows], coef, scale, gout, ws.scratch, ptrs, n, rank, bf16=True)

            def finish_bf16():
                hdl.barrier()
                eng.sum_slots_bf16(recv, d_txt, None if dot_slots is None else ws.txt_all[rows], dot_slots)
            return finish_bf16
        eng.gmat_gemm_push(gmat, shape, ws.img_all[rows], coef, scale, gout, ws.scratch, ptrs, n, rank)

        def finish():
            hdl.barrier()                  # every rank's tiles have landed in my slots (and mine in theirs)
            if dot_slots is None:
                eng.sum_slots(recv, d_txt)
            else:
                eng.sum_slots_dot(recv, d_txt, ws.txt_all[rows], dot_slots)
        return finish
    part = ws.dt_partial_buffer()
    eng.gmat_gemm(True, gmat, shape, ws.img_all[rows], coef, scale, gout, ws.scratch, part)
    dt32 = torch.empty((n, d), dtype=torch.float32, device=part.device)
    work = dist.reduce_scatter_tensor(dt32, part, op=dist.ReduceOp.SUM, async_op=True)

    def finish():
        work.wait()
        d_txt.copy_(dt32)
        if dot_slots is not None:
            dot_slots.zero_()
            dot_slots[0] = (dt32 * ws.txt_all[rows][:, :d].float()).sum()
    return finish


def _lse_forward(eng, ws, image_features, text_features, scale, shape, rank, world, keep_e, gather_images,
                 row_ent=False):
    """Pack + gather, the row-block tiles (keeping the bf16 exponentials when keep_e) and the statistics exchange:
    leaves lse2_row_all / lse2_col_all (every row / column of the global batch, log2 units) and diag2 in ws."""
    n, N = ws.n, ws.N
    rows = slice(rank * n, (rank + 1) * n)
    _gather_packed(eng, ws, image_features, text_features, rank, world, gather_images=gather_images)
    if keep_e and row_ent:
        eng.clip_fwd_tiles_eu(ws.img_all[rows], ws.txt_all, shape, scale, 0, N, ws.scratch, ws.gmat_buffer(eng))
        ws.has_emat = True
    elif keep_e:
        eng.clip_fwd_tiles_e(ws.img_all[rows], ws.txt_all, shape, scale, 0, N, ws.scratch, ws.gmat_buffer(eng))
        ws.has_emat = True
    else:
        eng.clip_fwd_tiles(ws.img_all[rows], ws.txt_all, shape, scale, 0, N, ws.scratch)
    col_m, col_l, row_lse = ws.stats_local[0], ws.stats_local[1], ws.stats_local[2]
    eng.clip_fwd_reduce(shape, ws.scratch, row_lse, col_m, col_l, ws.diag2)
    if keep_e and row_ent:   # R2(me, q) for every column owner q, from the forward's per-chunk row sums
        eng.row_ent_split(shape, ws.scratch, row_lse, n, world, ws.msums)
    if world > 1 and ws.sym is not None and N % 4 == 0:     # (16-byte granular peer stores)
        _, shdl, sptrs = ws.sym["stats"]
        ws.stats_all[rank].copy_(ws.stats_local)
        eng.push_copy(ws.stats_all[rank], sptrs, rank * 3 * N * 4, rank)
        shdl.barrier()
    elif world > 1:
        dist.all_gather_into_tensor(ws.stats_all.view(world * 3, N), ws.stats_local)
    if world > 1:
        eng.lse2_merge(ws.stats_all[0, 0], ws.stats_all[0, 1], world, 3 * N, N, ws.lse2_col_all)
        ws.lse2_row_all[:N].view(world, n).copy_(ws.stats_all[:, 2, :n])
    else:
        eng.lse2_merge(col_m, col_l, 1, N, N, ws.lse2_col_all)
        ws.lse2_row_all[:N].copy_(row_lse)


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    @_scoped
    def forward(ctx, image_features, text_features, logit_scale, module, raw=False):
        eng = _engine()
        device = image_features.device
        n, d = image_features.shape
        world, rank = (module.world_size, module.rank) if module.world_size > 1 else (1, 0)
        ws = module._pool.take(eng, device, n, world, d)
        N = ws.N
        scale = _scalar_f32(logit_scale, device, positive=not raw)
        shape = Shape(n, N, d, rank * n)
        rows = slice(rank * n, (rank + 1) * n)

        # (local_loss, not gather_with_grad) gives the two gradients different G matrices; it keeps the gmat path
        split_g = world > 1 and module.local_loss and not module.gather_with_grad
        split_g = split_g or (world > 1 and (n < 8 or world > 64))   # limits of the per-owner entropy sums
        use_emat = any(ctx.needs_input_grad) and _backend(eng, ws) == "emat" and not split_g
        # d logit_scale of the multi-rank local loss from forward-side row sums and <dT_r, T_r>, so that the rescale pass
        # carries no entropy arithmetic (MRCLIP_DS=entropy switches back to the entropy sums)
        need_grad = any(ctx.needs_input_grad)
        ctx.fast = None
        if (not split_g or not need_grad) and _step_path_ok(eng, ws, world, need_grad):
            # one C call launches the whole forward (csrc/mrclip_cabi.cu: mrclip_step_forward); on several ranks the text
            # all-gather rides on NVLink peer stores and overlaps the tiles on this rank's own columns
            plan = ws.step_plan(eng, KIND_CLIP, module.local_loss or world == 1, rank)
            if world > 1:
                ws.flip ^= 1
            loss = torch.empty((), dtype=torch.float32, device=device)   # 0-d, not a view: callers may modify it in place
            plan.forward(eng, ws.flip if world > 1 else 0, _rows_major(image_features), _rows_major(text_features), scale,
                         None, need_grad, loss, raw=raw)
            ws.has_emat = need_grad
            ctx.fast = (plan, ws.flip if world > 1 else 0)
            ctx.raw = raw
            ctx.ws, ctx.module, ctx.shape_args = ws, module, (n, N, d, rank, world)
            ctx.lease = module._pool.lease(ws)
            ctx.scale = plan.scale_buf if raw else scale      # raw: exp(log-scale), written by the pack pre-pass
            ctx.in_dtypes = (image_features.dtype, text_features.dtype)
            ctx.scale_meta = (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None
            if not need_grad:
                ctx.lease.release()
            return loss
        if raw:
            ws.in_use = False
            raise _RawUnavailable()
        ctx.raw = False
        ctx.fwd_ds = bool(use_emat and world > 1 and module.local_loss and module.gather_with_grad
                          and ctx.needs_input_grad[1] and ctx.needs_input_grad[2]
                          and os.environ.get("MRCLIP_DS", "fwd").lower() != "entropy"
                          and n * N >= _FWD_DS_MIN_PAIRS and d % 4 == 0 and eng.fwd_row_ent_ok(n, N, n))
        _lse_forward(eng, ws, image_features, text_features, scale, shape, rank, world, use_emat,
                     gather_images=not use_emat and any(ctx.needs_input_grad), row_ent=ctx.fwd_ds)
        loss = torch.empty((), dtype=torch.float32, device=device)   # 0-d, not a view: callers may modify it in place
        eng.clip_loss(ws.lse2_row_all[rows], ws.lse2_col_all, ws.diag2, n, rank * n, loss)
        ctx.loss_local = loss.clone()    # private: the caller may modify the returned loss in place (loss /= accum)
        if world > 1 and not module.local_loss:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            loss /= world

        ctx.ws, ctx.module, ctx.shape_args = ws, module, (n, N, d, rank, world)
        ctx.lease = module._pool.lease(ws)
        ctx.scale = scale
        ctx.in_dtypes = (image_features.dtype, text_features.dtype)
        ctx.scale_meta = (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None
        if not any(ctx.needs_input_grad):
            ctx.lease.release()          # inference / no_grad: nothing will come back for these buffers
        return loss

    @staticmethod
    @_scoped
    def backward(ctx, grad_output):
        eng = _engine()
        ctx.lease.check()
        ws, module = ctx.ws, ctx.module
        n, N, d, rank, world = ctx.shape_args
        device = ws.img_all.device
        rows = slice(rank * n, (rank + 1) * n)
        shape = Shape(n, N, d, rank * n)
        gout = grad_output.detach().reshape(1).to(torch.float32).contiguous()
        if world > 1 and not module.local_loss and not module.gather_with_grad:
            coef = 0.5 / N          # only the re-inserted local slot carries gradient (loss.py:58-61)
        else:
            coef = 0.5 / n          # 1/(2n): W x the global-mean gradient, or the local loss itself
        w_oth = 0.0 if (world > 1 and module.local_loss and not module.gather_with_grad) else 1.0

        need_i, need_t, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_img = d_txt = d_scale = None
        ds = torch.zeros((1,), dtype=torch.float32, device=device) if (need_s and ctx.fast is None) else None
        # d logit_scale always uses 1/(2n) per rank; global modes average it over ranks below
        d_img = torch.empty((n, d), dtype=ctx.in_dtypes[0], device=device)
        d_txt = torch.empty((n, d), dtype=ctx.in_dtypes[1], device=device)
        ds_done = False
        if ctx.fast is not None:
            plan, flip = ctx.fast
            ds = torch.empty((1,), dtype=torch.float32, device=device) if need_s else None
            plan.backward(flip, ctx.scale, gout, coef, d_img, d_txt, ds, None)
            if ctx.raw:       # chain to the un-normalised tower outputs and to the log-scale parameter
                plan.normalize_bwd(flip, d_img, d_txt)
                if need_s:
                    ds = ds * ctx.scale
            ds_done = True
        elif ws.has_emat:
            gmat = ws.gmat_buffer(eng)
            # d logit_scale: one rank with a large block takes <dI, I>/scale from the GEMM's reduce (free; the bf16
            # rounding of G averages out); otherwise the rescale pass accumulates the softmax entropies, which is
            # exact per rank and insensitive to that rounding
            use_dot = world == 1 and n * N >= (1 << 22)
            fwd_ds = ctx.fwd_ds and need_s and need_t
            rsplit = ws.msums.sum(0)[0] if fwd_ds else None     # before the rescale pass can reuse the buffer
            msums = ws.msums if (need_s and not use_dot and not fwd_ds) else None
            eng.emat_to_gmat(ws.img_all[rows], ws.txt_all, shape, ws.lse2_row_all[rows], ws.lse2_col_all, ws.diag2,
                             ctx.scale, 1.0, 1.0, ws.scratch, gmat, msums, n, world)
            if world == 1:
                eng.gmat_gemm_dot(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img,
                                  ws.img_all[rows] if (need_s and use_dot) else None, ds)
                eng.gmat_gemm(True, gmat, shape, ws.img_all[rows], coef, ctx.scale, gout, ws.scratch, d_txt)
                if need_s and not use_dot:
                    ds = (gout / ctx.scale) * (ctx.loss_local + (0.6931471805599453 * 0.5 / n) * ws.msums.sum())
                    ds_done = True
            elif fwd_ds:
                # s dL_r/ds = <dT_r, T_r> + ln2/(2n) (R2(r,*) - R2(*,r)),  R2(q, r) = sum_{i in q, j in r} Prow_ij S2_ij:
                # R2(r, .) is local (forward), R2(*, r) one W-float all-reduce, the dot rides on the slot sum
                colsum = rsplit.clone()
                work_cs = dist.all_reduce(colsum, op=dist.ReduceOp.SUM, async_op=True)
                finish_dt = _text_grad_scatter(eng, ws, gmat, shape, coef, ctx.scale, gout, rank, d_txt,
                                               dot_slots=ws.dot_slots)
                eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
                finish_dt()
                work_cs.wait()
                ds = ((ws.dot_slots.sum() + gout * ((0.6931471805599453 * 0.5 / n) * (rsplit.sum() - colsum[rank])))
                      / ctx.scale).reshape(1)
                ds_done = True
            else:
                colsum = work_cs = None
                if need_s:   # column-softmax entropies of my columns live on every rank: W-float all-reduce, async
                    ms = ws.msums.sum(0)
                    colsum = ms[1].clone()
                    work_cs = dist.all_reduce(colsum, op=dist.ReduceOp.SUM, async_op=True)
                # text gradient: every rank holds G_q^T . I_q for all N text rows; the owner sums them
                finish_dt = _text_grad_scatter(eng, ws, gmat, shape, coef, ctx.scale, gout, rank, d_txt)
                eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
                if need_s:
                    # scale * dL_r/dscale = L_r + ln2/(2n) * (sum P log2 P over my rows, row softmax, all columns
                    #                                         + over my columns, column softmax, all rows)
                    work_cs.wait()
                    ent = ms[0].sum() + colsum[rank]
                    ds = ((gout / ctx.scale) * (ctx.loss_local + (0.6931471805599453 * 0.5 / n) * ent)).reshape(1)
                    if not module.local_loss:
                        dist.all_reduce(ds, op=dist.ReduceOp.SUM)
                        ds = ds / world
                    ds_done = True
                finish_dt()
        elif _use_gmat(eng, ws):
            gmat = ws.gmat_buffer(eng)
            # image rows vs all texts: G block of this rank's rows -> dI_r
            eng.clip_gwrite(ws.img_all[rows], ws.txt_all, shape, ws.lse2_row_all[rows], ws.lse2_col_all, ctx.scale,
                            1.0, w_oth, coef, gout, ws.scratch, gmat, ds, True, world == 1)
            eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
            if world == 1:
                # one rank: the same block, contracted along its rows, is dT
                eng.gmat_gemm(True, gmat, shape, ws.img_all[rows], coef, ctx.scale, gout, ws.scratch, d_txt)
            else:
                eng.clip_gwrite(ws.txt_all[rows], ws.img_all, shape, ws.lse2_col_all[rows], ws.lse2_row_all,
                                ctx.scale, 1.0, w_oth, coef, gout, ws.scratch, gmat, ds, True, False)
                eng.gmat_gemm(False, gmat, shape, ws.img_all, coef, ctx.scale, gout, ws.scratch, d_txt)
        else:
            _ensure_transposed(eng, ws)
            eng.clip_bwd(ws.img_all[rows], ws.txt_all, ws.txt_t, shape, ws.lse2_row_all[rows], ws.lse2_col_all,
                         ctx.scale, 1.0, w_oth, coef, gout, ws.scratch, d_img, ds, True)
            eng.clip_bwd(ws.txt_all[rows], ws.img_all, ws.img_t, shape, ws.lse2_col_all[rows], ws.lse2_row_all,
                         ctx.scale, 1.0, w_oth, coef, gout, ws.scratch, d_txt, ds, True)
        if need_s:
            if not ds_done:
                if coef != 0.5 / n:
                    ds *= (0.5 / n) / coef
                if world > 1 and not module.local_loss:
                    dist.all_reduce(ds, op=dist.ReduceOp.SUM)
                    ds /= world
            shp, dt = ctx.scale_meta
            d_scale = ds.reshape(shp).to(dt)
        ctx.lease.release()
        return (d_img if need_i else None), (d_txt if need_t else None), d_scale, None, None


class ClipLoss(nn.Module):
    """Reference ``ClipLoss`` (loss.py:68-139) on the fused sm_100a path."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False):
        super().__init__()
        if use_horovod:
            raise NotImplementedError("Horovod is not supported by mrclip_b200 (NCCL over NVLink only)")
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod

        # cache state
        self.prev_num_logits = 0
        self.labels = {}
        self._pool = _WorkspacePool()

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        """loss.py:91-102, bit-exact: arange (+ num_logits*rank for local loss), cached per device."""
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def get_logits(self, image_features, text_features, logit_scale):
        """Materialised logits (loss.py:104-126) for subclasses that need them (CoCa / distill / multi-positive).

        Not used by ``forward``; plain tensor algebra with the reference's operand order.
        """
        if self.world_size > 1:
            all_image_features, all_text_features = gather_features(
                image_features, text_features, local_loss=self.local_loss, gather_with_grad=self.gather_with_grad,
                rank=self.rank, world_size=self.world_size, use_horovod=self.use_horovod)
            if self.local_loss:
                logits_per_image = logit_scale * image_features @ all_text_features.T
                logits_per_text = logit_scale * text_features @ all_image_features.T
            else:
                logits_per_image = logit_scale * all_image_features @ all_text_features.T
                logits_per_text = logits_per_image.T
        else:
            logits_per_image = logit_scale * image_features @ text_features.T
            logits_per_text = logit_scale * text_features @ image_features.T
        return logits_per_image, logits_per_text

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        total_loss = _ClipLossFn.apply(image_features, text_features, logit_scale, self)
        return {"contrastive_loss": total_loss} if output_dict else total_loss

    def forward_raw(self, image_embeds, text_embeds, log_logit_scale, output_dict=False):
        """Feature hand-off fusion: the same loss taken from the towers' UN-normalised outputs and the
        model's log-scale parameter -- ``forward(F.normalize(i), F.normalize(t), log_logit_scale.exp())`` of the reference
        (model.py:282-301, :324 feeding loss.py:128-139) with the two normalisations, the casts and the exponential folded
        into the pack pre-pass and their backward into one in-place pass over the gradients.  Falls back to exactly that
        composition when the whole-step path is unavailable (CPU stand-in engine, MRCLIP_STEP=py, NCCL transports)."""
        try:
            total_loss = _ClipLossFn.apply(image_embeds, text_embeds, log_logit_scale, self, True)
        except _RawUnavailable:
            return self.forward(torch.nn.functional.normalize(image_embeds, dim=-1),
                                torch.nn.functional.normalize(text_embeds, dim=-1), log_logit_scale.exp(), output_dict)
        return {"contrastive_loss": total_loss} if output_dict else total_loss


class _MultiPositiveFn(torch.autograd.Function):
    """MR-CLIP's multi-positive (SupCon Eq. 2) loss on the same tiles.

    With P(i) the samples of the global batch that share sample i's label and c = |P(i)|:
        loss_img = mean_i( lse_row_i - (1/c) sum_{j in P(i)} S_ij ),     loss_txt = the same along columns,
        L = delta * loss_img + (1 - delta) * loss_txt                                  (reference loss.py:626-644, :745)
    so the forward is the ClipLoss forward (row / column LSE) plus class means of the features, which are O(N*D):
        (1/c) sum_{j in P(i)} S_ij = s * <I_i, mean_{P(i)} T>.
    The gradient of the logits is  G = delta*Prow + (1-delta)*Pcol - [same label]/c  =  G_k + delta_ij - [same]/c with
    G_k what the emat pipeline already builds for weights (delta, 1-delta); the difference is again a class mean:
        dI_i = s/n * ( (G_k T)_i + T_i - mean_{P(i)} T ),      dT_j = s/n * ( (G_k^T I)_j + I_j - mean_{P(j)} I ).
    Nothing N x N beyond the bf16 E block is ever formed (the reference builds two [n, N] fp32 logit matrices, the
    [n, N] mask and their softmaxes)."""

    @staticmethod
    @_scoped
    def forward(ctx, image_features, text_features, logit_scale, labels, delta, module):
        eng = _engine()
        device = image_features.device
        n, d = image_features.shape
        world, rank = (module.world_size, module.rank) if module.world_size > 1 else (1, 0)
        if labels.dim() != 1 or labels.shape[0] != n:
            raise ValueError(f"tokenized_texts must be one label per sample ([{n}]), got {tuple(labels.shape)}")
        if world > 1 and not (module.local_loss and module.gather_with_grad):
            # (the reference's mask is [n, N]: it only fits the logits with local_loss, loss.py:707-712)
            raise NotImplementedError("MultiPositiveClipLoss on several ranks needs local_loss=True, gather_with_grad=True")
        if world > 1 and (n < 8 or world > 64):
            raise NotImplementedError("MultiPositiveClipLoss: per-rank batch must be >= 8 and world_size <= 64")
        ws = module._pool.take(eng, device, n, world, d)
        N = ws.N
        scale = _scalar_f32(logit_scale, device, positive=True)
        shape = Shape(n, N, d, rank * n)
        rows = slice(rank * n, (rank + 1) * n)
        keep_e = any(ctx.needs_input_grad)
        # the class means need every rank's rows of both modalities, and every rank's labels
        _lse_forward(eng, ws, image_features, text_features, scale, shape, rank, world, keep_e, gather_images=True)
        labels = labels.detach().to(device=device, dtype=torch.long).contiguous()
        if world > 1:
            labels_all = torch.empty((N,), dtype=torch.long, device=device)
            dist.all_gather_into_tensor(labels_all, labels)
        else:
            labels_all = labels
        # dense class ids in [0, N) without a data-dependent shape (torch.unique would synchronise the host): sort, flag the
        # first sample of every class, prefix-sum the flags; every class is then a contiguous group of `order`
        sorted_l, order = torch.sort(labels_all)
        first = torch.ones_like(sorted_l)
        first[1:] = (sorted_l[1:] != sorted_l[:-1]).to(sorted_l.dtype)
        cid_sorted = torch.cumsum(first, 0) - 1                      # class id of every sorted position
        inv_all = torch.empty_like(sorted_l)
        inv_all[order] = cid_sorted                                  # class id of every sample
        seg_start = torch.full((N,), N, dtype=torch.long, device=device).scatter_reduce_(
            0, cid_sorted, torch.arange(N, device=device), reduce="amin")
        seg_cnt = torch.zeros((N,), dtype=torch.long, device=device).index_add_(0, cid_sorted, torch.ones_like(cid_sorted))
        i32 = torch.int32
        # class means of the packed (bf16) features the kernels see, one launch per modality (csrc/aux_kernels.cuh)
        t_mean = torch.empty((N, ws.ld), dtype=torch.float32, device=device)
        i_mean = torch.empty((N, ws.ld), dtype=torch.float32, device=device)
        order32, start32, cnt32 = order.to(i32), seg_start.to(i32), seg_cnt.to(i32)
        eng.class_means(ws.txt_all, order32, start32, cnt32, t_mean)
        eng.class_means(ws.img_all, order32, start32, cnt32, i_mean)
        cls_r = inv_all[rows].to(i32).contiguous()
        # loss = delta * mean_i(lse_row_i - s <I_i, mean_{P(i)} T>) + (1 - delta) * mean_i(lse_col_i - s <T_i, mean_{P(i)} I>)
        loss = torch.empty((), dtype=torch.float32, device=device)   # 0-d, not a view: callers may modify it in place
        eng.mpos_forward(ws.img_all[rows], ws.txt_all[rows], d, cls_r, t_mean, i_mean, ws.lse2_row_all[rows],
                         ws.lse2_col_all[rows], scale, delta, loss)

        ctx.ws, ctx.module, ctx.shape_args = ws, module, (n, N, d, rank, world)
        ctx.lease = module._pool.lease(ws)
        ctx.scale, ctx.delta, ctx.loss_local = scale, float(delta), loss.clone()
        ctx.cls = (cls_r, t_mean, i_mean)
        ctx.in_dtypes = (image_features.dtype, text_features.dtype)
        ctx.scale_meta = (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None
        if not keep_e:
            ctx.lease.release()          # inference / no_grad: nothing will come back for these buffers
        return loss

    @staticmethod
    @_scoped
    def backward(ctx, grad_output):
        eng = _engine()
        ctx.lease.check()
        ws, module = ctx.ws, ctx.module
        n, N, d, rank, world = ctx.shape_args
        device = ws.img_all.device
        rows = slice(rank * n, (rank + 1) * n)
        shape = Shape(n, N, d, rank * n)
        gout = grad_output.detach().reshape(1).to(torch.float32).contiguous()
        coef, delta = 1.0 / n, ctx.delta
        need_i, need_t, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_img = torch.empty((n, d), dtype=ctx.in_dtypes[0], device=device)
        d_txt = torch.empty((n, d), dtype=ctx.in_dtypes[1], device=device)
        gmat = ws.gmat_buffer(eng)
        # G_k = delta*Prow + (1-delta)*Pcol - delta_ij  (exact positives); entropies weighted likewise
        eng.emat_to_gmat(ws.img_all[rows], ws.txt_all, shape, ws.lse2_row_all[rows], ws.lse2_col_all, ws.diag2,
                         ctx.scale, delta, 1.0 - delta, ws.scratch, gmat, ws.msums if need_s else None, n, world)
        ms = ws.msums.sum(0) if need_s else None
        if world == 1:
            eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
            eng.gmat_gemm(True, gmat, shape, ws.img_all[rows], coef, ctx.scale, gout, ws.scratch, d_txt)
            ent = ms.sum() if need_s else None
        else:
            colsum = work_cs = None
            if need_s:
                colsum = ms[1].clone()
                work_cs = dist.all_reduce(colsum, op=dist.ReduceOp.SUM, async_op=True)
            finish_dt = _text_grad_scatter(eng, ws, gmat, shape, coef, ctx.scale, gout, rank, d_txt)
            eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
            ent = None
            if need_s:
                work_cs.wait()
                ent = ms[0].sum() + colsum[rank]
            finish_dt()
        # class-mean corrections: + s/n * (T_i - mean_{P(i)} T)  and  + s/n * (I_j - mean_{P(j)} I), in place
        cls_r, t_mean, i_mean = ctx.cls
        eng.mpos_backward(d_img, d_txt, ws.img_all[rows], ws.txt_all[rows], d, cls_r, t_mean, i_mean, coef, ctx.scale, gout)
        d_scale = None
        if need_s:
            # scale * dL/dscale = L + ln2/n * (delta * sum Prow log2 Prow + (1-delta) * sum Pcol log2 Pcol)
            ds = (gout / ctx.scale) * (ctx.loss_local + (0.6931471805599453 * coef) * ent)
            shp, dt = ctx.scale_meta
            d_scale = ds.reshape(shp).to(dt)
        ctx.lease.release()
        return (d_img if need_i else None), (d_txt if need_t else None), d_scale, None, None, None


class MultiPositiveClipLoss(ClipLoss):
    """Reference ``MultiPositiveClipLoss`` (loss.py:671-747): samples whose ``tokenized_texts`` label (MR-CLIP's binned
    TE / TR / TI class, train.py:123) is equal are positives of each other.  Same constructor and call signature;
    returns ``{"multi contrastive_loss": loss}`` with ``output_dict=True`` like the reference."""

    def get_logits_custom(self, image_features, text_features, text_tokens, logit_scale):
        """Materialised logits plus the gathered labels (loss.py:672-694); not used by ``forward``."""
        if self.world_size > 1:
            logits_per_image, logits_per_text = self.get_logits(image_features, text_features, logit_scale)
            return logits_per_image, logits_per_text, _all_gather_tokens(text_tokens, self.world_size)
        logits_per_image, logits_per_text = self.get_logits(image_features, text_features, logit_scale)
        return logits_per_image, logits_per_text, text_tokens

    def forward(self, image_features, text_features, logit_scale, delta=0.5, tokenized_texts=None, output_dict=False):
        if tokenized_texts is None:
            raise ValueError("MultiPositiveClipLoss needs tokenized_texts (one integer label per sample)")
        total_loss = _MultiPositiveFn.apply(image_features, text_features, logit_scale, tokenized_texts, float(delta),
                                            self)
        return {"multi contrastive_loss": total_loss} if output_dict else total_loss


class _SigLipLossFn(torch.autograd.Function):
    @staticmethod
    @_scoped
    def forward(ctx, image_features, text_features, logit_scale, logit_bias, module, raw=False):
        eng = _engine()
        device = image_features.device
        n, d = image_features.shape
        world, rank = (module.world_size, module.rank) if module.world_size > 1 else (1, 0)
        ws = module._pool.take(eng, device, n, world, d)
        N = ws.N
        scale = _scalar_f32(logit_scale, device, positive=not raw)
        bias = _scalar_f32(logit_bias, device) if logit_bias is not None else None
        shape = Shape(n, N, d, rank * n)
        rows = slice(rank * n, (rank + 1) * n)
        use_emat = any(ctx.needs_input_grad) and _backend(eng, ws) == "emat"
        loss = torch.empty((), dtype=torch.float32, device=device)   # 0-d, not a view: callers may modify it in place
        ctx.fast = None
        if _step_path_ok(eng, ws, world, any(ctx.needs_input_grad)):
            plan = ws.step_plan(eng, KIND_SIGLIP, True, rank)
            if world > 1:
                ws.flip ^= 1
            plan.forward(eng, ws.flip if world > 1 else 0, _rows_major(image_features), _rows_major(text_features), scale,
                         bias, any(ctx.needs_input_grad), loss, raw=raw)
            ws.has_emat = any(ctx.needs_input_grad)
            ctx.fast = (plan, ws.flip if world > 1 else 0)
            if raw:
                scale = plan.scale_buf
        elif raw:
            ws.in_use = False
            raise _RawUnavailable()
        else:
            _gather_packed(eng, ws, image_features, text_features, rank, world,
                           gather_images=not use_emat and any(ctx.needs_input_grad))
        if ctx.fast is not None:
            pass
        elif use_emat:
            # no normaliser: the forward can store G = sigmoid(z) - delta itself
            eng.siglip_fwd_e(ws.img_all[rows], ws.txt_all, shape, scale, bias, ws.scratch, loss, ws.gmat_buffer(eng))
            ws.has_emat = True
        else:
            eng.siglip_fwd(ws.img_all[rows], ws.txt_all, shape, scale, bias, ws.scratch, loss)
        ctx.ws, ctx.module, ctx.shape_args = ws, module, (n, N, d, rank, world)
        ctx.lease = module._pool.lease(ws)
        ctx.scale, ctx.bias, ctx.raw = scale, bias, raw
        ctx.in_dtypes = (image_features.dtype, text_features.dtype)
        ctx.scale_meta = (logit_scale.shape, logit_scale.dtype) if torch.is_tensor(logit_scale) else None
        ctx.bias_meta = (logit_bias.shape, logit_bias.dtype) if torch.is_tensor(logit_bias) else None
        if not any(ctx.needs_input_grad):
            ctx.lease.release()          # inference / no_grad: nothing will come back for these buffers
        return loss

    @staticmethod
    @_scoped
    def backward(ctx, grad_output):
        eng = _engine()
        ctx.lease.check()
        ws, module = ctx.ws, ctx.module
        n, N, d, rank, world = ctx.shape_args
        device = ws.img_all.device
        rows = slice(rank * n, (rank + 1) * n)
        shape = Shape(n, N, d, rank * n)
        gout = grad_output.detach().reshape(1).to(torch.float32).contiguous()
        coef = 1.0 / n
        need_i, need_t, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_b = ctx.needs_input_grad[3] and ctx.bias_meta is not None
        d_img = d_txt = d_scale = d_bias = None
        mk = torch.empty if ctx.fast is not None else torch.zeros      # (the step entry overwrites, the legacy kernels add)
        ds = mk((1,), dtype=torch.float32, device=device) if need_s else None
        db = mk((1,), dtype=torch.float32, device=device) if need_b else None
        d_img = torch.empty((n, d), dtype=ctx.in_dtypes[0], device=device)
        d_txt = torch.empty((n, d), dtype=ctx.in_dtypes[1], device=device)
        # every rank's loss touches T_r: the column block gives the summed (W x) text gradient directly
        if ctx.fast is not None:
            plan, flip = ctx.fast
            plan.backward(flip, ctx.scale, gout, coef, d_img, d_txt, ds, db)
            if ctx.raw:
                plan.normalize_bwd(flip, d_img, d_txt)
                if need_s:
                    ds = ds * ctx.scale
        elif ws.has_emat:
            gmat = ws.gmat_buffer(eng)
            eng.siglip_e_scalars(shape, ws.scratch, coef, gout, ds, db, False)
            if world == 1:
                eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
                eng.gmat_gemm(True, gmat, shape, ws.img_all[rows], coef, ctx.scale, gout, ws.scratch, d_txt)
            else:
                finish_dt = _text_grad_scatter(eng, ws, gmat, shape, coef, ctx.scale, gout, rank, d_txt)
                eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
                finish_dt()
        elif _use_gmat(eng, ws):
            gmat = ws.gmat_buffer(eng)
            eng.siglip_gwrite(ws.img_all[rows], ws.txt_all, shape, ctx.scale, ctx.bias, coef, gout, ws.scratch, gmat,
                              ds, db, False)
            eng.gmat_gemm(False, gmat, shape, ws.txt_all, coef, ctx.scale, gout, ws.scratch, d_img)
            if world == 1:
                eng.gmat_gemm(True, gmat, shape, ws.img_all[rows], coef, ctx.scale, gout, ws.scratch, d_txt)
            else:
                eng.siglip_gwrite(ws.txt_all[rows], ws.img_all, shape, ctx.scale, ctx.bias, coef, gout, ws.scratch,
                                  gmat, None, None, False)
                eng.gmat_gemm(False, gmat, shape, ws.img_all, coef, ctx.scale, gout, ws.scratch, d_txt)
        else:
            _ensure_transposed(eng, ws)
            eng.siglip_bwd(ws.img_all[rows], ws.txt_all, ws.txt_t, shape, ctx.scale, ctx.bias, coef, gout, ws.scratch,
                           d_img, ds, db, False)
            eng.siglip_bwd(ws.txt_all[rows], ws.img_all, ws.img_t, shape, ctx.scale, ctx.bias, coef, gout, ws.scratch,
                           d_txt, None, None, False)
        if need_s:
            shp, dt = ctx.scale_meta
            d_scale = ds.reshape(shp).to(dt)
        if need_b:
            shp, dt = ctx.bias_meta
            d_bias = db.reshape(shp).to(dt)
        ctx.lease.release()
        return (d_img if need_i else None), (d_txt if need_t else None), d_scale, d_bias, None, None


class SigLipLoss(nn.Module):
    """Reference ``SigLipLoss`` (loss.py:314-448).

    All four ``dist_impl`` exchange schemes of the reference compute the same loss; on an NVSwitch
    domain a single all-gather of the bf16 features replaces the W-1 neighbour hops, so ``dist_impl``
    is validated and stored but does not change the communication pattern.
    """

    def __init__(self, cache_labels: bool = False, rank: int = 0, world_size: int = 1, dist_impl: Optional[str] = None):
        super().__init__()
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.dist_impl = dist_impl or 'bidir'
        assert self.dist_impl in ('bidir', 'shift', 'reduce', 'gather')

        # cache state (unused, as in the reference)
        self.prev_num_logits = 0
        self.labels = {}
        self._pool = _WorkspacePool()

    def get_ground_truth(self, device, dtype, num_logits, negative_only=False) -> torch.Tensor:
        """loss.py:338-342: -1 everywhere, +1 on the diagonal unless negative_only."""
        labels = -torch.ones((num_logits, num_logits), device=device, dtype=dtype)
        if not negative_only:
            labels = 2 * torch.eye(num_logits, device=device, dtype=dtype) + labels
        return labels

    def get_logits(self, image_features, text_features, logit_scale, logit_bias=None):
        """Materialised chunk logits (loss.py:344-348); not used by ``forward``."""
        logits = logit_scale * image_features @ text_features.T
        if logit_bias is not None:
            logits = logits + logit_bias
        return logits

    def _loss(self, image_features, text_features, logit_scale, logit_bias=None, negative_only=False):
        """One materialised chunk of the pairwise sigmoid loss (loss.py:354-363): -sum logsigmoid(label * logit) / n.
        Not used by ``forward`` (the tile kernel sums softplus over the row block without forming logits)."""
        logits = self.get_logits(image_features, text_features, logit_scale, logit_bias)
        labels = self.get_ground_truth(image_features.device, image_features.dtype, image_features.shape[0],
                                       negative_only=negative_only)
        return -torch.nn.functional.logsigmoid(labels * logits).sum() / image_features.shape[0]

    def forward(self, image_features, text_features, logit_scale, logit_bias, output_dict=False):
        loss = _SigLipLossFn.apply(image_features, text_features, logit_scale, logit_bias, self)
        return {"contrastive_loss": loss} if output_dict else loss

    def forward_raw(self, image_embeds, text_embeds, log_logit_scale, logit_bias, output_dict=False):
        """``forward(F.normalize(i), F.normalize(t), log_logit_scale.exp(), logit_bias)`` with the normalisations, casts and
        the exponential fused into the pack pre-pass (see ``ClipLoss.forward_raw``)."""
        try:
            loss = _SigLipLossFn.apply(image_embeds, text_embeds, log_logit_scale, logit_bias, self, True)
        except _RawUnavailable:
            return self.forward(torch.nn.functional.normalize(image_embeds, dim=-1),
                                torch.nn.functional.normalize(text_embeds, dim=-1), log_logit_scale.exp(), logit_bias,
                                output_dict)
        return {"contrastive_loss": loss} if output_dict else loss
