"""Ring exchanges of the reference's SigLIP path, kept for callers that import them.

Reference: ``src/open_clip/loss.py`` -- ``neighbour_exchange`` :226-244, ``neighbour_exchange_bidir`` :247-276,
``NeighbourExchange`` / ``neighbour_exchange_with_grad`` :279-293, ``NeighbourExchangeBidir`` /
``neighbour_exchange_bidir_with_grad`` :296-311.  Same names, argument order, return order and gradient rule
(the gradient of a hop is the opposite hop of the incoming gradient).

``SigLipLoss`` here does not use them: on an NVSwitch domain one all-gather of the packed text rows (peer stores over
NVLink) replaces the W-1 serialised hops, and the four ``dist_impl`` schemes give the same loss.  These
functions are plain ``torch.distributed`` point-to-point plumbing (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["neighbour_exchange", "neighbour_exchange_bidir", "neighbour_exchange_with_grad",
           "neighbour_exchange_bidir_with_grad", "NeighbourExchange", "NeighbourExchangeBidir"]


def _round(sends, recv_like, group):
    """One batched point-to-point round.  ``sends`` = [(tensor, dst)], ``recv_like`` = [(template, src)]; returns the
    received tensors in the order of ``recv_like``.  All sends are posted before the receives, as the reference does,
    so that two ranks talking to each other pair their operations identically."""
    inbox = [torch.zeros_like(t) for t, _ in recv_like]
    ops = [dist.P2POp(dist.isend, t, peer, group=group) for t, peer in sends]
    ops += [dist.P2POp(dist.irecv, buf, src, group=group) for buf, (_, src) in zip(inbox, recv_like)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return inbox


def neighbour_exchange(from_rank, to_rank, tensor, group=None):
    """Send ``tensor`` to ``to_rank`` and return what ``from_rank`` sent (same shape and dtype)."""
    return _round([(tensor, to_rank)], [(tensor, from_rank)], group)[0]


def neighbour_exchange_bidir(left_rank, right_rank, tensor_to_left, tensor_to_right, group=None):
    """Send one tensor to each neighbour; returns ``(tensor_from_right, tensor_from_left)``."""
    from_right, from_left = _round([(tensor_to_right, right_rank), (tensor_to_left, left_rank)],
                                   [(tensor_to_left, right_rank), (tensor_to_right, left_rank)], group)
    return from_right, from_left


class NeighbourExchange(torch.autograd.Function):
    """Differentiable hop: the gradient travels the hop backwards (from ``to_rank`` to ``from_rank``)."""

    @staticmethod
    def forward(ctx, from_rank, to_rank, group, tensor):
        ctx.hop = (from_rank, to_rank, group)
        return neighbour_exchange(from_rank, to_rank, tensor, group=group)

    @staticmethod
    def backward(ctx, grad_output):
        from_rank, to_rank, group = ctx.hop
        return None, None, None, NeighbourExchange.apply(to_rank, from_rank, group, grad_output)


def neighbour_exchange_with_grad(from_rank, to_rank, tensor, group=None):
    return NeighbourExchange.apply(from_rank, to_rank, group, tensor)


class NeighbourExchangeBidir(torch.autograd.Function):
    """Differentiable two-sided hop; the two incoming gradients go back with the neighbours swapped."""

    @staticmethod
    def forward(ctx, left_rank, right_rank, group, tensor_to_left, tensor_to_right):
        ctx.hop = (left_rank, right_rank, group)
        return neighbour_exchange_bidir(left_rank, right_rank, tensor_to_left, tensor_to_right, group=group)

    @staticmethod
    def backward(ctx, *grad_outputs):
        left_rank, right_rank, group = ctx.hop
        return (None, None, None) + NeighbourExchangeBidir.apply(right_rank, left_rank, group, *grad_outputs)


def neighbour_exchange_bidir_with_grad(left_rank, right_rank, tensor_to_left, tensor_to_right, group=None):
    return NeighbourExchangeBidir.apply(left_rank, right_rank, group, tensor_to_left, tensor_to_right)
