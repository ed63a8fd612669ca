"""Step drivers with the reference's signatures (train.py:19-51), on the B200 kernels.

``train_imagine_beam`` / ``train_nmt`` do zero_grad → forward → backward → clip → Adam exactly like the reference;
the clip and the Adam update are fused in ``ClipAdam.step`` (vag_nmt_b200/optim.py), which also all-reduces the
gradients first when torch.distributed is initialised (one process per GPU, NCCL over NVLink).
"""
from __future__ import annotations

import torch

from .losses import PairwiseRankingLoss
from .optim import ClipAdam

CLIP = 1.0


def train_nmt(input_variable, target_variable, input_lengths, model, criterion, optimizer: ClipAdam, teacher_force_ratio=0.5):
    """train.py:19-32"""
    model.train()
    optimizer.zero_grad()
    loss = model(input_variable, input_lengths, target_variable, teacher_force_ratio, criterion=criterion)
    with model.precision_scope():   # the backward contractions run in the same arithmetic mode as the forward ones
        loss.backward()
    optimizer.step(clip=CLIP)
    return loss.item()


def train_imagine_beam(input_variable, target_variable, im_variable, input_lengths, model, optimizer: ClipAdam, criterion_mt,
                       criterion_vse, loss_weight, teacher_force_ratio, max_length=40, clip=1, sync: bool = True):
    """train.py:36-51.  ``sync=False`` returns device scalars instead of Python floats (no host synchronisation)."""
    model.train()
    optimizer.zero_grad()
    loss, loss_mt, loss_vse = model(input_variable, input_lengths, target_variable, im_variable, teacher_force_ratio,
                                    criterion_mt=criterion_mt, criterion_vse=criterion_vse)
    with model.precision_scope():   # the backward contractions run in the same arithmetic mode as the forward ones
        loss.backward()
    optimizer.step(clip=clip)
    if not sync:
        return loss.detach(), loss_mt.detach(), loss_vse.detach() if torch.is_tensor(loss_vse) else loss_vse
    return loss.item(), loss_mt.item(), loss_vse.item() if torch.is_tensor(loss_vse) else loss_vse


class _GlobalRankLossFn(torch.autograd.Function):
    """Ranking loss over the GLOBAL batch under data parallelism: all-gather the local
    [B_local, S] embeddings, evaluate the full [B, B] hinge on every rank, keep the gradient rows of the local slice.
    The gradient is scaled by the world size so that the later all-reduce AVERAGE yields the single-process sum."""

    @staticmethod
    def forward(ctx, im, s, margin, one_direction, group):
        import torch.distributed as dist
        from . import ops
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        Bl = im.shape[0]
        im_all = torch.empty(world * Bl, im.shape[1], dtype=im.dtype, device=im.device)
        s_all = torch.empty_like(im_all)
        dist.all_gather_into_tensor(im_all, im.detach().contiguous(), group=group)
        dist.all_gather_into_tensor(s_all, s.detach().contiguous(), group=group)
        loss, g_im, g_s = ops.rank_loss(im_all, s_all, margin, one_direction, want_grad=True)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        ctx.save_for_backward(g_im[sl].clone(), g_s[sl].clone())
        ctx.scale = float(world)
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        g_im, g_s = ctx.saved_tensors
        return grad_out * ctx.scale * g_im, grad_out * ctx.scale * g_s, None, None, None


class DistributedPairwiseRankingLoss(PairwiseRankingLoss):
    """PairwiseRankingLoss whose B² pair sum spans all data-parallel ranks (identical to a single-process batch)."""

    def __init__(self, margin=1.0, group=None):
        super().__init__(margin)
        self.group = group

    def forward(self, im, s):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return super().forward(im, s)
        return _GlobalRankLossFn.apply(im, s, float(self.margin), self.one_direction, self.group)
