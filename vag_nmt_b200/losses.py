"""Shared-embedding ranking losses — same classes as machine_translation_vision/losses/*.py.

forward(im, s) → scalar tensor.  When autograd is recording, the backward pass uses the gradient kernel of
libvagnmt.so (vag_rank_loss_f32 produces dLoss/dim and dLoss/ds together with the loss).
"""
from __future__ import annotations

import torch

from . import ops


class _RankLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, s, margin, one_direction):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        loss, g_im, g_s = ops.rank_loss(im.detach(), s.detach(), margin, one_direction, want_grad=need)
        if need:
            ctx.save_for_backward(g_im, g_s)
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        g_im, g_s = ctx.saved_tensors
        return grad_out * g_im, grad_out * g_s, None, None


class PairwiseRankingLoss(torch.nn.Module):
    """Σ_{i≠j} max(0, m − s_jj + s_ij) + max(0, m − s_ii + s_ij).  losses/PairwiseRankingLoss.py:4-24."""

    one_direction = False

    def __init__(self, margin=1.0):
        super().__init__()
        self.margin = margin

    def forward(self, im, s):
        return _RankLossFn.apply(im, s, float(self.margin), self.one_direction)


class ImageRetrievalRankingLoss(PairwiseRankingLoss):
    """cost_s only.  losses/ImageRetrievalRankingLoss.py:4-21."""

    one_direction = True
