"""Retrieval metrics: text→image / image→text recall@1/5/10 and median rank.

Mirrors ``utils/im_retrieval_eval.py:4-58``.  The reference loops over the N queries issuing an ``mm`` + ``sort``
+ device→host copy each; here one contraction produces the [N, N] score matrix and one kernel counts, per query,
how many gallery items beat the true one (rank = #greater + #equal-with-lower-index), so there is a single
device→host copy of N int32 ranks.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _to_dev(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        x = x.cuda()
    return x.detach().to(torch.float32).contiguous()


def _metrics(ranks: np.ndarray):
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    medr = np.floor(np.median(ranks)) + 1
    return (r1, r5, r10, medr)


def t2i(images: torch.Tensor, captions: torch.Tensor):
    """Text → image.  images (N, K), captions (N, K) → (r1, r5, r10, medr).  im_retrieval_eval.py:4-30."""
    ranks = ops.recall_ranks(_to_dev(captions), _to_dev(images)).cpu().numpy().astype(np.float64)
    return _metrics(ranks)


def i2t(images: torch.Tensor, captions: torch.Tensor):
    """Image → text.  im_retrieval_eval.py:32-58."""
    ranks = ops.recall_ranks(_to_dev(images), _to_dev(captions)).cpu().numpy().astype(np.float64)
    return _metrics(ranks)
