"""Host batching, un-sorting and rank sharding around ``beamsearch_decode``.

The reference walks the test set in eval batches, sorts each batch by source length, decodes and restores the
corpus order (``data_generator_mtv`` preprocessing.py:234-306, ``translation_reorder_BPE`` :475-486, decode loop
nmt_multimodal_beam_DE.py:542-547).  Sentences are independent under decoding, so a
corpus shards across ranks with no data-path collective: every rank decodes a contiguous slice and the token
lists are gathered once at the end.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .synthetic import pad_and_sort


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def decode_corpus(decode_fn: Callable, sents: Sequence[List[int]], im: Optional[torch.Tensor], beam_size: int,
                  max_length: int, batch_size: Optional[int] = None) -> List[List[int]]:
    """Decode `sents` (token-id lists, any order) and return the translations in the SAME order.

    decode_fn(src [B, W] int64, lengths list, im [B, I] or None, beam_size, max_length) → list[B] of token lists,
    i.e. ``model.beamsearch_decode`` (text-only models ignore ``im``).  ``batch_size=None`` decodes everything in
    one call (the B200 path's preferred shape); 16 reproduces the reference's eval batching.
    """
    n = len(sents)
    out: List[Optional[List[int]]] = [None] * n
    step = n if not batch_size else batch_size
    for lo in range(0, n, max(step, 1)):
        chunk = sents[lo:lo + step]
        im_chunk = im[lo:lo + step] if im is not None else None
        src, lens, im_sorted, order = pad_and_sort(chunk, im_chunk)
        hyps = decode_fn(src, lens, im_sorted, beam_size, max_length)
        for r, c in enumerate(order):
            out[lo + c] = [int(t) for t in hyps[r]]
    return out  # type: ignore[return-value]


def decode_corpus_sharded(decode_fn: Callable, sents: Sequence[List[int]], im: Optional[torch.Tensor], beam_size: int,
                          max_length: int, batch_size: Optional[int] = None, group=None) -> List[List[int]]:
    """Every rank decodes its ``shard_range`` slice; the full list (corpus order) is returned on every rank.
    The only communication is one ``all_gather_object`` of the token lists after decoding."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return decode_corpus(decode_fn, sents, im, beam_size, max_length, batch_size)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(len(sents), world, rank)
    mine = decode_corpus(decode_fn, sents[lo:hi], im[lo:hi] if im is not None else None, beam_size, max_length, batch_size)
    parts: List[Optional[List[List[int]]]] = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    merged: List[List[int]] = []
    for p in parts:
        merged.extend(p)  # type: ignore[arg-type]
    return merged
