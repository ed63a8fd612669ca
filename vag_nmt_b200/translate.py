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


def decode_corpus_pipelined(model, sents: Sequence[List[int]], im: Optional[torch.Tensor], beam_size: int, max_length: int,
                            batch_size: int, lanes: int = 4) -> List[List[int]]:
    """``decode_corpus`` in eval batches (the reference's loop, nmt_multimodal_beam_DE.py:542-547) with up to `lanes` batches in
    flight: batch i runs on CUDA stream i mod lanes with private scratch buffers (``ops.lane``) and its own captured graphs,
    and its tokens are fetched only when the lane is needed again.  A batch of 16 sentences occupies a small part of the GPU
    in every launch and its 80 steps are one dependent chain, so consecutive batches overlap almost perfectly; the tokens
    are those of the sequential loop (batches are independent)."""
    from collections import deque
    from . import ops
    n = len(sents)
    out: List[Optional[List[int]]] = [None] * n
    dev = model._device()
    text_only = im is None
    lanes = max(1, int(lanes))
    streams = model._lane_streams(lanes, dev)
    cur = torch.cuda.current_stream(dev)
    with model.precision_scope():
        ops.decoder_weights(model.decoder, model.decoderini, prepare=not model.training)   # shared invariants, on the caller's stream
    pending = deque()

    def collect():
        lo, order, hyp, hyp_len, ev = pending.popleft()
        ev.synchronize()
        if hyp_len is None:
            hyps = [_cut_eos(r) for r in hyp.cpu().tolist()]
        else:
            hyps = model._hyp_lists(hyp, hyp_len)
        for r, c in enumerate(order):
            out[lo + c] = [int(t) for t in hyps[r]]

    for i, lo in enumerate(range(0, n, max(batch_size, 1))):
        chunk = sents[lo:lo + batch_size]
        src, lens, im_sorted, order = pad_and_sort(chunk, None if text_only else im[lo:lo + batch_size])
        k = i % lanes
        if len(pending) >= lanes:
            collect()
        streams[k].wait_stream(cur)
        with torch.cuda.stream(streams[k]), ops.lane(f"pipe{k}:"), model.precision_scope():
            hyp, hyp_len = model._decode_device_one(src, lens, None if text_only else im_sorted, beam_size, max_length)
            ev = torch.cuda.Event()
            ev.record(streams[k])
        pending.append((lo, order, hyp, hyp_len, ev))
    while pending:
        collect()
    for st in streams:
        cur.wait_stream(st)
    return out  # type: ignore[return-value]


def _cut_eos(row: List[int]) -> List[int]:
    from .models import EOS_token
    cut = []
    for t in row:
        if t == EOS_token:
            break
        cut.append(t)
    return cut


def shard_indices(lengths: Sequence[int], world_size: int, rank: int, balance: bool = True) -> List[int]:
    """Corpus positions `rank` works on.  balance=True (default): sort the corpus by length (descending, stable) and deal it
    round-robin, so every rank gets the same mix of long and short sentences ("sort globally by length
    then deal round-robin") — a contiguous split leaves the rank holding the longest sentences as the straggler.
    balance=False: the contiguous ``shard_range`` slice."""
    n = len(lengths)
    if not balance:
        lo, hi = shard_range(n, world_size, rank)
        return list(range(lo, hi))
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    order = sorted(range(n), key=lambda i: -int(lengths[i]))     # sorted() is stable: equal lengths keep corpus order
    return order[rank::world_size]


def decode_corpus_sharded(decode_fn: Callable, sents: Sequence[List[int]], im: Optional[torch.Tensor], beam_size: int,
                          max_length: int, batch_size: Optional[int] = None, group=None, balance: bool = True) -> List[List[int]]:
    """Every rank decodes its ``shard_indices`` share; the full list (corpus order) is returned on every rank.
    The only communication is one ``all_gather_object`` of (positions, token lists) after decoding."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return decode_corpus(decode_fn, sents, im, beam_size, max_length, batch_size)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    idx = shard_indices([len(s) for s in sents], world, rank, balance)
    mine = decode_corpus(decode_fn, [sents[i] for i in idx], im[idx] if im is not None else None, beam_size, max_length, batch_size)
    parts: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(parts, (idx, mine), group=group)
    merged: List[Optional[List[int]]] = [None] * len(sents)
    for pidx, ptoks in parts:  # type: ignore[misc]
        for i, t in zip(pidx, ptoks):
            merged[i] = t
    return merged  # type: ignore[return-value]


# ------------------------------------------------------------------ retrieval evaluation
def embed_corpus(embed_fn: Callable, sents: Sequence[List[int]], im: torch.Tensor, batch_size: Optional[int] = None):
    """Shared-space embeddings of a corpus in CORPUS order → (lim [n, S], ltxt [n, S]).

    embed_fn(src [B, W], lengths, im [B, I]) → (im_emb [B, S], txt_emb [B, S]), i.e. ``model.embed_sent_im_test``.  Mirrors the
    evaluation loop of nmt_multimodal_beam_DE.py:551-564: batches are length-sorted for the encoder and their rows scattered
    back to the corpus positions (``lim[index_reorder] = test_im_vecs``)."""
    n = len(sents)
    lim = ltxt = None
    step = n if not batch_size else batch_size
    for lo in range(0, n, max(step, 1)):
        chunk = sents[lo:lo + step]
        src, lens, im_sorted, order = pad_and_sort(chunk, im[lo:lo + step])
        e_im, e_txt = embed_fn(src, lens, im_sorted)
        if lim is None:
            lim = torch.empty(n, e_im.shape[1], dtype=e_im.dtype, device=e_im.device)
            ltxt = torch.empty_like(lim)
        pos = torch.as_tensor([lo + c for c in order], device=e_im.device)
        lim[pos] = e_im
        ltxt[pos] = e_txt
    return lim, ltxt


def embed_corpus_sharded(embed_fn: Callable, sents: Sequence[List[int]], im: torch.Tensor, batch_size: Optional[int] = None,
                         group=None, balance: bool = True):
    """Every rank embeds its ``shard_indices`` share; two ``all_gather_into_tensor`` calls of the (equal-size, padded)
    [ceil(n/G), S] blocks put the full ``lim`` / ``ltxt`` [n, S] — corpus order — on every rank (2 MB at n = 1000)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return embed_corpus(embed_fn, sents, im, batch_size)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = len(sents)
    lengths = [len(s) for s in sents]
    idx = shard_indices(lengths, world, rank, balance)
    e_im, e_txt = embed_corpus(embed_fn, [sents[i] for i in idx], im[idx], batch_size)
    per = -(-n // world)                                  # every rank contributes `per` rows; short shards are zero-padded
    S = e_im.shape[1]

    def gathered(local):
        block = torch.zeros(per, S, dtype=local.dtype, device=local.device)
        block[:local.shape[0]] = local
        out = torch.empty(world * per, S, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, block, group=group)
        return out

    g_im, g_txt = gathered(e_im), gathered(e_txt)
    lim = torch.empty(n, S, dtype=e_im.dtype, device=e_im.device)
    ltxt = torch.empty_like(lim)
    for r in range(world):                               # every rank can recompute every rank's positions: no index exchange
        ridx = shard_indices(lengths, world, r, balance)
        pos = torch.as_tensor(ridx, device=lim.device)
        lim[pos] = g_im[r * per:r * per + len(ridx)]
        ltxt[pos] = g_txt[r * per:r * per + len(ridx)]
    return lim, ltxt


def retrieval_eval_sharded(embed_fn: Callable, sents: Sequence[List[int]], im: torch.Tensor, batch_size: Optional[int] = None,
                           group=None, rank_fn: Optional[Callable] = None):
    """Sharded embedding + REPLICATED recall computation → (r1, r5, r10, medr), identical on every rank and to the
    single-process evaluation (nmt_multimodal_beam_DE.py:551-566 → im_retrieval_eval.t2i)."""
    lim, ltxt = embed_corpus_sharded(embed_fn, sents, im, batch_size, group)
    if rank_fn is None:
        from .retrieval import t2i as rank_fn
    return rank_fn(lim, ltxt)
