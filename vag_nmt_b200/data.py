"""Host batching on either side of the hot path.

* ``BucketBatchSampler`` — the training batch order of samplers/bucket.py:10-103: one bucket per TARGET length, so every
  training batch has a single target length and the teacher-forced loop carries no padding.  Same constructor, same
  ``__iter__`` / ``__len__`` contract and the same numpy-RNG call sequence (one permutation per bucket in insertion
  order, then one permutation of the bucket schedule), hence the same batches as the reference under the same
  ``np.random.seed``.  Extension for one-process-per-GPU data parallelism: ``world_size`` / ``rank`` cut every GLOBAL batch
  (``batch_size · world_size`` samples of one bucket) into equal per-rank slices, so all ranks run the same number of steps with
  the same local batch size (the ranking-loss all-gather and the gradient all-reduce need that).
* ``data_generator_tl_mtv`` / ``data_generator_mtv`` — the batch builders of preprocessing.py:308-384 / :234-306: pad with
  0, sort the batch by source length (descending, reference tie order), reorder targets and image rows alike.  Batches are
  built in pinned host memory and copied asynchronously to the device.
"""
from __future__ import annotations

import math
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch

PAD_token = 0


class BucketBatchSampler:
    """samplers/bucket.py:10-103.  ``lengths[i]`` is the target length of sample i.

    Data-parallel extension (``world_size`` > 1, one process per GPU): ``batch_size`` stays the PER-RANK batch.  The epoch's schedule
    is drawn over GLOBAL batches of ``batch_size · world_size`` samples of one bucket and every global batch is cut into
    ``world_size`` equal contiguous slices — every rank runs the same number of steps with the SAME local batch size, which the
    global-batch ranking loss (all-gather of [B_local, S] blocks) and the averaging gradient all-reduce both require.  A bucket's
    remainder that does not divide by ``world_size`` loses its last ``n mod world_size`` samples for that epoch (a remainder smaller
    than ``world_size`` is skipped by every rank).  All ranks must draw the same schedule: pass the same ``seed`` everywhere
    (a private ``RandomState`` advanced once per epoch); without a seed the global numpy RNG is used like the reference does, which
    is only correct when every rank seeded it identically."""

    def __init__(self, lengths: Sequence[int], batch_size: int, max_len: Optional[int] = None, world_size: int = 1, rank: int = 0,
                 seed: Optional[int] = None):
        if batch_size < 1:
            raise ValueError("batch_size must be positive")
        if not (0 <= rank < world_size):
            raise ValueError("rank outside the world")
        self.batch_size = int(batch_size)
        self.max_len = 10000 if max_len is None else max_len
        self.world_size, self.rank = int(world_size), int(rank)
        self._rng = np.random.RandomState(seed) if seed is not None else None
        members = {}                                   # length -> sample indices, in first-seen order (dict keeps insertion order)
        for idx, n in enumerate(lengths):
            n = int(n)
            if n <= self.max_len:
                members.setdefault(n, []).append(idx)
        self.buckets = {n: np.asarray(ix) for n, ix in members.items()}
        self.bucket_names = list(self.buckets)
        self._draw = self.batch_size * self.world_size  # samples per (global) batch
        schedule: List[int] = []                        # one entry per batch: the bucket it is drawn from
        for n, ix in self.buckets.items():
            schedule.extend([n] * math.ceil(ix.size / self._draw))
        self.bucket_idxs = np.asarray(schedule)
        if self.world_size == 1:
            self.n_batches = len(schedule)
        else:   # global batches that still hold at least one sample per rank
            self.n_batches = sum(1 for n, ix in self.buckets.items() for k in range(math.ceil(ix.size / self._draw))
                                 if min(self._draw, ix.size - k * self._draw) >= self.world_size)

    def _epoch(self) -> List[np.ndarray]:
        rng = self._rng if self._rng is not None else np.random
        views = {n: rng.permutation(len(ix)) for n, ix in self.buckets.items()}     # bucket.py:80-83
        cursor = dict.fromkeys(self.buckets, 0)
        out = []
        for n in rng.permutation(self.bucket_idxs):                                   # bucket.py:87
            n = int(n)
            pick = views[n][cursor[n]: cursor[n] + self._draw]
            cursor[n] += len(pick)
            out.append(self.buckets[n][pick])
        return out

    def __iter__(self) -> Iterator[np.ndarray]:
        batches = self._epoch()
        if self.world_size == 1:
            yield from batches
            return
        for g in batches:
            per = len(g) // self.world_size
            if per == 0:
                continue
            yield g[self.rank * per:(self.rank + 1) * per]

    def __len__(self) -> int:
        return self.n_batches


def _pad(rows: List[List[int]], width: int) -> torch.Tensor:
    out = torch.full((len(rows), width), PAD_token, dtype=torch.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = torch.as_tensor(r, dtype=torch.int64)
    return out


def _to(t: torch.Tensor, device) -> torch.Tensor:
    if device is None:
        return t
    if torch.device(device).type == "cuda":
        return t.pin_memory().to(device, non_blocking=True)
    return t.to(device)


def _sort_desc(lengths: List[int]) -> List[int]:
    """Row order of the reference: reversed ascending argsort (preprocessing.py:352-353), ties included."""
    return [int(i) for i in reversed(np.argsort(lengths))]


def data_generator_tl_mtv(data_pairs, data_im, batch_size: int, device=None, world_size: int = 1, rank: int = 0, seed: Optional[int] = None):
    """preprocessing.py:308-384 → (batch_x [B,Lx], batch_y [B,Ly], batch_im [B,I] float32, x_lengths desc, y_lengths)."""
    sampler = BucketBatchSampler([len(p[1]) for p in data_pairs], batch_size, world_size=world_size, rank=rank, seed=seed)
    for bidx in sampler:
        xs = [list(data_pairs[i][0]) for i in bidx]
        ys = [list(data_pairs[i][1]) for i in bidx]
        x_len = [len(x) for x in xs]
        order = _sort_desc(x_len)
        batch_x = _pad([xs[i] for i in order], max(x_len))
        batch_y = _pad([ys[i] for i in order], max(len(y) for y in ys))
        im = torch.as_tensor(np.asarray(data_im)[np.asarray(bidx)[order]]).float() if data_im is not None else None
        yield (_to(batch_x, device), _to(batch_y, device), _to(im, device) if im is not None else None,
               [x_len[i] for i in order], [len(ys[i]) for i in order])


def data_generator_mtv(data_pairs, data_im, batch_size: int, device=None):
    """preprocessing.py:234-306: consecutive evaluation batches → (batch_x, batch_y, batch_im, x_lengths desc, y_lengths,
    order), where ``order[i]`` is the position inside the batch of sorted row i (what translation_reorder undoes)."""
    n = len(data_pairs)
    for start in range(0, n, batch_size):
        chunk = data_pairs[start:start + batch_size]
        xs = [list(p[0]) for p in chunk]
        ys = [list(p[1]) for p in chunk]
        x_len = [len(x) for x in xs]
        order = _sort_desc(x_len)
        batch_x = _pad([xs[i] for i in order], max(x_len))
        batch_y = _pad([ys[i] for i in order], max(len(y) for y in ys))
        im = None
        if data_im is not None:
            im = torch.as_tensor(np.asarray(data_im)[start:start + len(chunk)][order]).float()
        yield (_to(batch_x, device), _to(batch_y, device), _to(im, device) if im is not None else None, [x_len[i] for i in order],
               [len(ys[i]) for i in order], order)


def translation_reorder(translations: List[List[int]], order: List[int]) -> List[List[int]]:
    """Undo the per-batch sort (preprocessing.py:475-486 without the BPE merge): sorted row i goes back to slot order[i]."""
    out: List[Optional[List[int]]] = [None] * len(translations)
    for i, slot in enumerate(order):
        out[slot] = translations[i]
    return out  # type: ignore[return-value]


def translation_reorder_BPE(translations: List[List[int]], order: List[int], id2word) -> List[List[str]]:
    """preprocessing.py:475-486: ids → sub-word strings (unknown ids → '<unk>'), undo the BPE split ('@@ ' joins a piece to
    its successor), re-tokenise on blanks and put sorted row i back into slot order[i]."""
    out: List[Optional[List[str]]] = [None] * len(translations)
    for slot, ids in zip(order, translations):
        text = " ".join(id2word.get(int(t), "<unk>") for t in ids)
        out[slot] = text.replace("@@ ", "").split()
    return out  # type: ignore[return-value]
