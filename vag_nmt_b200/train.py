"""Step drivers with the reference's signatures (train.py:19-51), on the B200 kernels.

``train_imagine_beam`` / ``train_nmt`` do zero_grad → forward → backward → clip → Adam exactly like the reference;
the clip and the Adam update are fused in ``ClipAdam.step`` (vag_nmt_b200/optim.py), which also all-reduces the
gradients first when torch.distributed is initialised (one process per GPU, NCCL over NVLink).
"""
from __future__ import annotations

import random
from typing import Dict, Optional

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


_checked_sizes = {}


def _check_equal_local_batch(Bl: int, group) -> None:
    """The all-gather below sizes its output as world·B_local and the averaging gradient all-reduce weights every rank equally:
    both are only right when all ranks hold the same number of sentences (data.BucketBatchSampler's data-parallel mode guarantees
    it).  Checked with one small collective the first time each local size is seen, then cached."""
    import torch.distributed as dist
    key = (id(group), Bl)
    if key in _checked_sizes:
        return
    sizes = [None] * dist.get_world_size(group)
    dist.all_gather_object(sizes, Bl, group=group)
    if len(set(sizes)) != 1:
        raise RuntimeError(f"data-parallel ranks hold different local batch sizes {sizes}: the global-batch ranking loss and the "
                           "averaging gradient all-reduce need equal slices (use data.BucketBatchSampler(world_size=…, seed=…))")
    _checked_sizes[key] = True


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
        _check_equal_local_batch(Bl, group)
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


# ---------------------------------------------------------------------------------------------------- CUDA-graph step
class _SurrogateRankLoss(torch.nn.Module):
    """Stand-in for the ranking loss inside a captured forward: ⟨im, G_im⟩ + ⟨s, G_s⟩ with G filled in later (its gradient with
    respect to the embeddings is G).  Remembers the embeddings it was called with — they are the inputs of the real loss."""

    def __init__(self, g_im, g_s):
        super().__init__()
        self.g_im, self.g_s = g_im, g_s
        self.im = self.s = None

    def forward(self, im, s):
        self.im, self.s = im.detach(), s.detach()
        return (im * self.g_im).sum() + (s * self.g_s).sum()


class GraphedTrainStep:
    """The optimisation step of train.py:36-51 with zero_grad → forward → backward replayed from a CUDA graph.

    A training step at 32 sentences per GPU is ~600 short dependent launches; issued one by one the host cannot keep the
    GPU fed.  Nothing in the launch sequence depends on the DATA of a batch — sentence lengths are read on the device
    (vag_encoder_train_fwd_f32 masks finished rows), the teacher-forcing coin is tossed before the graph is chosen —
    so one captured graph per batch SHAPE (B, Ts, Tt, teacher-forced?) serves every later batch of that shape.  Shapes are
    captured lazily the first time they appear (one eager warm-up pass + the capture ≈ three steps' worth of time) and
    share one memory pool.  Gradient all-reduce (data parallel), clipping and Adam run after the replay through
    ``ClipAdam.step``, exactly as in the eager path.  No collective is ever captured: with the global-batch ranking loss the
    step is two graphs (forward, backward) around the eager all-gather + ranking-loss kernel.

    ``step(src, lengths, tgt, im)`` takes the same arguments as ``train_imagine_beam`` (``im=None`` for the text-only
    model) and returns device scalars (loss, loss_mt, loss_vse) — no host synchronisation.
    """

    def __init__(self, model, optimizer: ClipAdam, criterion_mt, criterion_vse=None, clip: float = CLIP, enabled: Optional[bool] = None,
                 max_graphs: int = 64):
        from collections import OrderedDict
        self.model, self.optimizer = model, optimizer
        self.criterion_mt, self.criterion_vse = criterion_mt, criterion_vse
        self.clip = clip
        # One graph per batch shape, least-recently-used eviction beyond `max_graphs` (Multi30K's bucketed batches produce a few
        # hundred (Ts, Tt) pairs; an evicted shape is simply captured again).  The gradients are NOT part of a graph's private
        # memory: ClipAdam owns one persistent flat gradient buffer and the backward kernels write into it (autograd.set_grad_sink),
        # so every graph shares the same 64 MB and only the step's activations live in the graphs' pool.
        self.max_graphs = int(max_graphs)
        self._graphs: "OrderedDict[tuple, dict]" = OrderedDict()
        self._pool = None
        self._seeds = {}
        self.enabled = True if enabled is None else enabled
        # Data parallel with the global-batch ranking loss: its all-gather must stay OUTSIDE the graphs (ranks capture
        # independently, and a captured collective would have to be matched launch for launch) — the step is then two graphs,
        # forward and backward, with the eager collective + ranking-loss kernel between them (see _capture_split).
        import torch.distributed as dist
        self._world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self._split = self._world > 1 and isinstance(criterion_vse, DistributedPairwiseRankingLoss)
        import os
        # Opt-in.  Measured at 2 GPUs (bf16, 32 sentences per GPU): 2.72 ms per step with the overlap against 2.65 ms without —
        # NCCL's resident CTAs take SM slots from the one-wave recurrent kernels of the encoder BPTT (256 CTAs on 296 slots), which
        # costs more than the 31 MB exchange it hides; three collectives and three graph launches instead of one and two.
        self._overlap = os.environ.get("VAG_DP_OVERLAP", "0") != "0"

    # -- one forward + backward on given (static or caller) tensors
    def _fwd_bwd(self, src, lengths, tgt, im, ratio):
        model = self.model
        self.optimizer.zero_grad()
        # lengths None = graph flavour: the model recomputes them on the device from the padding (vag_src_mask_lengths)
        model._loss_vec = None
        if im is not None:
            loss, loss_mt, loss_vse = model(src, lengths, tgt, im, ratio, criterion_mt=self.criterion_mt, criterion_vse=self.criterion_vse)
        else:
            loss = model(src, lengths, tgt, ratio, criterion=self.criterion_mt)
            loss_mt, loss_vse = loss, None
        vec = getattr(model, "_loss_vec", None)
        model._loss_vec = None
        with model.precision_scope():
            if vec is not None:
                # the three losses come out of one kernel as one [3] tensor: seed the backward on it directly with a constant
                # one-hot instead of loss.backward() (which would add a select-backward, a fill and a stack to every step)
                out, which = vec
                torch.autograd.backward([out], [self._seed(out.device, which)])
                return out.detach()          # text-only: the kernel already wrote (loss_mt, loss_mt, 0)
            loss.backward()
        vse = loss_vse if torch.is_tensor(loss_vse) else torch.zeros((), device=loss.device)
        return torch.stack([loss.detach().reshape(()), loss_mt.detach().reshape(()), vse.detach().reshape(())])

    def _seed(self, dev, which):
        key = (str(dev), which)
        if key not in self._seeds:      # created once per device, outside any capture (the first call is the warm-up)
            e = torch.zeros(3, dtype=torch.float32, device=dev)
            e[which] = 1.0
            self._seeds[key] = e
        return self._seeds[key]

    @staticmethod
    def _device_lengths(src):
        return (src != 0).sum(1, dtype=torch.int32)

    @staticmethod
    def _check_lengths(lengths, width):
        ls = [int(x) for x in lengths]
        if width != max(ls):
            raise ValueError("the padded width must equal the longest sentence (pad_packed_sequence, Encoder.py:60)")
        if any(ls[i] < ls[i + 1] for i in range(len(ls) - 1)):
            raise RuntimeError("`lengths` array must be sorted in decreasing order (pack_padded_sequence, Encoder.py:55)")
        return ls

    def _capture(self, key, src, ls, tgt, im, ratio):
        from . import ops
        dev = src.device
        st = {"src": src.clone(), "tgt": tgt.clone(), "im": im.clone() if im is not None else None, "len": None}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):     # warm-up outside the capture: sizes the workspaces, sets kernel attributes
            # Ranks capture independently (a shape is new on one rank and known on another), so the warm-up must not execute a
            # collective the peers would not match: it uses the LOCAL ranking loss — its results are discarded anyway.
            crit = self.criterion_vse
            if isinstance(crit, DistributedPairwiseRankingLoss):
                self.criterion_vse = PairwiseRankingLoss(crit.margin)
            try:
                self._fwd_bwd(st["src"], st["len"], st["tgt"], st["im"], ratio)
            finally:
                self.criterion_vse = crit
        torch.cuda.current_stream(dev).wait_stream(side)
        self.optimizer.zero_grad()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self._pool):
            st["out"] = self._fwd_bwd(st["src"], st["len"], st["tgt"], st["im"], ratio)
        st["graph"] = g
        st["grads"] = [(p, p.grad) for p in self.optimizer._all_params() if p.grad is not None]
        st["keepalive"] = list(ops._workspaces.values())     # the graph holds raw pointers into these scratch buffers
        self._graphs[key] = st
        return st

    def _capture_split(self, key, src, ls, tgt, im, ratio):
        """Data-parallel flavour: graph A = zero_grad + forward with a SURROGATE ranking term ⟨im_emb, G_im⟩ + ⟨txt_emb, G_s⟩,
        graph B = backward.  Between the replays the real global ranking loss runs eagerly (all-gather of the embeddings,
        vag_rank_loss_f32 with gradients) and writes its local gradient rows into G_im / G_s — the surrogate's gradient with
        respect to the embeddings is exactly G, so graph B back-propagates the true loss."""
        from . import ops
        model, dev = self.model, src.device
        st = {"src": src.clone(), "tgt": tgt.clone(), "im": im.clone(), "len": None, "split": True}
        S = model.shared_embedding_size
        st["G_im"] = torch.zeros(src.shape[0], S, dtype=torch.float32, device=dev)
        st["G_s"] = torch.zeros_like(st["G_im"])
        sur = _SurrogateRankLoss(st["G_im"], st["G_s"])
        crit = self.criterion_vse
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):       # warm-up without collectives (local ranking loss; results are discarded)
            self.criterion_vse = PairwiseRankingLoss(crit.margin)
            try:
                self._fwd_bwd(st["src"], st["len"], st["tgt"], st["im"], ratio)
            finally:
                self.criterion_vse = crit
        torch.cuda.current_stream(dev).wait_stream(side)
        self.optimizer.zero_grad()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        ga, gb, gb2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), None
        # Overlap (VAG_DP_OVERLAP=1 switches it on): the backward is cut behind the decoder; the decoder's gradients — the
        # longest run of its slices in the flat buffer, 31 of 64 MB — are all-reduced while graph B2 runs the rest (ranking /
        # initialiser / encoder BPTT), the remainder afterwards.
        span = self.optimizer.flat_span(list(model.decoder.parameters())) if self._overlap else None
        with torch.cuda.graph(ga, pool=self._pool):
            self.optimizer.zero_grad()
            model._bwd_split = span is not None
            try:
                loss, loss_mt, _ = model(st["src"], None, st["tgt"], st["im"], ratio,
                                         criterion_mt=self.criterion_mt, criterion_vse=sur)
            finally:
                model._bwd_split = False
            st["mt"] = loss_mt.detach().reshape(())
            st["im_emb"], st["txt_emb"] = sur.im, sur.s
            vec, model._loss_vec = model._loss_vec, None
            cut, model._boundary = getattr(model, "_boundary", None), None
        with torch.cuda.graph(gb, pool=self._pool):
            with model.precision_scope():
                torch.autograd.backward([vec[0]], [self._seed(dev, vec[1])])
        if span is not None and cut is not None:
            gb2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb2, pool=self._pool):
                with model.precision_scope():
                    torch.autograd.backward(cut[0], [leaf.grad for leaf in cut[1]])
            st["span"] = span
        del loss, vec, cut
        st["graph"], st["graph_b"], st["graph_b2"] = ga, gb, gb2
        st["grads"] = [(p, p.grad) for p in self.optimizer._all_params() if p.grad is not None]
        st["keepalive"] = list(ops._workspaces.values())
        self._graphs[key] = st
        return st

    def _global_rank_loss(self, st):
        """Eager middle of the split step → loss_vse over the global batch; fills G_im / G_s (rows of this rank, × world so that
        the averaging gradient all-reduce yields the single-process sum, as in _GlobalRankLossFn)."""
        import torch.distributed as dist
        from . import ops
        crit = self.criterion_vse
        world, rank = self._world, dist.get_rank(crit.group)
        im, s = st["im_emb"], st["txt_emb"]
        Bl = im.shape[0]
        _check_equal_local_batch(Bl, crit.group)
        im_all = torch.empty(world * Bl, im.shape[1], dtype=im.dtype, device=im.device)
        s_all = torch.empty_like(im_all)
        dist.all_gather_into_tensor(im_all, im, group=crit.group)
        dist.all_gather_into_tensor(s_all, s, group=crit.group)
        loss, g_im, g_s = ops.rank_loss(im_all, s_all, float(crit.margin), crit.one_direction, want_grad=True)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        # .data: the surrogate saved G for its backward; filling it must not trip autograd's version check at capture time
        st["G_im"].data.copy_(g_im[sl]).mul_(float(world))
        st["G_s"].data.copy_(g_s[sl]).mul_(float(world))
        return loss

    def step(self, src, lengths, tgt, im=None, teacher_force_ratio: float = 1.0):
        model = self.model
        model.train()
        is_teacher = random.random() < teacher_force_ratio                        # V11:136 — decided before the graph is chosen
        ratio = 1.0 if is_teacher else 0.0
        dev = model._device()
        to_dev = lambda t: t.to(dev, non_blocking=True) if t is not None else None
        if not self.enabled or getattr(model, "_dropout_masks", None):
            out = self._fwd_bwd(to_dev(src), lengths, to_dev(tgt), to_dev(im), ratio)
        else:
            ls = self._check_lengths(lengths, src.shape[1])
            key = (tuple(src.shape), tuple(tgt.shape), is_teacher, im is not None, getattr(model, "precision", "fp32"))
            st = self._graphs.get(key)
            if st is None:
                while len(self._graphs) >= self.max_graphs:
                    self._graphs.popitem(last=False)          # least recently used shape; its memory returns to the shared pool
                split = self._split and im is not None
                st = (self._capture_split if split else self._capture)(key, to_dev(src), ls, to_dev(tgt), to_dev(im), ratio)
            else:
                self._graphs.move_to_end(key)
            # (pinned) host or device batch → the graph's static buffers; the lengths are recomputed from src inside the graph
            st["src"].copy_(src, non_blocking=True)
            st["tgt"].copy_(tgt, non_blocking=True)
            if im is not None:
                st["im"].copy_(im, non_blocking=True)
            st["graph"].replay()
            pre_reduced = False
            if st.get("split"):
                vse = self._global_rank_loss(st).reshape(())
                st["graph_b"].replay()
                if st.get("graph_b2") is not None:
                    opt = self.optimizer
                    lo, hi = st["span"]
                    works = opt.allreduce_range_async(lo, hi)        # decoder bucket, concurrent with graph B2
                    st["graph_b2"].replay()
                    n = opt._flat.numel()
                    for a, b in ((0, lo), (hi, n)):
                        if b > a:
                            works += opt.allreduce_range_async(a, b)
                    for wk in works:
                        wk.wait()
                    pre_reduced = True
                w = float(model.loss_w)
                out = torch.stack([w * st["mt"] + (1.0 - w) * vse, st["mt"], vse])
            else:
                out = st["out"].clone()    # the pool is shared between shapes: hand out a private copy of the three scalars
            for p, g in st["grads"]:       # the gradients live at fixed addresses inside the graph's pool
                p.grad = g
            if pre_reduced:
                self.optimizer.step(clip=self.clip, pre_reduced=True)
                self.last_out = (out[0], out[1], out[2])
                return out[0], out[1], out[2]
        self.optimizer.step(clip=self.clip)
        self.last_out = (out[0], out[1], out[2])
        return out[0], out[1], out[2]
