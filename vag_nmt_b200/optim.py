"""Optimiser side of the training step: gradient all-reduce (data parallel), global-norm clipping and Adam.

Mirrors ``torch.nn.utils.clip_grad_norm_(model.parameters(), clip)`` followed by ``optim.Adam(param_groups).step()``
with the parameter groups of nmt_multimodal_beam_DE.py:303-332 (weight decay — added to the gradient, not
decoupled — on every parameter whose name lacks 'bias').  All arithmetic runs in TWO launches for all parameter tensors
(vag_sumsq_multi_det_f32, vag_clip_adam_multi_f32 over a descriptor table); the clip coefficient is read on the device, so a
step never synchronises.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch

from . import train_ops as T


def named_param_groups(model: torch.nn.Module, weight_decay: float = 1e-5) -> List[dict]:
    """The two groups the reference builds (nmt_multimodal_beam_DE.py:303-312)."""
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    return [{"params": [p for n, p in named if "bias" not in n], "weight_decay": weight_decay},
            {"params": [p for n, p in named if "bias" in n], "weight_decay": 0.0}]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, flat: Optional[torch.Tensor] = None) -> None:
    """Average the gradients over the data-parallel ranks with ONE NCCL all-reduce.  `flat`: the persistent flat buffer the
    gradients are views of (ClipAdam keeps one) — reduced in place, no gather / scatter passes.  Without it the gradients are
    concatenated into a temporary bucket and copied back."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)

    def reduce_(t):
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)   # NCCL averages inside the collective: no extra pass over 64 MB
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.mul_(1.0 / world)

    if flat is not None:
        reduce_(flat)
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    bucket = torch.cat([g.reshape(-1) for g in grads])             # one launch
    reduce_(bucket)
    views, off = [], 0
    for g in grads:
        views.append(bucket[off:off + g.numel()].view_as(g))
        off += g.numel()
    torch._foreach_copy_(grads, views)                            # multi-tensor copy back: a couple of launches, not one per tensor


class _RawCuda:
    """Zero-copy view of a raw device allocation for torch.as_tensor (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None}


class PeerGradientExchange:
    """The data-parallel gradient all-reduce + Σ‖g‖² over NVLink peer memory (csrc/p2p.cu; one process per GPU, one node).

    Owns a peer-visible arena whose first ``n_floats`` floats are THE flat gradient buffer (``.flat``); the ranks exchange
    CUDA IPC handles once over torch.distributed (host objects) and map each other's arenas.  ``allreduce(sumsq)`` then runs
    barrier · reduce-scatter + norm partials · barrier · all-gather + norm · barrier as five kernels on the current stream: every
    rank ends with bit-identical averaged gradients and the same Σ‖g‖² — NCCL is not involved.  Must be called by every rank,
    once per optimisation step."""

    def __init__(self, n_floats: int, device: torch.device, group=None):
        import ctypes as C
        import socket
        import torch.distributed as dist
        from . import _cabi
        lib = _cabi.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) != 1 or self.world > 16:
            raise RuntimeError("peer-memory gradient exchange needs all ranks on one node (≤ 16 GPUs)")
        self.n = (int(n_floats) + 3) // 4 * 4
        nbytes = lib.vag_dp_arena_bytes(self.n)
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        with _cabi.on_device(device):
            _cabi.check(lib.vag_p2p_alloc(nbytes, C.byref(ptr), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.comm = _cabi.DpComm()
            self.comm.world, self.comm.rank = self.world, self.rank
            for p_, h in enumerate(handles):
                if p_ == self.rank:
                    self.comm.peers[p_] = ptr.value
                else:
                    q = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(h)
                    _cabi.check(lib.vag_p2p_open(buf, C.byref(q)))
                    self.comm.peers[p_] = q.value
        self._raw = _RawCuda(ptr.value, self.n * 4)
        self.flat = torch.as_tensor(self._raw, device=device).view(torch.float32)
        self.step_no = 0
        self.group = group
        dist.barrier(group=group)          # every rank has mapped every arena before anybody launches a kernel on it

    def allreduce(self, sumsq_out: torch.Tensor) -> None:
        import ctypes as C
        from . import _cabi
        lib = _cabi.lib()
        with _cabi.on_device(self.flat.device):
            _cabi.check(lib.vag_dp_allreduce_f32(C.byref(self.comm), self.n, self.step_no, sumsq_out.data_ptr(), _cabi.stream_ptr()))
        self.step_no += 1


class ClipAdam:
    """clip_grad_norm_ + Adam fused; ``param_groups`` entries carry 'params', 'weight_decay' and optionally 'lr'."""

    def __init__(self, param_groups, lr: float = 4e-4, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 clip: float = 1.0, group=None):
        if isinstance(param_groups, torch.nn.Module):
            param_groups = named_param_groups(param_groups)
        self.param_groups = [dict(g) for g in param_groups]
        for g in self.param_groups:
            g.setdefault("lr", lr)
            g.setdefault("weight_decay", 0.0)
            g["params"] = list(g["params"])
        self.betas, self.eps, self.clip, self.group = betas, eps, clip, group
        self.state = {}
        self.step_count = 0
        self._sumsq: Optional[torch.Tensor] = None
        self._flat: Optional[torch.Tensor] = None     # ONE persistent gradient buffer; every p.grad is a view of it
        self._offsets = {}                            # param data_ptr → (element offset, numel)

    # -- persistent flat gradient buffer ------------------------------------------------------------------------------------
    def _ensure_flat(self) -> bool:
        params = self._all_params()
        if not params or not params[0].is_cuda:
            return False
        key = tuple(p.data_ptr() for p in params)
        if self._flat is not None and self._flat_key == key:
            return True
        off, offsets = 0, {}
        for p in params:
            if p.data_ptr() in offsets:               # a tensor registered twice (tied weights): one slice
                continue
            offsets[p.data_ptr()] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4           # 16-byte aligned slices (vector loads in the optimiser kernels)
        self._peer = None
        import os
        import torch.distributed as dist
        if (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1 and dist.get_backend(self.group) == "nccl"
                and os.environ.get("VAG_DP_P2P", "0") != "0"):
            # Opt-in (VAG_DP_P2P=1): measured 204 µs for the 63.8 MB exchange + norm at 2 GPUs against 147 µs (+ ≈ 20 µs norm pass)
            # for NCCL — pull-style peer loads reach ≈ 45 % of the NVLink rate — so NCCL's all-reduce stays the default.
            try:    # the flat buffer lives in a peer-visible arena: the all-reduce becomes peer loads over NVLink (csrc/p2p.cu)
                self._peer = PeerGradientExchange(off, params[0].device, self.group)
            except Exception as exc:     # not one node / IPC unavailable: every rank takes the same branch (same exception class)
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable ({exc}); using the NCCL all-reduce")
                self._peer = None
        self._flat = self._peer.flat[:off] if self._peer is not None else torch.zeros(off, dtype=torch.float32, device=params[0].device)
        self._offsets, self._flat_key = offsets, key
        self._table_key = None
        return True

    def _grad_view(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        """A FRESH view of p's slice (autograd adopts an incoming gradient as p.grad without a copy only when nothing else holds
        the tensor object)."""
        hit = self._offsets.get(p.data_ptr())
        if hit is None or self._flat is None:
            return None
        return self._flat[hit[0]:hit[0] + hit[1]].view(p.shape)

    def zero_grad(self) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None
        if self._ensure_flat():
            from . import autograd
            autograd.set_grad_sink(self._grad_view)   # the backward kernels of the next pass write straight into the flat buffer

    def _all_params(self):
        return [p for g in self.param_groups for p in g["params"]]

    # -- bucketed exchange: the caller reduces slices of the flat buffer as soon as their gradients are final -------------------
    def flat_span(self, params) -> Optional[Tuple[int, int]]:
        """→ element range [lo, hi) of the LONGEST run of consecutive flat-buffer slices that all belong to `params` (None without
        a flat buffer, or with the peer-memory exchange, which reduces the whole arena in one kernel)."""
        if not self._ensure_flat() or getattr(self, "_peer", None) is not None:
            return None
        want = {p.data_ptr() for p in params}
        best, run = None, None
        for ptr, (off, n) in sorted(self._offsets.items(), key=lambda kv: kv[1][0]):
            end = off + (n + 3) // 4 * 4
            if ptr in want:
                run = (run[0], end) if run is not None else (off, end)
                if best is None or run[1] - run[0] > best[1] - best[0]:
                    best = run
            else:
                run = None
        return best

    def allreduce_range_async(self, lo: int, hi: int):
        """Average flat[lo:hi] over the data-parallel ranks; → the c10d work handle (the collective runs on the process group's own
        stream, ordered after what the current stream has enqueued so far: kernels launched afterwards overlap with it)."""
        import torch.distributed as dist
        view = self._flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            return [dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=True)]
        w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        w.wait()
        view.mul_(1.0 / dist.get_world_size(self.group))
        return []

    @torch.no_grad()
    def step(self, clip: Optional[float] = None, pre_reduced: bool = False) -> torch.Tensor:
        """→ device scalar Σ‖g‖² BEFORE clipping (sqrt of it is what clip_grad_norm_ returns).  pre_reduced: the caller has already
        averaged the whole flat buffer over the ranks (allreduce_range_async, overlapped with the backward)."""
        clip = self.clip if clip is None else clip
        params = [p for p in self._all_params() if p.grad is not None]
        flat = None
        if self._ensure_flat():
            # gradients that did not land in the flat buffer (a backward outside autograd.py's Functions, an accumulated or cloned
            # gradient) are moved there now; the common case finds every p.grad already in place
            base, in_place = self._flat.data_ptr(), True
            for p in params:
                off, n = self._offsets[p.data_ptr()]
                if p.grad.data_ptr() != base + 4 * off or not p.grad.is_contiguous():
                    view = self._grad_view(p)
                    view.copy_(p.grad)
                    p.grad = view
                    in_place = False
            self.grads_in_place = in_place            # diagnostic (tests / bench)
            if len(params) == len({p.data_ptr() for p in self._all_params()}):
                flat = self._flat                     # every parameter has a gradient: reduce the whole buffer in place
        dev = params[0].device
        if self._sumsq is None or self._sumsq.device != dev:
            self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        peer_done = False
        if pre_reduced:
            if flat is None or not self.grads_in_place:
                raise RuntimeError("pre_reduced=True needs every gradient in place in the flat buffer")
        elif flat is not None and getattr(self, "_peer", None) is not None:
            self._peer.allreduce(self._sumsq)          # averaged gradients in place + Σ‖g‖², bit-identical on every rank
            peer_done = True
        else:
            allreduce_gradients(params, self.group, flat)  # data parallel: clip must see the GLOBAL gradient
        self.step_count += 1
        b1, b2 = self.betas
        # ONE descriptor table (48 B per tensor) and two launches for all tensors: Σ‖g‖², then clip + Adam
        entries, max_n = [], 0
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = (torch.zeros_like(p), torch.zeros_like(p))
                entries.append((p.data, p.grad, st[0], st[1], float(g["weight_decay"]), float(g["lr"])))
                max_n = max(max_n, p.numel())
        # the descriptor table only changes when a tensor moves (graph-replayed steps keep every address fixed): cache it
        tkey = tuple((e[0].data_ptr(), e[1].data_ptr(), e[4], e[5]) for e in entries)
        if getattr(self, "_table_key", None) == tkey:
            table = self._table
        else:
            table = T.optim_table(entries).to(dev, non_blocking=True)
            self._table_key = tkey
        n_part = T._cabi.lib().vag_sumsq_multi_partials(len(entries), max_n)
        if getattr(self, "_partials", None) is None or self._partials.numel() < n_part or self._partials.device != dev:
            self._partials = torch.empty(max(int(n_part), 1), dtype=torch.float32, device=dev)
        if not peer_done:
            T.sumsq_multi_det_(self._sumsq, table, len(entries), max_n, self._partials)   # deterministic: replicas stay bit-identical
        T.clip_adam_multi_(table, len(entries), max_n, self._sumsq, clip if clip is not None else float("inf"), b1, b2, self.eps,
                           self.step_count)
        self._table = table   # keep the descriptors alive until the kernels have run
        from . import ops
        ops.invalidate_prepared()   # the update went through raw pointers: cached decode invariants are stale
        return self._sumsq
