"""Tensor-level wrappers over the training-side C ABI (backward kernels, generic contraction, clip+Adam)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi
from ._cabi import check, on_device, ptr, stream_ptr


def _f(n) -> float:
    return float(n)


def gemm(A: torch.Tensor, B: torch.Tensor, trans_a: bool = False, trans_b: bool = False, out: Optional[torch.Tensor] = None,
         alpha: float = 1.0, beta: float = 0.0) -> torch.Tensor:
    """out = alpha·op(A)·op(B) + beta·out with op = transpose when the flag is set; A, B 2-D fp32 (any strides)."""
    lib = _cabi.lib()
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    K2, N = (B.shape[1], B.shape[0]) if trans_b else B.shape
    assert K == K2, (A.shape, B.shape, trans_a, trans_b)
    sam, sak = (A.stride(1), A.stride(0)) if trans_a else (A.stride(0), A.stride(1))
    sbk, sbn = (B.stride(1), B.stride(0)) if trans_b else (B.stride(0), B.stride(1))
    if out is None:
        assert beta == 0.0
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    with on_device(A.device):
        if M >= 64 and N >= 64 and K >= 32:     # large enough for the tensor-core route: hand it a workspace for the operand planes
            from .ops import workspace
            ws = workspace(lib.vag_gemm_tc_workspace_bytes(M, N, K), A.device, slot="gemm")
            check(lib.vag_gemm_tc_f32(out.data_ptr(), out.stride(0), A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, M, N, K,
                                      _f(alpha), _f(beta), _cabi.precision(), ws.data_ptr(), ws.numel(), stream_ptr()))
        else:
            check(lib.vag_gemm_f32(out.data_ptr(), out.stride(0), A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, M, N, K,
                                   _f(alpha), _f(beta), _cabi.precision(), stream_ptr()))
    return out


def gru_gates_bwd(dh: torch.Tensor, gi: torch.Tensor, gh: torch.Tensor, h_prev: torch.Tensor):
    """→ dgi [rows,3H], dgh [rows,3H], dh_prev_part [rows,H] (= dh·z)"""
    lib = _cabi.lib()
    rows, H = h_prev.shape
    assert gi.is_contiguous() and gh.is_contiguous() and dh.stride(1) == 1 and h_prev.stride(1) == 1
    dgi = torch.empty(rows, 3 * H, dtype=torch.float32, device=dh.device)
    dgh = torch.empty_like(dgi)
    dhp = torch.empty(rows, H, dtype=torch.float32, device=dh.device)
    with on_device(dh.device):
        check(lib.vag_gru_gates_bwd_f32(dgi.data_ptr(), dgh.data_ptr(), dhp.data_ptr(), dh.data_ptr(), dh.stride(0), gi.data_ptr(),
                                        gh.data_ptr(), h_prev.data_ptr(), h_prev.stride(0), rows, H, stream_ptr()))
    return dgi, dgh, dhp


def attention_bwd(dc, alpha, q, keys, ctx, v, mask, dkeys, dctx, dv, mode: int) -> torch.Tensor:
    """Accumulates into dkeys / dctx / dv (may be None); → dq [B, C]."""
    lib = _cabi.lib()
    B, T, C = ctx.shape
    assert dc.stride(1) == 1 and q.stride(1) == 1 and alpha.is_contiguous() and keys.is_contiguous() and ctx.is_contiguous()
    dq = torch.empty(B, C, dtype=torch.float32, device=ctx.device)
    with on_device(ctx.device):
        check(lib.vag_attention_bwd_f32(dq.data_ptr(), C, dkeys.data_ptr(), ptr(dctx), ptr(dv), dc.data_ptr(), dc.stride(0),
                                        alpha.data_ptr(), q.data_ptr(), q.stride(0), keys.data_ptr(), ctx.data_ptr(), ptr(v),
                                        ptr(mask), B, T, C, mode, stream_ptr()))
    return dq


def nll_bwd(logits, lse, tgt, weight, grad_rows) -> torch.Tensor:
    lib = _cabi.lib()
    rows, V = logits.shape
    d = torch.empty(rows, V, dtype=torch.float32, device=logits.device)
    with on_device(logits.device):
        check(lib.vag_nll_bwd_f32(d.data_ptr(), V, logits.data_ptr(), logits.stride(0), lse.data_ptr(), tgt.data_ptr(), ptr(weight),
                                  grad_rows.data_ptr(), rows, V, stream_ptr()))
    return d


def tanh_bwd(dy: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    lib = _cabi.lib()
    dy, y = dy.contiguous(), y.contiguous()
    dx = torch.empty_like(y)
    with on_device(y.device):
        check(lib.vag_tanh_bwd_f32(dx.data_ptr(), dy.data_ptr(), y.data_ptr(), y.numel(), stream_ptr()))
    return dx


def axpby_(y: torch.Tensor, x: torch.Tensor, a: float = 1.0, b: float = 1.0) -> torch.Tensor:
    """y = a·x + b·y in place (contiguous tensors of equal size)."""
    lib = _cabi.lib()
    assert y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
    with on_device(y.device):
        check(lib.vag_axpby_f32(y.data_ptr(), x.data_ptr(), _f(a), _f(b), y.numel(), stream_ptr()))
    return y


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    lib = _cabi.lib()
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=x.device)
        accumulate = False
    with on_device(x.device):
        check(lib.vag_colsum_f32(out.data_ptr(), x.data_ptr(), x.stride(0), rows, cols, 1 if accumulate else 0, stream_ptr()))
    return out


def embed_bwd_(table_grad: torch.Tensor, g: torch.Tensor, ids: torch.Tensor) -> None:
    lib = _cabi.lib()
    assert table_grad.is_contiguous() and g.stride(1) == 1 and ids.is_contiguous() and ids.dtype == torch.int64
    with on_device(g.device):
        check(lib.vag_embed_bwd_f32(table_grad.data_ptr(), g.data_ptr(), g.stride(0), ids.data_ptr(), ids.numel(), g.shape[1],
                                    table_grad.shape[0], stream_ptr()))


def l2norm_bwd(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    lib = _cabi.lib()
    dy, x = dy.contiguous(), x.contiguous()
    dx = torch.empty_like(x)
    with on_device(x.device):
        check(lib.vag_l2norm_bwd_f32(dx.data_ptr(), dy.data_ptr(), x.data_ptr(), x.shape[0], x.shape[1], stream_ptr()))
    return dx


def init_mix_bwd(dz: torch.Tensor, mask: torch.Tensor, split: float, dctx: torch.Tensor, want_ctx_vec: bool):
    lib = _cabi.lib()
    B, T, C = dctx.shape
    dz = dz.contiguous()
    dvec = torch.empty(B, C, dtype=torch.float32, device=dz.device) if want_ctx_vec else None
    with on_device(dz.device):
        check(lib.vag_init_mix_bwd_f32(ptr(dvec), dctx.data_ptr(), dz.data_ptr(), mask.data_ptr(), _f(split), B, T, C, stream_ptr()))
    return dvec


def sumsq_(accum: torch.Tensor, g: torch.Tensor) -> None:
    lib = _cabi.lib()
    assert g.is_contiguous()
    with on_device(g.device):
        check(lib.vag_sumsq_f32(g.data_ptr(), g.numel(), accum.data_ptr(), stream_ptr()))


def clip_adam_(param, grad, exp_avg, exp_avg_sq, sumsq, clip, lr, beta1, beta2, eps, weight_decay, step) -> None:
    lib = _cabi.lib()
    assert param.is_contiguous() and grad.is_contiguous()
    with on_device(param.device):
        check(lib.vag_clip_adam_f32(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
                                    sumsq.data_ptr(), _f(clip), _f(lr), _f(beta1), _f(beta2), _f(eps), _f(weight_decay), int(step),
                                    stream_ptr()))


def optim_table(entries) -> torch.Tensor:
    """Descriptor table for the multi-tensor optimiser kernels: entries = [(param, grad, exp_avg, exp_avg_sq, wd, lr), …]
    → uint8 host tensor laid out as ``vag_optim_tensor[n]`` (include/vag_nmt.h: 4 pointers, int64 n, 2 floats = 48 B)."""
    import numpy as np
    dt = np.dtype([("param", "<u8"), ("grad", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("wd", "<f4"), ("lr", "<f4")])
    assert dt.itemsize == 48
    arr = np.zeros(len(entries), dtype=dt)
    for i, (p, g, m, v, wd, lr) in enumerate(entries):
        arr[i] = (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), wd, lr)
    return torch.from_numpy(arr.view(np.uint8).reshape(-1))


def sumsq_multi_(accum: torch.Tensor, table_dev: torch.Tensor, n_tensors: int, max_n: int) -> None:
    lib = _cabi.lib()
    with on_device(accum.device):
        check(lib.vag_sumsq_multi_f32(table_dev.data_ptr(), int(n_tensors), int(max_n), accum.data_ptr(), stream_ptr()))


def sumsq_multi_det_(out: torch.Tensor, table_dev: torch.Tensor, n_tensors: int, max_n: int, partials: torch.Tensor) -> None:
    """out[0] = Σ‖g‖² over all tensors, deterministic (block partials added in index order, no atomics)."""
    lib = _cabi.lib()
    with on_device(out.device):
        check(lib.vag_sumsq_multi_det_f32(table_dev.data_ptr(), int(n_tensors), int(max_n), out.data_ptr(), partials.data_ptr(),
                                          partials.numel(), stream_ptr()))


def clip_adam_multi_(table_dev: torch.Tensor, n_tensors: int, max_n: int, sumsq, clip, beta1, beta2, eps, step) -> None:
    lib = _cabi.lib()
    with on_device(sumsq.device):
        check(lib.vag_clip_adam_multi_f32(table_dev.data_ptr(), int(n_tensors), int(max_n), sumsq.data_ptr(), _f(clip), _f(beta1),
                                          _f(beta2), _f(eps), int(step), stream_ptr()))
