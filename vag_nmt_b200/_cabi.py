"""ctypes binding of libvagnmt.so (the C ABI declared in include/vag_nmt.h).

There is no fallback: if the library is missing, cannot be loaded, or the device is not a B200 (sm_100), every
operator raises.  The library is loaded lazily so that the package itself can be imported on a CPU-only box
(host logic, oracle tests, symbol checks).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
from pathlib import Path
from typing import Optional

import torch

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libvagnmt.so"

c_f32p = C.c_void_p
c_i64p = C.c_void_p
c_i32p = C.c_void_p


class EncoderWeights(C.Structure):
    _fields_ = [("E", C.c_int), ("H", C.c_int), ("precision", C.c_int), ("vocab", C.c_int64), ("emb", C.c_void_p),
                ("w_ih", C.c_void_p * 2), ("w_hh", C.c_void_p * 2), ("b_ih", C.c_void_p * 2), ("b_hh", C.c_void_p * 2)]


class VseWeights(C.Structure):
    _fields_ = [("I", C.c_int), ("C", C.c_int), ("S", C.c_int), ("method", C.c_int), ("activation", C.c_int), ("precision", C.c_int),
                ("im_w", C.c_void_p), ("im_b", C.c_void_p), ("txt_w", C.c_void_p), ("txt_b", C.c_void_p),
                ("ctx2ctx_w", C.c_void_p), ("emb2ctx_w", C.c_void_p), ("mlp_w", C.c_void_p)]


class DecoderWeights(C.Structure):
    _fields_ = [("E", C.c_int), ("H", C.c_int), ("C", C.c_int), ("precision", C.c_int), ("V", C.c_int64), ("emb", C.c_void_p),
                ("gru1_w_ih", C.c_void_p), ("gru1_w_hh", C.c_void_p), ("gru1_b_ih", C.c_void_p), ("gru1_b_hh", C.c_void_p),
                ("attn_h_w", C.c_void_p), ("attn_e_w", C.c_void_p), ("attn_v", C.c_void_p), ("c2h_w", C.c_void_p),
                ("gru2_w_ih", C.c_void_p), ("gru2_w_hh", C.c_void_p), ("gru2_b_ih", C.c_void_p), ("gru2_b_hh", C.c_void_p),
                ("w1_w", C.c_void_p), ("w1_b", C.c_void_p), ("w2_w", C.c_void_p), ("w2_b", C.c_void_p),
                ("w3_w", C.c_void_p), ("w3_b", C.c_void_p), ("out_w", C.c_void_p), ("out_b", C.c_void_p),
                ("ini_w", C.c_void_p), ("ini_b", C.c_void_p), ("prepared", C.c_void_p), ("prepared_bytes", C.c_size_t)]


class VseSaved(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("a_im", "iq", "pk", "a_txt")]


class VseGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("im_w", "im_b", "txt_w", "txt_b", "ctx2ctx_w", "emb2ctx_w", "mlp_w")]


class DpComm(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("peers", C.c_void_p * 16)]


class DecoderSeqSaved(C.Structure):
    _fields_ = [("ld_logits", C.c_int64)] + [(n, C.c_void_p) for n in (
        "keys", "e_all", "gi1_all", "gh1_all", "h1_all", "q_all", "alpha_all", "c_all", "x2_all", "gi2_all", "gh2_all", "h2_all",
        "t_all", "logits_all", "lse_all")]


class DecoderGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "emb", "gru1_w_ih", "gru1_w_hh", "gru1_b_ih", "gru1_b_hh", "attn_h_w", "attn_e_w", "attn_v", "c2h_w", "gru2_w_ih",
        "gru2_w_hh", "gru2_b_ih", "gru2_b_hh", "w1_w", "w1_b", "w2_w", "w2_b", "w3_w", "w3_b", "out_w", "out_b")]


I, I64, F, P, SZ = C.c_int, C.c_int64, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/vag_nmt.h declares (tests/test_cabi_symbols.py checks)
SIGNATURES = {
    "vag_last_error": (C.c_char_p, []),
    "vag_abi_version": (I, []),
    "vag_device_supported": (I, []),
    "vag_launch_count": (C.c_longlong, []),
    "vag_linear_f32": (I, [P, I64, P, I64, P, I64, P, I, I, I, I, P]),
    "vag_embed_rows_f32": (I, [P, I64, P, I, P, I, I64, P]),
    "vag_gru_gates_f32": (I, [P, I64, P, I64, P, I64, P, I64, P, I64, I, I, P]),
    "vag_attention_f32": (I, [P, I64, P, P, I64, P, P, P, P, I, I, I, I, I, P]),
    "vag_init_mix_f32": (I, [P, P, P, P, F, I, I, I, P]),
    "vag_l2norm_rows_f32": (I, [P, I64, I, I, P]),
    "vag_log_softmax_f32": (I, [P, P, I, I, P]),
    "vag_encoder_workspace_bytes": (SZ, [I, I, I, I]),
    "vag_encoder_fwd_f32": (I, [P, P, P, I, I, P, P, P, SZ, P]),
    "vag_vse_workspace_bytes": (SZ, [I, I, I, I, I]),
    "vag_vse_pool_fwd_f32": (I, [P, P, P, P, I, I, P, P, P, P, P, SZ, P]),
    "vag_vse_pool_train_fwd_f32": (I, [P, P, P, P, I, I, P, P, P, P, P, P, SZ, P]),
    "vag_vse_pool_bwd_workspace_bytes": (SZ, [I, I, I, I, I]),
    "vag_vse_pool_bwd_f32": (I, [P, P, P, P, I, I, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "vag_rank_loss_workspace_bytes": (SZ, [I, I]),
    "vag_rank_loss_f32": (I, [P, P, I, I, F, I, P, P, P, P, SZ, P]),
    "vag_recall_ranks_workspace_bytes": (SZ, [I, I]),
    "vag_recall_ranks_f32": (I, [P, P, I, I, P, P, SZ, P]),
    "vag_attn_keys_workspace_bytes": (SZ, [I, I, I]),
    "vag_attn_keys_f32": (I, [P, P, I, I, P, P, SZ, P]),
    "vag_tc_elem_bytes": (I, [I]),
    "vag_tc_set_debug": (I, [P]),
    "vag_tc_split_f32": (I, [P, I64, I, I, P, P, I64, I, P]),
    "vag_tc_gemm_f32": (I, [P, I64, P, P, I64, P, P, I64, P, I, I, I, I, P]),
    "vag_tc_gemm_top2_f32": (I, [P, P, P, I64, P, P, I64, P, I, I, I, I, P]),
    "vag_linear_tc_workspace_bytes": (SZ, [I, I, I]),
    "vag_linear_tc_f32": (I, [P, I64, P, I64, P, I64, P, I, I, I, I, P, SZ, P]),
    "vag_decoder_init_workspace_bytes": (SZ, [I, I, I]),
    "vag_decoder_init_f32": (I, [P, P, P, P, F, I, I, P, P, SZ, P]),
    "vag_decoder_step_workspace_bytes": (SZ, [I, I, I, I, I64]),
    "vag_decoder_step_f32": (I, [P, P, P, P, P, P, I, I, I, P, P, I, P, P, SZ, P]),
    "vag_beam_select_f32": (I, [P, I64, P, P, P, P, I, I, I64, I, I, P]),
    "vag_beam_decode_workspace_bytes": (SZ, [I, I, I, I, I, I, I, I64]),
    "vag_beam_decode_f32": (I, [P, P, P, P, P, I, I, I, I, I, P, P, P, P, P, P, P, SZ, P]),
    "vag_beam_decode_steps_f32": (I, [P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P, P, P, P, SZ, P]),
    "vag_beam_finalize_f32": (I, [P, P, P, P, I, I, I, P, P, P, P]),
    "vag_decoder_prepared_bytes": (SZ, [I, I, I, I64]),
    "vag_decoder_prepare_f32": (I, [P, P, SZ, P]),
    "vag_greedy_decode_f32": (I, [P, P, P, P, P, I, I, I, P, P, SZ, P]),
    "vag_nll_rows_f32": (I, [P, I64, P, P, I, I64, P, P, P]),
    "vag_row_argmax_f32": (I, [P, I64, I, I64, P, P]),
    "vag_translation_loss_f32": (I, [P, P, I, I, P, F, P, P]),
    "vag_translation_loss_bwd_f32": (I, [P, P, I, I, F, I, P, P, P]),
    "vag_src_mask_lengths": (I, [P, I, I, P, P, P]),
    "vag_gemm_f32": (I, [P, I64, P, I64, I64, P, I64, I64, I, I, I, F, F, I, P]),
    "vag_gemm_tc_workspace_bytes": (SZ, [I, I, I]),
    "vag_gemm_tc_f32": (I, [P, I64, P, I64, I64, P, I64, I64, I, I, I, F, F, I, P, SZ, P]),
    "vag_gru_gates_bwd_f32": (I, [P, P, P, P, I64, P, P, P, I64, I, I, P]),
    "vag_attention_bwd_f32": (I, [P, I64, P, P, P, P, I64, P, P, I64, P, P, P, P, I, I, I, I, P]),
    "vag_nll_bwd_f32": (I, [P, I64, P, I64, P, P, P, P, I, I64, P]),
    "vag_tanh_bwd_f32": (I, [P, P, P, I64, P]),
    "vag_axpby_f32": (I, [P, P, F, F, I64, P]),
    "vag_colsum_f32": (I, [P, P, I64, I, I, I, P]),
    "vag_embed_bwd_f32": (I, [P, P, I64, P, I, I, I64, P]),
    "vag_l2norm_bwd_f32": (I, [P, P, P, I, I, P]),
    "vag_init_mix_bwd_f32": (I, [P, P, P, P, F, I, I, I, P]),
    "vag_decoder_seq_workspace_bytes": (SZ, [I, I, I, I, I, I, I64]),
    "vag_decoder_seq_fwd_f32": (I, [P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, SZ, P]),
    "vag_decoder_seq_bwd_f32": (I, [P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, P, P, I, P, SZ, P]),
    "vag_mul_f32": (I, [P, P, I64, P]),
    "vag_encoder_train_workspace_bytes": (SZ, [I, I, I, I]),
    "vag_encoder_train_fwd_f32": (I, [P, P, P, P, I, I, P, P, P, P, P, P, P, SZ, P]),
    "vag_encoder_bwd_f32": (I, [P, P, P, I, I, P, P, P, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "vag_sumsq_f32": (I, [P, I64, P, P]),
    "vag_sumsq_multi_f32": (I, [P, I, I64, P, P]),
    "vag_sumsq_multi_partials": (SZ, [I, I64]),
    "vag_sumsq_multi_det_f32": (I, [P, I, I64, P, P, SZ, P]),
    "vag_clip_adam_multi_f32": (I, [P, I, I64, P, F, F, F, F, I, P]),
    "vag_clip_adam_f32": (I, [P, P, P, P, I64, P, F, F, F, F, F, F, I, P]),
    "vag_p2p_alloc": (I, [SZ, P, P]),
    "vag_p2p_free": (I, [P]),
    "vag_p2p_open": (I, [P, P]),
    "vag_p2p_close": (I, [P]),
    "vag_dp_arena_bytes": (SZ, [I64]),
    "vag_dp_allreduce_f32": (I, [P, I64, I, P, P]),
}

_lib: Optional[C.CDLL] = None

# ---- arithmetic mode.  The C ABI is stateless: every call carries its vag_precision (struct field, flag or argument).  What the
# Python layer keeps is WHICH precision the wrappers in ops.py / train_ops.py put into the calls they make on this thread.
PREC_FP32, PREC_BF16 = 0, 1
LIN_BF16 = 16
_PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, PREC_FP32: PREC_FP32, PREC_BF16: PREC_BF16}
_tls = threading.local()


def precision() -> int:
    """The vag_precision the wrappers pass right now on this thread (default VAG_PREC_FP32)."""
    return getattr(_tls, "precision", PREC_FP32)


@contextlib.contextmanager
def precision_scope(p):
    """``with precision_scope("bf16"):`` — calls made inside carry VAG_PREC_BF16.  Python-side and thread-local: autograd's worker
    thread does not inherit it, which is why the backward Functions re-enter the scope their forward ran in (autograd.py)."""
    new = _PRECISIONS[p]
    old = precision()
    _tls.precision = new
    try:
        yield
    finally:
        _tls.precision = old


class VagError(RuntimeError):
    pass


def load_library(path: Optional[os.PathLike] = None) -> C.CDLL:
    """dlopen the library and attach the signatures.  Raises if it is missing — there is no CPU path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise VagError(f"{p} not found: build it with `python -m vag_nmt_b200.build` "
                       "(the CUDA extension is the only implementation; there is no fallback)")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def lib() -> C.CDLL:
    """The library, checked against the current device."""
    l = load_library()
    if not torch.cuda.is_available():
        raise VagError("vag_nmt_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU path")
    return l


def check(status: int) -> None:
    if status != 0:
        msg = load_library().vag_last_error().decode(errors="replace")
        raise VagError(f"libvagnmt status {status}: {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class _NullCtx:
    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def on_device(device):
    """Context manager making `device` current — free when it already is (the common single-GPU-per-process case)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def f32c(t: torch.Tensor, device=None) -> torch.Tensor:
    """contiguous fp32 CUDA view/copy of t"""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return t.detach().to(device=device, dtype=torch.float32).contiguous()
