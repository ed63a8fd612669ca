"""Tensor-level wrappers over the C ABI (include/vag_nmt.h).

PyTorch supplies device memory and the current stream; every function here hands raw device pointers to
libvagnmt.so.  Nothing in this module computes on the CPU or through torch operators.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import DecoderWeights, EncoderWeights, VseWeights, check, on_device, ptr, stream_ptr

ATTN_MLP, ATTN_DOT = 0, 1
LIN_TANH, LIN_ACCUMULATE, LIN_FORCE_SIMT, LIN_FORCE_TC, LIN_BF16 = 1, 2, 4, 8, 16


def _pflag() -> int:
    """VAG_LIN_BF16 when the current precision scope is bf16"""
    return LIN_BF16 if _cabi.precision() == _cabi.PREC_BF16 else 0

_workspaces = {}


_lane = ""     # prefix of every workspace slot while a decode lane is active (see lane())


class lane:
    """Context manager: the calls inside use PRIVATE scratch buffers (workspace slots prefixed with `name`), so that several
    decodes can be in flight on different CUDA streams at once (models._decode_lanes, translate.decode_corpus_pipelined)."""

    def __init__(self, name: str):
        self.name, self.prev = name, ""

    def __enter__(self):
        global _lane
        self.prev, _lane = _lane, self.name
        return self

    def __exit__(self, *exc):
        global _lane
        _lane = self.prev
        return False


def workspace(nbytes: int, device: torch.device, slot: str = "main") -> torch.Tensor:
    """Grow-only scratch buffer per (device, lane + slot); the C side never allocates."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _lane + slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _dev(t: torch.Tensor) -> torch.device:
    if not t.is_cuda:
        raise _cabi.VagError("expected a CUDA tensor; vag_nmt_b200 has no CPU path")
    return t.device


def _chk_f32(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.float32 or not t.is_cuda):
            raise _cabi.VagError(f"expected fp32 CUDA tensors, got {t.dtype} on {t.device}")


# ------------------------------------------------------------------ primitives
def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, flags: int = 0,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(x Wᵀ + b [+ y]); x [..., in], w [out, in].  x/w/out may be row-strided 2-D views."""
    _chk_f32(x, w, bias, out)
    lib = _cabi.lib()
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if w.stride(-1) != 1:
        w = w.contiguous()
    rows, in_dim = x2.shape
    out_dim = w.shape[0]
    assert w.shape[1] == in_dim, (w.shape, x.shape)
    if out is None:
        assert not (flags & LIN_ACCUMULATE)
        out = torch.empty(rows, out_dim, dtype=torch.float32, device=x.device)
    out2 = out.reshape(-1, out_dim) if out.dim() != 2 else out
    assert out2.stride(-1) == 1 and out2.shape[0] == rows
    with on_device(x.device):
        check(lib.vag_linear_f32(out2.data_ptr(), out2.stride(0) if rows > 1 else out_dim, x2.data_ptr(),
                                 x2.stride(0) if rows > 1 else in_dim, w.data_ptr(), w.stride(0), ptr(bias), rows, in_dim,
                                 out_dim, flags | _pflag(), stream_ptr()))
    return out2.reshape(*lead, out_dim)


def linear_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, flags: int = 0,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same contract as ``linear`` on the tcgen05 split-precision kernel (raises VagError for ineligible shapes)."""
    _chk_f32(x, w, bias, out)
    lib = _cabi.lib()
    assert x.dim() == 2 and x.stride(1) == 1 and w.stride(1) == 1
    rows, in_dim = x.shape
    out_dim = w.shape[0]
    if out is None:
        assert not (flags & LIN_ACCUMULATE)
        out = torch.empty(rows, out_dim, dtype=torch.float32, device=x.device)
    ws = workspace(lib.vag_linear_tc_workspace_bytes(rows, in_dim, out_dim), x.device, slot="gemm")
    with on_device(x.device):
        check(lib.vag_linear_tc_f32(out.data_ptr(), out.stride(0), x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0),
                                    ptr(bias), rows, in_dim, out_dim, flags | _pflag(), ws.data_ptr(), ws.numel(), stream_ptr()))
    return out


def tc_split(x: torch.Tensor):
    """x [rows, K] fp32 → (hi, lo) split planes (uint8 storage, [rows, K] elements of vag_tc_elem_bytes() each)."""
    _chk_f32(x)
    lib = _cabi.lib()
    assert x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    esz = lib.vag_tc_elem_bytes(_cabi.precision())
    hi = torch.empty(rows * K * esz, dtype=torch.uint8, device=x.device)
    lo = torch.empty_like(hi)
    with on_device(x.device):
        check(lib.vag_tc_split_f32(x.data_ptr(), x.stride(0), rows, K, hi.data_ptr(), lo.data_ptr(), K, _cabi.precision(), stream_ptr()))
    return hi, lo


def tc_gemm(xs, ws, rows: int, in_dim: int, out_dim: int, bias: Optional[torch.Tensor] = None, flags: int = 0,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(x·Wᵀ + bias) from operands split by ``tc_split`` (xs = (hi, lo), ws = (hi, lo))."""
    lib = _cabi.lib()
    dev = xs[0].device
    if out is None:
        out = torch.empty(rows, out_dim, dtype=torch.float32, device=dev)
    with on_device(dev):
        check(lib.vag_tc_gemm_f32(out.data_ptr(), out.stride(0), xs[0].data_ptr(), xs[1].data_ptr(), in_dim, ws[0].data_ptr(),
                                  ws[1].data_ptr(), in_dim, ptr(bias), rows, in_dim, out_dim, flags | _pflag(), stream_ptr()))
    return out


def tc_gemm_top2(xs, ws, rows: int, in_dim: int, out_dim: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Summary-only vocabulary projection → [ceil(out_dim/32), rows, 4] (best, Σexp, second, columns as int bits)."""
    lib = _cabi.lib()
    dev = xs[0].device
    summ = torch.empty((out_dim + 31) // 32, rows, 4, dtype=torch.float32, device=dev)
    with on_device(dev):
        check(lib.vag_tc_gemm_top2_f32(summ.data_ptr(), xs[0].data_ptr(), xs[1].data_ptr(), in_dim, ws[0].data_ptr(), ws[1].data_ptr(),
                                       in_dim, ptr(bias), rows, in_dim, out_dim, _cabi.precision(), stream_ptr()))
    return summ


def embed_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    _chk_f32(table)
    lib = _cabi.lib()
    ids = ids.reshape(-1).to(device=table.device, dtype=torch.int64).contiguous()
    out = torch.empty(ids.numel(), table.shape[1], dtype=torch.float32, device=table.device)
    with on_device(table.device):
        check(lib.vag_embed_rows_f32(out.data_ptr(), table.shape[1], table.data_ptr(), table.shape[1], ids.data_ptr(),
                                     ids.numel(), table.shape[0], stream_ptr()))
    return out


def gru_gates(gi: torch.Tensor, gh: torch.Tensor, h_prev: torch.Tensor, out: Optional[torch.Tensor] = None,
              out2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """h' = GRU gates; all operands 2-D with unit inner stride; out2 optionally receives a second copy of h'."""
    _chk_f32(gi, gh, h_prev, out, out2)
    lib = _cabi.lib()
    rows, H = h_prev.shape
    for t in (gi, gh, h_prev):
        assert t.stride(1) == 1
    if out is None:
        out = torch.empty(rows, H, dtype=torch.float32, device=h_prev.device)
    with on_device(h_prev.device):
        check(lib.vag_gru_gates_f32(out.data_ptr(), out.stride(0), ptr(out2), out2.stride(0) if out2 is not None else 0,
                                    gi.data_ptr(), gi.stride(0), gh.data_ptr(), gh.stride(0), h_prev.data_ptr(),
                                    h_prev.stride(0), rows, H, stream_ptr()))
    return out


def attention(q: torch.Tensor, keys: torch.Tensor, ctx: torch.Tensor, v: Optional[torch.Tensor],
              mask: Optional[torch.Tensor], rows_per_sent: int = 1, mode: int = ATTN_MLP,
              want_alpha: bool = True, out_c: Optional[torch.Tensor] = None,
              out_alpha: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """q [N, C]; keys/ctx [B, T, C] sentence-major; mask [B, T] → (c [N, C], α [N, T])."""
    _chk_f32(q, keys, ctx, v, mask)
    lib = _cabi.lib()
    q, keys, ctx = q.contiguous(), keys.contiguous(), ctx.contiguous()
    mask = mask.contiguous() if mask is not None else None
    N, Cdim = q.shape
    B, T, _ = ctx.shape
    assert N == B * rows_per_sent
    c = out_c if out_c is not None else torch.empty(N, Cdim, dtype=torch.float32, device=q.device)
    alpha = out_alpha if out_alpha is not None else (torch.empty(N, T, dtype=torch.float32, device=q.device) if want_alpha else None)
    assert c.is_contiguous() and (alpha is None or alpha.is_contiguous())
    with on_device(q.device):
        check(lib.vag_attention_f32(c.data_ptr(), Cdim, ptr(alpha), q.data_ptr(), Cdim, keys.data_ptr(), ctx.data_ptr(),
                                    ptr(v), ptr(mask), N, rows_per_sent, T, Cdim, mode, stream_ptr()))
    return c, alpha


def l2norm_rows_(x: torch.Tensor) -> torch.Tensor:
    _chk_f32(x)
    lib = _cabi.lib()
    assert x.dim() == 2 and x.stride(1) == 1
    with on_device(x.device):
        check(lib.vag_l2norm_rows_f32(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], stream_ptr()))
    return x


def log_softmax(logits: torch.Tensor) -> torch.Tensor:
    _chk_f32(logits)
    lib = _cabi.lib()
    logits = logits.contiguous()
    out = torch.empty_like(logits)
    with on_device(logits.device):
        check(lib.vag_log_softmax_f32(out.data_ptr(), logits.data_ptr(), logits.shape[0], logits.shape[1], stream_ptr()))
    return out


def nll_rows(logits: torch.Tensor, tgt: torch.Tensor, weight: Optional[torch.Tensor], loss_rows: torch.Tensor,
             lse_out: Optional[torch.Tensor] = None) -> None:
    """loss_rows[r] += -weight[tgt[r]]·log_softmax(logits)[r, tgt[r]]"""
    _chk_f32(logits, weight, loss_rows, lse_out)
    lib = _cabi.lib()
    assert logits.stride(1) == 1 and tgt.dtype == torch.int64 and tgt.is_contiguous()
    with on_device(logits.device):
        check(lib.vag_nll_rows_f32(logits.data_ptr(), logits.stride(0), tgt.data_ptr(), ptr(weight), logits.shape[0],
                                   logits.shape[1], loss_rows.data_ptr(), ptr(lse_out), stream_ptr()))


# ------------------------------------------------------------------ weight structs
def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise _cabi.VagError("model parameters must be contiguous fp32 CUDA tensors (call model.cuda())")
    return t.data_ptr()


def encoder_weights(enc) -> EncoderWeights:
    g = enc.gru
    w = EncoderWeights()
    w.E, w.H, w.vocab = enc.embedding.weight.shape[1], g.hidden_size, enc.embedding.weight.shape[0]
    w.precision = _cabi.precision()
    w.emb = _p(enc.embedding.weight)
    for d, sfx in enumerate(("", "_reverse")):
        w.w_ih[d] = _p(getattr(g, "weight_ih_l0" + sfx))
        w.w_hh[d] = _p(getattr(g, "weight_hh_l0" + sfx))
        w.b_ih[d] = _p(getattr(g, "bias_ih_l0" + sfx))
        w.b_hh[d] = _p(getattr(g, "bias_hh_l0" + sfx))
    return w


def vse_weights(vse) -> VseWeights:
    w = VseWeights()
    w.I, w.C, w.S = vse.im_size, vse.hidden_size, vse.shared_embedding_size
    w.method = ATTN_DOT if vse.attn_type == "dot" else ATTN_MLP
    w.activation = 1 if vse.activation_vse else 0
    w.precision = _cabi.precision()
    w.im_w, w.im_b = _p(vse.im_embedding.weight), _p(vse.im_embedding.bias)
    w.txt_w, w.txt_b = _p(vse.text_embedding.weight), _p(vse.text_embedding.bias)
    w.ctx2ctx_w = _p(vse.imagine_attn.ctx2ctx.weight)
    w.emb2ctx_w = _p(vse.imagine_attn.emb2ctx.weight)
    w.mlp_w = _p(vse.imagine_attn.mlp.weight) if vse.attn_type == "mlp" else None
    return w


_prepared = {}        # (id(decoder module), precision) → (key, buffer): decode-call invariants of a weight set
_weights_epoch = 0


def invalidate_prepared() -> None:
    """Parameters were updated through raw pointers (ClipAdam's fused kernel does not bump tensor versions): drop every cached
    ``vag_decoder_prepare_f32`` buffer."""
    global _weights_epoch
    _weights_epoch += 1
    _prepared.clear()


def _decoder_params(dec, decoderini):
    ps = [dec.embedding.weight, dec.gru_1.weight_ih_l0, dec.gru_1.weight_hh_l0, dec.gru_1.bias_ih_l0, dec.gru_1.bias_hh_l0,
          dec.attn.attn_h.weight, dec.attn.attn_e.weight, dec.attn.v, dec.context2hid.weight, dec.gru_2.weight_ih_l0,
          dec.gru_2.weight_hh_l0, dec.gru_2.bias_ih_l0, dec.gru_2.bias_hh_l0, dec.W1.weight, dec.W1.bias, dec.W2.weight, dec.W2.bias,
          dec.W3.weight, dec.W3.bias, dec.out.weight, dec.out.bias]
    if decoderini is not None:
        ps += [decoderini.weight, decoderini.bias]
    return ps


def decoder_weights(dec, decoderini=None, prepare: bool = False) -> DecoderWeights:
    """The weight struct of one call.  prepare=True attaches the decode-call invariants (operand planes of every matrix, the
    per-token gru_1 table; vag_decoder_prepare_f32), computed once per (weights, precision) and cached until a parameter
    changes — what makes decoding in the reference's eval batches of 16 (nmt_multimodal_beam_DE.py:542-547) affordable."""
    w = DecoderWeights()
    w.E, w.H, w.C, w.V = dec.embedding_size, dec.hidden_size, dec.context_size, dec.embedding.weight.shape[0]
    w.precision = _cabi.precision()
    w.emb = _p(dec.embedding.weight)
    w.gru1_w_ih, w.gru1_w_hh = _p(dec.gru_1.weight_ih_l0), _p(dec.gru_1.weight_hh_l0)
    w.gru1_b_ih, w.gru1_b_hh = _p(dec.gru_1.bias_ih_l0), _p(dec.gru_1.bias_hh_l0)
    w.attn_h_w, w.attn_e_w, w.attn_v = _p(dec.attn.attn_h.weight), _p(dec.attn.attn_e.weight), _p(dec.attn.v)
    w.c2h_w = _p(dec.context2hid.weight)
    w.gru2_w_ih, w.gru2_w_hh = _p(dec.gru_2.weight_ih_l0), _p(dec.gru_2.weight_hh_l0)
    w.gru2_b_ih, w.gru2_b_hh = _p(dec.gru_2.bias_ih_l0), _p(dec.gru_2.bias_hh_l0)
    w.w1_w, w.w1_b = _p(dec.W1.weight), _p(dec.W1.bias)
    w.w2_w, w.w2_b = _p(dec.W2.weight), _p(dec.W2.bias)
    w.w3_w, w.w3_b = _p(dec.W3.weight), _p(dec.W3.bias)
    w.out_w, w.out_b = _p(dec.out.weight), _p(dec.out.bias)
    if decoderini is not None:
        w.ini_w, w.ini_b = _p(decoderini.weight), _p(decoderini.bias)
    if prepare and w.E % 8 == 0 and w.H % 8 == 0 and w.C % 8 == 0 and w.V >= 64:
        params = _decoder_params(dec, decoderini)
        key = (tuple((t.data_ptr(), t._version) for t in params), _weights_epoch)
        slot = (id(dec), w.precision)
        hit = _prepared.get(slot)
        if hit is None or hit[0] != key:
            lib = _cabi.lib()
            dev = dec.embedding.weight.device
            buf = torch.empty(lib.vag_decoder_prepared_bytes(w.E, w.H, w.C, w.V), dtype=torch.uint8, device=dev)
            with on_device(dev):
                status = lib.vag_decoder_prepare_f32(C.byref(w), buf.data_ptr(), buf.numel(), stream_ptr())
            if status == -4:          # VAG_ERR_UNSUPPORTED (VAG_GEMM=simt / tf32x3, unaligned weights): decode without it
                buf = None
            else:
                check(status)
            hit = _prepared[slot] = (key, buf)
        if hit[1] is not None:
            w.prepared, w.prepared_bytes = hit[1].data_ptr(), hit[1].numel()
            w._keepalive = hit[1]
    return w


# ------------------------------------------------------------------ composites
def encoder_fwd(w: EncoderWeights, src: torch.Tensor, lengths: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """→ ctx [B, T, 2H] sentence-major, mask [B, T]"""
    lib = _cabi.lib()
    dev = _dev(src)
    src = src.to(torch.int64).contiguous()
    B, T = src.shape
    lens = (C.c_int32 * B)(*[int(x) for x in lengths])
    ctx = torch.empty(B, T, 2 * w.H, dtype=torch.float32, device=dev)
    mask = torch.empty(B, T, dtype=torch.float32, device=dev)
    nbytes = lib.vag_encoder_workspace_bytes(B, T, w.E, w.H)
    ws = workspace(nbytes, dev)
    with on_device(dev):
        check(lib.vag_encoder_fwd_f32(C.byref(w), src.data_ptr(), lens, B, T, ctx.data_ptr(), mask.data_ptr(), ws.data_ptr(),
                                      ws.numel(), stream_ptr()))
    return ctx, mask


def vse_pool_fwd(w: VseWeights, im: torch.Tensor, ctx: torch.Tensor, mask: torch.Tensor, want_beta: bool = False):
    """→ im_emb [B,S], txt_emb [B,S], ctx_vec [B,C], beta [B,T] or None"""
    _chk_f32(im, ctx, mask)
    lib = _cabi.lib()
    dev = ctx.device
    im, ctx, mask = im.contiguous(), ctx.contiguous(), mask.contiguous()
    B, T, Cd = ctx.shape
    im_emb = torch.empty(B, w.S, dtype=torch.float32, device=dev)
    txt_emb = torch.empty(B, w.S, dtype=torch.float32, device=dev)
    ctx_vec = torch.empty(B, Cd, dtype=torch.float32, device=dev)
    beta = torch.empty(B, T, dtype=torch.float32, device=dev) if want_beta else None
    nbytes = lib.vag_vse_workspace_bytes(B, T, w.I, w.C, w.S)
    ws = workspace(nbytes, dev)
    with on_device(dev):
        check(lib.vag_vse_pool_fwd_f32(C.byref(w), im.data_ptr(), ctx.data_ptr(), mask.data_ptr(), B, T, im_emb.data_ptr(),
                                       txt_emb.data_ptr(), ctx_vec.data_ptr(), ptr(beta), ws.data_ptr(), ws.numel(),
                                       stream_ptr()))
    return im_emb, txt_emb, ctx_vec, beta


def attn_keys(w: DecoderWeights, ctx: torch.Tensor) -> torch.Tensor:
    lib = _cabi.lib()
    ctx = ctx.contiguous()
    B, T, Cd = ctx.shape
    keys = torch.empty_like(ctx)
    ws = workspace(lib.vag_attn_keys_workspace_bytes(B, T, Cd), ctx.device)
    with on_device(ctx.device):
        check(lib.vag_attn_keys_f32(C.byref(w), ctx.data_ptr(), B, T, keys.data_ptr(), ws.data_ptr(), ws.numel(),
                                    stream_ptr()))
    return keys


def decoder_init(w: DecoderWeights, ctx_vec: Optional[torch.Tensor], ctx: torch.Tensor, mask: torch.Tensor,
                 split: float) -> torch.Tensor:
    lib = _cabi.lib()
    B, T, Cd = ctx.shape
    h0 = torch.empty(B, w.H, dtype=torch.float32, device=ctx.device)
    ws = workspace(lib.vag_decoder_init_workspace_bytes(B, Cd, w.H), ctx.device)
    with on_device(ctx.device):
        check(lib.vag_decoder_init_f32(C.byref(w), ptr(ctx_vec), ctx.data_ptr(), mask.data_ptr(), float(split), B, T,
                                       h0.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()))
    return h0


def decoder_step(w: DecoderWeights, tokens: torch.Tensor, h_prev: torch.Tensor, keys: torch.Tensor, ctx: torch.Tensor,
                 mask: torch.Tensor, rows_per_sent: int = 1, want_logp: bool = True, want_alpha: bool = False):
    """→ (logp or logits [rows, V], h [rows, H], alpha or None)"""
    lib = _cabi.lib()
    dev = ctx.device
    tokens = tokens.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    rows = tokens.numel()
    B, T, Cd = ctx.shape
    assert rows == B * rows_per_sent, (rows, B, rows_per_sent)
    h_prev = h_prev.reshape(rows, w.H).contiguous()
    h_out = torch.empty_like(h_prev)
    out = torch.empty(rows, w.V, dtype=torch.float32, device=dev)
    alpha = torch.empty(rows, T, dtype=torch.float32, device=dev) if want_alpha else None
    nbytes = lib.vag_decoder_step_workspace_bytes(rows, w.E, w.H, w.C, w.V)
    ws = workspace(nbytes, dev)
    with on_device(dev):
        check(lib.vag_decoder_step_f32(C.byref(w), tokens.data_ptr(), h_prev.data_ptr(), keys.data_ptr(), ctx.data_ptr(),
                                       mask.data_ptr(), rows, rows_per_sent, T, h_out.data_ptr(), out.data_ptr(),
                                       1 if want_logp else 0, ptr(alpha), ws.data_ptr(), ws.numel(), stream_ptr()))
    return out, h_out, alpha


def beam_select(logp: torch.Tensor, prev_tokens: Optional[torch.Tensor], nll: torch.Tensor, B: int, K: int, step: int,
                avoid_double: bool = True):
    """In-place on nll [B,K]; → tokens int64 [B,K], parents int32 [B,K]"""
    _chk_f32(logp, nll)
    lib = _cabi.lib()
    dev = logp.device
    logp = logp.contiguous()
    V = logp.shape[1]
    tokens = torch.empty(B, K, dtype=torch.int64, device=dev)
    parents = torch.empty(B, K, dtype=torch.int32, device=dev)
    if prev_tokens is not None:
        prev_tokens = prev_tokens.to(torch.int64).contiguous()
    with on_device(dev):
        check(lib.vag_beam_select_f32(logp.data_ptr(), V, ptr(prev_tokens), nll.data_ptr(), tokens.data_ptr(),
                                      parents.data_ptr(), B, K, V, step, 1 if avoid_double else 0, stream_ptr()))
    return tokens, parents


_progress_words = {}


def _host_progress(dev: torch.device) -> Optional[torch.Tensor]:
    """One pinned (device-mapped) int32 per device for vag_beam_decode_f32's early stop; None while a CUDA graph is being
    captured (a captured call must not poll the host)."""
    if torch.cuda.is_current_stream_capturing():
        return None
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _lane)
    t = _progress_words.get(key)
    if t is None:
        t = _progress_words[key] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return t


def beam_decode(w: DecoderWeights, h0: torch.Tensor, keys: torch.Tensor, ctx: torch.Tensor, mask: torch.Tensor, K: int,
                L: int, avoid_double: bool = True, debug: bool = False, early_stop: bool = True):
    """→ hyp [B, L] int64, hyp_len [B] int32 (+ beam [L,B,K], nll [B,K], steps [1] when debug).  early_stop: the host follows
    the device's progress word and stops launching steps once every hypothesis has ended (V11:265-269)."""
    lib = _cabi.lib()
    dev = ctx.device
    B, T, Cd = ctx.shape
    hyp = torch.empty(B, L, dtype=torch.int64, device=dev)
    hyp_len = torch.empty(B, dtype=torch.int32, device=dev)
    beam = torch.empty(L, B, K, dtype=torch.int64, device=dev) if debug else None
    nll = torch.empty(B, K, dtype=torch.float32, device=dev) if debug else None
    steps = torch.empty(1, dtype=torch.int32, device=dev) if debug else None
    nbytes = lib.vag_beam_decode_workspace_bytes(B, K, T, L, w.E, w.H, w.C, w.V)
    ws = workspace(nbytes, dev)
    prog = _host_progress(dev) if early_stop else None
    with on_device(dev):
        check(lib.vag_beam_decode_f32(C.byref(w), h0.data_ptr(), keys.data_ptr(), ctx.data_ptr(), mask.data_ptr(), B, K, T,
                                      L, 1 if avoid_double else 0, hyp.data_ptr(), hyp_len.data_ptr(), ptr(beam), ptr(nll),
                                      ptr(steps), ptr(prog), ws.data_ptr(), ws.numel(), stream_ptr()))
    if debug:
        return hyp, hyp_len, beam, nll, steps
    return hyp, hyp_len


def beam_decode_steps(w: DecoderWeights, h0: torch.Tensor, keys: torch.Tensor, ctx: torch.Tensor, mask: torch.Tensor, K: int, L: int,
                      step_begin: int, step_end: int, hyp: torch.Tensor, hyp_len: torch.Tensor, done_host: Optional[torch.Tensor] = None,
                      avoid_double: bool = True) -> None:
    """Steps [step_begin, step_end) of the search ``beam_decode`` runs in one call, on the state the lane's workspace holds
    (vag_beam_decode_steps_f32): 0 initialises, step_end >= L appends the epilogue and fills hyp [B, L] / hyp_len [B].  done_host:
    one pinned int32 that receives the `done` flag after the range, in stream order.  Capturable in a CUDA graph."""
    lib = _cabi.lib()
    dev = ctx.device
    B, T, _ = ctx.shape
    nbytes = lib.vag_beam_decode_workspace_bytes(B, K, T, L, w.E, w.H, w.C, w.V)
    ws = workspace(nbytes, dev)
    with on_device(dev):
        check(lib.vag_beam_decode_steps_f32(C.byref(w), h0.data_ptr(), keys.data_ptr(), ctx.data_ptr(), mask.data_ptr(), B, K, T, L,
                                            1 if avoid_double else 0, step_begin, step_end, hyp.data_ptr(), hyp_len.data_ptr(), None, None,
                                            None, ptr(done_host), ws.data_ptr(), ws.numel(), stream_ptr()))


def greedy_decode(w: DecoderWeights, h0: torch.Tensor, keys: torch.Tensor, ctx: torch.Tensor, mask: torch.Tensor,
                  L: int) -> torch.Tensor:
    lib = _cabi.lib()
    dev = ctx.device
    B, T, Cd = ctx.shape
    toks = torch.empty(B, L, dtype=torch.int64, device=dev)
    nbytes = lib.vag_beam_decode_workspace_bytes(B, 1, T, L, w.E, w.H, w.C, w.V)
    ws = workspace(nbytes, dev)
    with on_device(dev):
        check(lib.vag_greedy_decode_f32(C.byref(w), h0.data_ptr(), keys.data_ptr(), ctx.data_ptr(), mask.data_ptr(), B, T, L,
                                        toks.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()))
    return toks


def rank_loss(im: torch.Tensor, s: torch.Tensor, margin: float, one_direction: bool = False, want_grad: bool = False):
    _chk_f32(im, s)
    lib = _cabi.lib()
    dev = im.device
    im, s = im.contiguous(), s.contiguous()
    B, S = im.shape
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    g_im = torch.empty_like(im) if want_grad else None
    g_s = torch.empty_like(s) if want_grad else None
    ws = workspace(lib.vag_rank_loss_workspace_bytes(B, S), dev)
    with on_device(dev):
        check(lib.vag_rank_loss_f32(im.data_ptr(), s.data_ptr(), B, S, float(margin), 1 if one_direction else 0,
                                    loss.data_ptr(), ptr(g_im), ptr(g_s), ws.data_ptr(), ws.numel(), stream_ptr()))
    return loss[0], g_im, g_s


def recall_ranks(queries: torch.Tensor, gallery: torch.Tensor) -> torch.Tensor:
    _chk_f32(queries, gallery)
    lib = _cabi.lib()
    dev = queries.device
    queries, gallery = queries.contiguous(), gallery.contiguous()
    n, S = queries.shape
    ranks = torch.empty(n, dtype=torch.int32, device=dev)
    ws = workspace(lib.vag_recall_ranks_workspace_bytes(n, S), dev)
    with on_device(dev):
        check(lib.vag_recall_ranks_f32(queries.data_ptr(), gallery.data_ptr(), n, S, ranks.data_ptr(), ws.data_ptr(),
                                       ws.numel(), stream_ptr()))
    return ranks


def row_argmax(logits: torch.Tensor) -> torch.Tensor:
    _chk_f32(logits)
    lib = _cabi.lib()
    assert logits.dim() == 2 and logits.stride(1) == 1
    out = torch.empty(logits.shape[0], dtype=torch.int64, device=logits.device)
    with on_device(logits.device):
        check(lib.vag_row_argmax_f32(logits.data_ptr(), logits.stride(0), logits.shape[0], logits.shape[1], out.data_ptr(),
                                     stream_ptr()))
    return out


def translation_loss_bwd(g: torch.Tensor, tgt: torch.Tensor, loss_w: float, has_vse: bool):
    """g [3] = d/d(loss, loss_mt, loss_vse) → (g_rows [B], g_vse [1] | None), one launch"""
    _chk_f32(g)
    lib = _cabi.lib()
    B, Tt = tgt.shape
    g_rows = torch.empty(B, dtype=torch.float32, device=g.device)
    g_vse = torch.empty(1, dtype=torch.float32, device=g.device) if has_vse else None
    with on_device(g.device):
        check(lib.vag_translation_loss_bwd_f32(g.data_ptr(), tgt.data_ptr(), B, Tt, float(loss_w), int(has_vse),
                                               g_rows.data_ptr(), ptr(g_vse), stream_ptr()))
    return g_rows, g_vse


def src_mask_lengths(src: torch.Tensor, want_lengths: bool = True):
    """src int64 [B,T] → (mask fp32 [B,T] = (src != 0), lengths int32 [B] | None), one launch (Encoder.py:47)"""
    lib = _cabi.lib()
    B, T = src.shape
    mask = torch.empty(B, T, dtype=torch.float32, device=src.device)
    lengths = torch.empty(B, dtype=torch.int32, device=src.device) if want_lengths else None
    with on_device(src.device):
        check(lib.vag_src_mask_lengths(src.data_ptr(), B, T, mask.data_ptr(), ptr(lengths), stream_ptr()))
    return mask, lengths


def translation_loss(loss_rows: torch.Tensor, tgt: torch.Tensor, loss_vse: Optional[torch.Tensor], loss_w: float) -> torch.Tensor:
    """→ fp32 [3] = (loss, loss_mt, loss_vse)"""
    _chk_f32(loss_rows, loss_vse)
    lib = _cabi.lib()
    B, Tt = tgt.shape
    out = torch.empty(3, dtype=torch.float32, device=loss_rows.device)
    with on_device(loss_rows.device):
        check(lib.vag_translation_loss_f32(loss_rows.data_ptr(), tgt.data_ptr(), B, Tt, ptr(loss_vse), float(loss_w),
                                           out.data_ptr(), stream_ptr()))
    return out
