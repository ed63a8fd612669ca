"""Autograd Functions of the training step (V11.forward + loss.backward(), train.py:36-51).

Each Function's forward AND backward are sequences of libvagnmt.so kernels; torch only records the graph between
them and owns the buffers.  Back-propagation through time is written out by hand:

  EncoderFn      packed bidirectional GRU (layers/Encoder.py:36-65)                       BPTT over Ts
  VsePoolFn      visual-attention pooling + shared-space embeddings (VSE_Imagine_Enc.py)  one shot
  DecoderInitFn  h0 = tanh(decoderini(mix))  (V11:118)                                    one shot
  DecoderSeqFn   the Tt-step conditional-GRU loop with NLL (V11:136-160)                  BPTT over Tt
  LossMixFn      per-sentence normalisation, batch mean, loss_w mix (V11:164-166)

Time-invariant contractions are batched over all steps (embedding → W_ih, the read-out W1/W2/W3, the vocabulary
projection, every weight gradient); only the five recurrent contractions per step stay inside the loops.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from ._cabi import on_device
from . import train_ops as T

SOS_token = 2


def _save_mode(fctx):
    """Remember the vag_precision the forward ran in.  loss.backward() executes the backward Functions on autograd's device
    worker thread, which does not inherit the caller's (Python-side, thread-local) precision scope."""
    fctx.precision = ops._cabi.precision()


def _in_forward_mode(fn):
    """Run a backward staticmethod under the precision its forward saved."""
    import functools

    @functools.wraps(fn)
    def wrapper(fctx, *grads):
        with ops._cabi.precision_scope(getattr(fctx, "precision", ops._cabi.PREC_FP32)):
            return fn(fctx, *grads)
    return wrapper


def _lin(x, w, b=None, flags=0, out=None):
    """y = act(x Wᵀ + b): tcgen05 split-precision kernel for eligible shapes, FP32 FFMA otherwise."""
    rows, K = x.shape
    N = w.shape[0]
    if (rows >= 64 and N >= 64 and K >= 32 and K % 8 == 0 and x.stride(0) % 4 == 0 and w.stride(0) % 4 == 0
            and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0 and x.stride(1) == 1 and w.stride(1) == 1):
        return ops.linear_tc(x, w, b, flags, out)
    return ops.linear(x, w, b, flags, out)


# Where parameter gradients are written.  By default every backward allocates its outputs; an optimiser that keeps ONE persistent
# flat gradient buffer (optim.ClipAdam) installs a sink mapping a parameter to its slice of that buffer, so that the backward
# kernels write the gradients in place: no torch.cat before the data-parallel all-reduce, no copy back after it, and every captured
# CUDA graph of the step writes to the same addresses instead of pinning a private 64 MB gradient set in its pool.
_grad_sink = None


def set_grad_sink(fn) -> None:
    global _grad_sink
    _grad_sink = fn


def _gbuf(p: torch.Tensor) -> torch.Tensor:
    """Uninitialised buffer for the gradient of parameter `p` (the backward kernels overwrite every element)."""
    if _grad_sink is not None:
        v = _grad_sink(p)
        if v is not None:
            return v
    return torch.empty_like(p)


_side_streams = {}
_pending_join = {}


def _overlap_weight_gradients() -> bool:
    import os
    return os.environ.get("VAG_TRAIN_OVERLAP", "1") != "0"


def _side_stream(dev: torch.device):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=dev)
    return st


def _join_after_backward(dev: torch.device, side) -> None:
    """Make the stream the backward pass runs on wait for `side` when the pass ends (once per pass)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if _pending_join.get(key):
        return
    _pending_join[key] = True

    def join():
        _pending_join[key] = False
        torch.cuda.current_stream(dev).wait_stream(side)

    torch.autograd.Variable._execution_engine.queue_callback(join)


def _zeros(*shape, like):
    return torch.zeros(*shape, dtype=torch.float32, device=like.device)


def _empty(*shape, like):
    return torch.empty(*shape, dtype=torch.float32, device=like.device)


# ---------------------------------------------------------------------------------------------------- encoder
class EncoderFn(torch.autograd.Function):
    """src int64 [B, T] (sorted by length desc), lengths list → ctx [B, T, 2H] (sentence-major).
    Forward (with saved pre-activations) and BPTT are one C call each."""

    @staticmethod
    def forward(fctx, src, lengths, emb_mask, emb, *gru):
        _save_mode(fctx)
        # gru = (w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r); emb_mask: [T·B, E] time-major dropout mask or None
        import ctypes as C
        from ._cabi import EncoderWeights
        lib = ops._cabi.lib()
        B, Tn = src.shape
        E, H = emb.shape[1], gru[1].shape[1]
        dev = emb.device
        w = EncoderWeights()
        w.E, w.H, w.vocab, w.emb = E, H, emb.shape[0], ops._p(emb.detach())
        w.precision = ops._cabi.precision()
        for d in range(2):
            w.w_ih[d], w.w_hh[d] = ops._p(gru[4 * d].detach()), ops._p(gru[4 * d + 1].detach())
            w.b_ih[d], w.b_hh[d] = ops._p(gru[4 * d + 2].detach()), ops._p(gru[4 * d + 3].detach())
        # lengths: python list (validated on the host, copied to the device) or an int32 CUDA tensor [B] (graph-captured step:
        # the caller refreshes that buffer between replays; nothing about the launch sequence depends on its contents)
        if torch.is_tensor(lengths):
            lens, lens_dev = None, lengths.to(device=dev, dtype=torch.int32).contiguous()
        else:
            lens = (C.c_int32 * B)(*[int(x) for x in lengths])
            lens_dev = torch.tensor([int(x) for x in lengths], dtype=torch.int32).to(dev, non_blocking=True)
        ctx = torch.empty(B, Tn, 2 * H, dtype=torch.float32, device=dev)
        x = torch.empty(Tn * B, E, dtype=torch.float32, device=dev)
        ids_tm = torch.empty(Tn * B, dtype=torch.int64, device=dev)
        gi = torch.empty(2, Tn, B, 3 * H, dtype=torch.float32, device=dev)
        gh = torch.empty(2, Tn, B, 3 * H, dtype=torch.float32, device=dev)
        ws = ops.workspace(lib.vag_encoder_train_workspace_bytes(B, Tn, E, H), dev)
        with on_device(dev):
            ops.check(lib.vag_encoder_train_fwd_f32(C.byref(w), src.data_ptr(), lens, lens_dev.data_ptr(), B, Tn, ctx.data_ptr(), x.data_ptr(),
                                                    ids_tm.data_ptr(), gi.data_ptr(), gh.data_ptr(), ops.ptr(emb_mask), ws.data_ptr(),
                                                    ws.numel(), ops.stream_ptr()))
        fctx.save_for_backward(x, ctx, ids_tm, gi, gh, emb_mask if emb_mask is not None else ids_tm, lens_dev, emb, *gru)
        fctx.meta = (B, Tn, E, H, emb_mask is not None)
        return ctx

    @staticmethod
    @_in_forward_mode
    def backward(fctx, dctx):
        import ctypes as C
        from ._cabi import EncoderWeights
        lib = ops._cabi.lib()
        x, ctx, ids_tm, gi, gh, emb_mask, lens_dev, emb, *gru = fctx.saved_tensors
        B, Tn, E, H, has_mask = fctx.meta
        dev = emb.device
        w = EncoderWeights()
        w.E, w.H, w.vocab, w.emb = E, H, emb.shape[0], ops._p(emb.detach())
        w.precision = ops._cabi.precision()
        for d in range(2):
            w.w_ih[d], w.w_hh[d] = ops._p(gru[4 * d].detach()), ops._p(gru[4 * d + 1].detach())
            w.b_ih[d], w.b_hh[d] = ops._p(gru[4 * d + 2].detach()), ops._p(gru[4 * d + 3].detach())
        demb = _gbuf(emb)
        grads = [_gbuf(p) for p in gru]
        arr = lambda idx: (C.c_void_p * 2)(grads[idx].data_ptr(), grads[4 + idx].data_ptr())
        ws = ops.workspace(lib.vag_encoder_train_workspace_bytes(B, Tn, E, H), dev)
        with on_device(dev):
            ops.check(lib.vag_encoder_bwd_f32(C.byref(w), None, lens_dev.data_ptr(), B, Tn, ctx.data_ptr(), dctx.contiguous().data_ptr(), x.data_ptr(),
                                              ids_tm.data_ptr(), gi.data_ptr(), gh.data_ptr(), demb.data_ptr(), arr(0), arr(1), arr(2),
                                              arr(3), emb_mask.data_ptr() if has_mask else None, ws.data_ptr(), ws.numel(),
                                              ops.stream_ptr()))
        return (None, None, None, demb, *grads)


class MaskMulFn(torch.autograd.Function):
    """y = x ⊙ m with a fixed mask (dropout with a pre-drawn, pre-scaled mask): dx = dy ⊙ m."""

    @staticmethod
    def forward(fctx, x, m):
        lib = ops._cabi.lib()
        y = x.contiguous().clone()
        with on_device(x.device):
            ops.check(lib.vag_mul_f32(y.data_ptr(), m.data_ptr(), y.numel(), ops.stream_ptr()))
        fctx.save_for_backward(m)
        return y

    @staticmethod
    def backward(fctx, dy):
        (m,) = fctx.saved_tensors
        lib = ops._cabi.lib()
        dx = dy.contiguous().clone()
        with on_device(dx.device):
            ops.check(lib.vag_mul_f32(dx.data_ptr(), m.data_ptr(), dx.numel(), ops.stream_ptr()))
        return dx, None


# ---------------------------------------------------------------------------------------------------- VSE pooling
class VsePoolFn(torch.autograd.Function):
    """(im [B,I], ctx [B,T,C], mask [B,T]) → im_emb [B,S], txt_emb [B,S], ctx_vec [B,C].
    Forward (with saved activations) and backward are ONE C call each: vag_vse_pool_train_fwd_f32 / vag_vse_pool_bwd_f32."""

    @staticmethod
    def _weights(method, activation, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w, mlp_w):
        from ._cabi import VseWeights
        w = VseWeights()
        w.I, w.C, w.S = im_w.shape[1], ctx2ctx_w.shape[0], im_w.shape[0]
        w.method = ops.ATTN_DOT if method == "dot" else ops.ATTN_MLP
        w.activation = 1 if activation else 0
        w.precision = ops._cabi.precision()
        w.im_w, w.im_b, w.txt_w, w.txt_b = ops._p(im_w.detach()), ops._p(im_b.detach()), ops._p(txt_w.detach()), ops._p(txt_b.detach())
        w.ctx2ctx_w, w.emb2ctx_w = ops._p(ctx2ctx_w.detach()), ops._p(emb2ctx_w.detach())
        w.mlp_w = ops._p(mlp_w.detach().reshape(-1)) if mlp_w is not None else None
        return w

    @staticmethod
    def forward(fctx, im, ctx, mask, method, activation, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w, mlp_w):
        _save_mode(fctx)
        import ctypes as C
        from ._cabi import VseSaved
        lib = ops._cabi.lib()
        B, Tn, Cd = ctx.shape
        S, dev = im_w.shape[0], ctx.device
        im, ctx, mask = im.contiguous(), ctx.contiguous(), mask.contiguous()
        w = VsePoolFn._weights(method, activation, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w, mlp_w)
        im_emb, txt_emb = _empty(B, S, like=ctx), _empty(B, S, like=ctx)
        ctx_vec, beta = _empty(B, Cd, like=ctx), _empty(B, Tn, like=ctx)
        a_im, a_txt, iq, pk = _empty(B, S, like=ctx), _empty(B, S, like=ctx), _empty(B, Cd, like=ctx), _empty(B, Tn, Cd, like=ctx)
        sv = VseSaved()
        sv.a_im, sv.iq, sv.pk, sv.a_txt = a_im.data_ptr(), iq.data_ptr(), pk.data_ptr(), a_txt.data_ptr()
        ws = ops.workspace(lib.vag_vse_workspace_bytes(B, Tn, w.I, w.C, w.S), dev)
        with on_device(dev):
            ops.check(lib.vag_vse_pool_train_fwd_f32(C.byref(w), im.data_ptr(), ctx.data_ptr(), mask.data_ptr(), B, Tn, im_emb.data_ptr(),
                                                     txt_emb.data_ptr(), ctx_vec.data_ptr(), beta.data_ptr(), C.byref(sv), ws.data_ptr(),
                                                     ws.numel(), ops.stream_ptr()))
        fctx.save_for_backward(im, ctx, mask, a_im, im_emb, iq, pk, beta, ctx_vec, a_txt, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w,
                               mlp_w if mlp_w is not None else im_b)
        fctx.meta = (method, activation, mlp_w is not None)
        return im_emb, txt_emb, ctx_vec

    @staticmethod
    @_in_forward_mode
    def backward(fctx, d_im_emb, d_txt_emb, d_ctx_vec):
        import ctypes as C
        from ._cabi import VseGrads, VseSaved
        lib = ops._cabi.lib()
        (im, ctx, mask, a_im, im_emb, iq, pk, beta, ctx_vec, a_txt, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w, mlp_w) = fctx.saved_tensors
        method, activation, has_mlp = fctx.meta
        B, Tn, Cd = ctx.shape
        dev = ctx.device
        w = VsePoolFn._weights(method, activation, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w, mlp_w if has_mlp else None)
        sv = VseSaved()
        sv.a_im, sv.iq, sv.pk, sv.a_txt = a_im.data_ptr(), iq.data_ptr(), pk.data_ptr(), a_txt.data_ptr()
        grads = [_gbuf(t) for t in (im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w)]
        d_mlp = _gbuf(mlp_w) if has_mlp else None
        g = VseGrads()
        g.im_w, g.im_b, g.txt_w, g.txt_b, g.ctx2ctx_w, g.emb2ctx_w = [t.data_ptr() for t in grads]
        g.mlp_w = d_mlp.data_ptr() if has_mlp else None
        dctx = _empty(B, Tn, Cd, like=ctx)
        cont = lambda t: None if t is None else t.contiguous()
        d_im_emb, d_txt_emb, d_ctx_vec = cont(d_im_emb), cont(d_txt_emb), cont(d_ctx_vec)
        ws = ops.workspace(lib.vag_vse_pool_bwd_workspace_bytes(B, Tn, w.I, w.C, w.S), dev)
        with on_device(dev):
            ops.check(lib.vag_vse_pool_bwd_f32(C.byref(w), im.data_ptr(), ctx.data_ptr(), mask.data_ptr(), B, Tn, C.byref(sv), im_emb.data_ptr(),
                                               beta.data_ptr(), ctx_vec.data_ptr(), ops.ptr(d_im_emb), ops.ptr(d_txt_emb), ops.ptr(d_ctx_vec),
                                               C.byref(g), dctx.data_ptr(), ws.data_ptr(), ws.numel(), ops.stream_ptr()))
        return (None, dctx, None, None, None, *grads, d_mlp)


# ---------------------------------------------------------------------------------------------------- decoder init
class DecoderInitFn(torch.autograd.Function):
    """h0 = tanh(decoderini(split·ctx_vec + (1-split)·mean_t ctx))   (ctx_vec None ⇒ text-only model)."""

    @staticmethod
    def forward(fctx, ctx_vec, ctx, mask, split, ini_w, ini_b):
        _save_mode(fctx)
        lib_mix = ops._cabi.lib()
        B, Tn, C = ctx.shape
        z = _empty(B, C, like=ctx)
        with on_device(ctx.device):
            ops.check(lib_mix.vag_init_mix_f32(z.data_ptr(), ops.ptr(ctx_vec), ctx.data_ptr(), mask.data_ptr(), float(split), B, Tn, C,
                                               ops.stream_ptr()))
        h0 = _lin(z, ini_w, ini_b, ops.LIN_TANH)
        fctx.save_for_backward(z, h0, mask, ini_w)
        fctx.meta = (float(split), ctx_vec is not None, (B, Tn, C))
        fctx.bias_ref = ini_b
        return h0

    @staticmethod
    @_in_forward_mode
    def backward(fctx, dh0):
        z, h0, mask, ini_w = fctx.saved_tensors
        split, has_vec, (B, Tn, C) = fctx.meta
        du = T.tanh_bwd(dh0, h0)
        d_w = T.gemm(du, z, trans_a=True, out=_gbuf(ini_w))
        d_b = T.colsum(du, out=_gbuf(fctx.bias_ref))
        dz = T.gemm(du, ini_w)
        dctx = _zeros(B, Tn, C, like=z)
        dvec = T.init_mix_bwd(dz, mask, split, dctx, has_vec)
        return dvec, dctx, None, None, d_w, d_b


# ---------------------------------------------------------------------------------------------------- decoder loop
_PARAM_FIELDS = ("emb", "gru1_w_ih", "gru1_w_hh", "gru1_b_ih", "gru1_b_hh", "attn_h_w", "attn_e_w", "attn_v", "c2h_w", "gru2_w_ih",
                 "gru2_w_hh", "gru2_b_ih", "gru2_b_hh", "w1_w", "w1_b", "w2_w", "w2_b", "w3_w", "w3_b", "out_w", "out_b")


class DecoderSeqFn(torch.autograd.Function):
    """(h0 [B,H], ctx [B,T,C], mask [B,T], tgt [B,Tt]) → Σ_t NLL rows [B]   (V11:136-160, NMT_Decoder.py:109-145).

    Forward and backward are ONE C call each (vag_decoder_seq_fwd_f32 / vag_decoder_seq_bwd_f32): the Tt-step loop,
    the batched read-out / vocabulary projection / NLL and the whole BPTT run inside libvagnmt.so.
    """

    @staticmethod
    def forward(fctx, h0, enc, mask, tgt, weight, teacher, tied, out_mask, *params):
        _save_mode(fctx)
        import ctypes as C
        from ._cabi import DecoderGrads, DecoderSeqSaved, DecoderWeights
        lib = ops._cabi.lib()
        B, Tt = tgt.shape
        _, Tn, Cd = enc.shape
        emb, out_w = params[0], params[19]
        H, E, V = h0.shape[1], emb.shape[1], out_w.shape[0]
        dev = h0.device
        h0, enc, mask = h0.contiguous(), enc.contiguous(), mask.contiguous()
        tgt_t = tgt.t().contiguous()
        tok_in = torch.empty(Tt, B, dtype=torch.int64, device=dev)
        tok_in[0].fill_(SOS_token)
        if teacher and Tt > 1:
            tok_in[1:].copy_(tgt_t[:-1])
        w = DecoderWeights()
        w.E, w.H, w.C, w.V = E, H, Cd, V
        w.precision = ops._cabi.precision()
        for name, p in zip(_PARAM_FIELDS, params):
            setattr(w, name, ops._p(p.detach()))
        ldl = (V + 3) // 4 * 4
        sizes = dict(keys=B * Tn * Cd, e_all=Tt * B * E, gi1_all=Tt * B * 3 * H, gh1_all=Tt * B * 3 * H, h1_all=Tt * B * H,
                     q_all=Tt * B * Cd, alpha_all=Tt * B * Tn, c_all=Tt * B * Cd, x2_all=Tt * B * H, gi2_all=Tt * B * 3 * H,
                     gh2_all=Tt * B * 3 * H, h2_all=Tt * B * H, t_all=Tt * B * E, logits_all=Tt * B * ldl, lse_all=Tt * B)
        offs, total = {}, 0
        for k, n in sizes.items():
            offs[k] = total
            total += (n + 63) // 64 * 64
        store = torch.empty(total, dtype=torch.float32, device=dev)       # one allocation for every saved activation
        saved = DecoderSeqSaved()
        saved.ld_logits = ldl
        for k in sizes:
            setattr(saved, k, store.data_ptr() + 4 * offs[k])
        loss_rows = torch.empty(B, dtype=torch.float32, device=dev)
        ws = ops.workspace(lib.vag_decoder_seq_workspace_bytes(B, Tn, Tt, E, H, Cd, V), dev)
        with on_device(dev):
            ops.check(lib.vag_decoder_seq_fwd_f32(C.byref(w), h0.data_ptr(), enc.data_ptr(), mask.data_ptr(), tok_in.data_ptr(),
                                                  tgt_t.data_ptr(), ops.ptr(weight), B, Tn, Tt, 1 if teacher else 0, C.byref(saved),
                                                  ops.ptr(out_mask), loss_rows.data_ptr(), ws.data_ptr(), ws.numel(), ops.stream_ptr()))
        fctx.save_for_backward(h0, enc, mask, tgt_t, tok_in, store, weight if weight is not None else loss_rows,
                               out_mask if out_mask is not None else loss_rows, *params)
        fctx.meta = (weight is not None, bool(tied), offs, ldl, (B, Tn, Tt, E, H, Cd, V), out_mask is not None)
        return loss_rows

    @staticmethod
    @_in_forward_mode
    def backward(fctx, dloss_rows):
        import ctypes as C
        from ._cabi import DecoderGrads, DecoderSeqSaved, DecoderWeights
        lib = ops._cabi.lib()
        h0, enc, mask, tgt_t, tok_in, store, weight, out_mask, *params = fctx.saved_tensors
        has_weight, tied, offs, ldl, (B, Tn, Tt, E, H, Cd, V), has_out_mask = fctx.meta
        dev = h0.device
        w = DecoderWeights()
        w.E, w.H, w.C, w.V = E, H, Cd, V
        w.precision = ops._cabi.precision()
        for name, p in zip(_PARAM_FIELDS, params):
            setattr(w, name, ops._p(p.detach()))
        saved = DecoderSeqSaved()
        saved.ld_logits = ldl
        for k, o in offs.items():
            setattr(saved, k, store.data_ptr() + 4 * o)
        grads = [_gbuf(p) if not (tied and i == 19) else torch.empty(0, device=p.device) for i, p in enumerate(params)]
        g = DecoderGrads()
        for name, t in zip(_PARAM_FIELDS, grads):
            setattr(g, name, t.data_ptr())
        d_h0 = torch.empty_like(h0)
        d_enc = torch.empty_like(enc)
        # its own scratch slot: the weight-gradient phase may still be reading it while later backward Functions use the main one
        ws = ops.workspace(lib.vag_decoder_seq_workspace_bytes(B, Tn, Tt, E, H, Cd, V), dev, slot="dec_bwd")
        dl = dloss_rows.contiguous()

        def run(phases):
            with on_device(dev):
                ops.check(lib.vag_decoder_seq_bwd_f32(C.byref(w), h0.data_ptr(), enc.data_ptr(), mask.data_ptr(), tok_in.data_ptr(),
                                                      tgt_t.data_ptr(), weight.data_ptr() if has_weight else None, B, Tn, Tt,
                                                      1 if tied else 0, C.byref(saved), out_mask.data_ptr() if has_out_mask else None,
                                                      dl.data_ptr(), C.byref(g), d_h0.data_ptr(), d_enc.data_ptr(), phases,
                                                      ws.data_ptr(), ws.numel(), ops.stream_ptr()))

        # (only for a fresh pass: with gradients being ACCUMULATED autograd adds into p.grad on this stream right after this node)
        if not _overlap_weight_gradients() or any(getattr(p, "grad", None) is not None for p in params):
            run(15)
        else:
            # Everything d_h0 / d_enc depend on runs on this stream; the time-batched weight-gradient contractions — ~0.4 ms of
            # tensor-core work nothing downstream reads before the optimiser — run on a second stream: the vocabulary / read-out
            # ones beside the decoder's own latency-bound recurrent part, the rest beside the back-propagation of the encoder /
            # pooling that autograd runs next.  The streams re-join when the backward pass ends (engine callback), i.e. before
            # anybody can look at a gradient; inside a CUDA-graph capture the forks and the join become parallel graph branches.
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev)
            run(1)                          # head: d logits, d read-out
            side.wait_stream(main)
            with torch.cuda.stream(side):
                run(4)                      # vocabulary / read-out weight gradients, beside the recurrent part
            run(2)                          # recurrent part → d_h0, d_enc
            side.wait_stream(main)
            with torch.cuda.stream(side):
                run(8)                      # remaining weight gradients, beside whatever autograd runs next
            for t in (store, tok_in, h0, enc, tgt_t, dl):     # freed when this node is done, read by the side stream until the join
                t.record_stream(side)
            _join_after_backward(dev, side)
        if tied:
            grads[19] = None          # out.weight IS the embedding: its gradient was accumulated into grads[0]
        return (d_h0, d_enc, None, None, None, None, None, None, *grads)


# ---------------------------------------------------------------------------------------------------- loss epilogue
class LossMixFn(torch.autograd.Function):
    """→ [3] = (loss, loss_mt, loss_vse) with loss_mt = mean_b(loss_rows/#non-pad), loss = w·mt + (1-w)·vse."""

    @staticmethod
    def forward(fctx, loss_rows, tgt, loss_vse, loss_w):
        out = ops.translation_loss(loss_rows, tgt, loss_vse, loss_w if loss_vse is not None else 1.0)
        fctx.save_for_backward(tgt)
        fctx.meta = (float(loss_w), loss_vse is not None)
        return out

    @staticmethod
    def backward(fctx, g):
        (tgt,) = fctx.saved_tensors
        loss_w, has_vse = fctx.meta
        # d loss_rows[b] = (g_loss·w + g_mt) / (B · count_b);   d loss_vse = g_loss·(1-w) + g_vse
        g_rows, g_vse = ops.translation_loss_bwd(g.to(torch.float32).contiguous(), tgt, loss_w, has_vse)
        return g_rows, None, g_vse, None
