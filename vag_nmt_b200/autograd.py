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


def _lin(x, w, b=None, flags=0, out=None):
    """y = act(x Wᵀ + b): tcgen05 split-precision kernel for eligible shapes, FP32 FFMA otherwise."""
    rows, K = x.shape
    N = w.shape[0]
    if (rows >= 64 and N >= 64 and K >= 32 and K % 8 == 0 and x.stride(0) % 4 == 0 and w.stride(0) % 4 == 0
            and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0 and x.stride(1) == 1 and w.stride(1) == 1):
        return ops.linear_tc(x, w, b, flags, out)
    return ops.linear(x, w, b, flags, out)


def _zeros(*shape, like):
    return torch.zeros(*shape, dtype=torch.float32, device=like.device)


def _empty(*shape, like):
    return torch.empty(*shape, dtype=torch.float32, device=like.device)


# ---------------------------------------------------------------------------------------------------- encoder
class EncoderFn(torch.autograd.Function):
    """src int64 [B, T] (sorted by length desc), lengths list → ctx [B, T, 2H] (sentence-major)."""

    @staticmethod
    def forward(fctx, src, lengths, emb, *gru):
        # gru = (w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r)
        B, Tn = src.shape
        E = emb.shape[1]
        H = gru[1].shape[1]
        lens = [int(x) for x in lengths]
        n_act = [sum(1 for l in lens if l > t) for t in range(Tn)]
        ids_tm = src.t().contiguous().reshape(-1)                      # time-major token ids [T·B]
        x = ops.embed_rows(emb, ids_tm)                                # [T·B, E]
        ctx = _zeros(B, Tn, 2 * H, like=emb)
        gi, gh_all = [], []
        for d in range(2):
            w_ih, w_hh, b_ih, b_hh = gru[4 * d:4 * d + 4]
            gi_d = _lin(x, w_ih, b_ih).view(Tn, B, 3 * H)
            gh_d = _zeros(Tn, B, 3 * H, like=emb)
            h = _zeros(B, H, like=emb)
            steps = range(Tn) if d == 0 else range(Tn - 1, -1, -1)
            for t in steps:
                n = n_act[t]
                if n == 0:
                    continue
                ops.linear(h[:n], w_hh, b_hh, out=gh_d[t, :n])
                ops.gru_gates(gi_d[t, :n], gh_d[t, :n], h[:n], out=h[:n], out2=ctx[:n, t, d * H:(d + 1) * H])
            gi.append(gi_d)
            gh_all.append(gh_d)
        fctx.save_for_backward(x, ctx, ids_tm, emb, gi[0], gi[1], gh_all[0], gh_all[1], *gru)
        fctx.meta = (B, Tn, E, H, n_act)
        return ctx

    @staticmethod
    def backward(fctx, dctx):
        x, ctx, ids_tm, emb, gi0, gi1, gh0, gh1, *gru = fctx.saved_tensors
        B, Tn, E, H, n_act = fctx.meta
        dctx = dctx.contiguous()
        grads: List[Optional[torch.Tensor]] = []
        dx = _zeros(Tn * B, E, like=x)
        zeros_h = _zeros(B, H, like=x)
        for d in range(2):
            w_ih, w_hh, b_ih, b_hh = gru[4 * d:4 * d + 4]
            gi_d, gh_d = (gi0, gh0) if d == 0 else (gi1, gh1)
            dgi_all = _zeros(Tn, B, 3 * H, like=x)
            dgh_all = _zeros(Tn, B, 3 * H, like=x)
            hprev_all = _zeros(Tn, B, H, like=x)
            carry = _zeros(B, H, like=x)
            steps = range(Tn - 1, -1, -1) if d == 0 else range(Tn)    # reverse of the forward order
            for t in steps:
                n = n_act[t]
                if n == 0:
                    continue
                tp = t - 1 if d == 0 else t + 1                        # where h_prev of this step was produced
                if 0 <= tp < Tn and n_act[tp] > 0:
                    # rows whose chain starts at this step (reverse direction) had h_prev = 0: ctx is 0 there already
                    hp = ctx[:n, tp, d * H:(d + 1) * H]
                else:
                    hp = zeros_h[:n]
                hprev_all[t, :n].copy_(hp)
                dh = _empty(n, H, like=x)
                dh.copy_(dctx[:n, t, d * H:(d + 1) * H])
                T.axpby_(dh, carry[:n].contiguous(), 1.0, 1.0)
                dgi, dgh, dhp = T.gru_gates_bwd(dh, gi_d[t, :n].contiguous(), gh_d[t, :n].contiguous(), hprev_all[t, :n])
                dgi_all[t, :n].copy_(dgi)
                dgh_all[t, :n].copy_(dgh)
                T.gemm(dgh, w_hh, out=dhp, beta=1.0)                   # dh_prev = dh·z + dgh·W_hh
                carry.zero_()
                carry[:n].copy_(dhp)
            dgi_f = dgi_all.view(Tn * B, 3 * H)
            dgh_f = dgh_all.view(Tn * B, 3 * H)
            T.gemm(dgi_f, w_ih, out=dx, beta=1.0)                       # dx += dgi·W_ih
            grads += [T.gemm(dgi_f, x, trans_a=True), T.gemm(dgh_f, hprev_all.view(Tn * B, H), trans_a=True),
                      T.colsum(dgi_f), T.colsum(dgh_f)]
        demb = _zeros(*emb.shape, like=x)
        T.embed_bwd_(demb, dx, ids_tm)
        return (None, None, demb, *grads)


# ---------------------------------------------------------------------------------------------------- VSE pooling
class VsePoolFn(torch.autograd.Function):
    """(im [B,I], ctx [B,T,C], mask [B,T]) → im_emb [B,S], txt_emb [B,S], ctx_vec [B,C]."""

    @staticmethod
    def forward(fctx, im, ctx, mask, method, activation, im_w, im_b, txt_w, txt_b, ctx2ctx_w, emb2ctx_w, mlp_w):
        B, Tn, C = ctx.shape
        act = ops.LIN_TANH if activation else 0
        a_im = _lin(im, im_w, im_b, act)                               # VSE_Imagine_Enc.py:123-127
        im_emb = ops.l2norm_rows_(a_im.clone())                        # :132
        iq = _lin(im_emb, emb2ctx_w)                                   # :58
        pk = _lin(ctx.view(B * Tn, C), ctx2ctx_w).view(B, Tn, C)       # :57
        mode = ops.ATTN_DOT if method == "dot" else ops.ATTN_MLP
        v = mlp_w.reshape(-1) if mode == ops.ATTN_MLP else None
        ctx_vec, beta = ops.attention(iq, pk, ctx, v, mask, 1, mode)   # :135-137
        a_txt = _lin(ctx_vec, txt_w, txt_b, act)                       # :138-140
        txt_emb = ops.l2norm_rows_(a_txt.clone())                      # :145
        fctx.save_for_backward(im, ctx, mask, a_im, im_emb, iq, pk, beta, ctx_vec, a_txt, im_w, txt_w, ctx2ctx_w, emb2ctx_w,
                               mlp_w if mlp_w is not None else im_b)
        fctx.meta = (mode, activation, mlp_w is not None)
        return im_emb, txt_emb, ctx_vec

    @staticmethod
    def backward(fctx, d_im_emb, d_txt_emb, d_ctx_vec):
        im, ctx, mask, a_im, im_emb, iq, pk, beta, ctx_vec, a_txt, im_w, txt_w, ctx2ctx_w, emb2ctx_w, mlp_w = fctx.saved_tensors
        mode, activation, has_mlp = fctx.meta
        B, Tn, C = ctx.shape
        # text branch
        du_t = T.l2norm_bwd(d_txt_emb, a_txt)
        if activation:
            du_t = T.tanh_bwd(du_t, a_txt)
        d_txt_w = T.gemm(du_t, ctx_vec, trans_a=True)
        d_txt_b = T.colsum(du_t)
        dcv = d_ctx_vec.contiguous().clone()
        T.gemm(du_t, txt_w, out=dcv, beta=1.0)
        # pooling attention
        dctx = _zeros(B, Tn, C, like=ctx)
        dpk = _zeros(B, Tn, C, like=ctx)
        dv = _zeros(C, like=ctx) if has_mlp else None
        d_iq = T.attention_bwd(dcv, beta, iq, pk, ctx, mlp_w.reshape(-1) if has_mlp else None, mask, dpk, dctx, dv, mode)
        T.gemm(dpk.view(B * Tn, C), ctx2ctx_w, out=dctx.view(B * Tn, C), beta=1.0)
        d_ctx2ctx = T.gemm(dpk.view(B * Tn, C), ctx.view(B * Tn, C), trans_a=True)
        d_emb2ctx = T.gemm(d_iq, im_emb, trans_a=True)
        # image branch
        d_ie = d_im_emb.contiguous().clone()
        T.gemm(d_iq, emb2ctx_w, out=d_ie, beta=1.0)
        du_i = T.l2norm_bwd(d_ie, a_im)
        if activation:
            du_i = T.tanh_bwd(du_i, a_im)
        d_im_w = T.gemm(du_i, im, trans_a=True)
        d_im_b = T.colsum(du_i)
        d_mlp = dv.view(1, C) if has_mlp else None
        return None, dctx, None, None, None, d_im_w, d_im_b, d_txt_w, d_txt_b, d_ctx2ctx, d_emb2ctx, d_mlp


# ---------------------------------------------------------------------------------------------------- decoder init
class DecoderInitFn(torch.autograd.Function):
    """h0 = tanh(decoderini(split·ctx_vec + (1-split)·mean_t ctx))   (ctx_vec None ⇒ text-only model)."""

    @staticmethod
    def forward(fctx, ctx_vec, ctx, mask, split, ini_w, ini_b):
        lib_mix = ops._cabi.lib()
        B, Tn, C = ctx.shape
        z = _empty(B, C, like=ctx)
        with on_device(ctx.device):
            ops.check(lib_mix.vag_init_mix_f32(z.data_ptr(), ops.ptr(ctx_vec), ctx.data_ptr(), mask.data_ptr(), float(split), B, Tn, C,
                                               ops.stream_ptr()))
        h0 = _lin(z, ini_w, ini_b, ops.LIN_TANH)
        fctx.save_for_backward(z, h0, mask, ini_w)
        fctx.meta = (float(split), ctx_vec is not None, (B, Tn, C))
        return h0

    @staticmethod
    def backward(fctx, dh0):
        z, h0, mask, ini_w = fctx.saved_tensors
        split, has_vec, (B, Tn, C) = fctx.meta
        du = T.tanh_bwd(dh0, h0)
        d_w = T.gemm(du, z, trans_a=True)
        d_b = T.colsum(du)
        dz = T.gemm(du, ini_w)
        dctx = _zeros(B, Tn, C, like=z)
        dvec = T.init_mix_bwd(dz, mask, split, dctx, has_vec)
        return dvec, dctx, None, None, d_w, d_b


# ---------------------------------------------------------------------------------------------------- decoder loop
class DecoderSeqFn(torch.autograd.Function):
    """(h0 [B,H], ctx [B,T,C], mask [B,T], tgt [B,Tt]) → Σ_t NLL rows [B]   (V11:136-160, NMT_Decoder.py:109-145)."""

    @staticmethod
    def forward(fctx, h0, enc, mask, tgt, weight, teacher, tied, emb, g1_wih, g1_whh, g1_bih, g1_bhh, attn_h, attn_e, v, c2h,
                g2_wih, g2_whh, g2_bih, g2_bhh, w1, b1, w2, b2, w3, b3, out_w, out_b):
        B, Tt = tgt.shape
        _, Tn, C = enc.shape
        H, E, V = h0.shape[1], emb.shape[1], out_w.shape[0]
        dev = h0.device
        tgt_t = tgt.t().contiguous()                                   # [Tt, B]
        tok_in = torch.empty(Tt, B, dtype=torch.int64, device=dev)
        tok_in[0].fill_(SOS_token)
        keys = _lin(enc.view(B * Tn, C), attn_e).view(B, Tn, C)        # hoisted attn_e(ctx), NMT_Decoder.py:47
        gh1_all, gi2_all, gh2_all = (_empty(Tt, B, 3 * H, like=h0) for _ in range(3))
        h1_all, x2_all, h2_all = (_empty(Tt, B, H, like=h0) for _ in range(3))
        q_all, c_all = _empty(Tt, B, C, like=h0), _empty(Tt, B, C, like=h0)
        alpha_all = _empty(Tt, B, Tn, like=h0)
        loss_rows = _zeros(B, like=h0)
        lse_all = _empty(Tt, B, like=h0)
        ldl = (V + 3) // 4 * 4
        logits_all = _empty(Tt * B, ldl, like=h0)[:, :V]
        if teacher:
            tok_in[1:].copy_(tgt_t[:-1])
            e_all = ops.embed_rows(emb, tok_in.reshape(-1))            # [Tt·B, E]
            gi1_all = _lin(e_all, g1_wih, g1_bih).view(Tt, B, 3 * H)
        else:
            e_all = _empty(Tt * B, E, like=h0)
            gi1_all = _empty(Tt, B, 3 * H, like=h0)
            t_all = _empty(Tt * B, E, like=h0)
        h = h0
        for s in range(Tt):
            if not teacher:
                e_s = e_all[s * B:(s + 1) * B]
                e_s.copy_(ops.embed_rows(emb, tok_in[s]))
                ops.linear(e_s, g1_wih, g1_bih, out=gi1_all[s])
            ops.linear(h, g1_whh, g1_bhh, out=gh1_all[s])
            ops.gru_gates(gi1_all[s], gh1_all[s], h, out=h1_all[s])
            ops.linear(h1_all[s], attn_h, out=q_all[s])
            ops.attention(q_all[s], keys, enc, v, mask, 1, ops.ATTN_MLP, out_c=c_all[s], out_alpha=alpha_all[s])
            ops.linear(c_all[s], c2h, out=x2_all[s])
            ops.linear(x2_all[s], g2_wih, g2_bih, out=gi2_all[s])
            ops.linear(h1_all[s], g2_whh, g2_bhh, out=gh2_all[s])
            ops.gru_gates(gi2_all[s], gh2_all[s], h1_all[s], out=h2_all[s])
            h = h2_all[s]
            if not teacher:                                             # free running: the next input is this step's argmax
                t_s = t_all[s * B:(s + 1) * B]
                ops.linear(h, w1, b1, out=t_s)
                ops.linear(e_all[s * B:(s + 1) * B], w3, b3, flags=ops.LIN_ACCUMULATE, out=t_s)
                ops.linear(c_all[s], w2, b2, flags=ops.LIN_ACCUMULATE | ops.LIN_TANH, out=t_s)
                lg = logits_all[s * B:(s + 1) * B]
                ops.linear(t_s, out_w, out_b, out=lg)
                if s + 1 < Tt:
                    tok_in[s + 1].copy_(ops.row_argmax(lg))
        if teacher:                                                     # read-out + vocabulary projection for all steps at once
            t_all = _lin(h2_all.view(Tt * B, H), w1, b1)
            _lin(e_all, w3, b3, flags=ops.LIN_ACCUMULATE, out=t_all)
            _lin(c_all.view(Tt * B, C), w2, b2, flags=ops.LIN_ACCUMULATE | ops.LIN_TANH, out=t_all)
            _lin(t_all, out_w, out_b, out=logits_all)
        for s in range(Tt):
            ops.nll_rows(logits_all[s * B:(s + 1) * B], tgt_t[s], weight, loss_rows, lse_all[s])
        fctx.save_for_backward(h0, enc, mask, tgt_t, tok_in, keys, e_all, gi1_all, gh1_all, h1_all, q_all, alpha_all, c_all, x2_all,
                               gi2_all, gh2_all, h2_all, t_all, logits_all, lse_all, emb, g1_wih, g1_whh, attn_h, attn_e, v, c2h,
                               g2_wih, g2_whh, w1, w2, w3, out_w, weight if weight is not None else lse_all)
        fctx.meta = (weight is not None, tied)
        return loss_rows

    @staticmethod
    def backward(fctx, dloss_rows):
        (h0, enc, mask, tgt_t, tok_in, keys, e_all, gi1_all, gh1_all, h1_all, q_all, alpha_all, c_all, x2_all, gi2_all, gh2_all,
         h2_all, t_all, logits_all, lse_all, emb, g1_wih, g1_whh, attn_h, attn_e, v, c2h, g2_wih, g2_whh, w1, w2, w3, out_w,
         weight) = fctx.saved_tensors
        has_weight, tied = fctx.meta
        weight = weight if has_weight else None
        Tt, B = tgt_t.shape
        _, Tn, C = enc.shape
        H, E, V = h0.shape[1], emb.shape[1], out_w.shape[0]
        g_rows = dloss_rows.contiguous()
        # ---- batched over all steps: vocabulary projection and read-out
        dlogits = _empty(Tt * B, V, like=h0)
        for s in range(Tt):
            dlogits[s * B:(s + 1) * B].copy_(T.nll_bwd(logits_all[s * B:(s + 1) * B], lse_all[s], tgt_t[s], weight, g_rows))
        d_t = T.gemm(dlogits, out_w)                                    # [Tt·B, E]
        d_out_w = T.gemm(dlogits, t_all, trans_a=True)
        d_out_b = T.colsum(dlogits)
        du = T.tanh_bwd(d_t, t_all)
        h2_f, c_f = h2_all.view(Tt * B, H), c_all.view(Tt * B, C)
        d_h2_dir = T.gemm(du, w1).view(Tt, B, H)
        d_e = T.gemm(du, w3)                                            # [Tt·B, E]
        d_c_dir = T.gemm(du, w2).view(Tt, B, C)
        d_w1, d_w3, d_w2 = T.gemm(du, h2_f, trans_a=True), T.gemm(du, e_all, trans_a=True), T.gemm(du, c_f, trans_a=True)
        d_b = T.colsum(du)                                              # b1, b2, b3 all receive Σ du
        # ---- recurrent part, reverse time
        dgi2_all, dgh2_all, dgi1_all, dgh1_all = (_empty(Tt, B, 3 * H, like=h0) for _ in range(4))
        dx2_all, dq_all = _empty(Tt, B, H, like=h0), _empty(Tt, B, C, like=h0)
        dkeys, dctx, dv = _zeros(B, Tn, C, like=h0), _zeros(B, Tn, C, like=h0), _zeros(C, like=h0)
        dh_next = _zeros(B, H, like=h0)
        for s in range(Tt - 1, -1, -1):
            dh2 = d_h2_dir[s]
            T.axpby_(dh2, dh_next, 1.0, 1.0)
            dgi2, dgh2, dh1 = T.gru_gates_bwd(dh2, gi2_all[s], gh2_all[s], h1_all[s])
            dgi2_all[s].copy_(dgi2)
            dgh2_all[s].copy_(dgh2)
            T.gemm(dgi2, g2_wih, out=dx2_all[s])                         # dx2 = dgi2·W_ih2
            dc = d_c_dir[s]
            T.gemm(dx2_all[s], c2h, out=dc, beta=1.0)                    # dc += dx2·W_c2h
            dq = T.attention_bwd(dc, alpha_all[s], q_all[s], keys, enc, v, mask, dkeys, dctx, dv, ops.ATTN_MLP)
            dq_all[s].copy_(dq)
            T.gemm(dgh2, g2_whh, out=dh1, beta=1.0)                      # dh1 += dgh2·W_hh2 + dq·W_attn_h
            T.gemm(dq, attn_h, out=dh1, beta=1.0)
            h_prev = h0 if s == 0 else h2_all[s - 1]
            dgi1, dgh1, dhp = T.gru_gates_bwd(dh1, gi1_all[s], gh1_all[s], h_prev)
            dgi1_all[s].copy_(dgi1)
            dgh1_all[s].copy_(dgh1)
            T.gemm(dgh1, g1_whh, out=dhp, beta=1.0)                      # dh_prev = dh1·z + dgh1·W_hh1
            dh_next = dhp
        d_h0 = dh_next
        # ---- weight gradients, batched over steps
        dgi2_f, dgh2_f = dgi2_all.view(Tt * B, 3 * H), dgh2_all.view(Tt * B, 3 * H)
        dgi1_f, dgh1_f = dgi1_all.view(Tt * B, 3 * H), dgh1_all.view(Tt * B, 3 * H)
        h1_f = h1_all.view(Tt * B, H)
        d_g2_wih, d_g2_bih = T.gemm(dgi2_f, x2_all.view(Tt * B, H), trans_a=True), T.colsum(dgi2_f)
        d_g2_whh, d_g2_bhh = T.gemm(dgh2_f, h1_f, trans_a=True), T.colsum(dgh2_f)
        d_c2h = T.gemm(dx2_all.view(Tt * B, H), c_f, trans_a=True)
        d_attn_h = T.gemm(dq_all.view(Tt * B, C), h1_f, trans_a=True)
        d_g1_wih, d_g1_bih = T.gemm(dgi1_f, e_all, trans_a=True), T.colsum(dgi1_f)
        d_g1_whh = T.gemm(dgh1_all[0], h0, trans_a=True)
        if Tt > 1:
            T.gemm(dgh1_all[1:].reshape((Tt - 1) * B, 3 * H), h2_all[:-1].reshape((Tt - 1) * B, H), trans_a=True, out=d_g1_whh, beta=1.0)
        d_g1_bhh = T.colsum(dgh1_f)
        T.gemm(dgi1_f, g1_wih, out=d_e, beta=1.0)                        # de += dgi1·W_ih1
        d_emb = _zeros(*emb.shape, like=h0)
        T.embed_bwd_(d_emb, d_e, tok_in.reshape(-1))
        if tied:                                                         # out.weight IS the embedding (NMT_Decoder.py:105-106)
            T.axpby_(d_emb, d_out_w, 1.0, 1.0)
            d_out_w = None
        # ---- hoisted keys: dW_attn_e and the path back into the encoder context
        dk_f, enc_f = dkeys.view(B * Tn, C), enc.view(B * Tn, C)
        d_attn_e = T.gemm(dk_f, enc_f, trans_a=True)
        T.gemm(dk_f, attn_e, out=dctx.view(B * Tn, C), beta=1.0)
        return (d_h0, dctx, None, None, None, None, None, d_emb, d_g1_wih, d_g1_whh, d_g1_bih, d_g1_bhh, d_attn_h, d_attn_e, dv,
                d_c2h, d_g2_wih, d_g2_whh, d_g2_bih, d_g2_bhh, d_w1, d_b, d_w2, d_b.clone(), d_w3, d_b.clone(), d_out_w, d_out_b)


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
        B = tgt.shape[0]
        # d loss_rows[b] = (g_loss·w + g_mt) / (B · count_b);   d loss_vse = g_loss·(1-w) + g_vse
        counts = (tgt != 0).sum(-1).to(torch.float32)
        g = g.to(torch.float32)
        w_eff = loss_w if has_vse else 1.0
        g_rows = (g[0] * w_eff + g[1]) / (B * counts)
        g_vse = (g[0] * (1.0 - loss_w) + g[2]).reshape(1) if has_vse else None
        return g_rows, None, g_vse, None
