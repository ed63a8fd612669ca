"""CPU oracle for the VAG-NMT per-timestep translation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vag_nmt_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or as the timed CPU baseline.

What it is: a functional, CPU-only restatement (torch CPU tensor arithmetic,
fp32 or fp64) of the algorithm that the reference implements with
``nn.GRU``/``nn.Linear``/``bmm``/``topk``.  Every function cites the reference
file:line (relative to the upstream repository root) it follows.  Parameters
are passed as a plain ``dict`` keyed by the reference's ``state_dict`` names
(SURVEY.md section 8b), so a reference checkpoint feeds the oracle unchanged.

Third-party arithmetic: the reference pins ``torch==0.4.1``
(requirements.txt:162) which is not vendored.  The GRU cell equations restated
here are the published PyTorch ones (gate order r, z, n).

Pinning status: the reference ships no golden vectors or known-answer tests
for this path.  The oracle is pinned instead against
the reference's own code executed in the build container
(``oracle/make_golden.py`` imports ``/root/reference`` with the three-pattern
torch-2.x compat patch of SURVEY.md appendix B and writes
``tests/golden/*.pt``); ``tests/test_oracle_golden.py`` re-checks the oracle
against those committed fixtures on every run.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

PAD_token = 0
UNK_token = 1
SOS_token = 2
EOS_token = 3

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
_OPERAND_ROUNDING = None


def set_operand_rounding(kind: Optional[str]) -> None:
    """None: plain arithmetic of the dtype.  "bf16": every contraction (nn.Linear / GRU matrix) sees its two operands
    rounded to bfloat16 and accumulates in the working dtype — the reference run the bf16 mode is compared with
    (SURVEY.md section 7 hard part 2: a reference that rounds at the same points)."""
    global _OPERAND_ROUNDING
    assert kind in (None, "bf16")
    _OPERAND_ROUNDING = kind


def _round_operand(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype) if _OPERAND_ROUNDING == "bf16" else t


class _RoundedLinearFn(torch.autograd.Function):
    """y = r(x)·r(W)ᵀ with r = round-to-bfloat16, for the bf16-mode comparison runs.  The backward rounds the operands of ITS
    two contractions the same way (dx = r(dy)·r(W), dW = r(dy)ᵀ·r(x)) and accumulates in the working dtype — the points at
    which the B200 path rounds in bf16 mode ("every contraction, forward and backward, rounds its operands").  Plain autograd
    through ``x.to(bf16).to(dtype)`` would instead round the RESULTS dx / dW, which no implementation does."""

    @staticmethod
    def forward(ctx, x, w):
        xr, wr = _round_operand(x), _round_operand(w)
        ctx.save_for_backward(xr, wr)
        return xr.matmul(wr.t())

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        dyr = _round_operand(dy)
        dx = dyr.matmul(wr)
        dw = dyr.reshape(-1, dyr.shape[-1]).t().matmul(xr.reshape(-1, xr.shape[-1]))
        return dx, dw


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x Wᵀ (+ b) — what every nn.Linear on the path computes."""
    if _OPERAND_ROUNDING == "bf16" and (x.requires_grad or w.requires_grad):
        y = _RoundedLinearFn.apply(x, w)
    else:
        y = _round_operand(x).matmul(_round_operand(w).t())
    if b is not None:
        y = y + b
    return y


def l2norm(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """Row-wise L2 normalisation.  utils/utils.py:6-10."""
    return x / x.norm(2, 1).clamp(min=eps).unsqueeze(1)


def gru_cell(x: torch.Tensor, h: torch.Tensor, w_ih, w_hh, b_ih, b_hh) -> torch.Tensor:
    """One PyTorch GRU cell, gate order (r, z, n).

    This is the arithmetic behind ``nn.GRU`` as used at layers/Encoder.py:34,58
    and layers/NMT_Decoder.py:83,86,121,129.
    """
    H = h.shape[-1]
    gi = linear(x, w_ih, b_ih)
    gh = linear(h, w_hh, b_hh)
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


# --------------------------------------------------------------------------
# a1  encoder
# --------------------------------------------------------------------------
def encoder_forward(p: Params, src: torch.Tensor, lengths: Sequence[int], prefix: str = "encoder.",
                    emb_mask: Optional[torch.Tensor] = None, ctx_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """LIUMCVC_Encoder.forward in eval mode (no dropout).  layers/Encoder.py:36-65.

    src      int64 [B, T], 0-padded, rows sorted by length descending
    lengths  python list, lengths[b] real tokens of row b
    returns  ctx [T, B, 2H] (forward half ‖ backward half, exact zeros at
             padded positions, as pad_packed_sequence produces) and the float
             mask [T, B] built from ``src != 0`` (Encoder.py:47).
    """
    B, T = src.shape
    lengths = [int(x) for x in lengths]
    if any(lengths[i] < lengths[i + 1] for i in range(B - 1)):
        # pack_padded_sequence(enforce_sorted=True) raises (Encoder.py:55)
        raise RuntimeError("encoder_forward: lengths must be sorted in decreasing order")
    T = max(lengths)
    emb = p[prefix + "embedding.weight"]
    dt = emb.dtype
    x = emb[src[:, :T]].transpose(0, 1)  # [T, B, E]   Encoder.py:50
    if emb_mask is not None:  # training-mode embedding dropout with a given mask (0 or 1/(1-p)), Encoder.py:51-52
        x = x * emb_mask.reshape(T, B, -1)
    H = p[prefix + "gru.weight_hh_l0"].shape[1]
    ctx = torch.zeros(T, B, 2 * H, dtype=dt)
    lens = torch.tensor(lengths)
    for direction, sfx in ((0, ""), (1, "_reverse")):
        w_ih = p[prefix + "gru.weight_ih_l0" + sfx]
        w_hh = p[prefix + "gru.weight_hh_l0" + sfx]
        b_ih = p[prefix + "gru.bias_ih_l0" + sfx]
        b_hh = p[prefix + "gru.bias_hh_l0" + sfx]
        h = torch.zeros(B, H, dtype=dt)
        steps = range(T) if direction == 0 else range(T - 1, -1, -1)
        for t in steps:
            n_act = int((lens > t).sum())  # sorted desc ⇒ active rows are a prefix
            if n_act == 0:
                continue
            h_new = gru_cell(x[t, :n_act], h[:n_act], w_ih, w_hh, b_ih, b_hh)
            h = torch.cat([h_new, h[n_act:]], 0)
            ctx[t, :n_act, direction * H:(direction + 1) * H] = h_new
    if ctx_mask is not None:  # context dropout, Encoder.py:62-63; ctx_mask is sentence-major [B, T, 2H]
        ctx = ctx * ctx_mask.transpose(0, 1)
    mask = (src[:, :T] != 0).long().transpose(0, 1).to(dt)  # Encoder.py:47,65
    return ctx, mask


# --------------------------------------------------------------------------
# a2/a3  Bahdanau attention + conditional-GRU decoder step
# --------------------------------------------------------------------------
def bahdanau_attention(p: Params, h1: torch.Tensor, ctx: torch.Tensor, mask: Optional[torch.Tensor],
                       keys: Optional[torch.Tensor] = None, prefix: str = "decoder.attn.") -> torch.Tensor:
    """BahdanauAttn.forward.  layers/NMT_Decoder.py:27-51.

    h1 [N, H]; ctx [T, N, C]; mask [T, N] float or None → α [N, T].
    ``keys`` may carry a pre-computed attn_e(ctx) ([N, T, C]); the reference
    recomputes it at every step (NMT_Decoder.py:47).
    """
    enc = ctx.transpose(0, 1)  # [N, T, C]
    if keys is None:
        keys = linear(enc, p[prefix + "attn_e.weight"])
    q = linear(h1, p[prefix + "attn_h.weight"]).unsqueeze(1)  # [N, 1, C]
    energy = torch.tanh(q + keys)  # NMT_Decoder.py:47
    scores = energy.matmul(p[prefix + "v"])  # [N, T]     NMT_Decoder.py:49-51
    if mask is not None:
        scores = scores.masked_fill(mask.transpose(0, 1) == 0, -float("inf"))  # :41-43
    return torch.softmax(scores, dim=1)


def decoder_step(p: Params, tok: torch.Tensor, h: torch.Tensor, ctx: torch.Tensor,
                 mask: Optional[torch.Tensor], keys: Optional[torch.Tensor] = None,
                 prefix: str = "decoder.", return_parts: bool = False, out_mask: Optional[torch.Tensor] = None):
    """NMT_Decoder.forward in eval mode.  layers/NMT_Decoder.py:109-145.

    tok int64 [N]; h [N, H]; ctx [T, N, C]; mask [T, N] → logp [N, V], h2 [N, H].
    """
    tok = tok.reshape(-1)
    emb_w = p[prefix + "embedding.weight"]
    e = emb_w[tok]  # :118
    h1 = gru_cell(e, h, p[prefix + "gru_1.weight_ih_l0"], p[prefix + "gru_1.weight_hh_l0"],
                  p[prefix + "gru_1.bias_ih_l0"], p[prefix + "gru_1.bias_hh_l0"])  # :121
    alpha = bahdanau_attention(p, h1, ctx, mask, keys, prefix + "attn.")  # :124
    c = alpha.unsqueeze(1).bmm(ctx.transpose(0, 1)).squeeze(1)  # :126
    x2 = linear(c, p[prefix + "context2hid.weight"])  # :127
    h2 = gru_cell(x2, h1, p[prefix + "gru_2.weight_ih_l0"], p[prefix + "gru_2.weight_hh_l0"],
                  p[prefix + "gru_2.bias_ih_l0"], p[prefix + "gru_2.bias_hh_l0"])  # :129
    t = torch.tanh(linear(h2, p[prefix + "W1.weight"], p[prefix + "W1.bias"])
                   + linear(e, p[prefix + "W3.weight"], p[prefix + "W3.bias"])
                   + linear(c, p[prefix + "W2.weight"], p[prefix + "W2.bias"]))  # :137
    if out_mask is not None:  # output dropout with a given mask, NMT_Decoder.py:140-141
        t = t * out_mask
    out_w = p.get(prefix + "out.weight", emb_w)  # tied: NMT_Decoder.py:105-106
    logits = linear(t, out_w, p[prefix + "out.bias"])
    logp = torch.log_softmax(logits, dim=-1)  # :143
    if return_parts:
        return logp, h2, dict(e=e, h1=h1, alpha=alpha, c=c, x2=x2, t=t, logits=logits)
    return logp, h2


# --------------------------------------------------------------------------
# a4/a5/a6  visual-attention text pooling
# --------------------------------------------------------------------------
def imagine_attention(p: Params, im_emb: torch.Tensor, ctx: torch.Tensor, mask: Optional[torch.Tensor],
                      method: str = "dot", prefix: str = "vse_imagine.imagine_attn.") -> torch.Tensor:
    """ImagineAttn.forward → β [B, T].  layers/VSE_Imagine_Enc.py:29-79."""
    enc = ctx.transpose(0, 1)  # [B, T, C]
    pk = linear(enc, p[prefix + "ctx2ctx.weight"])  # :57 / :76
    iq = linear(im_emb, p[prefix + "emb2ctx.weight"])  # :58 / :77   [B, C]
    if method == "dot":
        energies = pk.matmul(iq.unsqueeze(2)).squeeze(2)  # :64  [B, T]
    elif method == "mlp":
        energies = linear(torch.tanh(pk + iq.unsqueeze(1)), p[prefix + "mlp.weight"]).squeeze(2)  # :79
    else:
        raise ValueError(method)
    if mask is not None:
        energies = energies.masked_fill(mask.transpose(0, 1) == 0, -float("inf"))  # :42-44
    return torch.softmax(energies, dim=-1)


def vse_pool(p: Params, im: torch.Tensor, ctx: torch.Tensor, mask: Optional[torch.Tensor],
             method: str = "dot", activation: bool = True, prefix: str = "vse_imagine."):
    """VSE_Imagine_Enc.forward / get_emb_vec without the loss.  layers/VSE_Imagine_Enc.py:110-172.

    returns im_emb [B, S], txt_emb [B, S], ctx_vec [B, C], β [B, T]
    """
    im_emb = linear(im, p[prefix + "im_embedding.weight"], p[prefix + "im_embedding.bias"])  # :123
    if activation:
        im_emb = torch.tanh(im_emb)
    im_emb = l2norm(im_emb)  # :132
    beta = imagine_attention(p, im_emb, ctx, mask, method, prefix + "imagine_attn.")  # :135
    ctx_vec = beta.unsqueeze(1).bmm(ctx.transpose(0, 1)).squeeze(1)  # :137
    txt = linear(ctx_vec, p[prefix + "text_embedding.weight"], p[prefix + "text_embedding.bias"])  # :138
    if activation:
        txt = torch.tanh(txt)
    txt = l2norm(txt)  # :145
    return im_emb, txt, ctx_vec, beta


# --------------------------------------------------------------------------
# a7  ranking losses
# --------------------------------------------------------------------------
def pairwise_ranking_loss(im: torch.Tensor, s: torch.Tensor, margin: float) -> torch.Tensor:
    """PairwiseRankingLoss.forward.  losses/PairwiseRankingLoss.py:9-24."""
    scores = im.matmul(s.t())
    d = scores.diag()
    cost_s = (margin - d).unsqueeze(0).expand_as(scores) + scores  # column j uses d_j   :16
    cost_im = (margin - d).unsqueeze(1).expand_as(scores) + scores  # row i uses d_i      :18
    cost_s = cost_s.clamp(min=0)
    cost_im = cost_im.clamp(min=0)
    eye = torch.eye(scores.shape[0], dtype=torch.bool)
    cost_s = cost_s.masked_fill(eye, 0)  # :20-22
    cost_im = cost_im.masked_fill(eye, 0)
    return cost_s.sum() + cost_im.sum()


def image_retrieval_ranking_loss(im: torch.Tensor, s: torch.Tensor, margin: float) -> torch.Tensor:
    """ImageRetrievalRankingLoss.forward (cost_s only).  losses/ImageRetrievalRankingLoss.py:9-21."""
    scores = im.matmul(s.t())
    d = scores.diag()
    cost_s = ((margin - d).unsqueeze(0).expand_as(scores) + scores).clamp(min=0)
    cost_s = cost_s.masked_fill(torch.eye(scores.shape[0], dtype=torch.bool), 0)
    return cost_s.sum()


# --------------------------------------------------------------------------
# a8/a11  training forward (loss)
# --------------------------------------------------------------------------
def decoder_init(p: Params, ctx: torch.Tensor, mask: torch.Tensor, ctx_vec: Optional[torch.Tensor],
                 init_split: float) -> torch.Tensor:
    """h0 = tanh(decoderini(...)).  V11:118,201 ; NMT_Seq2Seq_Beam_V2.py:85,142."""
    mean_ctx = ctx.sum(0) / mask.sum(0).unsqueeze(1)
    if ctx_vec is None:
        z = mean_ctx
    else:
        z = init_split * ctx_vec + (1 - init_split) * mean_ctx
    return torch.tanh(linear(z, p["decoderini.weight"], p["decoderini.bias"]))


def _nll_rows(logp: torch.Tensor, tgt: torch.Tensor, weight: Optional[torch.Tensor]) -> torch.Tensor:
    """nn.NLLLoss(weight, reduce=False)(logp, tgt) → [B].  nmt_multimodal_beam_DE.py:286-291."""
    picked = -logp.gather(1, tgt.unsqueeze(1)).squeeze(1)
    if weight is not None:
        picked = picked * weight[tgt]
    return picked


def translation_loss(p: Params, ctx, mask, h0, tgt: torch.Tensor, teacher_force: bool,
                     nll_weight: Optional[torch.Tensor], out_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The Tt-step loop of forward.  V11:136-164 ; NMT_Seq2Seq_Beam_V2.py:91-111."""
    B, Tt = tgt.shape
    inp = torch.full((B,), SOS_token, dtype=torch.long)
    h = h0
    loss_rows = torch.zeros(B, dtype=h0.dtype)
    for di in range(Tt):
        logp, h = decoder_step(p, inp, h, ctx, mask, out_mask=None if out_mask is None else out_mask.reshape(Tt, B, -1)[di])
        loss_rows = loss_rows + _nll_rows(logp, tgt[:, di], nll_weight)
        inp = tgt[:, di] if teacher_force else logp.argmax(1)
    tgt_mask = (tgt != 0).to(h0.dtype)
    return (loss_rows / tgt_mask.sum(-1)).mean()  # V11:164


def multimodal_forward(p: Params, src, lengths, tgt, im, teacher_force: bool = True,
                       nll_weight: Optional[torch.Tensor] = None, vse_loss: Optional[str] = "pairwise",
                       margin: float = 0.1, loss_w: float = 0.99, init_split: float = 0.5,
                       attn_model: str = "dot", activation_vse: bool = True, dropout_masks: Optional[dict] = None):
    """NMT_AttentionImagine_Seq2Seq_Beam_V11.forward.  V11:82-168.  dropout_masks (optional) = {"emb": [T·B, E] time-major,
    "ctx": [B, T, 2H], "out": [Tt·B, E]} with entries 0 or 1/(1-p): the training-mode dropouts with given masks."""
    dm = dropout_masks or {}
    ctx, mask = encoder_forward(p, src, lengths, emb_mask=dm.get("emb"), ctx_mask=dm.get("ctx"))  # :111
    im_emb, txt_emb, ctx_vec, _ = vse_pool(p, im, ctx, mask, attn_model, activation_vse)  # :114
    if vse_loss == "pairwise":
        loss_vse = pairwise_ranking_loss(im_emb, txt_emb, margin)
    elif vse_loss == "imageretrieval":
        loss_vse = image_retrieval_ranking_loss(im_emb, txt_emb, margin)
    else:
        loss_vse = torch.zeros((), dtype=ctx.dtype)
    h0 = decoder_init(p, ctx, mask, ctx_vec, init_split)  # :118
    loss_mt = translation_loss(p, ctx, mask, h0, tgt, teacher_force, nll_weight, out_mask=dm.get("out"))
    loss = loss_w * loss_mt + (1 - loss_w) * loss_vse  # :166
    return loss, loss_mt, loss_vse


def text_forward(p: Params, src, lengths, tgt, teacher_force: bool = True,
                 nll_weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NMT_Seq2Seq_Beam_V2.forward.  models/NMT_Seq2Seq_Beam_V2.py:58-113."""
    ctx, mask = encoder_forward(p, src, lengths)
    h0 = decoder_init(p, ctx, mask, None, 0.0)
    return translation_loss(p, ctx, mask, h0, tgt, teacher_force, nll_weight)


# --------------------------------------------------------------------------
# a9  beam search
# --------------------------------------------------------------------------
def topk_canonical(x: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """topk(k, sorted=False) with a DEFINED order: value descending, ties by lowest index.

    The reference's ``sorted=False`` order is implementation-defined;
    the final hypotheses do not depend on it.  Both the oracle and the CUDA path use
    this canonical order so that intermediate beams can be compared.
    """
    v, i = torch.sort(x, dim=-1, descending=True, stable=True)
    return v[..., :k].contiguous(), i[..., :k].contiguous()


def beamsearch(step_fn, n_vocab: int, B: int, h0: torch.Tensor, beam_size: int, max_length: int,
               avoid_double: bool = True, return_state: bool = False):
    """The batched tensorised beam search.  V11:233-337 (== NMT_Seq2Seq_Beam_V2.py:173-276).

    ``step_fn(tokens[N], hidden[N,H], tile[N] or None) -> (logp[N,V], hidden[N,H])``
    runs one decoder step; ``tile`` maps each of the N rows to its sentence
    (None at step 0 where N == B).
    """
    K = beam_size
    nk = torch.arange(B * K)
    pdxs_mask = (nk // K) * K  # :242
    tile = nk // K  # :245
    sent_of_row = nk // K
    beam = torch.zeros(max_length, B, K, dtype=torch.long)  # :248
    inf = -1e5  # :257
    h = h0
    nll = None
    steps_run = 0
    for di in range(max_length):
        if di == 0:
            logp, h = step_fn(torch.full((B,), SOS_token, dtype=torch.long), h, None)  # :261
            nll, topi = topk_canonical(logp, K)  # :262
            beam[0] = topi
        else:
            cur = beam[di - 1].reshape(-1)  # :265
            fini = (cur == EOS_token).nonzero()  # :266
            n_fini = fini.numel()
            if n_fini == B * K:  # :268
                break
            h = h[tile]  # :273
            logp, h = step_fn(cur, h, sent_of_row)  # :275
            logp = logp.clone()
            if avoid_double:
                logp.view(-1).index_fill_(0, cur + nk * n_vocab, inf)  # :280
            if n_fini > 0:
                fidx = fini[:, 0]
                logp.index_fill_(0, fidx, inf)  # :293
                logp.view(-1).index_fill_(0, fidx * n_vocab + EOS_token, 0)  # :294
            cand = (nll.unsqueeze(2) + logp.view(B, K, n_vocab)).view(B, -1)  # :297
            nll, idxs = topk_canonical(cand, K)  # :300
            pdxs = idxs // n_vocab  # :303
            beam[di] = idxs % n_vocab  # :306
            beam[:di] = beam[:di].gather(2, pdxs.unsqueeze(0).expand(di, B, K))  # :309
            tile = pdxs.reshape(-1) + pdxs_mask  # :313
        steps_run = di + 1
    beam[max_length - 1] = EOS_token  # :315
    lens = (beam.transpose(0, 2) > 3).sum(-1).t().to(nll.dtype).clamp(min=1)  # :318
    nll_norm = nll / lens  # :321
    best = topk_canonical(nll_norm, 1)[1].squeeze(1)  # :322
    hyps = beam[:, torch.arange(B), best].cpu().numpy().T  # :324 (.cpu(): a no-op here; lets the same code run on CUDA tensors for bench.py's reference-eager datum)
    final = []
    for b in range(B):
        cur_list = []
        for i in range(max_length):
            tok = int(hyps[b][i])
            if tok == EOS_token:
                break
            cur_list.append(tok)
        final.append(cur_list)
    if return_state:
        return final, dict(beam=beam, nll=nll, nll_norm=nll_norm, best=best, steps_run=steps_run)
    return final


def _greedy(p: Params, ctx, mask, h0, max_length: int) -> List[List[int]]:
    """beam_size == 1 branch.  V11:207-226."""
    B = ctx.shape[1]
    inp = torch.full((B,), SOS_token, dtype=torch.long)
    h = h0
    toks = []
    for _ in range(max_length):
        logp, h = decoder_step(p, inp, h, ctx, mask)
        inp = logp.argmax(1)
        toks.append(inp)
    toks = torch.stack(toks, 1).tolist()
    out = []
    for b in range(B):
        row = []
        for t in toks[b]:
            if t == EOS_token:
                break
            row.append(t)
        out.append(row)
    return out


def _decode_from_encoder(p: Params, ctx, mask, h0, beam_size: int, max_length: int, hoist_keys: bool,
                         return_state: bool = False):
    B = ctx.shape[1]
    if beam_size == 1:
        return _greedy(p, ctx, mask, h0, max_length)
    n_vocab = p["decoder.embedding.weight"].shape[0]
    keys_b = linear(ctx.transpose(0, 1), p["decoder.attn.attn_e.weight"]) if hoist_keys else None

    def step_fn(tok, h, sent_of_row):
        if sent_of_row is None:
            return decoder_step(p, tok, h, ctx, mask, keys_b)
        # the reference materialises the K-times tiled context (V11:253-254)
        return decoder_step(p, tok, h, ctx[:, sent_of_row], mask[:, sent_of_row],
                            None if keys_b is None else keys_b[sent_of_row])

    return beamsearch(step_fn, n_vocab, B, h0, beam_size, max_length, return_state=return_state)


def multimodal_beamsearch_decode(p: Params, src, lengths, im, beam_size: int = 1, max_length: int = 80,
                                 init_split: float = 0.5, attn_model: str = "dot", activation_vse: bool = True,
                                 hoist_keys: bool = False, return_state: bool = False):
    """NMT_AttentionImagine_Seq2Seq_Beam_V11.beamsearch_decode.  V11:179-231."""
    ctx, mask = encoder_forward(p, src, lengths)  # :193
    _, _, ctx_vec, _ = vse_pool(p, im, ctx, mask, attn_model, activation_vse)  # :196
    h0 = decoder_init(p, ctx, mask, ctx_vec, init_split)  # :201
    return _decode_from_encoder(p, ctx, mask, h0, beam_size, max_length, hoist_keys, return_state)


def text_beamsearch_decode(p: Params, src, lengths, beam_size: int = 1, max_length: int = 80,
                           hoist_keys: bool = False, return_state: bool = False):
    """NMT_Seq2Seq_Beam_V2.beamsearch_decode.  models/NMT_Seq2Seq_Beam_V2.py:124-171."""
    ctx, mask = encoder_forward(p, src, lengths)
    h0 = decoder_init(p, ctx, mask, None, 0.0)
    return _decode_from_encoder(p, ctx, mask, h0, beam_size, max_length, hoist_keys, return_state)


def embed_sent_im(p: Params, src, lengths, im, attn_model: str = "dot", activation_vse: bool = True):
    """V11.embed_sent_im_test / embed_sent_im_eval.  V11:341-397."""
    ctx, mask = encoder_forward(p, src, lengths)
    im_emb, txt_emb, _, _ = vse_pool(p, im, ctx, mask, attn_model, activation_vse)
    return im_emb, txt_emb


# --------------------------------------------------------------------------
# a12  retrieval metrics
# --------------------------------------------------------------------------
def retrieval_ranks(queries: torch.Tensor, gallery: torch.Tensor) -> np.ndarray:
    """Rank of gallery[i] for query i under a descending STABLE sort of the scores.

    utils/im_retrieval_eval.py:15-22 uses ``torch.sort(descending=True)`` whose tie
    order is unspecified; ties are measure-zero for real embeddings.  The canonical
    rule used here and by the CUDA path: rank = #(score > true) + #(score == true
    with a lower index).
    """
    n = queries.shape[0]
    ranks = np.zeros(n)
    for i in range(n):
        d = gallery.matmul(queries[i])
        true = d[i]
        ranks[i] = int((d > true).sum()) + int((d[:i] == true).sum())
    return ranks


def _recall_from_ranks(ranks: np.ndarray):
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    medr = np.floor(np.median(ranks)) + 1
    return (r1, r5, r10, medr)


def t2i(images: torch.Tensor, captions: torch.Tensor):
    """Text→image recall@1/5/10 and median rank.  utils/im_retrieval_eval.py:4-30."""
    return _recall_from_ranks(retrieval_ranks(captions, images))


def i2t(images: torch.Tensor, captions: torch.Tensor):
    """Image→text recall.  utils/im_retrieval_eval.py:32-58."""
    return _recall_from_ranks(retrieval_ranks(images, captions))


# --------------------------------------------------------------------------
# a13  optimiser step restated (clip_grad_norm_ + Adam with L2-in-grad weight decay)
# --------------------------------------------------------------------------
def clip_adam_step(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Dict[str, dict],
                   lr: float, clip: float = 1.0, weight_decay: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8):
    """train.py:46-49 with the optimizer of nmt_multimodal_beam_DE.py:303-332.

    clip_grad_norm_(all params, clip) then torch.optim.Adam: weight_decay (added to
    the gradient, not decoupled) on every parameter whose name lacks 'bias'.
    Updates ``params`` / ``state`` in place; returns the pre-clip global norm.
    """
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads.values()))
    coef = min(1.0, clip / (total + 1e-6))
    b1, b2 = betas
    for name, w in params.items():
        g = grads[name] * coef
        if "bias" not in name:
            g = g + weight_decay * w
        st = state.setdefault(name, dict(step=0, m=torch.zeros_like(w), v=torch.zeros_like(w)))
        st["step"] += 1
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** st["step"]
        bc2 = 1 - b2 ** st["step"]
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        w.addcdiv_(st["m"], denom, value=-lr / bc1)
    return total
