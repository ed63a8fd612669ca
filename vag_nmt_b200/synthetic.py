"""Multi30K-shaped synthetic inputs.

No dataset ships with the reference, so every parity test and the bench build
their batches here: source/target lengths ~ clip(round(N(14, 4.5)), 4, 40)
including the final EOS, tokens uniform in [4, V), EOS = 3, pad = 0, rows of a
batch sorted by source length descending (what ``data_generator_mtv`` does,
preprocessing.py:234-306), image features ``torch.rand`` (non-negative like
ResNet pool5).  Everything is generated on the CPU from explicit generators so
the build container and the GPU box see identical bytes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

# vocabulary / model shapes of the shipped experiments (nohup.out:14-28, +1 for pad)
DE = dict(src_size=8507, tgt_size=9391, im_feats_size=2048, src_embedding_size=256,
          tgt_embedding_size=256, hidden_size=512, shared_embedding_size=512)
FR = dict(DE, tgt_size=8748)
TINY = dict(src_size=41, tgt_size=50, im_feats_size=24, src_embedding_size=8,
            tgt_embedding_size=8, hidden_size=16, shared_embedding_size=12)


@dataclass
class Batch:
    src: torch.Tensor            # int64 [B, Ts], 0-padded, rows sorted by length desc
    src_lengths: List[int]
    tgt: Optional[torch.Tensor]  # int64 [B, Tt]
    im: Optional[torch.Tensor]   # fp32 [B, I]
    order: List[int]             # position in the unsorted draw of every sorted row


def draw_lengths(n: int, gen: torch.Generator, lo: int = 4, hi: int = 40, mean: float = 14.0,
                 std: float = 4.5) -> torch.Tensor:
    x = torch.randn(n, generator=gen) * std + mean
    return x.round().clamp(lo, hi).long()


def _sentences(lengths: torch.Tensor, vocab: int, gen: torch.Generator, width: Optional[int] = None) -> torch.Tensor:
    n = lengths.numel()
    width = int(lengths.max()) if width is None else width
    toks = torch.randint(4, vocab, (n, width), generator=gen)
    pos = torch.arange(width).unsqueeze(0)
    L = lengths.unsqueeze(1)
    toks = torch.where(pos < L - 1, toks, torch.zeros_like(toks))
    toks = torch.where(pos == L - 1, torch.full_like(toks, 3), toks)
    return toks


def make_batch(batch_size: int, src_vocab: int, tgt_vocab: Optional[int] = None, im_size: Optional[int] = None,
               seed: int = 7, common_tgt_len: bool = True, max_len: int = 40, min_len: int = 4, mean: float = 14.0,
               std: float = 4.5) -> Batch:
    """One batch the way the reference's generators hand it to the model."""
    gen = torch.Generator().manual_seed(seed)
    ls = draw_lengths(batch_size, gen, lo=min_len, hi=max_len, mean=mean, std=std)
    src = _sentences(ls, src_vocab, gen)
    tgt = None
    if tgt_vocab is not None:
        if common_tgt_len:  # BucketBatchSampler: one target length per train batch (samplers/bucket.py:10-103)
            lt = draw_lengths(1, gen, lo=min_len, hi=max_len, mean=mean, std=std).expand(batch_size).contiguous()
        else:
            lt = draw_lengths(batch_size, gen, lo=min_len, hi=max_len, mean=mean, std=std)
        tgt = _sentences(lt, tgt_vocab, gen)
    im = torch.rand(batch_size, im_size, generator=gen) if im_size is not None else None
    order = torch.argsort(ls, descending=True, stable=True)
    src = src[order]
    if tgt is not None:
        tgt = tgt[order]
    if im is not None:
        im = im[order]
    return Batch(src.contiguous(), [int(x) for x in ls[order]], tgt, im, [int(x) for x in order])


def make_corpus(n_sent: int, src_vocab: int, im_size: Optional[int], seed: int = 7, max_len: int = 40,
                min_len: int = 4, mean: float = 14.0, std: float = 4.5):
    """An unsorted test-set-shaped corpus: list of token lists + image matrix."""
    gen = torch.Generator().manual_seed(seed)
    ls = draw_lengths(n_sent, gen, lo=min_len, hi=max_len, mean=mean, std=std)
    toks = _sentences(ls, src_vocab, gen)
    sents = [toks[i, :int(ls[i])].tolist() for i in range(n_sent)]
    im = torch.rand(n_sent, im_size, generator=gen) if im_size is not None else None
    return sents, im


def pad_and_sort(sents: List[List[int]], im: Optional[torch.Tensor] = None):
    """Pad a list of sentences to one width and sort by length descending.

    Mirrors the per-batch preparation of ``data_generator_mtv``
    (preprocessing.py:262-296); returns (src, lengths, im_sorted, order) where
    ``order[i]`` is the corpus index of sorted row i (the reference's
    ``x_reverse_sorted_index``).
    """
    lens = torch.tensor([len(s) for s in sents])
    order = torch.argsort(lens, descending=True, stable=True)
    width = int(lens.max())
    src = torch.zeros(len(sents), width, dtype=torch.long)
    for r, i in enumerate(order.tolist()):
        src[r, :len(sents[i])] = torch.tensor(sents[i])
    im_s = im[order] if im is not None else None
    return src, [int(x) for x in lens[order]], im_s, order.tolist()


def dropout_masks(seed: int, B: int, Ts: int, Tt: int, src_emb: int, hidden: int, tgt_emb: int, p_emb: float, p_ctx: float,
                  p_out: float) -> dict:
    """Pre-drawn training-mode dropout masks (entries 0 or 1/(1-p)) in the layouts the drop-in injects through
    ``model._dropout_masks``: "emb" [Ts·B, E] time-major (Encoder.py:51-52), "ctx" [B, Ts, 2H] (Encoder.py:62-63),
    "out" [Tt·B, E] (NMT_Decoder.py:140-141).  CPU generator ⇒ the same bytes in the build container and on the GPU box."""
    gen = torch.Generator().manual_seed(seed)

    def draw(shape, p):
        return torch.empty(shape).bernoulli_(1.0 - p, generator=gen) / (1.0 - p)

    out = {}
    if p_emb > 0:
        out["emb"] = draw((Ts * B, src_emb), p_emb)
    if p_ctx > 0:
        out["ctx"] = draw((B, Ts, 2 * hidden), p_ctx)
    if p_out > 0:
        out["out"] = draw((Tt * B, tgt_emb), p_out)
    return out


def install_eos_clock(model, mean_len: float = 15.0, gain: float = 8.0, slope: float = 3.0) -> None:
    """Make a RANDOM-INIT model end its hypotheses after ≈ mean_len tokens, the way a trained Multi30K model does (targets average
    ≈ 14 tokens) — random-init weights never emit <eos>, so without this every decode runs max_length
    steps and the early stop (V11:265-269) cannot be measured.  One hidden unit j = H-1 of the decoder is rewired into a clock:
      h0_j = -0.9 (decoderini row j), gru_1 passes it through (z_j ≈ 1), gru_2 leaks it towards 1 (h' = ε + (1-ε)·h),
      read-out channel d = E-1 is tanh(slope·h_j) and only <eos> (id 3) reads that channel with weight `gain`.
    The <eos> logit so rises by ≈ gain·slope·ε per step and overtakes the random logits of the other tokens around step mean_len.
    Everything else keeps its random initialisation; the oracle sees the same state_dict, so parity checks still apply."""
    import math
    dec = model.decoder
    H, E = dec.hidden_size, dec.embedding_size
    j, d = H - 1, E - 1
    # crossing level of the clock: <eos> wins once gain·tanh(slope·h) ≈ 1.3 (measured transition of the random logits)
    h_cross = math.atanh(min(1.3 / gain, 0.99)) / slope
    eps = 1.0 - ((1.0 - h_cross) / 1.9) ** (1.0 / mean_len)
    with torch.no_grad():
        model.decoderini.weight[j].zero_()
        model.decoderini.bias[j] = math.atanh(-0.9)
        for gru, z_bias, n_bias in ((dec.gru_1, 20.0, 0.0), (dec.gru_2, math.log((1.0 - eps) / eps), 4.0)):
            for base in (0, H, 2 * H):                         # rows r_j, z_j, n_j see no input and no state
                gru.weight_ih_l0[base + j].zero_()
                gru.weight_hh_l0[base + j].zero_()
                gru.bias_ih_l0[base + j] = 0.0
                gru.bias_hh_l0[base + j] = 0.0
            gru.bias_ih_l0[H + j] = z_bias                     # z_j = σ(z_bias)
            gru.bias_ih_l0[2 * H + j] = n_bias                 # n_j = tanh(n_bias)
        for lin in (dec.W1, dec.W2, dec.W3):                   # read-out channel d = tanh(slope·h2_j)
            lin.weight[d].zero_()
            lin.bias[d] = 0.0
        dec.W1.weight[d, j] = slope
        dec.embedding.weight[:, d] = 0.0                       # tied with out.weight: only <eos> reads channel d
        dec.embedding.weight[3, d] = gain
        if dec.out.weight.data_ptr() != dec.embedding.weight.data_ptr():
            dec.out.weight[:, d] = 0.0
            dec.out.weight[3, d] = gain
