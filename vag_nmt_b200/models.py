"""Drop-in mirrors of the two model classes the reference's drivers instantiate.

  NMT_AttentionImagine_Seq2Seq_Beam_V11   models/NMT_AttentionImagine_Seq2Seq_Beam_V11.py   (multimodal)
  NMT_Seq2Seq_Beam_V2                     models/NMT_Seq2Seq_Beam_V2.py                     (text-only)

Same constructors, methods, ``state_dict`` keys and return values.  The bodies enqueue
sm_100a kernels through libvagnmt.so; the beam search runs entirely on the device (one C call for the L-step
loop, no per-step host synchronisation) and never tiles the encoder context K times.
"""
from __future__ import annotations

import random
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .layers import LIUMCVC_Encoder, NMT_Decoder, VSE_Imagine_Enc
from .losses import ImageRetrievalRankingLoss, PairwiseRankingLoss

SOS_token = 2
EOS_token = 3
UNK_token = 1


def _nll_weight(criterion, device) -> Optional[torch.Tensor]:
    """The caller passes nn.NLLLoss(weight=vocab_mask, reduce=False) (nmt_multimodal_beam_DE.py:286-291)."""
    if criterion is None:
        return None
    if not isinstance(criterion, nn.NLLLoss):
        raise TypeError("criterion_mt must be an nn.NLLLoss(weight=..., reduce=False) like the reference driver builds")
    if getattr(criterion, "reduction", "none") != "none":
        raise ValueError("criterion_mt must be built with reduce=False / reduction='none' (the model normalises per sentence)")
    w = criterion.weight
    return None if w is None else w.detach().to(device=device, dtype=torch.float32).contiguous()


def _dropout_mask(model, name: str, p: float, shape, device) -> Optional[torch.Tensor]:
    """Mask of nn.Dropout(p) in training mode: 0 with probability p, else 1/(1-p).  Drawn with torch's CUDA generator;
    tests inject fixed masks through ``model._dropout_masks[name]`` to compare against a reference run."""
    inject = getattr(model, "_dropout_masks", None)
    if inject is not None and name in inject:
        return inject[name].to(device=device, dtype=torch.float32).contiguous()
    if not model.training or p <= 0.0:
        return None
    return torch.empty(shape, dtype=torch.float32, device=device).bernoulli_(1.0 - p).div_(1.0 - p)


def _with_precision(fn):
    """Run a public model method under the model's arithmetic mode (model.precision = "fp32" | "bf16"): every C-ABI call the
    method makes carries that vag_precision (the C library itself has no mode switch)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        from . import _cabi
        with _cabi.precision_scope(getattr(self, "precision", "fp32")):
            return fn(self, *args, **kwargs)
    return wrapper


class _Seq2SeqBase(nn.Module):
    """Pieces shared by the two models: encoder → h0 → decoder loop / beam search.

    ``precision``: "fp32" (default) keeps FP32-level accuracy on the tensor cores (error-compensated FP16 split,
    token-exact decoding); "bf16" rounds every contraction operand to bfloat16 and issues a single product with
    FP32 accumulation — state, soft-max, attention scores and losses stay FP32 (north_star "bf16 mode")."""

    precision = "fp32"

    def precision_scope(self):
        """Context manager: run the enclosed calls under the model's arithmetic mode."""
        from . import _cabi
        return _cabi.precision_scope(getattr(self, "precision", "fp32"))

    def _reset_like_reference(self):
        # V11:77-80 / V2:53-56: kaiming-normal on EVERY ≥2-D non-bias parameter (embeddings and GRU matrices too)
        for name, param in self.named_parameters():
            if param.requires_grad and 'bias' not in name and param.data.dim() > 1:
                nn.init.kaiming_normal_(param.data)

    # -- helpers ------------------------------------------------------------------------------------------
    def _device(self) -> torch.device:
        dev = self.decoderini.weight.device
        if dev.type != "cuda":
            raise RuntimeError("vag_nmt_b200 models run on a B200 only: call .cuda() first (there is no CPU path)")
        return dev

    def _encode(self, src_var, src_lengths):
        return self.encoder.forward_sentence_major(src_var, src_lengths)

    def _decode_tokens(self, w, h0, keys, ctx, mask, beam_size: int, tgt_l: int) -> List[List[int]]:
        B = ctx.shape[0]
        if beam_size == 1:
            toks = ops.greedy_decode(w, h0, keys, ctx, mask, tgt_l).cpu().tolist()  # V11:207-226
            out = []
            for row in toks:
                cut = []
                for t in row:
                    if t == EOS_token:
                        break
                    cut.append(t)
                out.append(cut)
            return out
        hyp, hyp_len = self._beam_decode(w, h0, keys, ctx, mask, beam_size, tgt_l)  # V11:229 → 233-337
        return self._hyp_lists(hyp, hyp_len)

    @staticmethod
    def _hyp_lists(hyp, hyp_len) -> List[List[int]]:
        rows = hyp.cpu().numpy()             # one device→host copy; numpy row slices convert ~2x faster than a whole-tensor tolist()
        lens = hyp_len.cpu().tolist()
        return [rows[b, :lens[b]].tolist() for b in range(rows.shape[0])]

    # The L-step loop of one (B, T, K, L) shape is captured in CUDA graphs once and replayed: a replayed loop loses ≈ 10 µs less per
    # step between its dependent launches than an enqueued one (tools/graph_rows_ab.py: 125 / 250 / 1000 sentences 13.6 → 12.5,
    # 18.3 → 17.5, 47.5 → 47.0 ms), and at the reference's eval batch of 16 (nmt_multimodal_beam_DE.py:542-547) the host could not even
    # enqueue 12 launches per step fast enough.  A captured loop cannot poll the host, so it is captured in CHUNKS of
    # `_GRAPH_CHUNK` steps (vag_beam_decode_steps_f32): every chunk ends with an asynchronous copy of the device's `done` flag into
    # pinned host memory, the host replays chunk c + 1, waits for the event behind chunk c − 1 and stops replaying when that chunk
    # reported `done` — the reference's per-step test (V11:265-269) at chunk granularity, at most two chunks of returned-at-once
    # kernels after the search has ended.  Inside a decode lane (several batches in flight from one host thread) the whole loop stays
    # ONE graph: waiting for a lane's events would serialise the lanes.  Above `_GRAPH_ROWS_MAX` rows (VAG_DECODE_GRAPH_ROWS) the
    # loop is enqueued from one C call that follows the device's progress word (ops.beam_decode).
    _GRAPH_ROWS_MAX = 16384
    _GRAPH_CACHE_MAX = 192
    _GRAPH_BIG_MAX = 4           # graphs of more than 2048 rows hold hundreds of MB of private scratch each: keep few
    _GRAPH_CHUNK = 8

    def _capture_large_now(self, key) -> bool:
        """A shape of more than 2048 rows is captured into graphs only when the large decode just before it had the SAME shape.
        Capturing costs two extra decodes and hundreds of MB of static inputs: a one-off call (a whole test set in one batch) would
        never recover them, and a caller that cycles through more distinct large shapes than `_GRAPH_BIG_MAX` would capture, evict
        and capture again on every call — both keep the enqueued loop, which is only ≈ 1-4 % slower than the replayed one there."""
        last = self.__dict__.get("_decode_last_large")
        self.__dict__["_decode_last_large"] = key
        return last == key

    def _beam_decode(self, w, h0, keys, ctx, mask, K, L):
        import os
        B, T, _ = ctx.shape
        rows_max = int(os.environ.get("VAG_DECODE_GRAPH_ROWS", self._GRAPH_ROWS_MAX))
        if B * K > rows_max or not w.prepared or os.environ.get("VAG_DECODE_GRAPH", "1") == "0" \
                or torch.cuda.is_current_stream_capturing():
            return ops.beam_decode(w, h0, keys, ctx, mask, K, L)
        from collections import OrderedDict
        cache = self.__dict__.setdefault("_decode_graphs", OrderedDict())
        # chunk length: a chunk boundary costs ≈ 20 µs (flag copy, event, graph-to-graph launch), a returned-at-once step ≈ 20 µs too
        chunk = 0 if ops._lane else int(os.environ.get("VAG_DECODE_CHUNK", self._GRAPH_CHUNK * (1 if B * K > 2048 else 2)))
        key = (ops._lane, B, T, K, L, w.precision, w.prepared, chunk, ops._weights_epoch)
        st = cache.get(key)
        if st is None and B * K > 2048 and not self._capture_large_now(key):
            return ops.beam_decode(w, h0, keys, ctx, mask, K, L)
        if st is None:
            for k in [k for k in cache if k[-1] != ops._weights_epoch]:     # graphs of an older weight set hold stale pointers
                del cache[k]
            big = [k for k in cache if k[1] * k[3] > 2048]
            while B * K > 2048 and len(big) >= self._GRAPH_BIG_MAX:
                del cache[big.pop(0)]
            while len(cache) >= self._GRAPH_CACHE_MAX:
                cache.popitem(last=False)
            dev = ctx.device
            st = {"h0": h0.clone(), "keys": keys.clone(), "ctx": ctx.clone(), "mask": mask.clone()}
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):      # warm-up outside the capture: sizes the workspace, sets kernel attributes
                ops.beam_decode(w, st["h0"], st["keys"], st["ctx"], st["mask"], K, L, early_stop=False)
            torch.cuda.current_stream(dev).wait_stream(side)
            if chunk <= 0 or chunk >= L:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    st["hyp"], st["hyp_len"] = ops.beam_decode(w, st["h0"], st["keys"], st["ctx"], st["mask"], K, L, early_stop=False)
                st["graph"] = g
            else:
                st["hyp"] = torch.empty(B, L, dtype=torch.int64, device=dev)
                st["hyp_len"] = torch.empty(B, dtype=torch.int32, device=dev)
                bounds = list(range(0, L, chunk)) + [L]
                st["done_host"] = torch.zeros(len(bounds), dtype=torch.int32).pin_memory()
                st["chunks"] = []
                pool = None
                for c in range(len(bounds)):        # the last "chunk" is the epilogue alone: [L, L)
                    lo, hi = bounds[c], (bounds[c + 1] if c + 1 < len(bounds) else L)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        ops.beam_decode_steps(w, st["h0"], st["keys"], st["ctx"], st["mask"], K, L, lo, hi, st["hyp"], st["hyp_len"],
                                              st["done_host"][c:c + 1] if hi < L or lo < L else None)
                    pool = pool or g.pool()
                    st["chunks"].append(g)
            st["keepalive"] = (list(ops._workspaces.values()), getattr(w, "_keepalive", None))   # raw pointers inside the graph
            cache[key] = st
        else:
            cache.move_to_end(key)
        st["h0"].copy_(h0)
        st["keys"].copy_(keys)
        st["ctx"].copy_(ctx)
        st["mask"].copy_(mask)
        if "graph" in st:
            st["graph"].replay()
        else:
            graphs, flags, events, stopped = st["chunks"], st["done_host"], [], False
            for c in range(len(graphs) - 1):
                if c >= 2:
                    events[c - 2].synchronize()        # chunk c − 1 is queued behind it: the device stays busy while the host looks
                    if int(flags[c - 2]) != 0:
                        stopped = True
                        break
                graphs[c].replay()
                ev = torch.cuda.Event()
                ev.record()
                events.append(ev)
            if stopped:
                graphs[-1].replay()                    # the epilogue alone (the chunk that reaches step L carries its own)
        return st["hyp"], st["hyp_len"]      # static buffers: consume (or clone) before the next decode of this shape

    # Decode lanes (opt-in, VAG_DECODE_LANES=n): one call's batch cut into n independent sub-batches (sentences do not interact),
    # each decoded by its own CUDA graph on its own stream with private scratch buffers (ops.lane).  Measured on one B200
    # (tools/decode_lanes_bench.py), tokens identical, but SLOWER for every batch a call would be cut from — 125 sentences 14.7 →
    # 18.5 ms with 2 lanes, 250: 22.0 → 23.1, 500: 29.8 → 36.3 — because the persistent contraction kernels of each lane claim all
    # SMs and the lanes' steps serialise.  The same machinery pays where the batches are small by definition: the reference's eval
    # batches of 16 with several batches in flight (translate.decode_corpus_pipelined: 1.46 k → 3.5 k sentences/s).
    def _lane_plan(self, B: int, K: int):
        import os
        env = os.environ.get("VAG_DECODE_LANES", "1")
        if K <= 1 or self.training or torch.cuda.is_current_stream_capturing() or ops._lane or not env.isdigit() or int(env) < 2:
            return None
        n = min(int(env), 8, B)
        if n < 2 or -(-B // n) * K > 2048:      # lanes only pay for batches that leave most of the GPU idle
            return None
        per = -(-B // n)
        return [(lo, min(B, lo + per)) for lo in range(0, B, per)]

    def _lane_streams(self, n: int, dev):
        pool = self.__dict__.setdefault("_lane_stream_pool", [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(device=dev))
        return pool[:n]

    def _decode_lanes(self, plan, src_var, src_lengths, im_var, beam_size, max_length):
        dev = self._device()
        cur = torch.cuda.current_stream(dev)
        src = src_var.to(device=dev, dtype=torch.int64)
        im = im_var.to(dev) if im_var is not None else None
        lens = [int(x) for x in src_lengths]
        ops.decoder_weights(self.decoder, self.decoderini, prepare=True)    # the shared invariants, once, on the caller's stream
        streams = self._lane_streams(len(plan), dev)
        outs = []
        for i, (lo, hi) in enumerate(plan):
            streams[i].wait_stream(cur)
            with torch.cuda.stream(streams[i]), ops.lane(f"lane{i}:"):
                sub = src[lo:hi, :lens[lo]].contiguous()                    # rows are sorted by length: lens[lo] is the slice's width
                outs.append(self._decode_device_one(sub, lens[lo:hi], im[lo:hi] if im is not None else None, beam_size, max_length))
        for st in streams:
            cur.wait_stream(st)
        if beam_size == 1:
            return torch.cat([o[0] for o in outs]), None
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    @_with_precision
    def decode_device(self, src_var, src_lengths, im_var=None, beam_size=12, max_length=80):
        """Device-resident beam search: same work as ``beamsearch_decode`` but returns CUDA tensors
        (hyp int64 [B, L], hyp_len int32 [B]) without the final device→host copy / list building."""
        plan = self._lane_plan(len(src_lengths), beam_size)
        if plan is not None:
            return self._decode_lanes(plan, src_var, src_lengths, im_var, beam_size, max_length)
        return self._decode_device_one(src_var, src_lengths, im_var, beam_size, max_length)

    def _decode_device_one(self, src_var, src_lengths, im_var, beam_size, max_length):
        if isinstance(self, NMT_AttentionImagine_Seq2Seq_Beam_V11):
            w, ctx, mask, keys, h0, _, _ = self._prepare(src_var, src_lengths, im_var)
        else:
            w, ctx, mask, keys, h0 = self._prepare(src_var, src_lengths)
        if beam_size == 1:
            return ops.greedy_decode(w, h0, keys, ctx, mask, max_length), None
        hyp, hyp_len = self._beam_decode(w, h0, keys, ctx, mask, beam_size, max_length)
        return hyp.clone(), hyp_len.clone()

    # -- training path (autograd through hand-written backward kernels) ------------------------------------
    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _decoder_param_list(self):
        d = self.decoder
        return [d.embedding.weight, d.gru_1.weight_ih_l0, d.gru_1.weight_hh_l0, d.gru_1.bias_ih_l0, d.gru_1.bias_hh_l0,
                d.attn.attn_h.weight, d.attn.attn_e.weight, d.attn.v, d.context2hid.weight, d.gru_2.weight_ih_l0,
                d.gru_2.weight_hh_l0, d.gru_2.bias_ih_l0, d.gru_2.bias_hh_l0, d.W1.weight, d.W1.bias, d.W2.weight, d.W2.bias,
                d.W3.weight, d.W3.bias, d.out.weight, d.out.bias]

    def _encode_train(self, src_var, src_lengths):
        from .autograd import EncoderFn
        enc = self.encoder
        dev = self._device()
        src = src_var.to(device=dev, dtype=torch.int64).contiguous()
        mask = None
        if src_lengths is None:
            # graph-captured step: the caller validated the lengths on the host; the device recomputes them from the padding
            mask, lengths = ops.src_mask_lengths(src)
        elif torch.is_tensor(src_lengths) and src_lengths.is_cuda:
            lengths = src_lengths
        else:
            lengths = [int(x) for x in src_lengths]
            if src.shape[1] != max(lengths):
                raise ValueError("the padded width must equal the longest sentence (pad_packed_sequence, Encoder.py:60)")
            if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
                raise RuntimeError("`lengths` array must be sorted in decreasing order (pack_padded_sequence, Encoder.py:55)")
        g = enc.gru
        B, Tn = src.shape
        E, H = enc.embedding.weight.shape[1], enc.hidden_size
        emb_mask = _dropout_mask(self, "emb", enc.dropout_emb, (Tn * B, E), dev)        # time-major rows, Encoder.py:51-52
        ctx = EncoderFn.apply(src, lengths, emb_mask, enc.embedding.weight, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0,
                              g.bias_hh_l0, g.weight_ih_l0_reverse, g.weight_hh_l0_reverse, g.bias_ih_l0_reverse,
                              g.bias_hh_l0_reverse)
        ctx_mask = _dropout_mask(self, "ctx", enc.dropout_ctx, (B, Tn, 2 * H), dev)     # Encoder.py:62-63
        if ctx_mask is not None:
            from .autograd import MaskMulFn
            ctx = MaskMulFn.apply(ctx, ctx_mask)
        if mask is None:
            mask, _ = ops.src_mask_lengths(src, want_lengths=False)   # Encoder.py:47
        return ctx, mask

    def _decoder_loss_train(self, h0, ctx, mask, tgt, teacher_force_ratio, weight):
        from .autograd import DecoderSeqFn
        is_teacher = random.random() < teacher_force_ratio          # V11:136
        B, Tt = tgt.shape
        out_mask = _dropout_mask(self, "out", self.decoder.dropout_out, (Tt * B, self.decoder.embedding_size), h0.device)
        return DecoderSeqFn.apply(h0, ctx, mask, tgt, weight, is_teacher, bool(self.decoder.tied_emb), out_mask,
                                  *self._decoder_param_list())

    def _translation_loss_rows(self, w, h0, keys, ctx, mask, tgt, teacher_force_ratio, weight):
        """The Tt-step loop of forward (V11:136-160): Σ_t NLL rows [B]."""
        B, Tt = tgt.shape
        dev = ctx.device
        loss_rows = torch.zeros(B, dtype=torch.float32, device=dev)
        inp = torch.full((B,), SOS_token, dtype=torch.int64, device=dev)
        is_teacher = random.random() < teacher_force_ratio  # V11:136 (python RNG, once per batch)
        h = h0
        tgt_t = tgt.t().contiguous()  # [Tt, B] so that a step's targets are one contiguous row
        for di in range(Tt):
            logits, h, _ = ops.decoder_step(w, inp, h, keys, ctx, mask, 1, want_logp=False)
            ops.nll_rows(logits, tgt_t[di], weight, loss_rows)
            inp = tgt_t[di] if is_teacher else ops.row_argmax(logits)
        return loss_rows


class NMT_AttentionImagine_Seq2Seq_Beam_V11(_Seq2SeqBase):
    def __init__(self, src_size, tgt_size, im_feats_size, src_embedding_size, tgt_embedding_size, hidden_size,
                 shared_embedding_size, loss_w, beam_size=1, attn_model='dot', n_layers=1, dropout_ctx=0.0,
                 dropout_emb=0.0, dropout_out=0.0, dropout_rnn_enc=0.0, dropout_rnn_dec=0.0, dropout_im_emb=0.0,
                 dropout_txt_emb=0.0, activation_vse=True, tied_emb=False, init_split=0.5):
        super().__init__()
        self.src_size = src_size
        self.tgt_size = tgt_size
        self.im_feats_size = im_feats_size
        self.src_embedding_size = src_embedding_size
        self.tgt_embedding_size = tgt_embedding_size
        self.hidden_size = hidden_size
        self.n_layers = n_layers
        self.shared_embedding_size = shared_embedding_size
        self.beam_size = beam_size
        self.loss_w = loss_w
        self.tied_emb = tied_emb
        self.dropout_im_emb = dropout_im_emb
        self.dropout_txt_emb = dropout_txt_emb
        self.activation_vse = activation_vse
        self.attn_model = attn_model
        self.init_split = init_split
        self.encoder = LIUMCVC_Encoder(src_size, src_embedding_size, hidden_size, n_layers, dropout_rnn=dropout_rnn_enc,
                                       dropout_ctx=dropout_ctx, dropout_emb=dropout_emb)
        self.decoder = NMT_Decoder(tgt_size, tgt_embedding_size, hidden_size, 2 * hidden_size, n_layers,
                                   dropout_rnn=dropout_rnn_dec, dropout_out=dropout_out, dropout_emb=0.0, tied_emb=tied_emb)
        self.vse_imagine = VSE_Imagine_Enc(self.attn_model, self.im_feats_size, 2 * hidden_size, self.shared_embedding_size,
                                           self.dropout_im_emb, self.dropout_txt_emb, self.activation_vse)
        self.decoderini = nn.Linear(2 * hidden_size, hidden_size)
        self.reset_parameters()

    def reset_parameters(self):
        self._reset_like_reference()

    def _prepare(self, src_var, src_lengths, im_var, want_embeddings=False):
        dev = self._device()
        ctx, mask = self._encode(src_var, src_lengths)                                   # V11:111 / :193
        im_emb, txt_emb, ctx_vec, _ = self.vse_imagine.pool_sentence_major(im_var.to(dev), ctx, mask)  # :114 / :196
        w = ops.decoder_weights(self.decoder, self.decoderini, prepare=not self.training)
        h0 = ops.decoder_init(w, ctx_vec, ctx, mask, self.init_split)                    # :118 / :201
        keys = ops.attn_keys(w, ctx)
        return w, ctx, mask, keys, h0, im_emb, txt_emb

    @_with_precision
    def forward(self, src_var, src_lengths, tgt_var, im_var, teacher_force_ratio=1.0, max_length=80, criterion_mt=None,
                criterion_vse=None):
        """→ (loss, loss_mt, loss_vse), V11:82-168."""
        dev = self._device()
        self.tgt_l = tgt_var.size()[1]
        if self._wants_grad():
            return self._forward_train(src_var, src_lengths, tgt_var, im_var, teacher_force_ratio, criterion_mt, criterion_vse)
        w, ctx, mask, keys, h0, im_emb, txt_emb = self._prepare(src_var, src_lengths, im_var)
        loss_vse = None
        if criterion_vse is not None:
            loss_vse = criterion_vse(im_emb, txt_emb)                                    # VSE_Imagine_Enc.py:149-150
        tgt = tgt_var.to(device=dev, dtype=torch.int64).contiguous()
        weight = _nll_weight(criterion_mt, dev)
        loss_rows = self._translation_loss_rows(w, h0, keys, ctx, mask, tgt, teacher_force_ratio, weight)
        vse_dev = None
        if loss_vse is not None:
            vse_dev = loss_vse.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        out = ops.translation_loss(loss_rows, tgt, vse_dev, self.loss_w if vse_dev is not None else 1.0)
        if loss_vse is None:
            # the reference mixes with the python int 0 (V11:91,166): loss = loss_w * loss_mt
            return self.loss_w * out[1], out[1], 0
        return out[0], out[1], out[2]

    def _forward_train(self, src_var, src_lengths, tgt_var, im_var, teacher_force_ratio, criterion_mt, criterion_vse):
        """Same values as the inference-mode forward, recorded for autograd (hand-written backward kernels)."""
        from .autograd import DecoderInitFn, LossMixFn, VsePoolFn
        dev = self._device()
        ctx, mask = self._encode_train(src_var, src_lengths)
        vse = self.vse_imagine
        im = im_var.to(device=dev, dtype=torch.float32).contiguous()
        mlp_w = vse.imagine_attn.mlp.weight if vse.attn_type == "mlp" else None
        im_emb, txt_emb, ctx_vec = VsePoolFn.apply(im, ctx, mask, vse.attn_type, bool(vse.activation_vse), vse.im_embedding.weight,
                                                   vse.im_embedding.bias, vse.text_embedding.weight, vse.text_embedding.bias,
                                                   vse.imagine_attn.ctx2ctx.weight, vse.imagine_attn.emb2ctx.weight, mlp_w)
        loss_vse = criterion_vse(im_emb, txt_emb) if criterion_vse is not None else None
        h0 = DecoderInitFn.apply(ctx_vec, ctx, mask, float(self.init_split), self.decoderini.weight, self.decoderini.bias)
        tgt = tgt_var.to(device=dev, dtype=torch.int64).contiguous()
        weight = _nll_weight(criterion_mt, dev)
        vse_in = loss_vse.reshape(1) if loss_vse is not None else None
        self._boundary = None
        if getattr(self, "_bwd_split", False) and vse_in is not None:
            # Data-parallel step (train.GraphedTrainStep): the backward runs in two pieces — decoder first, so that the
            # all-reduce of the decoder's gradients overlaps with the encoder's back-propagation through time.  The cut is made of
            # detached leaves; piece 2 starts from (roots, the leaves' gradients).
            roots = [ctx, h0, vse_in]
            ctx, h0, vse_in = (t.detach().requires_grad_() for t in roots)
            self._boundary = (roots, [ctx, h0, vse_in])
        loss_rows = self._decoder_loss_train(h0, ctx, mask, tgt, teacher_force_ratio, weight)
        out = LossMixFn.apply(loss_rows, tgt, vse_in, float(self.loss_w))
        if loss_vse is None:
            self._loss_vec = None
            return self.loss_w * out[1], out[1], 0
        self._loss_vec = (out, 0)       # GraphedTrainStep seeds backward on the whole vector (no select/stack kernels)
        return out[0], out[1], out[2]

    @_with_precision
    def beamsearch_decode(self, src_var, src_lengths, im_var, beam_size=1, max_length=80, tgt_var=None):
        """→ list[B] of token-id lists (EOS excluded), V11:179-231."""
        tgt_l = max_length if tgt_var is None else tgt_var.size()[1]
        self.tgt_l = tgt_l
        self.beam_size = beam_size
        plan = self._lane_plan(len(src_lengths), beam_size)
        if plan is not None:
            hyp, hyp_len = self._decode_lanes(plan, src_var, src_lengths, im_var, beam_size, tgt_l)
            self.final_sample = self._hyp_lists(hyp, hyp_len)
            return self.final_sample
        w, ctx, mask, keys, h0, _, _ = self._prepare(src_var, src_lengths, im_var)
        self.final_sample = self._decode_tokens(w, h0, keys, ctx, mask, beam_size, tgt_l)
        return self.final_sample

    @_with_precision
    def embed_sent_im_eval(self, src_var, src_lengths, tgt_var, im_feats):
        """→ (im_embedding, text_embedding), V11:341-368."""
        self.tgt_l = tgt_var.size()[1]
        return self._embed(src_var, src_lengths, im_feats)

    @_with_precision
    def embed_sent_im_test(self, src_var, src_lengths, im_feats, max_length=80):
        """→ (im_embedding, text_embedding), V11:370-397."""
        self.tgt_l = max_length
        return self._embed(src_var, src_lengths, im_feats)

    def _embed(self, src_var, src_lengths, im_feats):
        dev = self._device()
        ctx, mask = self._encode(src_var, src_lengths)
        im_emb, txt_emb, _, _ = self.vse_imagine.pool_sentence_major(im_feats.to(dev), ctx, mask)
        return im_emb, txt_emb

    @_with_precision
    def get_imagine_attention_eval(self, src_var, src_lengths, tgt_var, im_feats):
        """→ attention weights [B, 1, T], V11:399-424."""
        self.tgt_l = tgt_var.size()[1]
        return self._imagine_attention(src_var, src_lengths, im_feats)

    @_with_precision
    def get_imagine_attention_test(self, src_var, src_lengths, im_feats, max_length=80):
        """→ attention weights [B, 1, T], V11:426-450."""
        self.tgt_l = max_length
        return self._imagine_attention(src_var, src_lengths, im_feats)

    def _imagine_attention(self, src_var, src_lengths, im_feats):
        dev = self._device()
        ctx, mask = self._encode(src_var, src_lengths)
        _, _, _, beta = self.vse_imagine.pool_sentence_major(im_feats.to(dev), ctx, mask, want_beta=True)
        return beta.unsqueeze(1)


class NMT_Seq2Seq_Beam_V2(_Seq2SeqBase):
    def __init__(self, src_size, tgt_size, src_embedding_size, tgt_embedding_size, hidden_size, beam_size=1, n_layers=1,
                 dropout_ctx=0.0, dropout_emb=0.0, dropout_out=0.0, dropout_rnn=0.0, tied_emb=False):
        super().__init__()
        self.src_size = src_size
        self.tgt_size = tgt_size
        self.src_embedding_size = src_embedding_size
        self.tgt_embedding_size = tgt_embedding_size
        self.hidden_size = hidden_size
        self.n_layers = n_layers
        self.beam_size = beam_size
        self.tied_emb = tied_emb
        self.encoder = LIUMCVC_Encoder(src_size, src_embedding_size, hidden_size, n_layers, dropout_rnn=dropout_rnn,
                                       dropout_ctx=dropout_ctx, dropout_emb=dropout_emb)
        self.decoder = NMT_Decoder(tgt_size, tgt_embedding_size, hidden_size, 2 * hidden_size, n_layers,
                                   dropout_out=dropout_out, tied_emb=tied_emb)
        self.decoderini = nn.Linear(2 * hidden_size, hidden_size)
        self.reset_parameters()

    def reset_parameters(self):
        self._reset_like_reference()

    def _prepare(self, src_var, src_lengths):
        self._device()
        ctx, mask = self._encode(src_var, src_lengths)
        w = ops.decoder_weights(self.decoder, self.decoderini, prepare=not self.training)
        h0 = ops.decoder_init(w, None, ctx, mask, 0.0)                                   # V2:85,142
        keys = ops.attn_keys(w, ctx)
        return w, ctx, mask, keys, h0

    @_with_precision
    def forward(self, src_var, src_lengths, tgt_var, teacher_force_ratio=1.0, max_length=80, criterion=None):
        """→ loss, models/NMT_Seq2Seq_Beam_V2.py:58-113."""
        dev = self._device()
        self.tgt_l = tgt_var.size()[1]
        if self._wants_grad():
            from .autograd import DecoderInitFn, LossMixFn
            ctx, mask = self._encode_train(src_var, src_lengths)
            h0 = DecoderInitFn.apply(None, ctx, mask, 0.0, self.decoderini.weight, self.decoderini.bias)
            tgt = tgt_var.to(device=dev, dtype=torch.int64).contiguous()
            loss_rows = self._decoder_loss_train(h0, ctx, mask, tgt, teacher_force_ratio, _nll_weight(criterion, dev))
            out = LossMixFn.apply(loss_rows, tgt, None, 1.0)
            self._loss_vec = (out, 1)
            return out[1]
        w, ctx, mask, keys, h0 = self._prepare(src_var, src_lengths)
        tgt = tgt_var.to(device=dev, dtype=torch.int64).contiguous()
        weight = _nll_weight(criterion, dev)
        loss_rows = self._translation_loss_rows(w, h0, keys, ctx, mask, tgt, teacher_force_ratio, weight)
        return ops.translation_loss(loss_rows, tgt, None, 1.0)[1]

    @_with_precision
    def beamsearch_decode(self, src_var, src_lengths, beam_size=1, max_length=80, tgt_var=None):
        """→ list[B] of token-id lists, models/NMT_Seq2Seq_Beam_V2.py:124-171."""
        tgt_l = max_length if tgt_var is None else tgt_var.size()[1]
        self.tgt_l = tgt_l
        self.beam_size = beam_size
        plan = self._lane_plan(len(src_lengths), beam_size)
        if plan is not None:
            hyp, hyp_len = self._decode_lanes(plan, src_var, src_lengths, None, beam_size, tgt_l)
            self.final_sample = self._hyp_lists(hyp, hyp_len)
            return self.final_sample
        w, ctx, mask, keys, h0 = self._prepare(src_var, src_lengths)
        self.final_sample = self._decode_tokens(w, h0, keys, ctx, mask, beam_size, tgt_l)
        return self.final_sample


__all__ = ["NMT_AttentionImagine_Seq2Seq_Beam_V11", "NMT_Seq2Seq_Beam_V2", "PairwiseRankingLoss",
           "ImageRetrievalRankingLoss", "SOS_token", "EOS_token", "UNK_token"]
