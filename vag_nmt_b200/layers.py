"""Drop-in mirrors of the reference's layer classes, bodies dispatching to libvagnmt.so.

Same class names, constructor arguments, ``state_dict`` keys and return shapes as
``machine_translation_vision/layers/{Encoder,NMT_Decoder,VSE_Imagine_Enc}.py``.
The torch ``nn.Embedding`` / ``nn.GRU`` / ``nn.Linear`` objects are PARAMETER CONTAINERS ONLY (they give the
reference's parameter names, shapes, default initialisation and RNG order); their ``forward`` is never called.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import ops


def _no_training_dropout(module: nn.Module, *rates: float) -> None:
    """The per-layer forward methods are the inference path (no autograd); dropout applies only in the models'
    training path (models.py::_forward_train), so a layer called directly in train mode with p > 0 refuses."""
    if module.training and torch.is_grad_enabled() and any(r > 0 for r in rates):
        raise NotImplementedError(
            "layer-level forward is inference-only; train through the model's forward (dropout + autograd live there)")


class LIUMCVC_Encoder(nn.Module):
    """Embedding + packed bidirectional GRU.  layers/Encoder.py:11-65."""

    def __init__(self, input_size, embedding_size, hidden_size, n_layers=1, dropout_rnn=0, dropout_emb=0, dropout_ctx=0):
        super().__init__()
        if n_layers != 1:
            raise NotImplementedError("only n_layers=1 is exercised by the reference drivers (SURVEY.md section 8b)")
        self.n_layers = n_layers
        self.hidden_size = hidden_size
        self.n_direction = 2
        self.dropout_rnn = dropout_rnn
        self.dropout_emb = dropout_emb
        self.dropout_ctx = dropout_ctx
        self.embedding = nn.Embedding(input_size, embedding_size, padding_idx=0)
        self.gru = nn.GRU(embedding_size, hidden_size, num_layers=n_layers, bidirectional=True, dropout=dropout_rnn)

    def forward_sentence_major(self, input_var, input_lengths):
        """→ ctx [B, T, 2H], mask [B, T] — the layout the kernels use."""
        _no_training_dropout(self, self.dropout_emb, self.dropout_ctx)
        dev = self.embedding.weight.device
        src = input_var.to(dev)
        lengths = [int(x) for x in input_lengths]
        if src.dim() != 2 or len(lengths) != src.shape[0]:
            raise ValueError("input_var must be [B, W] with one length per row")
        if src.shape[1] != max(lengths):
            raise ValueError("the padded width must equal the longest sentence (pad_packed_sequence, Encoder.py:60)")
        if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
            raise RuntimeError("`lengths` array must be sorted in decreasing order (pack_padded_sequence, Encoder.py:55)")
        return ops.encoder_fwd(ops.encoder_weights(self), src, lengths)

    def forward(self, input_var, input_lengths):
        """→ output [W, B, 2H], ctx_mask [W, B] float, like Encoder.py:36-65 (views of the sentence-major buffers)."""
        ctx, mask = self.forward_sentence_major(input_var, input_lengths)
        return ctx.transpose(0, 1), mask.transpose(0, 1)


def _sentence_major(encoder_outputs: torch.Tensor) -> torch.Tensor:
    """[T, N, C] (reference layout) → contiguous [N, T, C]; free when it is a view made by LIUMCVC_Encoder.forward."""
    return encoder_outputs.transpose(0, 1).contiguous()


class BahdanauAttn(nn.Module):
    """MLP attention.  layers/NMT_Decoder.py:9-51."""

    def __init__(self, context_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.context_size = context_size
        self.attn_h = nn.Linear(self.hidden_size, self.context_size, bias=False)
        self.attn_e = nn.Linear(self.context_size, self.context_size, bias=False)
        self.v = nn.Parameter(torch.rand(self.context_size))
        stdv = 1. / math.sqrt(self.v.size(0))
        self.v.data.normal_(mean=0, std=stdv)

    def forward(self, hidden, encoder_outputs, ctx_mask=None):
        """hidden [1, B, H], encoder_outputs [S, B, C], ctx_mask [S, B] → weights [B, 1, S]."""
        ctx = _sentence_major(encoder_outputs)
        keys = ops.linear(ctx, self.attn_e.weight)
        q = ops.linear(hidden.reshape(-1, self.hidden_size), self.attn_h.weight)
        mask = ctx_mask.transpose(0, 1).contiguous() if ctx_mask is not None else None
        _, alpha = ops.attention(q, keys, ctx, self.v, mask, 1, ops.ATTN_MLP, want_alpha=True)
        return alpha.unsqueeze(1)


class NMT_Decoder(nn.Module):
    """Conditional-GRU attention decoder, one step per call.  layers/NMT_Decoder.py:54-145."""

    def __init__(self, output_size, embedding_size, hidden_size, context_size, n_layers=1, dropout_emb=0.0,
                 dropout_rnn=0.0, dropout_out=0.0, bias_zero=True, tied_emb=False):
        super().__init__()
        if n_layers != 1:
            raise NotImplementedError("only n_layers=1 is exercised by the reference drivers")
        self.embedding_size = embedding_size
        self.hidden_size = hidden_size
        self.context_size = context_size
        self.n_layers = n_layers
        self.dropout_emb = dropout_emb
        self.dropout_out = dropout_out
        self.bias_zero = bias_zero
        self.tied_emb = tied_emb
        self.embedding = nn.Embedding(output_size, embedding_size, padding_idx=0)
        self.gru_1 = nn.GRU(embedding_size, hidden_size, num_layers=n_layers, dropout=dropout_rnn)
        self.attn = BahdanauAttn(context_size, hidden_size)
        self.context2hid = nn.Linear(context_size, hidden_size, bias=False)
        self.gru_2 = nn.GRU(hidden_size, hidden_size, num_layers=n_layers, dropout=dropout_rnn)
        self.W1 = nn.Linear(hidden_size, embedding_size)
        self.W2 = nn.Linear(context_size, embedding_size)
        self.W3 = nn.Linear(embedding_size, embedding_size)
        self.out = nn.Linear(embedding_size, output_size)
        if self.bias_zero:
            for lin in (self.W1, self.W2, self.W3, self.out):
                torch.nn.init.constant_(lin.bias.data, 0.0)
        if self.tied_emb:
            self.out.weight = self.embedding.weight
        self._keys_cache = None

    def _keys_for(self, ctx_sm: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
        """attn_e(ctx) is step-invariant: compute once per encoder output (the reference redoes it every step)."""
        tag = (src.data_ptr(), src._version, tuple(src.shape), self.attn.attn_e.weight._version)
        if self._keys_cache is None or self._keys_cache[0] != tag:
            self._keys_cache = (tag, ops.linear(ctx_sm, self.attn.attn_e.weight))
        return self._keys_cache[1]

    def forward(self, word_input, last_hidden, encoder_outputs, ctx_mask=None):
        """word_input [B] or [B,1]; last_hidden [1,B,H]; encoder_outputs [T,B,C]; ctx_mask [T,B]
        → (log-probabilities [B, V], hidden [1, B, H])."""
        _no_training_dropout(self, self.dropout_out, self.dropout_emb)
        ctx = _sentence_major(encoder_outputs)
        keys = self._keys_for(ctx, encoder_outputs)
        B, T, _ = ctx.shape
        mask = (ctx_mask.transpose(0, 1).contiguous() if ctx_mask is not None
                else torch.ones(B, T, dtype=torch.float32, device=ctx.device))
        w = ops.decoder_weights(self)
        logp, h, _ = ops.decoder_step(w, word_input, last_hidden, keys, ctx, mask, 1, want_logp=True)
        return logp, h.unsqueeze(0)


class ImagineAttn(nn.Module):
    """Image-conditioned attention over the encoder states.  layers/VSE_Imagine_Enc.py:9-79."""

    def __init__(self, method, context_size, shared_embedding_size):
        super().__init__()
        self.method = method
        self.embedding_size = shared_embedding_size
        self.context_size = context_size
        self.mid_dim = self.context_size
        self.ctx2ctx = nn.Linear(self.context_size, self.context_size, bias=False)
        self.emb2ctx = nn.Linear(self.embedding_size, self.context_size, bias=False)
        if self.method == 'mlp':
            self.mlp = nn.Linear(self.mid_dim, 1, bias=False)
        elif self.method != 'dot':
            raise ValueError("imagine_attn must be 'dot' or 'mlp'")

    def forward(self, image_vec, decoder_hidden, ctx_mask=None):
        """image_vec [B, S], decoder_hidden [T, B, C], ctx_mask [T, B] → weights [B, 1, T]."""
        ctx = _sentence_major(decoder_hidden)
        pk = ops.linear(ctx, self.ctx2ctx.weight)
        iq = ops.linear(image_vec, self.emb2ctx.weight)
        mask = ctx_mask.transpose(0, 1).contiguous() if ctx_mask is not None else None
        if self.method == 'dot':
            _, beta = ops.attention(iq, pk, ctx, None, mask, 1, ops.ATTN_DOT)
        else:
            _, beta = ops.attention(iq, pk, ctx, self.mlp.weight.reshape(-1), mask, 1, ops.ATTN_MLP)
        return beta.unsqueeze(1)


class VSE_Imagine_Enc(nn.Module):
    """Visual-attention text pooling (+ optional ranking loss).  layers/VSE_Imagine_Enc.py:81-185."""

    def __init__(self, attn_type, im_size, hidden_size, shared_embedding_size, dropout_im_emb=0.0, dropout_txt_emb=0.0,
                 activation_vse=True):
        super().__init__()
        self.attn_type = attn_type
        self.im_size = im_size
        self.hidden_size = hidden_size
        self.shared_embedding_size = shared_embedding_size
        # the reference overwrites both rates with 0.0 (VSE_Imagine_Enc.py:95-96): the flags are dead
        self.dropout_im_emb = 0.0
        self.dropout_txt_emb = 0.0
        self.activation_vse = activation_vse
        self.imagine_attn = ImagineAttn(self.attn_type, self.hidden_size, self.shared_embedding_size)
        self.im_embedding = nn.Linear(self.im_size, self.shared_embedding_size)
        self.text_embedding = nn.Linear(self.hidden_size, self.shared_embedding_size)

    def pool_sentence_major(self, im_var, ctx, mask, want_beta=False):
        """ctx [B,T,C], mask [B,T] → im_emb, txt_emb, ctx_vec, beta"""
        im = im_var.to(device=ctx.device, dtype=torch.float32)
        return ops.vse_pool_fwd(ops.vse_weights(self), im, ctx, mask, want_beta)

    def forward(self, im_var, decoder_hiddens, criterion_vse=None, context_mask=None):
        """→ (loss_vse, context_vec [B, C]) like VSE_Imagine_Enc.py:110-152."""
        ctx = _sentence_major(decoder_hiddens)
        mask = self._mask(context_mask, ctx)
        im_emb, txt_emb, ctx_vec, _ = self.pool_sentence_major(im_var, ctx, mask)
        loss_vse = 0
        if criterion_vse is not None:
            loss_vse = criterion_vse(im_emb, txt_emb)
        return loss_vse, ctx_vec

    def get_emb_vec(self, im_var, decoder_hiddens, ctx_mask=None):
        ctx = _sentence_major(decoder_hiddens)
        im_emb, txt_emb, _, _ = self.pool_sentence_major(im_var, ctx, self._mask(ctx_mask, ctx))
        return im_emb, txt_emb

    def get_imagine_weights(self, im_var, decoder_hiddens, ctx_mask=None):
        ctx = _sentence_major(decoder_hiddens)
        _, _, _, beta = self.pool_sentence_major(im_var, ctx, self._mask(ctx_mask, ctx), want_beta=True)
        return beta.unsqueeze(1)

    @staticmethod
    def _mask(ctx_mask: Optional[torch.Tensor], ctx: torch.Tensor) -> torch.Tensor:
        if ctx_mask is None:
            return torch.ones(ctx.shape[0], ctx.shape[1], dtype=torch.float32, device=ctx.device)
        return ctx_mask.transpose(0, 1).contiguous()
