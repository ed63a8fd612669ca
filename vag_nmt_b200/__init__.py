"""vag_nmt_b200 — B200-native (sm_100a) implementation of VAG-NMT's per-timestep translation hot path.

Public API = the reference's own model/loss classes:

    from vag_nmt_b200 import NMT_AttentionImagine_Seq2Seq_Beam_V11, NMT_Seq2Seq_Beam_V2
    from vag_nmt_b200 import PairwiseRankingLoss, ImageRetrievalRankingLoss, t2i, i2t

All arithmetic runs in hand-written CUDA kernels behind the C ABI of ``include/vag_nmt.h``
(``vag_nmt_b200/csrc/libvagnmt.so``).  There is no CPU or torch-operator fallback: without the built
library and a B200 every operator raises.
"""
from .models import (EOS_token, SOS_token, UNK_token, NMT_AttentionImagine_Seq2Seq_Beam_V11, NMT_Seq2Seq_Beam_V2)
from .losses import ImageRetrievalRankingLoss, PairwiseRankingLoss
from .layers import BahdanauAttn, ImagineAttn, LIUMCVC_Encoder, NMT_Decoder, VSE_Imagine_Enc
from .retrieval import i2t, t2i

__version__ = "0.1.0"
