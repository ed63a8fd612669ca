"""Generate tests/golden/*.pt from the REAL reference code and check the oracle restatement against it.

Runs only in the build container (needs /root/reference).  The reference is copied to a temp dir and given the
three-pattern torch-2.x compat patch of SURVEY.md appendix B (`.byte()`→`.bool()`, integer `/`→`//`,
hard-coded `.cuda()`→`.to(scores.device)`); no semantic change.  Nothing here is imported by the product.

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference deviations
"""
from __future__ import annotations

import json
import os
import re
import shutil
import sys
import tempfile
import warnings
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path(os.environ.get("VAG_REFERENCE", "/root/reference"))
GOLD = ROOT / "tests" / "golden"

warnings.filterwarnings("ignore")


def import_reference():
    tmp = Path(tempfile.mkdtemp(prefix="vag_ref_"))
    shutil.copytree(REF / "machine_translation_vision", tmp / "machine_translation_vision")
    pk = tmp / "machine_translation_vision"

    def sub(path, pairs):
        s = path.read_text()
        for a, b in pairs:
            s = s.replace(a, b)
        path.write_text(s)

    for f in ("layers/NMT_Decoder.py", "layers/VSE_Imagine_Enc.py"):
        sub(pk / f, [(".byte()", ".bool()")])
    for f in ("models/NMT_AttentionImagine_Seq2Seq_Beam_V11.py", "models/NMT_Seq2Seq_Beam_V2.py"):
        sub(pk / f, [("(nk_mask/beam_size)*beam_size", "(nk_mask//beam_size)*beam_size"),
                     ("tile = nk_mask / beam_size", "tile = nk_mask // beam_size"),
                     ("pdxs = idxs / n_vocab", "pdxs = idxs // n_vocab")])
    for f in (pk / "losses").glob("*.py"):
        sub(f, [(".cuda()),", ".to(scores.device)),")])
    sys.path.insert(0, str(tmp))
    import machine_translation_vision  # noqa: F401
    from machine_translation_vision.models.NMT_AttentionImagine_Seq2Seq_Beam_V11 import NMT_AttentionImagine_Seq2Seq_Beam_V11
    from machine_translation_vision.models.NMT_Seq2Seq_Beam_V2 import NMT_Seq2Seq_Beam_V2
    from machine_translation_vision.losses.PairwiseRankingLoss import PairwiseRankingLoss
    from machine_translation_vision.losses.ImageRetrievalRankingLoss import ImageRetrievalRankingLoss
    from machine_translation_vision.utils.im_retrieval_eval import t2i, i2t
    return dict(V11=NMT_AttentionImagine_Seq2Seq_Beam_V11, V2=NMT_Seq2Seq_Beam_V2, Pairwise=PairwiseRankingLoss,
                ImageRetrieval=ImageRetrievalRankingLoss, t2i=t2i, i2t=i2t)



class InjectedDropout(torch.nn.Module):
    """Stands in for an nn.Dropout INSTANCE of the reference model (the reference's sources stay untouched): in training mode
    multiplies by pre-drawn masks in call order (one mask per call, wrapping around), identity in eval mode."""

    def __init__(self, masks):
        super().__init__()
        self.masks, self.calls = list(masks), 0

    def forward(self, x):
        if not self.training:
            return x
        m = self.masks[self.calls % len(self.masks)]
        self.calls += 1
        assert m.shape == x.shape, (m.shape, x.shape)
        return x * m.to(x.dtype)


def inject_dropout(mm, masks, B, Ts, Tt):
    """masks in the drop-in's layouts (vag_nmt_b200.synthetic.dropout_masks) → the reference's call shapes: embedding dropout
    sees [T, B, E] (Encoder.py:50-52), context dropout [T, B, 2H] (:62-64), output dropout one [B, E] per step (NMT_Decoder.py:140-141)."""
    if "emb" in masks:
        mm.encoder.embedding_dropout = InjectedDropout([masks["emb"].reshape(Ts, B, -1)])
    if "ctx" in masks:
        mm.encoder.context_dropout = InjectedDropout([masks["ctx"].transpose(0, 1).contiguous()])
    if "out" in masks:
        mm.decoder.output_dropout = InjectedDropout(list(masks["out"].reshape(Tt, B, -1)))


def grad_probe(g):
    """Compact record of one gradient tensor: norm, sum, and 64 probed entries (the 32 largest + 32 evenly spaced)."""
    flat = g.detach().double().reshape(-1)
    top = flat.abs().topk(min(32, flat.numel())).indices
    idx = torch.cat([top, torch.linspace(0, flat.numel() - 1, 32).long()])
    return dict(shape=list(g.shape), l2=float(flat.norm()), sum=float(flat.sum()), idx=idx, vals=flat[idx].clone())


def ref_train_gradients(ref, mm, batch, cfg, margin=0.1, masks=None, vse="Pairwise"):
    """loss.backward() through the REAL reference in training mode (teacher forced): the three losses + a probe of every
    parameter gradient.  With `masks` the model's dropout instances are replaced by InjectedDropout."""
    src, lens, tgt, im = batch.src, batch.src_lengths, batch.tgt, batch.im
    dt = mm.decoderini.weight.dtype
    vocab_mask = torch.ones(cfg["tgt_size"], dtype=dt)
    vocab_mask[0] = 0
    crit_mt = torch.nn.NLLLoss(weight=vocab_mask, reduce=False)
    mm.train()
    if masks:
        inject_dropout(mm, masks, src.shape[0], src.shape[1], tgt.shape[1])
    mm.zero_grad()
    l, lmt, lv = mm(src, lens, tgt, im.to(dt), 1.0, criterion_mt=crit_mt, criterion_vse=ref[vse](margin=margin))
    l.backward()
    out = dict(losses=torch.stack([l.detach(), lmt.detach(), lv.detach()]).clone(),
               grads={n: grad_probe(p.grad) for n, p in mm.named_parameters()})
    full = {n: p.grad.detach().clone() for n, p in mm.named_parameters()}
    mm.zero_grad()
    mm.eval()
    return out, full


def oracle_train_gradients(sd, batch, cfg, margin=0.1, masks=None):
    """The same through autograd over oracle/vag_oracle.py (validates the restatement's training path here)."""
    from oracle import vag_oracle as O
    dt = sd["decoderini.weight"].dtype
    p = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    p["decoder.out.weight"] = p["decoder.embedding.weight"]
    w = torch.ones(cfg["tgt_size"], dtype=dt)
    w[0] = 0
    dm = {k: v.to(dt) for k, v in masks.items()} if masks else None
    l, lmt, lv = O.multimodal_forward(p, batch.src, batch.src_lengths, batch.tgt, batch.im.to(dt), True, w, "pairwise", margin,
                                      dropout_masks=dm)
    l.backward()
    return torch.stack([l.detach(), lmt.detach(), lv.detach()]), {k: v.grad for k, v in p.items() if v.grad is not None}


def compare_grads(ref_full, ora_full, label, tol):
    worst = 0.0
    for n, g in ref_full.items():
        o = ora_full[n]
        err = float((g.double() - o.double()).norm() / g.double().norm().clamp(min=1e-30))
        worst = max(worst, err)
        assert err < tol, f"{label}: gradient of {n} differs from the reference: {err:.3e}"
    print(f"  oracle autograd == reference autograd on {len(ref_full)} parameter gradients ({label}); worst ‖Δ‖/‖g‖ {worst:.2e}")


def param_checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}


def build_models(ref, cfg, seed, mirror=True):
    """reference V11 + V2 under the seed; optionally check that the drop-in's constructor draws the same weights."""
    import vag_nmt_b200 as vag
    torch.manual_seed(seed)
    mm = ref["V11"](cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"],
                    cfg["tgt_embedding_size"], cfg["hidden_size"], cfg["shared_embedding_size"], 0.99,
                    attn_model=cfg.get("attn_model", "dot"), tied_emb=True, init_split=0.5).eval()
    torch.manual_seed(seed + 1)
    tm = ref["V2"](cfg["src_size"], cfg["tgt_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
                   cfg["hidden_size"], tied_emb=True).eval()
    if mirror:
        torch.manual_seed(seed)
        mine = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(
            cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
            cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, attn_model=cfg.get("attn_model", "dot"), tied_emb=True,
            init_split=0.5)
        a, b = mm.state_dict(), mine.state_dict()
        assert list(a.keys()) == list(b.keys()), "state_dict keys differ from the reference"
        for k in a:
            assert torch.equal(a[k], b[k]), f"mirror init differs at {k}"
        torch.manual_seed(seed + 1)
        mine_t = vag.NMT_Seq2Seq_Beam_V2(cfg["src_size"], cfg["tgt_size"], cfg["src_embedding_size"],
                                         cfg["tgt_embedding_size"], cfg["hidden_size"], tied_emb=True)
        a, b = tm.state_dict(), mine_t.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.equal(a[k], b[k]), f"text-only mirror init differs at {k}"
        print(f"  drop-in constructors reproduce the reference's init bit-for-bit (seed {seed})")
    return mm, tm


def ref_outputs(ref, mm, tm, batch, cfg, beams, max_length, margin=0.1, with_parts=True):
    """Everything the parity tests compare, computed by the reference itself."""
    out = {}
    src, lens, tgt, im = batch.src, batch.src_lengths, batch.tgt, batch.im
    V = cfg["tgt_size"]
    vocab_mask = torch.ones(V, dtype=mm.decoderini.weight.dtype)
    vocab_mask[0] = 0
    crit_mt = torch.nn.NLLLoss(weight=vocab_mask, reduce=False)
    with torch.no_grad():
        ctx, mask = mm.encoder(src, lens)
        out["enc_ctx"], out["enc_mask"] = ctx.clone(), mask.clone()
        im_t = im.to(ctx.dtype)
        loss_vse, ctx_vec = mm.vse_imagine(im_t, ctx, criterion_vse=ref["Pairwise"](margin=margin), context_mask=mask)
        im_emb, txt_emb = mm.vse_imagine.get_emb_vec(im_t, ctx, ctx_mask=mask)
        out["vse_ctx_vec"], out["vse_im_emb"], out["vse_txt_emb"] = ctx_vec.clone(), im_emb.clone(), txt_emb.clone()
        out["vse_beta"] = mm.vse_imagine.get_imagine_weights(im_t, ctx, ctx_mask=mask).squeeze(1).clone()
        out["loss_pairwise"] = loss_vse.clone()
        out["loss_imageretrieval"] = ref["ImageRetrieval"](margin=margin)(im_emb, txt_emb).clone()
        h0 = torch.tanh(mm.decoderini(0.5 * ctx_vec + 0.5 * ctx.sum(0) / mask.sum(0).unsqueeze(1))).unsqueeze(0)
        out["h0"] = h0.squeeze(0).clone()
        if with_parts:
            tok = torch.full((src.shape[0],), 2, dtype=torch.long)
            logp, h1 = mm.decoder(tok, h0, ctx, ctx_mask=mask)
            out["step0_logp"], out["step0_h"] = logp.clone(), h1.squeeze(0).clone()
            tok2 = logp.argmax(1)
            logp2, h2 = mm.decoder(tok2, h1, ctx, ctx_mask=mask)
            out["step1_tok"], out["step1_logp"], out["step1_h"] = tok2.clone(), logp2.clone(), h2.squeeze(0).clone()
        # training forward, teacher forced and free running
        for name, tf in (("tf", 1.0), ("free", 0.0)):
            l, lmt, lv = mm(src, lens, tgt, im_t, tf, criterion_mt=crit_mt, criterion_vse=ref["Pairwise"](margin=margin))
            out[f"fwd_{name}"] = torch.stack([l, lmt, lv]).clone()
            out[f"fwd_text_{name}"] = tm(src, lens, tgt, tf, criterion=crit_mt).clone()
        l, lmt, lv = mm(src, lens, tgt, im_t, 1.0, criterion_mt=crit_mt, criterion_vse=ref["ImageRetrieval"](margin=margin))
        out["fwd_tf_imageretrieval"] = torch.stack([l, lmt, lv]).clone()
        # decoding
        for K in beams:
            out[f"decode_k{K}"] = [[int(t) for t in s] for s in mm.beamsearch_decode(src, lens, im_t, beam_size=K, max_length=max_length)]
            out[f"decode_text_k{K}"] = [[int(t) for t in s] for s in tm.beamsearch_decode(src, lens, beam_size=K, max_length=max_length)]
        e_im, e_txt = mm.embed_sent_im_test(src, lens, im_t)
        out["embed_im"], out["embed_txt"] = e_im.clone(), e_txt.clone()
        out["t2i"] = [float(x) for x in ref["t2i"](e_im, e_txt)]
        out["i2t"] = [float(x) for x in ref["i2t"](e_im, e_txt)]
    return out


def oracle_outputs(sd_mm, sd_tm, batch, cfg, beams, max_length, margin=0.1, with_parts=True):
    """The same quantities from oracle/vag_oracle.py (used to validate the restatement here)."""
    from oracle import vag_oracle as O
    out = {}
    src, lens, tgt, im = batch.src, batch.src_lengths, batch.tgt, batch.im
    dt = sd_mm["decoderini.weight"].dtype
    im = im.to(dt)
    V = cfg["tgt_size"]
    w = torch.ones(V, dtype=dt)
    w[0] = 0
    method = cfg.get("attn_model", "dot")
    ctx, mask = O.encoder_forward(sd_mm, src, lens)
    out["enc_ctx"], out["enc_mask"] = ctx, mask
    im_emb, txt_emb, ctx_vec, beta = O.vse_pool(sd_mm, im, ctx, mask, method)
    out["vse_ctx_vec"], out["vse_im_emb"], out["vse_txt_emb"], out["vse_beta"] = ctx_vec, im_emb, txt_emb, beta
    out["loss_pairwise"] = O.pairwise_ranking_loss(im_emb, txt_emb, margin)
    out["loss_imageretrieval"] = O.image_retrieval_ranking_loss(im_emb, txt_emb, margin)
    h0 = O.decoder_init(sd_mm, ctx, mask, ctx_vec, 0.5)
    out["h0"] = h0
    if with_parts:
        tok = torch.full((src.shape[0],), 2, dtype=torch.long)
        logp, h1 = O.decoder_step(sd_mm, tok, h0, ctx, mask)
        out["step0_logp"], out["step0_h"] = logp, h1
        tok2 = logp.argmax(1)
        logp2, h2 = O.decoder_step(sd_mm, tok2, h1, ctx, mask)
        out["step1_tok"], out["step1_logp"], out["step1_h"] = tok2, logp2, h2
    for name, tf in (("tf", True), ("free", False)):
        out[f"fwd_{name}"] = torch.stack(O.multimodal_forward(sd_mm, src, lens, tgt, im, tf, w, "pairwise", margin, attn_model=method))
        out[f"fwd_text_{name}"] = O.text_forward(sd_tm, src, lens, tgt, tf, w)
    out["fwd_tf_imageretrieval"] = torch.stack(O.multimodal_forward(sd_mm, src, lens, tgt, im, True, w, "imageretrieval", margin, attn_model=method))
    for K in beams:
        out[f"decode_k{K}"] = O.multimodal_beamsearch_decode(sd_mm, src, lens, im, K, max_length, attn_model=method)
        out[f"decode_text_k{K}"] = O.text_beamsearch_decode(sd_tm, src, lens, K, max_length)
    e_im, e_txt = O.embed_sent_im(sd_mm, src, lens, im, method)
    out["embed_im"], out["embed_txt"] = e_im, e_txt
    out["t2i"] = [float(x) for x in O.t2i(e_im, e_txt)]
    out["i2t"] = [float(x) for x in O.i2t(e_im, e_txt)]
    return out


def compare(a, b, label):
    worst = 0.0
    for k in a:
        x, y = a[k], b[k]
        if isinstance(x, torch.Tensor):
            if x.dtype in (torch.long, torch.int32):
                assert torch.equal(x, y), f"{label}: {k} differs"
                continue
            denom = max(float(x.abs().max()), 1e-30)
            err = float((x - y).abs().max()) / denom
            worst = max(worst, err)
            tol = 2e-4 if x.dtype == torch.float32 else 1e-9
            assert err < tol, f"{label}: {k} rel err {err:.3e}"
        else:
            assert x == y, f"{label}: {k} differs:\n ref    {x}\n oracle {y}"
    print(f"  oracle == reference on {len(a)} quantities ({label}); worst float deviation {worst:.2e}")


def beam_kat(ref):
    """the reference's beamsearch driven by a table look-up decoder."""
    import math
    P = torch.full((8, 8), 0.02)
    for r, cols in {2: {4: .5, 5: .3, 6: .1}, 4: {4: .6, 5: .2, 3: .1}, 5: {3: .7, 6: .2}, 6: {7: .5, 3: .3}, 7: {3: .9},
                    3: {3: .9}}.items():
        for c, v in cols.items():
            P[r, c] = v
        P[r] /= P[r].sum()
    LP = P.log()
    model = ref["V2"](8, 8, 4, 4, 8, tied_emb=True).eval()

    class TableDecoder(torch.nn.Module):
        def forward(self, word_input, last_hidden, encoder_outputs, ctx_mask=None):
            return LP[word_input.view(-1)].clone(), last_hidden

    model.decoder = TableDecoder()
    cases = []
    for K, L in ((2, 6), (3, 6), (3, 3), (2, 2), (4, 7)):
        with torch.no_grad():
            res = model.beamsearch(torch.zeros(3, 2, 16), torch.ones(3, 2), torch.full((2, 1), 2, dtype=torch.long),
                                   torch.zeros(1, 2, 8), K, L)
        cases.append(dict(K=K, L=L, expected=[[int(t) for t in s] for s in res]))
    return dict(P=P.tolist(), cases=cases)



FR_DROPOUT = dict(dropout_emb=0.2, dropout_ctx=0.4, dropout_out=0.4, dropout_im_emb=0.2)   # nmt_multimodal_beam_FR.py:55-61
# data seed 12: the first of 11, 12, … whose beam-5 AND beam-12 tokens are the same under every arithmetic tried here (reference
# fp32 / fp64, oracle fp32 with and without hoisted keys, oracle fp64).  Seed 11 has one near-tie (sentence 28 at beam 5: the fp32
# oracle and the fp64 runs pick different hypotheses), i.e. tokens the reference itself only reproduces up to rounding order.
FR_SEED, FR_DATA_SEED, FR_MASK_SEED = 4321, 12, 5


def build_fr(ref, cfg, dtype=torch.float32):
    torch.manual_seed(FR_SEED)
    mm = ref["V11"](cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
                    cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, attn_model="dot", tied_emb=True, init_split=0.5,
                    **FR_DROPOUT).eval()
    return mm.to(dtype)


def fr_fixture(ref):
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.FR)
    print("full EN→FR config (V=8748, dropout 0.2/0.4/0.4), B=32")
    mm = build_fr(ref, cfg)
    torch.manual_seed(FR_SEED)
    mine = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(
        cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
        cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, tied_emb=True, init_split=0.5, **FR_DROPOUT)
    a, b = mm.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), f"FR mirror init differs at {k}"
    batch = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=FR_DATA_SEED)
    B, Ts = batch.src.shape
    Tt = batch.tgt.shape[1]
    masks = synthetic.dropout_masks(FR_MASK_SEED, B, Ts, Tt, cfg["src_embedding_size"], cfg["hidden_size"], cfg["tgt_embedding_size"],
                                    FR_DROPOUT["dropout_emb"], FR_DROPOUT["dropout_ctx"], FR_DROPOUT["dropout_out"])
    from oracle import vag_oracle as O
    out = {}
    for name, dt in (("fp32", torch.float32), ("fp64", torch.float64)):
        m = build_fr(ref, cfg, dt)
        vocab_mask = torch.ones(cfg["tgt_size"], dtype=dt)
        vocab_mask[0] = 0
        crit_mt = torch.nn.NLLLoss(weight=vocab_mask, reduce=False)
        im = batch.im.to(dt)
        with torch.no_grad():
            ev = torch.stack(m(batch.src, batch.src_lengths, batch.tgt, im, 1.0, criterion_mt=crit_mt,
                               criterion_vse=ref["Pairwise"](margin=0.1))).clone()
            toks = [[int(t) for t in s] for s in m.beamsearch_decode(batch.src, batch.src_lengths, im, beam_size=12, max_length=40)]
            toks5 = [[int(t) for t in s] for s in m.beamsearch_decode(batch.src, batch.src_lengths, im, beam_size=5, max_length=40)]
            e_im, e_txt = m.embed_sent_im_test(batch.src, batch.src_lengths, im)
            recall = [float(x) for x in ref["t2i"](e_im, e_txt)]
        tr, full = ref_train_gradients(ref, m, batch, cfg, masks=masks)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        ol, og = oracle_train_gradients(sd, batch, cfg, masks=masks)
        assert float(((ol - tr["losses"]).abs() / tr["losses"].abs()).max()) < (2e-5 if dt == torch.float32 else 1e-10)
        compare_grads(full, og, f"EN→FR {name}, dropout masks injected", 2e-4 if dt == torch.float32 else 1e-9)
        with torch.no_grad():
            for hoist in (False, True):
                assert O.multimodal_beamsearch_decode(sd, batch.src, batch.src_lengths, im, 12, 40, hoist_keys=hoist) == toks
                assert O.multimodal_beamsearch_decode(sd, batch.src, batch.src_lengths, im, 5, 40, hoist_keys=hoist) == toks5
        out[name] = dict(fwd_eval=ev, decode_k12=toks, decode_k5=toks5, t2i=recall, train=tr)
    print("  fp32 vs fp64 reference token agreement (beam 12 / 5):", out["fp32"]["decode_k12"] == out["fp64"]["decode_k12"],
          out["fp32"]["decode_k5"] == out["fp64"]["decode_k5"])
    assert out["fp32"]["decode_k12"] == out["fp64"]["decode_k12"] and out["fp32"]["decode_k5"] == out["fp64"]["decode_k5"]
    torch.save(dict(cfg=cfg, seed=FR_SEED, data_seed=FR_DATA_SEED, mask_seed=FR_MASK_SEED, dropout=FR_DROPOUT, batch_size=32,
                    max_length=40, param_checksums=param_checksums(build_fr(ref, cfg).state_dict()), ref_fp32=out["fp32"],
                    ref_fp64=out["fp64"], tokens_stable=out["fp32"]["decode_k12"] == out["fp64"]["decode_k12"]),
               GOLD / "full_fr_b32.pt")


def main():
    from vag_nmt_b200 import synthetic
    ref = import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)
    if "--only-fr" in sys.argv:
        fr_fixture(ref)
        return

    # ---------------- tiny config: full tensors, fp32 + fp64
    for attn in ("dot", "mlp"):
        cfg = dict(synthetic.TINY, attn_model=attn)
        print(f"tiny config ({attn})")
        mm, tm = build_models(ref, cfg, seed=11, mirror=(attn == "dot"))
        batch = synthetic.make_batch(6, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=3, max_len=9, min_len=2,
                                     mean=5.0, std=2.5, common_tgt_len=False)
        assert len(set(batch.src_lengths)) > 2 and (batch.tgt == 0).any()   # ragged source AND padded targets
        fix = dict(cfg=cfg, seed=11, batch=dict(src=batch.src, src_lengths=batch.src_lengths, tgt=batch.tgt, im=batch.im),
                   params_mm={k: v.clone() for k, v in mm.state_dict().items()},
                   params_tm={k: v.clone() for k, v in tm.state_dict().items()}, beams=[1, 3, 4], max_length=12)
        r32 = ref_outputs(ref, mm, tm, batch, cfg, fix["beams"], fix["max_length"])
        o32 = oracle_outputs(mm.state_dict(), tm.state_dict(), batch, cfg, fix["beams"], fix["max_length"])
        compare(r32, o32, "fp32")
        mm64, tm64 = mm.double(), tm.double()
        r64 = ref_outputs(ref, mm64, tm64, batch, cfg, fix["beams"], fix["max_length"])
        o64 = oracle_outputs(mm64.state_dict(), tm64.state_dict(), batch, cfg, fix["beams"], fix["max_length"])
        compare(r64, o64, "fp64")
        fix["ref_fp32"], fix["ref_fp64"] = r32, r64
        torch.save(fix, GOLD / f"tiny_{attn}.pt")
        if attn == "dot":
            # whole-module pickles the way the reference's driver writes its checkpoints (nmt_multimodal_beam_DE.py:491-519):
            # the fixtures of vag_nmt_b200.checkpoint_compat (the fp64 cast above is undone first: same weights as params_mm / params_tm)
            torch.save(mm.float(), GOLD / "ref_module_tiny_mm.pt")
            torch.save(tm.float(), GOLD / "ref_module_tiny_tm.pt")

    if "--only-tiny" in sys.argv:
        return
    # ---------------- full EN→DE shapes, B = 32: weights by seed, outputs (small) stored
    cfg = dict(synthetic.DE)
    print("full EN→DE config, B=32")
    mm, tm = build_models(ref, cfg, seed=1234)
    batch = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=7)
    beams, L = [1, 5, 12], 40
    r32 = ref_outputs(ref, mm, tm, batch, cfg, beams, L)
    o32 = oracle_outputs(mm.state_dict(), tm.state_dict(), batch, cfg, beams, L)
    compare(r32, o32, "fp32")
    checks = dict(mm=param_checksums(mm.state_dict()), tm=param_checksums(tm.state_dict()))
    g32, g32_full = ref_train_gradients(ref, mm, batch, cfg)
    compare_grads(g32_full, oracle_train_gradients(mm.state_dict(), batch, cfg)[1], "EN→DE fp32", 2e-4)
    mm64, tm64 = mm.double(), tm.double()
    r64 = ref_outputs(ref, mm64, tm64, batch, cfg, beams, L)
    g64, g64_full = ref_train_gradients(ref, mm64, batch, cfg)
    compare_grads(g64_full, oracle_train_gradients(mm64.state_dict(), batch, cfg)[1], "EN→DE fp64", 1e-9)
    stable = {k: r32[k] == r64[k] for k in r32 if k.startswith("decode")}
    print("  fp32 vs fp64 reference token agreement:", stable)
    keep = ("loss_pairwise", "loss_imageretrieval", "fwd_tf", "fwd_free", "fwd_text_tf", "fwd_text_free",
            "fwd_tf_imageretrieval", "t2i", "i2t", "step1_tok")
    small = {k: v for k, v in r32.items() if k.startswith("decode") or k in keep}
    small64 = {k: v for k, v in r64.items() if k.startswith("decode") or k in keep}
    # compact numeric probes of the big tensors (fp64 reference): first rows / norms
    probes = {}
    for k in ("enc_ctx", "vse_ctx_vec", "vse_im_emb", "vse_txt_emb", "h0", "step0_logp", "step0_h", "step1_logp", "step1_h",
              "embed_im", "embed_txt", "vse_beta"):
        t = r64[k]
        flat = t.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, 64).long()
        probes[k] = dict(shape=list(t.shape), l2=float(t.norm()), sum=float(t.sum()), idx=idx, vals=flat[idx].clone())
    torch.save(dict(cfg=cfg, seed=1234, data_seed=7, batch_size=32, beams=beams, max_length=L, param_checksums=checks,
                    ref_fp32=small, ref_fp64=small64, probes_fp64=probes, fp32_fp64_token_agreement=stable,
                    train_fp32=g32, train_fp64=g64),
               GOLD / "full_de_b32.pt")

    # ---------------- full EN→FR shapes (BASELINE configs[3]: V = 8748, dropout emb 0.2 / ctx 0.4 / out 0.4,
    #                  nmt_multimodal_beam_FR.py:55-67), B = 32: beam-12 tokens (eval mode), training-mode losses and every
    #                  parameter gradient under injected dropout masks, eval-mode forward
    fr_fixture(ref)

    # ---------------- beam-search known-answer cases
    kat = beam_kat(ref)
    (GOLD / "beam_kat.json").write_text(json.dumps(kat, indent=1))
    print("beam KATs:", [(c["K"], c["L"], c["expected"][0]) for c in kat["cases"]])
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
