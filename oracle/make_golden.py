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


def main():
    from vag_nmt_b200 import synthetic
    ref = import_reference()
    GOLD.mkdir(parents=True, exist_ok=True)

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
    mm64, tm64 = mm.double(), tm.double()
    r64 = ref_outputs(ref, mm64, tm64, batch, cfg, beams, L)
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
                    ref_fp32=small, ref_fp64=small64, probes_fp64=probes, fp32_fp64_token_agreement=stable),
               GOLD / "full_de_b32.pt")

    # ---------------- beam-search known-answer cases
    kat = beam_kat(ref)
    (GOLD / "beam_kat.json").write_text(json.dumps(kat, indent=1))
    print("beam KATs:", [(c["K"], c["L"], c["expected"][0]) for c in kat["cases"]])
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
