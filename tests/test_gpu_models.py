"""Model-level GPU parity: the drop-in classes against (a) the committed golden fixtures produced by the real
reference (oracle/make_golden.py) and (b) the CPU oracle on fresh seeded inputs.

Bars: decoded token sequences bit-exact; FP32 scalars (losses) within 1e-4 relative; FP32 tensors within
1e-4 of max|ref| (the reference itself differs by 5e-7 between its FP32 and FP64 runs on these inputs).
"""
import json

import pytest
import torch

from conftest import build_mm, build_tm, cpu_params, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _crit(V, dev):
    w = torch.ones(V)
    w[0] = 0
    return torch.nn.NLLLoss(weight=w.to(dev), reduce=False)


def _load_into(model, sd):
    model.load_state_dict(sd)
    return model.cuda().eval()


@pytest.mark.parametrize("attn", ["dot", "mlp"])
def test_tiny_against_reference_golden(attn, tiny_dot, tiny_mlp):
    import vag_nmt_b200 as vag
    fix = tiny_dot if attn == "dot" else tiny_mlp
    cfg, b, ref, ref64 = fix["cfg"], fix["batch"], fix["ref_fp32"], fix["ref_fp64"]
    mm = _load_into(build_mm(cfg, 0, attn_model=attn), fix["params_mm"])
    tm = _load_into(build_tm(cfg, 0), fix["params_tm"])
    src, lens, tgt, im = b["src"], b["src_lengths"], b["tgt"], b["im"]

    ctx, mask = mm.encoder(src, lens)
    assert ctx.shape == ref["enc_ctx"].shape and mask.shape == ref["enc_mask"].shape
    assert rel_err(ctx, ref64["enc_ctx"]) < TOL
    assert torch.equal(mask.cpu(), ref["enc_mask"])
    pads = ctx.cpu()[ref["enc_mask"] == 0]
    assert pads.numel() > 0 and float(pads.abs().max()) == 0.0             # exact zeros at pads
    loss_vse, ctx_vec = mm.vse_imagine(im.cuda(), ctx, criterion_vse=vag.PairwiseRankingLoss(margin=0.1), context_mask=mask)
    assert rel_err(ctx_vec, ref64["vse_ctx_vec"]) < TOL
    assert abs(float(loss_vse) - float(ref64["loss_pairwise"])) < TOL * abs(float(ref64["loss_pairwise"]))
    im_emb, txt_emb = mm.vse_imagine.get_emb_vec(im.cuda(), ctx, ctx_mask=mask)
    assert rel_err(im_emb, ref64["vse_im_emb"]) < TOL and rel_err(txt_emb, ref64["vse_txt_emb"]) < TOL
    beta = mm.vse_imagine.get_imagine_weights(im.cuda(), ctx, ctx_mask=mask)
    assert rel_err(beta.squeeze(1), ref64["vse_beta"]) < TOL
    l_ir = vag.ImageRetrievalRankingLoss(margin=0.1)(im_emb, txt_emb)
    assert abs(float(l_ir) - float(ref64["loss_imageretrieval"])) < TOL * abs(float(ref64["loss_imageretrieval"]))

    # per-step decoder API (NMT_Decoder.forward) on the reference's layouts
    h0 = ref["h0"].cuda().unsqueeze(0)
    tok = torch.full((src.shape[0],), 2, dtype=torch.long).cuda()
    logp, h1 = mm.decoder(tok, h0, ctx, ctx_mask=mask)
    assert h1.shape == (1, src.shape[0], cfg["hidden_size"])
    assert float((logp.cpu().double() - ref64["step0_logp"]).abs().max()) < 2e-5
    assert rel_err(h1.squeeze(0), ref64["step0_h"]) < TOL
    logp2, h2 = mm.decoder(ref["step1_tok"].cuda().unsqueeze(1), h1, ctx, ctx_mask=mask)
    assert float((logp2.cpu().double() - ref64["step1_logp"]).abs().max()) < 2e-5
    assert rel_err(h2.squeeze(0), ref64["step1_h"]) < TOL
    att = mm.decoder.attn(h0, ctx, ctx_mask=mask)
    assert att.shape == (src.shape[0], 1, ctx.shape[0])

    # training forward: teacher forced / free running, multimodal + text-only, both ranking losses
    crit = _crit(cfg["tgt_size"], "cuda")
    for name, tf in (("tf", 1.0), ("free", 0.0)):
        out = mm(src, lens, tgt, im, tf, criterion_mt=crit, criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
        got = torch.stack([x.reshape(()) for x in out]).cpu().double()
        assert float(((got - ref64[f"fwd_{name}"]).abs() / ref64[f"fwd_{name}"].abs()).max()) < TOL, name
        lt = tm(src, lens, tgt, tf, criterion=crit)
        assert abs(float(lt) - float(ref64[f"fwd_text_{name}"])) < TOL * abs(float(ref64[f"fwd_text_{name}"]))
    out = mm(src, lens, tgt, im, 1.0, criterion_mt=crit, criterion_vse=vag.ImageRetrievalRankingLoss(margin=0.1))
    got = torch.stack([x.reshape(()) for x in out]).cpu().double()
    assert float(((got - ref64["fwd_tf_imageretrieval"]).abs() / ref64["fwd_tf_imageretrieval"].abs()).max()) < TOL

    # decoding: greedy and beams, token-exact against the reference
    for K in fix["beams"]:
        assert mm.beamsearch_decode(src, lens, im, beam_size=K, max_length=fix["max_length"]) == ref[f"decode_k{K}"], K
        assert tm.beamsearch_decode(src, lens, beam_size=K, max_length=fix["max_length"]) == ref[f"decode_text_k{K}"], K
    e_im, e_txt = mm.embed_sent_im_test(src, lens, im)
    assert rel_err(e_im, ref64["embed_im"]) < TOL and rel_err(e_txt, ref64["embed_txt"]) < TOL
    assert list(vag.t2i(e_im, e_txt)) == ref["t2i"] and list(vag.i2t(e_im, e_txt)) == ref["i2t"]


def test_full_de_b32_against_reference_golden(full_de):
    """EN→DE shapes, B=32 (BASELINE configs[0] at beam 5, and beam 12): weights regenerated from the seed."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    fix = full_de
    cfg = fix["cfg"]
    mm = build_mm(cfg, fix["seed"])
    tm = build_tm(cfg, fix["seed"] + 1)
    for model, key in ((mm, "mm"), (tm, "tm")):      # the regenerated weights are the ones the golden run used
        for k, v in model.state_dict().items():
            s, a = fix["param_checksums"][key][k]
            assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
    mm, tm = mm.cuda(), tm.cuda()
    batch = synthetic.make_batch(fix["batch_size"], cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=fix["data_seed"])
    src, lens, tgt, im = batch.src, batch.src_lengths, batch.tgt, batch.im
    ref, ref64, probes = fix["ref_fp32"], fix["ref_fp64"], fix["probes_fp64"]
    assert all(fix["fp32_fp64_token_agreement"].values())   # the golden tokens are stable under the reference's own precision

    def probe_ok(t, name, tol=TOL):
        p = probes[name]
        assert list(t.shape) == p["shape"], name
        got = t.detach().cpu().double().reshape(-1)[p["idx"]]
        scale = max(float(p["vals"].abs().max()), 1e-30)
        assert float((got - p["vals"]).abs().max()) / scale < tol, name
        assert abs(float(t.detach().cpu().double().norm()) - p["l2"]) < tol * p["l2"], name

    ctx, mask = mm.encoder(src, lens)
    probe_ok(ctx, "enc_ctx")
    e_im, e_txt = mm.embed_sent_im_test(src, lens, im)
    probe_ok(e_im, "embed_im")
    probe_ok(e_txt, "embed_txt")
    assert list(vag.t2i(e_im, e_txt)) == ref["t2i"] and list(vag.i2t(e_im, e_txt)) == ref["i2t"]   # recall@1/5/10 exact

    crit = _crit(cfg["tgt_size"], "cuda")
    for name, tf in (("tf", 1.0), ("free", 0.0)):
        out = mm(src, lens, tgt, im, tf, criterion_mt=crit, criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
        got = torch.stack([x.reshape(()) for x in out]).cpu().double()
        assert float(((got - ref64[f"fwd_{name}"]).abs() / ref64[f"fwd_{name}"].abs()).max()) < TOL, name
        lt = tm(src, lens, tgt, tf, criterion=crit)
        assert abs(float(lt) - float(ref64[f"fwd_text_{name}"])) < TOL * abs(float(ref64[f"fwd_text_{name}"]))

    for K in fix["beams"]:   # 1 (greedy), 5 (configs[0]), 12 (headline)
        got = mm.beamsearch_decode(src, lens, im, beam_size=K, max_length=fix["max_length"])
        assert got == ref[f"decode_k{K}"], f"multimodal beam {K}"
        got = tm.beamsearch_decode(src, lens, beam_size=K, max_length=fix["max_length"])
        assert got == ref[f"decode_text_k{K}"], f"text-only beam {K}"


def test_beam_known_answers(beam_kat):
    """SURVEY appendix A known answers through the real fused loop: a decoder whose weights make
    logp = LP[prev_token] (one-hot embedding, identity W3, out.weight = LPᵀ)."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import ops
    LP = torch.tensor(beam_kat["P"]).log()
    torch.manual_seed(0)
    m = vag.NMT_Seq2Seq_Beam_V2(8, 8, 8, 8, 8, tied_emb=False)
    with torch.no_grad():
        m.decoder.embedding.weight.copy_(10.0 * torch.eye(8))   # tanh(10) == 1.0f
        m.decoder.W3.weight.copy_(torch.eye(8))
        m.decoder.W1.weight.zero_()
        m.decoder.W2.weight.zero_()
        m.decoder.out.weight.copy_(LP.t())                      # logits[v] = LP[prev, v]
        m.decoder.out.bias.zero_()
    m = m.cuda().eval()
    w = ops.decoder_weights(m.decoder, m.decoderini)
    ctx = torch.zeros(2, 3, 16).cuda()
    mask = torch.ones(2, 3).cuda()
    keys = ops.attn_keys(w, ctx)
    h0 = torch.zeros(2, 8).cuda()
    for case in beam_kat["cases"]:
        hyp, hyp_len, beam, nll, steps = ops.beam_decode(w, h0, keys, ctx, mask, case["K"], case["L"], debug=True)
        got = [hyp[b, :int(hyp_len[b])].tolist() for b in range(2)]
        assert got == case["expected"], (case, got, beam.tolist())
    # (K=2, L=6): every beam has emitted EOS after 3 steps → the device-side early stop must fire (V11:268-269)
    hyp, hyp_len, beam, nll, steps = ops.beam_decode(w, h0, keys, ctx, mask, 2, 6, debug=True)
    assert int(steps) == 3
    assert beam[3:5].abs().sum().item() == 0 and (beam[5] == 3).all()   # rows after the break stay 0, last row forced EOS


def test_fresh_seeds_against_oracle():
    """Unsorted-order corpus → pad_and_sort → decode, on seeds the fixtures do not cover; oracle in FP32 on CPU."""
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY)
    for seed in (5, 6):
        mm = build_mm(cfg, seed).cuda()
        p = cpu_params(mm)
        sents, im = synthetic.make_corpus(9, cfg["src_size"], cfg["im_feats_size"], seed=seed, max_len=11, min_len=1, mean=5.0, std=3.0)
        src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
        for K in (1, 2, 5):
            assert mm.beamsearch_decode(src, lens, im_s, beam_size=K, max_length=15) == \
                O.multimodal_beamsearch_decode(p, src, lens, im_s, K, 15)


@pytest.mark.parametrize("precision,n_sent", [("fp32", 300), ("bf16", 300), ("fp32", 100)])
def test_encoder_fused_cell_path_ragged_batch(precision, n_sent):
    """More than 32 active rows: the encoder runs a GRU step as ONE launch (cell fused into the hidden contraction, two state
    buffers); the ragged batch crosses the 32-row boundary in both directions (forward: fused steps, then the three-kernel
    steps on the current buffer; backward: the reverse, with sentences joining at their last token on a zero state).
    Against the CPU oracle of Encoder.py:36-65: values, exact zeros at the padded positions, mask."""
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.DE)
    from vag_nmt_b200 import _cabi
    mm = build_mm(cfg, 3).cuda()
    p = cpu_params(mm)
    lens = sorted([2 + (7 * i) % 23 for i in range(n_sent)], reverse=True)    # 2..24 tokens: every row active at t = 0, fewer than
    g = torch.Generator().manual_seed(5)                                      # 128 / 32 rows towards the end (three-kernel steps)
    src = torch.zeros(n_sent, lens[0], dtype=torch.int64)
    for b, n in enumerate(lens):
        src[b, :n] = torch.randint(4, cfg["src_size"], (n,), generator=g)
    with _cabi.precision_scope(precision):
        ctx, mask = mm.encoder(src.cuda(), lens)                              # reference layout [T, B, 2H]
    if precision == "bf16":
        O.set_operand_rounding("bf16")
    try:
        want, want_mask = O.encoder_forward(p, src, lens)
    finally:
        O.set_operand_rounding(None)
    assert torch.equal(mask.cpu(), want_mask)
    got = ctx.cpu()
    assert rel_err(got, want) < (2e-5 if precision == "fp32" else 2e-3)
    for b in (0, 57, n_sent // 2, n_sent - 1):
        assert got[lens[b]:, b].abs().sum().item() == 0.0                     # pad_packed_sequence: exact zeros


def test_decode_160_sentences_token_exact_vs_oracle():
    """1920 rows at full Multi30K shapes: the encoder's fused-cell steps (more than 32 active rows), the tensor-core decoder step with
    every fused kernel, 6-row attention CTAs, the summary-based selection — token for token against the CPU oracle (beam 12, 10
    steps), multimodal and text-only."""
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.DE)
    sents, im = synthetic.make_corpus(160, cfg["src_size"], cfg["im_feats_size"], seed=41)
    src, lens, im_s, _ = synthetic.pad_and_sort(sents, im)
    mm = build_mm(cfg, 77).cuda().eval()
    assert mm.beamsearch_decode(src, lens, im_s, beam_size=12, max_length=10) == \
        O.multimodal_beamsearch_decode(cpu_params(mm), src, lens, im_s, 12, 10)
    tm = build_tm(cfg, 78).cuda().eval()
    assert tm.beamsearch_decode(src, lens, beam_size=12, max_length=10) == O.text_beamsearch_decode(cpu_params(tm), src, lens, 12, 10)


def test_error_behaviour():
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY)
    mm = build_mm(cfg, 1).cuda()
    src = torch.tensor([[5, 6, 3, 0], [5, 6, 7, 3]])
    with pytest.raises(RuntimeError):            # unsorted lengths: pack_padded_sequence raises in the reference
        mm.encoder(src, [3, 4])
    cpu_model = build_mm(cfg, 1)
    with pytest.raises(RuntimeError):            # no CPU path
        cpu_model.beamsearch_decode(src.flip(0), [4, 3], torch.rand(2, cfg["im_feats_size"]), beam_size=2, max_length=4)


@pytest.mark.parametrize("env", [{}, {"VAG_SELECT_RECOMPUTE": "1"}, {"VAG_KEEP_LOGITS": "1"}])
def test_fused_step_selection_variants_token_exact(env, monkeypatch, full_de):
    """The fused beam loop selects from per-tile top-2 summaries without logits; VAG_SELECT_RECOMPUTE=1 forces the rare
    'third candidate of a tile' path (tile recomputed from the operand planes) on every refill, VAG_KEEP_LOGITS=1 keeps
    the logits + rescan selection.  All three must reproduce the reference's tokens (golden fixture, beams 5 and 12)
    and, at a batch large enough for step 0 to run fused as well (B = 160 > 128), the CPU oracle's."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    fix = full_de
    cfg = fix["cfg"]
    mm = build_mm(cfg, fix["seed"]).cuda().eval()
    batch = synthetic.make_batch(fix["batch_size"], cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=fix["data_seed"])
    for K in (5, 12):
        got = mm.beamsearch_decode(batch.src, batch.src_lengths, batch.im, beam_size=K, max_length=fix["max_length"])
        assert got == fix["ref_fp32"][f"decode_k{K}"], f"beam {K} differs under {env}"
    big = synthetic.make_batch(160, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=21)
    got = mm.beamsearch_decode(big.src, big.src_lengths, big.im, beam_size=12, max_length=12)
    with torch.no_grad():
        want = O.multimodal_beamsearch_decode(cpu_params(mm), big.src, big.src_lengths, big.im, 12, 12)
    assert got == want


def test_reference_whole_module_checkpoint_decodes_like_the_reference(tiny_dot):
    """A torch.save(model) file written by the real reference (nmt_multimodal_beam_DE.py:491-519 format), loaded through
    checkpoint_compat with the reference package absent, reproduces the reference's tokens and losses."""
    import vag_nmt_b200 as vag
    from conftest import GOLD
    from vag_nmt_b200.checkpoint_compat import load_reference_module
    fix = tiny_dot
    b, ref, ref64 = fix["batch"], fix["ref_fp32"], fix["ref_fp64"]
    mm = load_reference_module(GOLD / "ref_module_tiny_mm.pt", "cuda")
    tm = load_reference_module(GOLD / "ref_module_tiny_tm.pt", "cuda")
    for K in fix["beams"]:
        assert mm.beamsearch_decode(b["src"], b["src_lengths"], b["im"], beam_size=K, max_length=fix["max_length"]) == ref[f"decode_k{K}"]
        assert tm.beamsearch_decode(b["src"], b["src_lengths"], beam_size=K, max_length=fix["max_length"]) == ref[f"decode_text_k{K}"]
    out = mm(b["src"], b["src_lengths"], b["tgt"], b["im"], 1.0, criterion_mt=_crit(fix["cfg"]["tgt_size"], "cuda"),
             criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
    got = torch.stack([x.reshape(()) for x in out]).cpu().double()
    assert float(((got - ref64["fwd_tf"]).abs() / ref64["fwd_tf"].abs()).max()) < TOL
