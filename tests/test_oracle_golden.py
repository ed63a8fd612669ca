"""CPU: the oracle restatement against the fixtures the REAL reference produced (oracle/make_golden.py)."""
import math

import pytest
import torch

from oracle import vag_oracle as O


def _check(ref, got, tol):
    if isinstance(ref, torch.Tensor):
        if not ref.dtype.is_floating_point:
            assert torch.equal(ref, got)
            return
        scale = max(float(ref.abs().max()), 1e-30)
        assert float((ref - got).abs().max()) / scale < tol
    else:
        assert ref == got


@pytest.mark.parametrize("attn", ["dot", "mlp"])
@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_tiny_every_quantity(attn, prec, tiny_dot, tiny_mlp):
    fix = tiny_dot if attn == "dot" else tiny_mlp
    dt = torch.float32 if prec == "fp32" else torch.float64
    tol = 2e-5 if prec == "fp32" else 1e-10
    ref = fix["ref_" + prec]
    cfg, b = fix["cfg"], fix["batch"]
    p = {k: v.to(dt) for k, v in fix["params_mm"].items()}
    pt = {k: v.to(dt) for k, v in fix["params_tm"].items()}
    src, lens, tgt, im = b["src"], b["src_lengths"], b["tgt"], b["im"].to(dt)
    w = torch.ones(cfg["tgt_size"], dtype=dt)
    w[0] = 0
    ctx, mask = O.encoder_forward(p, src, lens)
    _check(ref["enc_ctx"], ctx, tol)
    _check(ref["enc_mask"], mask, tol)
    assert float(ctx[mask == 0].abs().max()) == 0.0
    im_emb, txt_emb, ctx_vec, beta = O.vse_pool(p, im, ctx, mask, attn)
    for k, v in (("vse_im_emb", im_emb), ("vse_txt_emb", txt_emb), ("vse_ctx_vec", ctx_vec), ("vse_beta", beta)):
        _check(ref[k], v, tol)
    _check(ref["loss_pairwise"], O.pairwise_ranking_loss(im_emb, txt_emb, 0.1), tol)
    _check(ref["loss_imageretrieval"], O.image_retrieval_ranking_loss(im_emb, txt_emb, 0.1), tol)
    h0 = O.decoder_init(p, ctx, mask, ctx_vec, 0.5)
    _check(ref["h0"], h0, tol)
    logp, h1 = O.decoder_step(p, torch.full((src.shape[0],), 2), h0, ctx, mask)
    _check(ref["step0_logp"], logp, tol)
    _check(ref["step0_h"], h1, tol)
    for name, tf in (("tf", True), ("free", False)):
        _check(ref[f"fwd_{name}"], torch.stack(O.multimodal_forward(p, src, lens, tgt, im, tf, w, "pairwise", 0.1, attn_model=attn)), tol)
        _check(ref[f"fwd_text_{name}"], O.text_forward(pt, src, lens, tgt, tf, w), tol)
    for K in fix["beams"]:
        assert O.multimodal_beamsearch_decode(p, src, lens, im, K, fix["max_length"], attn_model=attn) == ref[f"decode_k{K}"]
        assert O.multimodal_beamsearch_decode(p, src, lens, im, K, fix["max_length"], attn_model=attn, hoist_keys=True) == ref[f"decode_k{K}"]
        assert O.text_beamsearch_decode(pt, src, lens, K, fix["max_length"]) == ref[f"decode_text_k{K}"]
    e_im, e_txt = O.embed_sent_im(p, src, lens, im, attn)
    assert [float(x) for x in O.t2i(e_im, e_txt)] == ref["t2i"]
    assert [float(x) for x in O.i2t(e_im, e_txt)] == ref["i2t"]


def test_full_shapes_forward_and_beam5(full_de):
    """EN→DE shapes (BASELINE configs[0]: B=32, beam 5, CPU): oracle on seed-regenerated weights vs the reference."""
    from conftest import build_mm, cpu_params
    from vag_nmt_b200 import synthetic
    fix = full_de
    cfg = fix["cfg"]
    mm = build_mm(cfg, fix["seed"])
    for k, v in mm.state_dict().items():
        s, a = fix["param_checksums"]["mm"][k]
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
    p = cpu_params(mm)
    b = synthetic.make_batch(fix["batch_size"], cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=fix["data_seed"])
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    with torch.no_grad():
        got = torch.stack(O.multimodal_forward(p, b.src, b.src_lengths, b.tgt, b.im, True, w, "pairwise", 0.1))
        _check(fix["ref_fp32"]["fwd_tf"], got, 2e-5)
        assert O.multimodal_beamsearch_decode(p, b.src, b.src_lengths, b.im, 5, fix["max_length"], hoist_keys=True) == \
            fix["ref_fp32"]["decode_k5"]


def test_full_fr_shapes_losses_gradients_and_beam5():
    """EN→FR shapes (BASELINE configs[3]: V = 8748, dropout 0.2 / 0.4 / 0.4): the oracle — forward, autograd with the injected
    dropout masks, beam 5 — against the fixture the real reference produced (oracle/make_golden.py:fr_fixture)."""
    import vag_nmt_b200 as vag
    from conftest import cpu_params, load_golden
    from vag_nmt_b200 import synthetic
    fix = load_golden("full_fr_b32.pt")
    cfg = fix["cfg"]
    torch.manual_seed(fix["seed"])
    m = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"],
                                                  cfg["tgt_embedding_size"], cfg["hidden_size"], cfg["shared_embedding_size"], 0.99,
                                                  tied_emb=True, init_split=0.5, **fix["dropout"])
    for k, v in m.state_dict().items():
        s_, a = fix["param_checksums"][k]
        assert abs(float(v.double().sum()) - s_) <= 1e-9 * max(1.0, abs(a)), k
    b = synthetic.make_batch(fix["batch_size"], cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=fix["data_seed"])
    B, Ts = b.src.shape
    d = fix["dropout"]
    masks = synthetic.dropout_masks(fix["mask_seed"], B, Ts, b.tgt.shape[1], cfg["src_embedding_size"], cfg["hidden_size"],
                                    cfg["tgt_embedding_size"], d["dropout_emb"], d["dropout_ctx"], d["dropout_out"])
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    p = {k: v.clone().requires_grad_(True) for k, v in cpu_params(m).items()}
    p["decoder.out.weight"] = p["decoder.embedding.weight"]
    ref = fix["ref_fp32"]
    with torch.no_grad():
        _check(ref["fwd_eval"], torch.stack(O.multimodal_forward(p, b.src, b.src_lengths, b.tgt, b.im, True, w, "pairwise", 0.1)), 2e-5)
        assert O.multimodal_beamsearch_decode(p, b.src, b.src_lengths, b.im, 5, fix["max_length"], hoist_keys=True) == ref["decode_k5"]
    out = O.multimodal_forward(p, b.src, b.src_lengths, b.tgt, b.im, True, w, "pairwise", 0.1, dropout_masks=masks)
    _check(ref["train"]["losses"], torch.stack([x.detach() for x in out]), 2e-5)
    out[0].backward()
    for name, pr in fix["ref_fp64"]["train"]["grads"].items():
        g = p[name].grad.double()
        assert list(g.shape) == pr["shape"]
        if pr["l2"] > 0:
            assert float((g.reshape(-1)[pr["idx"]] - pr["vals"]).norm() / pr["vals"].norm()) < 2e-4, name
            assert abs(float(g.norm()) - pr["l2"]) < 2e-4 * pr["l2"], name


def test_bf16_rounded_linear_backward_rounds_its_operands():
    """The comparison arithmetic of the bf16 mode: forward AND backward contractions see bf16-rounded operands."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 16, generator=g, requires_grad=True)
    wt = torch.randn(5, 16, generator=g, requires_grad=True)
    dy = torch.randn(7, 5, generator=g)
    r = lambda t: t.to(torch.bfloat16).to(torch.float32)
    O.set_operand_rounding("bf16")
    try:
        y = O.linear(x, wt)
        y.backward(dy)
    finally:
        O.set_operand_rounding(None)
    assert torch.equal(y.detach(), r(x.detach()) @ r(wt.detach()).t())
    assert torch.allclose(x.grad, r(dy) @ r(wt.detach()), atol=1e-6)
    assert torch.allclose(wt.grad, r(dy).t() @ r(x.detach()), atol=1e-6)


def test_beam_known_answers(beam_kat):
    """SURVEY.md appendix A cases (generated by the reference's beamsearch with a table decoder)."""
    LP = torch.tensor(beam_kat["P"]).log()
    for case in beam_kat["cases"]:
        got, st = O.beamsearch(lambda tok, h, tile: (LP[tok].clone(), h), 8, 2, torch.zeros(2, 8), case["K"], case["L"],
                               return_state=True)
        assert got == case["expected"], case
    by_kl = {(c["K"], c["L"]): c["expected"][0] for c in beam_kat["cases"]}
    assert by_kl[(2, 6)] == [4, 5] and by_kl[(3, 6)] == [5, 6, 7] and by_kl[(3, 3)] == [4, 5]   # the survey's table
    _, st = O.beamsearch(lambda tok, h, tile: (LP[tok].clone(), h), 8, 2, torch.zeros(2, 8), 2, 6, return_state=True)
    assert st["steps_run"] == 3 and st["beam"][3:5].abs().sum() == 0 and (st["beam"][5] == 3).all()


def test_ranking_loss_edge_cases():
    im = O.l2norm(torch.randn(1, 7))
    assert float(O.pairwise_ranking_loss(im, im, 0.1)) == 0.0        # B = 1: only the (zeroed) diagonal
    im = O.l2norm(torch.randn(4, 7, generator=torch.Generator().manual_seed(0)))
    s = O.l2norm(torch.randn(4, 7, generator=torch.Generator().manual_seed(1)))
    both = float(O.pairwise_ranking_loss(im, s, 0.1))
    one = float(O.image_retrieval_ranking_loss(im, s, 0.1))
    other = float(O.image_retrieval_ranking_loss(s, im, 0.1))          # cost_im == cost_s of the transposed problem
    assert abs(both - (one + other)) < 1e-6


def test_clip_adam_matches_torch():
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    params = {n: p.detach().clone() for n, p in lin.named_parameters()}
    opt = torch.optim.Adam([{"params": [lin.weight], "weight_decay": 1e-5}, {"params": [lin.bias]}], lr=4e-4)
    state = {}
    for it in range(3):
        x = torch.randn(8, 5)
        opt.zero_grad()
        (lin(x).pow(2).sum() * 10).backward()
        grads = {n: p.grad.detach().clone() for n, p in lin.named_parameters()}
        torch.nn.utils.clip_grad_norm_(lin.parameters(), 1.0)
        opt.step()
        O.clip_adam_step(params, grads, state, lr=4e-4, clip=1.0, weight_decay=1e-5)
        for n, p in lin.named_parameters():
            assert float((p.detach() - params[n]).abs().max()) < 1e-6
