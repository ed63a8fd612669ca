"""GPU: EN→FR shapes (BASELINE configs[3]: V = 8748, dropout emb 0.2 / ctx 0.4 / out 0.4, nmt_multimodal_beam_FR.py:55-67)
against the fixture the REAL reference produced (oracle/make_golden.py:fr_fixture → tests/golden/full_fr_b32.pt):
beam-12 / beam-5 tokens exact, eval-mode losses, training-mode losses and every parameter gradient under the same injected
dropout masks (losses ≤ 1e-4 FP32 / 1e-3 bf16; gradients norm-wise and on the probed entries against the reference's fp64 run:
≤ 2e-4 in FP32 mode, ≤ 2e-2 in bf16 mode — there the bound is the mode's own rounding against an UNROUNDED reference).
V = 8748 = 68·128 + 44 exercises a different last vocabulary tile from EN→DE's 9391 = 73·128 + 47.
"""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fr():
    return load_golden("full_fr_b32.pt")


def _build(fix, train=False):
    import vag_nmt_b200 as vag
    cfg = fix["cfg"]
    torch.manual_seed(fix["seed"])
    m = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(
        cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
        cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, tied_emb=True, init_split=0.5, **fix["dropout"])
    for k, v in m.state_dict().items():
        s, a = fix["param_checksums"][k]
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
    return (m.train() if train else m.eval()).cuda()


def _batch(fix):
    from vag_nmt_b200 import synthetic
    cfg = fix["cfg"]
    b = synthetic.make_batch(fix["batch_size"], cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=fix["data_seed"])
    B, Ts = b.src.shape
    d = fix["dropout"]
    masks = synthetic.dropout_masks(fix["mask_seed"], B, Ts, b.tgt.shape[1], cfg["src_embedding_size"], cfg["hidden_size"],
                                    cfg["tgt_embedding_size"], d["dropout_emb"], d["dropout_ctx"], d["dropout_out"])
    return b, masks


def _crit(V):
    w = torch.ones(V)
    w[0] = 0
    return torch.nn.NLLLoss(weight=w.cuda(), reduction="none")


def test_fr_decode_tokens_and_eval_forward(fr):
    import vag_nmt_b200 as vag
    assert fr["tokens_stable"]          # the reference's fp32 and fp64 runs agree on these tokens (no near-ties)
    model = _build(fr)
    b, _ = _batch(fr)
    ref, ref64 = fr["ref_fp32"], fr["ref_fp64"]
    assert model.beamsearch_decode(b.src, b.src_lengths, b.im, beam_size=12, max_length=fr["max_length"]) == ref["decode_k12"]
    assert model.beamsearch_decode(b.src, b.src_lengths, b.im, beam_size=5, max_length=fr["max_length"]) == ref["decode_k5"]
    with torch.no_grad():
        out = model(b.src, b.src_lengths, b.tgt, b.im, 1.0, criterion_mt=_crit(fr["cfg"]["tgt_size"]),
                    criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
    got = torch.stack([x.reshape(()) for x in out]).cpu().double()
    assert float(((got - ref64["fwd_eval"]).abs() / ref64["fwd_eval"].abs()).max()) < 1e-4
    e_im, e_txt = model.embed_sent_im_test(b.src, b.src_lengths, b.im)
    assert list(vag.t2i(e_im, e_txt)) == ref["t2i"]
    # a larger batch of the same shapes runs step 0 fused as well (B > 128) and crosses the last vocabulary tile at full width
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    cfg = fr["cfg"]
    big = synthetic.make_batch(160, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=23)
    got_big = model.beamsearch_decode(big.src, big.src_lengths, big.im, beam_size=12, max_length=10)
    with torch.no_grad():
        want = O.multimodal_beamsearch_decode({k: v.detach().cpu() for k, v in model.state_dict().items()}, big.src, big.src_lengths,
                                              big.im, 12, 10)
    assert got_big == want


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fr_training_losses_and_gradients_with_injected_dropout(fr, precision):
    import vag_nmt_b200 as vag
    model = _build(fr, train=True)
    model.precision = precision
    b, masks = _batch(fr)
    model._dropout_masks = masks
    ref = fr["ref_fp64"]["train"]
    out = model(b.src, b.src_lengths, b.tgt, b.im, 1.0, criterion_mt=_crit(fr["cfg"]["tgt_size"]),
                criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
    got = torch.stack([x.reshape(()) for x in out]).detach().cpu().double()
    loss_tol = 1e-4 if precision == "fp32" else 1e-3
    assert float(((got - ref["losses"]).abs() / ref["losses"].abs()).max()) < loss_tol, (got, ref["losses"])
    with model.precision_scope():
        out[0].backward()
    tol = 2e-4 if precision == "fp32" else 2e-2       # bf16 here is against the UNROUNDED fp64 reference: the mode's own error
    names = [n for n, _ in model.named_parameters()]
    assert sorted(names) == sorted(ref["grads"].keys())
    worst = 0.0
    for name, prm in model.named_parameters():
        pr = ref["grads"][name]
        g = prm.grad.detach().cpu().double()
        assert list(g.shape) == pr["shape"], name
        if pr["l2"] == 0.0:
            assert float(g.abs().max()) < 1e-9, name
            continue
        probe = float((g.reshape(-1)[pr["idx"]] - pr["vals"]).norm() / pr["vals"].norm())
        norm = abs(float(g.norm()) - pr["l2"]) / pr["l2"]
        worst = max(worst, probe, norm)
        assert probe < tol and norm < tol, f"{name}: probed entries {probe:.3e}, norm {norm:.3e}"
    print(f"EN→FR {precision} gradients vs the reference's fp64 autograd: worst {worst:.3e}")
