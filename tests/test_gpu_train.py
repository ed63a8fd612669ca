"""GPU: the training step — loss values AND every parameter gradient against autograd through the CPU oracle,
then the fused clip + Adam update against torch's clip_grad_norm_ + optim.Adam.

Tolerances: gradients are compared per parameter by  max|g − g_ref| / max|g_ref|  < 2e-3 against an FP64 oracle
(FP32 kernels through ~10-40 chained steps of BPTT; the observed deviation is ~1e-5) and the losses to 1e-4.
"""
import pytest
import torch

from conftest import build_mm, build_tm, cpu_params

pytestmark = pytest.mark.gpu


def _oracle_grads(kind, model, batch, teacher, dtype=torch.float64, vse="pairwise"):
    from oracle import vag_oracle as O
    p = {k: v.clone().requires_grad_(True) for k, v in cpu_params(model, dtype).items()}
    if "decoder.out.weight" in p:
        p["decoder.out.weight"] = p["decoder.embedding.weight"]       # tied
    w = torch.ones(p["decoder.embedding.weight"].shape[0], dtype=dtype)
    w[0] = 0
    if kind == "mm":
        loss, lmt, lvse = O.multimodal_forward(p, batch.src, batch.src_lengths, batch.tgt, batch.im.to(dtype), teacher, w, vse, 0.1)
    else:
        loss = O.text_forward(p, batch.src, batch.src_lengths, batch.tgt, teacher, w)
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if v.grad is not None}
    return float(loss), grads


def _check_grads(model, ref, tol=2e-3):
    worst = 0.0
    for name, prm in model.named_parameters():
        assert name in ref, name
        g = prm.grad
        assert g is not None, f"no gradient for {name}"
        r = ref[name]
        scale = float(r.abs().max())
        err = float((g.detach().cpu().double() - r.double()).abs().max())
        if scale == 0:
            assert err < 1e-7, name
            continue
        worst = max(worst, err / scale)
        assert err / scale < tol, f"{name}: rel err {err / scale:.3e}"
    return worst


@pytest.mark.parametrize("teacher", [True, False])
@pytest.mark.parametrize("vse", ["pairwise", "imageretrieval"])
def test_tiny_multimodal_loss_and_gradients(teacher, vse):
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY)
    model = build_mm(cfg, 21).cuda().train()
    batch = synthetic.make_batch(6, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=5, max_len=9, min_len=2,
                                 mean=5.0, std=2.5, common_tgt_len=False)
    ref_loss, ref = _oracle_grads("mm", model, batch, teacher, vse=vse)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    cv = vag.PairwiseRankingLoss(margin=0.1) if vse == "pairwise" else vag.ImageRetrievalRankingLoss(margin=0.1)
    loss, lmt, lvse = model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0 if teacher else 0.0, criterion_mt=crit,
                            criterion_vse=cv)
    assert abs(float(loss) - ref_loss) < 1e-4 * abs(ref_loss)
    loss.backward()
    _check_grads(model, ref)


def test_tiny_mlp_attention_and_text_only_gradients():
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY, attn_model="mlp")
    batch = synthetic.make_batch(5, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=8, max_len=8, min_len=2,
                                 mean=5.0, std=2.0, common_tgt_len=False)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    tm = build_tm(cfg, 22).cuda().train()
    ref_loss, ref = _oracle_grads("tm", tm, batch, True)
    loss = tm(batch.src, batch.src_lengths, batch.tgt, 1.0, criterion=crit)
    assert abs(float(loss) - ref_loss) < 1e-4 * abs(ref_loss)
    loss.backward()
    _check_grads(tm, ref)


def test_full_shape_gradients_b32_reference_golden_and_oracle(full_de):
    """EN→DE shapes, B = 32 (BASELINE configs[1]): losses and EVERY parameter gradient against (a) the fixture written by the
    REAL reference's autograd in fp64 (oracle/make_golden.py:ref_train_gradients — norm and 64 probed entries per tensor) and
    (b) full tensors from autograd through the FP32 CPU oracle on the same batch."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    fix = full_de
    cfg = fix["cfg"]
    model = build_mm(cfg, fix["seed"]).cuda().train()
    batch = synthetic.make_batch(fix["batch_size"], cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=fix["data_seed"])
    ref_loss, ref = _oracle_grads("mm", model, batch, True, dtype=torch.float32)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    out = model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0, criterion_mt=crit,
                criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
    assert abs(float(out[0]) - ref_loss) < 1e-4 * abs(ref_loss)
    gold = fix["train_fp64"]
    got = torch.stack([x.reshape(()) for x in out]).detach().cpu().double()
    assert float(((got - gold["losses"]).abs() / gold["losses"].abs()).max()) < 1e-4
    out[0].backward()
    worst = _check_grads(model, ref, tol=2e-3)
    assert worst < 2e-3
    for name, prm in model.named_parameters():
        pr = gold["grads"][name]
        g = prm.grad.detach().cpu().double()
        assert list(g.shape) == pr["shape"], name
        if pr["l2"] == 0.0:
            continue
        assert float((g.reshape(-1)[pr["idx"]] - pr["vals"]).norm() / pr["vals"].norm()) < 2e-3, name
        assert abs(float(g.norm()) - pr["l2"]) < 2e-3 * pr["l2"], name


def test_training_mode_dropout_with_injected_masks():
    """Embedding / context / output dropout (DE defaults 0.3 / 0.5 / 0.5) with the SAME masks on both sides."""
    import vag_nmt_b200 as vag
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY)
    torch.manual_seed(31)
    model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(
        cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
        cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, dropout_ctx=0.5, dropout_emb=0.3, dropout_out=0.5, tied_emb=True).cuda().train()
    batch = synthetic.make_batch(6, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=5, max_len=9, min_len=2,
                                 mean=5.0, std=2.5, common_tgt_len=False)
    B, Ts = batch.src.shape
    Tt = batch.tgt.shape[1]
    gen = torch.Generator().manual_seed(3)
    mk = lambda shape, p: torch.empty(shape).bernoulli_(1 - p, generator=gen) / (1 - p)
    masks = {"emb": mk((Ts * B, cfg["src_embedding_size"]), 0.3), "ctx": mk((B, Ts, 2 * cfg["hidden_size"]), 0.5),
             "out": mk((Tt * B, cfg["tgt_embedding_size"]), 0.5)}
    model._dropout_masks = masks
    p = {k: v.clone().double().requires_grad_(True) for k, v in cpu_params(model).items()}
    p["decoder.out.weight"] = p["decoder.embedding.weight"]
    w = torch.ones(cfg["tgt_size"], dtype=torch.float64)
    w[0] = 0
    ref, _, _ = O.multimodal_forward(p, batch.src, batch.src_lengths, batch.tgt, batch.im.double(), True, w, "pairwise", 0.1,
                                     dropout_masks={k: v.double() for k, v in masks.items()})
    ref.backward()
    crit = torch.nn.NLLLoss(weight=w.float().cuda(), reduce=False)
    loss, _, _ = model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0, criterion_mt=crit, criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    loss.backward()
    _check_grads(model, {k: v.grad for k, v in p.items() if v.grad is not None})
    # without injected masks the draw is random: two passes differ, eval mode is deterministic
    model._dropout_masks = None
    l1 = float(model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0, criterion_mt=crit)[0])
    l2 = float(model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0, criterion_mt=crit)[0])
    assert l1 != l2


def test_clip_adam_matches_torch():
    from vag_nmt_b200.optim import ClipAdam, named_param_groups
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(40, 30), torch.nn.Tanh(), torch.nn.Linear(30, 7)).cuda()
    ref = torch.nn.Sequential(torch.nn.Linear(40, 30), torch.nn.Tanh(), torch.nn.Linear(30, 7)).cuda()
    ref.load_state_dict(net.state_dict())
    opt = ClipAdam(named_param_groups(net, 1e-5), lr=4e-4)
    named = list(ref.named_parameters())
    topt = torch.optim.Adam([{"params": [p for n, p in named if "bias" not in n], "weight_decay": 1e-5},
                             {"params": [p for n, p in named if "bias" in n]}], lr=4e-4)
    for it in range(4):
        x = torch.randn(16, 40, device="cuda")
        for m, o in ((net, opt), (ref, topt)):
            o.zero_grad()
            (m(x).pow(2).sum() * (50.0 if it % 2 == 0 else 1e-3)).backward()     # one clipped, one unclipped step
        sumsq = opt.step(clip=1.0)
        norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        topt.step()
        assert abs(float(sumsq.sqrt()) - float(norm)) < 1e-4 * float(norm)
        for a, b in zip(net.parameters(), ref.parameters()):
            assert float((a - b).abs().max()) < 2e-6


def test_training_steps_reduce_the_loss():
    """Ten optimiser steps through the reference-signature driver on a fixed batch must drive the loss down."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import train_imagine_beam
    cfg = dict(synthetic.TINY)
    model = build_mm(cfg, 3).cuda()
    batch = synthetic.make_batch(8, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=2, max_len=8, min_len=2, mean=5, std=2)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    opt = ClipAdam(model, lr=1e-2)
    losses = [train_imagine_beam(batch.src, batch.tgt, batch.im, batch.src_lengths, model, opt, crit,
                                 vag.PairwiseRankingLoss(margin=0.1), 0.99, 1.0)[0] for _ in range(10)]
    assert losses[-1] < 0.7 * losses[0], losses


def _same_shape_batches(cfg, n, B, seed0):
    """n batches with the SAME padded shapes but different sentence lengths / tokens (what a CUDA graph must survive)."""
    from vag_nmt_b200 import synthetic
    out, seed = [], seed0
    while len(out) < n:
        bt = synthetic.make_batch(B, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=seed, max_len=9, min_len=2, mean=6.0, std=2.5)
        seed += 1
        if not out or (bt.src.shape == out[0].src.shape and bt.tgt.shape == out[0].tgt.shape and bt.src_lengths != out[0].src_lengths):
            out.append(bt)
    return out


@pytest.mark.parametrize("kind", ["mm", "tm"])
def test_graphed_step_matches_eager_step(kind):
    """GraphedTrainStep (zero_grad → forward → backward replayed from one CUDA graph per batch shape, lengths read on the
    device) must walk the parameters exactly where the eager step driver walks them, over batches whose lengths differ."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import GraphedTrainStep, train_imagine_beam, train_nmt
    cfg = dict(synthetic.TINY)
    build = build_mm if kind == "mm" else build_tm
    batches = _same_shape_batches(cfg, 3, 8, 40)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    results = []
    for graphed in (False, True):
        model = build(cfg, 11).cuda()
        crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
        cv = vag.PairwiseRankingLoss(margin=0.1) if kind == "mm" else None
        opt = ClipAdam(model, lr=1e-2)
        stepper = GraphedTrainStep(model, opt, crit, cv, clip=1.0, enabled=True)
        losses = []
        for it in range(6):
            bt = batches[it % len(batches)]
            if graphed:
                out = stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im if kind == "mm" else None, 1.0)
                losses.append(float(out[0]))
            elif kind == "mm":
                losses.append(train_imagine_beam(bt.src, bt.tgt, bt.im, bt.src_lengths, model, opt, crit, cv, 0.99, 1.0)[0])
            else:
                losses.append(train_nmt(bt.src, bt.tgt, bt.src_lengths, model, crit, opt, 1.0))
        if graphed:
            assert len(stepper._graphs) == 1          # one shape → one capture, replayed for every batch
        results.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    (l0, p0), (l1, p1) = results
    for a, b in zip(l0, l1):
        assert abs(a - b) < 1e-5 * max(1.0, abs(a)), (l0, l1)
    for k in p0:
        assert float((p0[k] - p1[k]).abs().max()) < 1e-5, k


def test_global_norm_is_deterministic_and_matches_torch():
    """vag_sumsq_multi_det_f32: bit-identical from call to call (no float atomics) and equal to torch's norm."""
    from vag_nmt_b200 import train_ops as T
    g = torch.Generator().manual_seed(3)
    shapes = [(9391, 256), (1536, 512), (1536,), (7,), (1024, 1024)]
    ps = [torch.randn(*s, generator=g).cuda() for s in shapes]
    gs = [torch.randn(*s, generator=g).cuda() for s in shapes]
    entries = [(p, gr, torch.zeros_like(p), torch.zeros_like(p), 0.0, 1e-3) for p, gr in zip(ps, gs)]
    table = T.optim_table(entries).cuda()
    n_part = T._cabi.lib().vag_sumsq_multi_partials(len(entries), max(p.numel() for p in ps))
    partials = torch.empty(n_part, device="cuda")
    outs = []
    for _ in range(5):
        out = torch.full((1,), -1.0, device="cuda")
        T.sumsq_multi_det_(out, table, len(entries), max(p.numel() for p in ps), partials)
        outs.append(float(out))
    assert len(set(outs)) == 1
    ref = float(sum((gr.double() ** 2).sum() for gr in gs))
    assert abs(outs[0] - ref) < 1e-5 * ref


def test_split_graph_step_with_eager_global_ranking_loss_matches_eager_step():
    """The data-parallel flavour of GraphedTrainStep (forward graph → eager all-gather + ranking-loss kernel → backward graph)
    on a one-rank NCCL group: same parameter walk as the eager driver with DistributedPairwiseRankingLoss."""
    import socket
    import torch.distributed as dist
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import DistributedPairwiseRankingLoss, GraphedTrainStep, train_imagine_beam
    own_group = not dist.is_initialized()
    if own_group:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        cfg = dict(synthetic.TINY)
        batches = _same_shape_batches(cfg, 3, 8, 40)
        w = torch.ones(cfg["tgt_size"])
        w[0] = 0
        results = []
        for graphed in (False, True):
            model = build_mm(cfg, 11).cuda()
            crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
            cv = DistributedPairwiseRankingLoss(margin=0.1)
            opt = ClipAdam(model, lr=1e-2)
            stepper = GraphedTrainStep(model, opt, crit, cv, clip=1.0, enabled=True)
            stepper._split = True           # force the two-graph path although the world has one rank
            losses = []
            for it in range(6):
                bt = batches[it % len(batches)]
                if graphed:
                    out = stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im, 1.0)
                    losses.append([float(v) for v in out])
                else:
                    losses.append(list(train_imagine_beam(bt.src, bt.tgt, bt.im, bt.src_lengths, model, opt, crit, cv, 0.99, 1.0)))
            if graphed:
                assert len(stepper._graphs) == 1 and next(iter(stepper._graphs.values())).get("split")
            results.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
        (l0, p0), (l1, p1) = results
        for a, b in zip(l0, l1):
            for x, y in zip(a, b):
                assert abs(x - y) < 1e-5 * max(1.0, abs(x)), (l0, l1)
        for k in p0:
            assert float((p0[k] - p1[k]).abs().max()) < 1e-5, k
    finally:
        if own_group:
            dist.destroy_process_group()


def test_graphed_step_draws_fresh_dropout_masks_on_every_replay():
    """Dropout masks come from torch's CUDA generator INSIDE the captured step: every replay must draw new ones (Philox offsets
    advance through the graph), the loss must stay finite and one graph must serve the shape."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import GraphedTrainStep
    cfg = dict(synthetic.TINY)
    torch.manual_seed(5)
    model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(
        cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
        cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, dropout_emb=0.2, dropout_ctx=0.4, dropout_out=0.4, tied_emb=True).cuda()
    bt = synthetic.make_batch(8, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=2, max_len=8, min_len=2, mean=5, std=2)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    opt = ClipAdam(model, lr=0.0)          # frozen weights: only the masks change from step to step
    stepper = GraphedTrainStep(model, opt, crit, vag.PairwiseRankingLoss(margin=0.1), enabled=True)
    losses = [float(stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im, 1.0)[1]) for _ in range(5)]
    assert len(stepper._graphs) == 1
    assert all(torch.isfinite(torch.tensor(losses)))
    assert len({round(v, 6) for v in losses}) >= 4, losses


def test_flat_gradient_buffer_shared_by_all_graphs_and_lru_eviction():
    """ClipAdam keeps ONE persistent flat gradient buffer: the backward kernels write into it (no autograd copy), every captured
    graph of GraphedTrainStep uses the same addresses (nothing per graph but activations), the graph cache evicts least-recently
    used shapes, and an evicted shape is captured again with identical results."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import GraphedTrainStep, train_imagine_beam
    cfg = dict(synthetic.TINY)
    w = torch.ones(cfg["tgt_size"], device="cuda")
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w, reduction="none")
    cv = vag.PairwiseRankingLoss(margin=0.1)
    mk = lambda seed, n: synthetic.make_batch(n, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=seed, max_len=9, min_len=2,
                                              mean=5.0, std=2.5)
    batches = [mk(41, 6), mk(42, 5), mk(43, 4), mk(41, 6)]       # three shapes, the first one again at the end
    results = {}
    for mode in ("eager", "graph"):
        model = build_mm(cfg, 77).cuda()
        opt = ClipAdam(model, lr=1e-2)
        stepper = GraphedTrainStep(model, opt, crit, cv, clip=1.0, enabled=(mode == "graph"), max_graphs=2)
        losses = []
        for bt in batches:
            if mode == "eager":
                losses.append(train_imagine_beam(bt.src, bt.tgt, bt.im, bt.src_lengths, model, opt, crit, cv, 0.99, 1.0)[0])
            else:
                losses.append(float(stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im, 1.0)[0]))
            assert opt.grads_in_place, "a gradient was produced outside the flat buffer and had to be copied"
        base, size = opt._flat.data_ptr(), opt._flat.numel() * 4
        for p in model.parameters():
            assert base <= p.grad.data_ptr() < base + size
        if mode == "graph":
            assert len(stepper._graphs) == 2                      # LRU: three shapes seen, two kept
            ptrs = [[g.data_ptr() for _, g in st["grads"]] for st in stepper._graphs.values()]
            assert ptrs[0] == ptrs[1]                             # both graphs write the same gradient addresses
        results[mode] = (losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu())
    for a, b in zip(results["eager"][0], results["graph"][0]):
        assert abs(a - b) < 1e-5 * abs(a)
    assert float((results["eager"][1] - results["graph"][1]).abs().max()) < 1e-5


def test_loss_mix_backward_and_mask_lengths_kernels():
    """vag_translation_loss_bwd_f32 against autograd of the same formula, vag_src_mask_lengths against (src != 0) — exact."""
    from vag_nmt_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, Tt, Ts = 37, 19, 23
    tgt = torch.randint(1, 50, (B, Tt), generator=g)
    src = torch.randint(1, 50, (B, Ts), generator=g)
    for b in range(B):
        tgt[b, 1 + b % (Tt - 1):] = 0
        src[b, 1 + (b * 7) % Ts:] = 0
    rows = torch.rand(B, generator=g).double().requires_grad_(True)
    vse = torch.rand(1, generator=g).double().requires_grad_(True)
    gvec = torch.tensor([0.7, -0.3, 0.2], dtype=torch.float64)
    for has_vse in (True, False):
        w = 0.99
        cnt = (tgt != 0).sum(1).double()
        mt = (rows / cnt).mean()
        out = torch.stack([w * mt + (1 - w) * vse[0], mt, vse[0]]) if has_vse else torch.stack([mt, mt, 0 * mt])
        rows.grad = vse.grad = None
        out.backward(gvec)
        g_rows, g_vse = ops.translation_loss_bwd(gvec.float().cuda(), tgt.cuda(), w, has_vse)
        assert torch.allclose(g_rows.cpu().double(), rows.grad, rtol=1e-6, atol=1e-9)
        if has_vse:
            assert torch.allclose(g_vse.cpu().double(), vse.grad, rtol=1e-6)
        else:
            assert g_vse is None
    mask, lens = ops.src_mask_lengths(src.cuda())
    assert torch.equal(mask.cpu(), (src != 0).float())
    assert torch.equal(lens.cpu(), (src != 0).sum(1, dtype=torch.int32))
    mask2, none = ops.src_mask_lengths(src.cuda(), want_lengths=False)
    assert none is None and torch.equal(mask2, mask)


@pytest.mark.parametrize("hidden,batch_size", [(24, 3), (40, 7), (72, 32)])
def test_persistent_encoder_loops_odd_widths_and_ragged_batches(hidden, batch_size):
    """csrc/enc_seq.cu (one launch for the whole bidirectional GRU time loop, forward and BPTT) at widths where the warps' contraction
    ranges do not tile H evenly (72: the last working warp holds one k-step, 24 / 40: most warps idle), with fewer than 32 rows and
    ragged lengths — loss and every gradient against the fp64 oracle."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY, hidden_size=hidden, shared_embedding_size=16)
    model = build_mm(cfg, 33).cuda().train()
    batch = synthetic.make_batch(batch_size, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=9, max_len=11, min_len=1,
                                 mean=5.0, std=3.0, common_tgt_len=False)
    ref_loss, ref = _oracle_grads("mm", model, batch, True)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    loss, _, _ = model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0, criterion_mt=crit,
                       criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
    assert abs(float(loss) - ref_loss) < 1e-4 * abs(ref_loss)
    loss.backward()
    _check_grads(model, ref)


def test_persistent_decoder_forward_loop_opt_in(monkeypatch):
    """csrc/dec_seq.cu (VAG_DEC_SEQ=1: one launch for the teacher-forced decoder time loop — five exchange phases per step around
    the attention) writes what the per-step kernels write: loss and every gradient against the fp64 oracle, FP32 and bf16 modes
    against each other's tolerance."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    cfg = dict(synthetic.TINY, hidden_size=128, shared_embedding_size=32, src_embedding_size=16, tgt_embedding_size=16)
    model = build_mm(cfg, 41).cuda().train()
    batch = synthetic.make_batch(9, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=11, max_len=12, min_len=2,
                                 mean=6.0, std=3.0, common_tgt_len=False)
    ref_loss, ref = _oracle_grads("mm", model, batch, True)
    w = torch.ones(cfg["tgt_size"])
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w.cuda(), reduce=False)
    losses = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VAG_DEC_SEQ", flag)
        for p in model.parameters():
            p.grad = None
        loss, _, _ = model(batch.src, batch.src_lengths, batch.tgt, batch.im, 1.0, criterion_mt=crit,
                           criterion_vse=vag.PairwiseRankingLoss(margin=0.1))
        assert abs(float(loss) - ref_loss) < 1e-4 * abs(ref_loss)
        loss.backward()
        _check_grads(model, ref)
        losses[flag] = float(loss)
    assert abs(losses["0"] - losses["1"]) < 2e-6 * abs(ref_loss)
