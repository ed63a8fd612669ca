"""GPU parity of every C-ABI operator against the CPU oracle (oracle/vag_oracle.py) on seeded inputs.

Floating-point tolerance (stated per test): the kernels compute in FP32 (SIMT FFMA, or three error-compensated tensor-core products of
FP16 hi/lo operand planes with FP32 accumulation) while the oracle is evaluated in FP64 here, so the bound is a few FP32 ulps of the
largest magnitude: 2e-5 relative to max|ref| for single contractions, 1e-4 for chained operators.
Integer outputs (tokens, parents, ranks) must be bit-exact.
"""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_GEMM = 2e-5
TOL_CHAIN = 1e-4


@pytest.fixture(scope="module")
def ops():
    from vag_nmt_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def O():
    from oracle import vag_oracle
    return vag_oracle


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("rows,K,N", [(1, 8, 8), (5, 24, 12), (7, 13, 50), (32, 256, 1536), (192, 512, 1024),
                                      (33, 1024, 512), (130, 256, 9391), (640, 1792, 256), (1, 2048, 512), (300, 100, 77)])
@pytest.mark.parametrize("flags", [0, 4])
def test_linear(ops, rows, K, N, flags):
    x = torch.randn(rows, K, generator=g(1))
    w = torch.randn(N, K, generator=g(2)) / math.sqrt(K)
    b = torch.randn(N, generator=g(3))
    ref = x.double() @ w.double().t() + b.double()
    y = ops.linear(x.cuda(), w.cuda(), b.cuda(), flags=flags)
    assert rel_err(y, ref) < TOL_GEMM
    # accumulate + tanh epilogue, no bias
    y0 = torch.randn(rows, N, generator=g(4))
    ref2 = torch.tanh(y0.double() + x.double() @ w.double().t())
    y2 = ops.linear(x.cuda(), w.cuda(), None, flags=flags | ops.LIN_TANH | ops.LIN_ACCUMULATE, out=y0.cuda())
    assert rel_err(y2, ref2) < TOL_GEMM


def test_linear_strided_views(ops):
    big = torch.randn(40, 300, generator=g(5)).cuda()
    x = big[:, 20:148]           # ld 300, 128 columns, base offset not 16B-multiple-safe
    w = torch.randn(64, 128, generator=g(6)).cuda()
    out_big = torch.zeros(40, 100).cuda()
    y = ops.linear(x, w, None, out=out_big[:, 10:74])
    ref = x.double().cpu() @ w.double().cpu().t()
    assert rel_err(y, ref) < TOL_GEMM
    assert float(out_big[:, :10].abs().max()) == 0 and float(out_big[:, 74:].abs().max()) == 0


def test_embed_rows(ops):
    table = torch.randn(50, 24, generator=g(1))
    ids = torch.randint(0, 50, (37,), generator=g(2))
    out = ops.embed_rows(table.cuda(), ids.cuda())
    assert torch.equal(out.cpu(), table[ids])


@pytest.mark.parametrize("rows,H", [(3, 16), (192, 512), (5, 6)])
def test_gru_gates(ops, O, rows, H):
    gi = torch.randn(rows, 3 * H, generator=g(1))
    gh = torch.randn(rows, 3 * H, generator=g(2))
    h = torch.randn(rows, H, generator=g(3))
    r = torch.sigmoid(gi[:, :H].double() + gh[:, :H].double())
    z = torch.sigmoid(gi[:, H:2 * H].double() + gh[:, H:2 * H].double())
    n = torch.tanh(gi[:, 2 * H:].double() + r * gh[:, 2 * H:].double())
    ref = (1 - z) * n + z * h.double()
    out = ops.gru_gates(gi.cuda(), gh.cuda(), h.cuda())
    assert rel_err(out, ref) < 1e-6


@pytest.mark.parametrize("B,R,T,C", [(4, 1, 7, 32), (3, 3, 9, 32), (16, 12, 20, 1024), (2, 5, 40, 1024), (2, 16, 5, 20),
                                     (1, 1, 1, 8)])
@pytest.mark.parametrize("mode", [0, 1])
def test_attention(ops, B, R, T, C, mode):
    q = torch.randn(B * R, C, generator=g(1))
    keys = torch.randn(B, T, C, generator=g(2))
    ctx = torch.randn(B, T, C, generator=g(3))
    v = torch.randn(C, generator=g(4)) / math.sqrt(C)
    lens = torch.randint(1, T + 1, (B,), generator=g(5))
    lens[0] = T
    mask = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)).float()
    qd, kd, cd, vd = q.double(), keys.double(), ctx.double(), v.double()
    kk = kd.repeat_interleave(R, 0)
    if mode == 0:
        s = torch.tanh(qd.unsqueeze(1) + kk).matmul(vd)
    else:
        s = (qd.unsqueeze(1) * kk).sum(-1)
    s = s.masked_fill(mask.repeat_interleave(R, 0) == 0, -float("inf"))
    a_ref = torch.softmax(s, 1)
    c_ref = a_ref.unsqueeze(1).bmm(cd.repeat_interleave(R, 0)).squeeze(1)
    c, a = ops.attention(q.cuda(), keys.cuda(), ctx.cuda(), v.cuda() if mode == 0 else None, mask.cuda(), R, mode)
    assert rel_err(a, a_ref) < 2e-5
    assert rel_err(c, c_ref) < 2e-5
    assert float(a.cpu()[mask.repeat_interleave(R, 0) == 0].abs().max() if (mask == 0).any() else 0.0) == 0.0


def test_l2norm_logsoftmax_nll(ops):
    x = torch.randn(9, 512, generator=g(1))
    y = ops.l2norm_rows_(x.clone().cuda())
    assert rel_err(y, x.double() / x.double().norm(2, 1, keepdim=True)) < 1e-6
    z = torch.zeros(2, 12).cuda()
    assert torch.equal(ops.l2norm_rows_(z.clone()).cpu(), torch.zeros(2, 12))  # eps clamp: 0/1e-12 = 0
    logits = torch.randn(33, 9391, generator=g(2)) * 3
    lp = ops.log_softmax(logits.cuda())
    ref = torch.log_softmax(logits.double(), -1)
    assert float((lp.cpu().double() - ref).abs().max()) < 5e-6
    tgt = torch.randint(0, 9391, (33,), generator=g(3))
    tgt[:3] = 0
    w = torch.ones(9391)
    w[0] = 0
    rows = torch.zeros(33).cuda()
    ops.nll_rows(logits.cuda(), tgt.cuda(), w.cuda(), rows)
    ops.nll_rows(logits.cuda(), tgt.cuda(), w.cuda(), rows)  # accumulates
    ref_rows = -2 * ref.gather(1, tgt.unsqueeze(1)).squeeze(1) * w.double()[tgt]
    assert float((rows.cpu().double() - ref_rows).abs().max()) < 1e-4
    assert float(rows[:3].abs().max()) == 0.0
    am = ops.row_argmax(logits.cuda())
    assert torch.equal(am.cpu(), logits.argmax(1))


@pytest.mark.parametrize("B,S,margin", [(1, 12, 0.1), (5, 12, 0.1), (32, 512, 0.1), (256, 512, 0.1), (7, 33, 1.0)])
@pytest.mark.parametrize("one_dir", [False, True])
def test_rank_loss_value_and_grad(ops, O, B, S, margin, one_dir):
    im = O.l2norm(torch.randn(B, S, generator=g(1)))
    s = O.l2norm(torch.randn(B, S, generator=g(2)) + 0.5 * im)
    imd = im.double().requires_grad_(True)
    sd = s.double().requires_grad_(True)
    fn = O.image_retrieval_ranking_loss if one_dir else O.pairwise_ranking_loss
    ref = fn(imd, sd, margin)
    ref.backward()
    loss, g_im, g_s = ops.rank_loss(im.cuda(), s.cuda(), margin, one_dir, want_grad=True)
    assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))   # fp32 sum of ≤ 2·B² hinge terms
    assert rel_err(g_im, imd.grad) < 1e-5 or float(imd.grad.abs().max()) == 0
    assert rel_err(g_s, sd.grad) < 1e-5 or float(sd.grad.abs().max()) == 0


def test_rank_loss_module_autograd(ops, O):
    import vag_nmt_b200 as vag
    im = O.l2norm(torch.randn(6, 12, generator=g(1))).cuda().requires_grad_(True)
    s = O.l2norm(torch.randn(6, 12, generator=g(2))).cuda().requires_grad_(True)
    loss = vag.PairwiseRankingLoss(margin=0.1)(im, s)
    (2.0 * loss).backward()
    imd = im.detach().cpu().double().requires_grad_(True)
    sd = s.detach().cpu().double().requires_grad_(True)
    (2.0 * O.pairwise_ranking_loss(imd, sd, 0.1)).backward()
    assert rel_err(im.grad, imd.grad) < 1e-5 and rel_err(s.grad, sd.grad) < 1e-5


@pytest.mark.parametrize("n", [1, 17, 1000])
def test_recall_ranks_exact(ops, O, n):
    im = O.l2norm(torch.randn(n, 512, generator=g(1)))
    cap = O.l2norm(im + 0.9 * torch.randn(n, 512, generator=g(2)))
    ranks = ops.recall_ranks(cap.cuda(), im.cuda()).cpu().numpy()
    ref = O.retrieval_ranks(cap, im)
    # scores are fp32 dot products on both sides; a rank can only differ where two scores tie to the last ulp
    assert (ranks == ref).mean() >= 0.995
    import vag_nmt_b200 as vag
    assert vag.t2i(im, cap)[:3] == O.t2i(im, cap)[:3]
    assert vag.i2t(im, cap)[:3] == O.i2t(im, cap)[:3]


# ---------------------------------------------------------------------------------------------- beam selection
def _select_oracle(O, logp, prev, nll, B, K, V, step, avoid_double=True):
    """one selection step exactly as oracle.beamsearch does it"""
    if step == 0:
        return O.topk_canonical(logp, K) + (torch.zeros(B, K, dtype=torch.long),)
    nk = torch.arange(B * K)
    lp = logp.clone()
    cur = prev.reshape(-1)
    if avoid_double:
        lp.view(-1).index_fill_(0, cur + nk * V, -1e5)
    fin = (cur == 3).nonzero()
    if fin.numel() > 0:
        lp.index_fill_(0, fin[:, 0], -1e5)
        lp.view(-1).index_fill_(0, fin[:, 0] * V + 3, 0)
    cand = (nll.unsqueeze(2) + lp.view(B, K, V)).view(B, -1)
    v, i = O.topk_canonical(cand, K)
    return v, i % V, i // V


@pytest.mark.parametrize("B,K,V", [(3, 3, 8), (4, 5, 50), (16, 12, 9391), (2, 16, 1000), (5, 2, 6)])
def test_beam_select_bit_exact(ops, O, B, K, V):
    gen = g(B * 100 + K)
    logp0 = torch.log_softmax(torch.randn(B, V, generator=gen) * 2, -1)
    nll = torch.zeros(B, K)
    v_ref, t_ref, _ = _select_oracle(O, logp0, None, None, B, K, V, 0)
    nll_d = nll.cuda()
    t, p = ops.beam_select(logp0.cuda(), None, nll_d, B, K, 0)
    assert torch.equal(t.cpu(), t_ref) and torch.equal(nll_d.cpu(), v_ref)
    prev, nll = t_ref, v_ref
    for step in range(1, 5):
        logp = torch.log_softmax(torch.randn(B * K, V, generator=gen) * 2, -1)
        if step == 2:   # force some hypotheses to have finished and others to want a repeat
            prev = prev.clone()
            prev[0, 0] = 3
            prev[-1, -1] = 3
        if step == 3:
            logp[torch.arange(B * K), prev.reshape(-1)] = 0.0  # the repeated token would win without suppression
        v_ref, t_ref, p_ref = _select_oracle(O, logp, prev, nll, B, K, V, step)
        nll_d = nll.cuda()
        t, p = ops.beam_select(logp.cuda(), prev.cuda(), nll_d, B, K, step)
        assert torch.equal(t.cpu(), t_ref), f"tokens differ at step {step}"
        assert torch.equal(p.cpu().long(), p_ref), f"parents differ at step {step}"
        assert torch.equal(nll_d.cpu(), v_ref), f"scores differ at step {step}"
        prev, nll = t_ref, v_ref


def test_beam_select_ties_and_all_finished(ops, O):
    B, K, V = 2, 3, 6
    logp = torch.full((B * K, V), -2.0)          # every candidate ties: canonical order = lowest flat index
    prev = torch.tensor([[4, 5, 4], [3, 3, 3]])   # second sentence: all hypotheses finished
    nll = torch.tensor([[-1.0, -1.0, -3.0], [-0.5, -0.7, -0.9]])
    v_ref, t_ref, p_ref = _select_oracle(O, logp, prev, nll, B, K, V, 1)
    nll_d = nll.cuda()
    t, p = ops.beam_select(logp.cuda(), prev.cuda(), nll_d, B, K, 1)
    assert torch.equal(t.cpu(), t_ref) and torch.equal(p.cpu().long(), p_ref) and torch.equal(nll_d.cpu(), v_ref)
    assert t_ref[1].tolist() == [3, 3, 3] and torch.equal(v_ref[1], nll[1])   # finished hyps keep their score (+0)
