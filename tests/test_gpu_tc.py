"""GPU: the tcgen05 3xTF32 contraction against an FP64 reference.

Bar: max |y - ref| / max|ref| < 3e-6 — two orders of magnitude tighter than a plain TF32 product (~1e-3) could
meet, i.e. the error-compensated split is doing its job — and within 4x of what the FP32 FFMA kernel achieves."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("rows,K,N", [(128, 32, 64), (64, 64, 256), (300, 512, 1536), (1000, 1024, 512), (2000, 1792, 256),
                                      (200, 104, 128), (12000, 256, 9391), (129, 2048, 512), (20000, 1024, 1024), (777, 40, 65)])
def test_linear_tc_matches_fp64(rows, K, N):
    from vag_nmt_b200 import ops
    x = torch.randn(rows, K, generator=g(1))
    w = torch.randn(N, K, generator=g(2)) / math.sqrt(K)
    b = torch.randn(N, generator=g(3))
    xd, wd = x.cuda().double(), w.cuda().double()
    ref = xd @ wd.t() + b.cuda().double()
    y = ops.linear_tc(x.cuda(), w.cuda(), b.cuda())
    e_tc = rel_err(y, ref)
    e_simt = rel_err(ops.linear(x.cuda(), w.cuda(), b.cuda()), ref)
    assert e_tc < 3e-6, (e_tc, e_simt)
    assert e_tc < 3 * e_simt + 2e-7, (e_tc, e_simt)
    y0 = torch.randn(rows, N, generator=g(4)).cuda()
    ref2 = torch.tanh(y0.double() + xd @ wd.t())
    y2 = ops.linear_tc(x.cuda(), w.cuda(), None, flags=ops.LIN_TANH | ops.LIN_ACCUMULATE, out=y0.clone())
    assert rel_err(y2, ref2) < 2e-5       # tanh output in (-1, 1), pre-activations up to ~6


def test_linear_tc_strided_and_rejects():
    from vag_nmt_b200 import _cabi, ops
    big = torch.randn(500, 1792, generator=g(5)).cuda()
    x = big[:, 768:1792]                      # ld 1792, K = 1024 (the context slice of a concatenated buffer)
    w = torch.randn(512, 1024, generator=g(6)).cuda() / 32
    out_big = torch.zeros(500, 600).cuda()
    y = ops.linear_tc(x, w, None, out=out_big[:, 40:552])
    assert rel_err(y, x.double() @ w.double().t()) < 3e-6
    assert float(out_big[:, :40].abs().max()) == 0 and float(out_big[:, 552:].abs().max()) == 0
    with pytest.raises(_cabi.VagError):
        ops.linear_tc(torch.randn(8, 64).cuda(), torch.randn(64, 64).cuda())       # too few rows
    with pytest.raises(_cabi.VagError):
        ops.linear_tc(torch.randn(128, 36).cuda(), torch.randn(64, 36).cuda())     # K not a multiple of 8
