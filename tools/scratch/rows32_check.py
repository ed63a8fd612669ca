"""Correctness + timing of the <=32-row contraction kernel (linear_rows.cu) through vag_linear_f32 / vag_gemm_f32."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vag_nmt_b200 import ops, train_ops as T

torch.manual_seed(0)
def timeit(f, n=40):
    """n back-to-back calls captured in one CUDA graph: GPU time per call without host launch overhead"""
    for _ in range(3): f()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        f()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for rows in (32, 17):
    for K, N in ((512, 1536), (256, 1536), (1024, 512), (512, 1024), (1536, 512), (1024, 1024), (512, 260), (68, 36)):
        x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
        ref = (x.double() @ w.double().t() + b.double())
        y = ops.linear(x, w, b)
        yo = torch.empty_like(y); dxo = torch.empty(rows, K, device='cuda')
        e1 = float((y.double() - ref).abs().max())
        y2 = ops.linear(x, w, b, ops.LIN_TANH)
        e2 = float((y2.double() - ref.tanh()).abs().max())
        # backward orientation: dx = dy · W
        dy = torch.randn(rows, N, device="cuda")
        refd = dy.double() @ w.double()
        dx = T.gemm(dy, w)
        e3 = float((dx.double() - refd).abs().max())
        base = torch.randn(rows, K, device="cuda")
        dx2 = base.clone(); T.gemm(dy, w, out=dx2, beta=1.0)
        e4 = float((dx2.double() - (refd + base.double())).abs().max())
        t1 = timeit(lambda: ops.linear(x, w, b, out=yo)); t2 = timeit(lambda: T.gemm(dy, w, out=dxo))
        t3 = timeit(lambda: torch.nn.functional.linear(x, w, b))
        print(f"rows {rows:2d} K {K:4d} N {N:4d}: err fwd {e1:.2e} tanh {e2:.2e} bwd {e3:.2e} acc {e4:.2e} | fwd {t1:5.1f} us bwd {t2:5.1f} us torch {t3:5.1f} us")
        assert max(e1, e2) < 2e-5 and max(e3, e4) < 5e-5 * max(1, N / 512) ** 0.5 * 4
print("ok")
