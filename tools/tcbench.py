"""Micro-benchmark of the tcgen05 split-precision contraction vs the FP32 FFMA kernel on the decode-step shapes."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vag_nmt_b200 import ops


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


shapes = [(12000, 256, 9391), (12000, 512, 1536), (12000, 1024, 512), (12000, 512, 1024), (12000, 1792, 256),
          (12000, 256, 1536), (20000, 1024, 1024), (1500, 256, 9391), (1500, 512, 1536)]
for rows, K, N in shapes:
    x = torch.randn(rows, K, device='cuda')
    w = torch.randn(N, K, device='cuda') / math.sqrt(K)
    b = torch.randn(N, device='cuda')
    ldy = (N + 3) // 4 * 4
    ybuf = torch.empty(rows, ldy, device='cuda')
    y = ybuf[:, :N]
    ref = x.double() @ w.double().t() + b.double()
    ops.linear_tc(x, w, b, out=y)
    err = rel(y, ref)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        ops.linear_tc(x, w, b, out=y)
    e[0].record()
    for _ in range(10):
        ops.linear_tc(x, w, b, out=y)
    e[1].record()
    torch.cuda.synchronize()
    t = e[0].elapsed_time(e[1]) / 10
    fl = 2.0 * rows * K * N
    print(f"{rows}x{K}x{N}: tc+splits {t * 1e3:8.1f} us {fl / t / 1e9:7.1f} TF/s err {err:.2e}  cfg={os.environ.get('VAG_TC_CFG', 'auto')}")
