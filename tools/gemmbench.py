"""GEMM-only micro-benchmark of the tensor-core contraction on pre-split operands (the launch the decode loop makes).

    python tools/gemmbench.py [--shapes vocab|all] [--reps 20] [--mode 1|2]

Prints per shape: microseconds per launch, effective TFLOP/s (algorithmic 2*M*K*N), max error vs float64.
`--shapes vocab` runs the dominant kernel alone (vocabulary projection 12000 x 256 x 9391): the command ncu wraps.
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vag_nmt_b200 import _cabi, ops

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="all")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--mode", type=int, default=-1)
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
lib = _cabi.lib()
# the arithmetic mode travels with each call now (ops wrappers read _cabi.precision_scope)
STEP = [(12000, 256, 1536), (12000, 512, 1536), (12000, 512, 1024), (12000, 1024, 512), (12000, 1792, 256), (12000, 256, 9391)]
shapes = [(12000, 256, 9391)] if a.shapes == "vocab" else STEP + [(1000, 256, 9391), (384, 512, 1536), (12000, 1024, 1024)]
tot = 0.0
for rows, K, N in shapes:
    g = torch.Generator(device="cuda").manual_seed(rows + K + N)
    x = torch.randn(rows, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    ldy = (N + 3) // 4 * 4
    y = torch.empty(rows, ldy, device="cuda")[:, :N]
    xs, ws = ops.tc_split(x), ops.tc_split(w)
    for _ in range(3):
        ops.tc_gemm(xs, ws, rows, K, N, b, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        ops.tc_gemm(xs, ws, rows, K, N, b, out=y)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / a.reps
    err = float("nan")
    if a.check:
        if a.mode == 2:
            ref = x.bfloat16().double() @ w.bfloat16().double().t() + b.double()
        else:
            ref = x.double() @ w.double().t() + b.double()
        err = float((y.double() - ref).abs().max() / ref.abs().max())
    if (rows, K, N) in STEP:
        tot += t
    print(f"{rows:6d} x {K:5d} x {N:5d}: {t * 1e3:8.1f} us  {2.0 * rows * K * N / t / 1e9:7.1f} TF/s  err {err:.2e}", flush=True)
print(f"sum over the six decode-step shapes: {tot * 1e3:.1f} us")
