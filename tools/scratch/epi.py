import math, os, sys, torch
sys.path.insert(0, '/root/repo')
from vag_nmt_b200 import _cabi, ops
lib = _cabi.lib()
for mode in (-1, 2):
    lib.vag_set_gemm_mode(mode)
    for rows, K, N in [(12000, 32, 1536), (12000, 64, 1536), (12000, 128, 1536), (12000, 256, 1536), (12000, 512, 1536), (12000, 32, 9391), (12000, 64, 9391), (12000,128,9391),(12000, 32, 256), (12000, 32, 512)]:
        x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
        ldy = (N + 3) // 4 * 4
        y = torch.empty(rows, ldy, device="cuda")[:, :N]
        xs, ws = ops.tc_split(x), ops.tc_split(w)
        for _ in range(3): ops.tc_gemm(xs, ws, rows, K, N, b, out=y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): ops.tc_gemm(xs, ws, rows, K, N, b, out=y)
        e1.record(); torch.cuda.synchronize()
        print(mode, rows, K, N, f"{e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
