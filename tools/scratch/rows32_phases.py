import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vag_nmt_b200 import ops, train_ops as T
def timeit(f, n=40):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): f()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
out = []
for K, N in ((512, 1536), (512, 3072), (512, 1024), (1024, 512), (1536, 512), (68, 36)):
    x = torch.randn(32, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    yo = torch.empty(32, N, device="cuda"); dy = torch.randn(32, N, device="cuda"); dxo = torch.empty(32, K, device="cuda")
    out.append(f"K{K} N{N}: fwd {timeit(lambda: ops.linear(x, w, b, out=yo)):.1f} bwd {timeit(lambda: T.gemm(dy, w, out=dxo)):.1f}")
print(os.environ.get("TAG", ""), " | ".join(out))
