import math, os, sys, torch
sys.path.insert(0, '/root/repo')
from vag_nmt_b200 import _cabi, ops
lib = _cabi.lib()
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.vag_tc_set_debug(dbg.data_ptr())
mode = int(os.environ.get("MODE", "-1"))
lib.vag_set_gemm_mode(mode)
for rows, K, N in [(12000, 512, 1536), (12000, 1792, 256)]:
    x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    ldy = (N + 3) // 4 * 4
    y = torch.empty(rows, ldy, device="cuda")[:, :N]
    xs, ws = ops.tc_split(x), ops.tc_split(w)
    for _ in range(3): ops.tc_gemm(xs, ws, rows, K, N, b, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): ops.tc_gemm(xs, ws, rows, K, N, b, out=y)
    e1.record(); torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    print(f"{rows}x{K}x{N}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
    for r in (0, 1):
        o = r * 16
        print(f"  cta{r} producer total {d[o+0]} wait_empty {d[o+1]} kblocks {d[o+2]} | epi(w2) total {d[o+8]} wait_tfull {d[o+9]} tmem_ld {d[o+10]} wait_store {d[o+11]} math {d[o+12]} stage {d[o+13]} tma {d[o+14]}")
    print(f"  mma total {d[4]} wait_tempty {d[5]} wait_full {d[6]} tiles {d[7]}")
