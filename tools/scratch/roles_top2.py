import math, os, sys, torch
sys.path.insert(0, '/root/repo')
from vag_nmt_b200 import _cabi, ops
lib = _cabi.lib()
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.vag_tc_set_debug(dbg.data_ptr())
rows, K, N = 12000, 256, 9391
x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 16; b = torch.randn(N, device="cuda")
xs, ws = ops.tc_split(x), ops.tc_split(w)
for _ in range(3): summ = ops.tc_gemm_top2(xs, ws, rows, K, N, b)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): summ = ops.tc_gemm_top2(xs, ws, rows, K, N, b)
e1.record(); torch.cuda.synchronize()
d = dbg.cpu().tolist()
print(f"top2 {rows}x{K}x{N}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
for r in (0, 1):
    o = r * 16
    print(f"  cta{r} producer total {d[o+0]} wait_empty {d[o+1]} kblocks {d[o+2]} | epi(w2) total {d[o+8]} wait_tfull {d[o+9]} tmem_ld {d[o+10]} wait_store {d[o+11]} math {d[o+12]} stage {d[o+13]} tma {d[o+14]}")
print(f"  mma total {d[4]} wait_tempty {d[5]} wait_full {d[6]} tiles {d[7]}")
# check vs fp64
ref = (x.double() @ w.double().t() + b.double())
tiles = (N + 31) // 32
pad = torch.full((rows, tiles * 32 - N), -float("inf"), device="cuda", dtype=torch.float64)
rt = torch.cat([ref, pad], 1).view(rows, tiles, 32)
top = rt.topk(2, dim=2)
summ = summ.permute(1, 0, 2).contiguous()
bits = summ[..., 3].contiguous().view(torch.int32)
i1 = (bits & 0xFFFF).long(); i2 = ((bits >> 16) & 0xFFFF).long()
base = (torch.arange(tiles, device="cuda") * 32).view(1, tiles, 1)
gi = top.indices + base
print("best idx match", float((i1 == gi[..., 0]).float().mean()), "second idx match", float((i2 == gi[..., 1]).float().mean()))
print("best val err", float((summ[..., 0].double() - top.values[..., 0]).abs().max()), "second val err", float((summ[..., 2].double() - top.values[..., 1]).abs().max()))
lse_ref = torch.logsumexp(ref, 1)
m = summ[..., 0].max(1).values
lse = m + torch.log((summ[..., 1] * torch.exp(summ[..., 0] - m[:, None])).sum(1))
print("lse err", float((lse.double() - lse_ref).abs().max()))
