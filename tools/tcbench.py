import torch, math, sys
sys.path.insert(0, '/root/repo')
from vag_nmt_b200 import ops
def rel(a,b): return float((a.double()-b).abs().max()/b.abs().max())
for rows,K,N in [(12000,256,9391),(12000,512,1536),(12000,1024,512),(12000,512,1024),(12000,1024,256),(12000,256,1536),(20000,1024,1024),(12000,512,256)]:
    x=torch.randn(rows,K,device='cuda'); w=torch.randn(N,K,device='cuda')/math.sqrt(K); b=torch.randn(N,device='cuda')
    ref=x.double()@w.double().t()+b.double()
    y=ops.linear_tc(x,w,b); ys=ops.linear(x,w,b)
    e=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(3): ops.linear_tc(x,w,b,out=y)
    e[0].record()
    for _ in range(10): ops.linear_tc(x,w,b,out=y)
    e[1].record()
    for _ in range(3): ops.linear(x,w,b,out=ys)
    e[2].record()
    for _ in range(10): ops.linear(x,w,b,out=ys)
    e[3].record(); torch.cuda.synchronize()
    t_tc=e[0].elapsed_time(e[1])/10; t_s=e[2].elapsed_time(e[3])/10
    fl=2.0*rows*K*N
    print(f"{rows}x{K}x{N}: tc {t_tc*1e3:8.1f} us {fl/t_tc/1e9:7.1f} TF/s err {rel(y,ref):.2e} | simt {t_s*1e3:8.1f} us {fl/t_s/1e9:6.1f} TF/s err {rel(ys,ref):.2e}")
