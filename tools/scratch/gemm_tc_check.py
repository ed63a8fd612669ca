import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vag_nmt_b200 import ops, train_ops as T
torch.manual_seed(0)
def check(M, N, K, ta, tb, beta):
    A = torch.randn(K, M, device="cuda") if ta else torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") if tb else torch.randn(K, N, device="cuda")
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    if beta:
        C0 = torch.randn(M, N, device="cuda"); out = C0.clone(); ref = ref + C0.double()
        T.gemm(A, B, trans_a=ta, trans_b=tb, out=out, beta=1.0)
    else:
        out = T.gemm(A, B, trans_a=ta, trans_b=tb)
    err = float((out.double() - ref).abs().max()) / float(ref.abs().max())
    print(f"M {M} N {N} K {K} ta {ta} tb {tb} beta {beta}: rel err {err:.2e}")
    assert err < 5e-6, err
for (M, N, K) in ((1536, 512, 352), (352, 256, 1536), (832, 1024, 1024), (9391, 256, 352), (1024, 1024, 830), (100, 64, 36)):
    for ta in (False, True):
        for tb in (False, True):
            for beta in (0, 1):
                check(M, N, K, ta, tb, beta)
print("ok")
