"""Where does a training step spend its time? (host-side wall clock with device syncs between phases)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vag_nmt_b200 as vag
from vag_nmt_b200 import synthetic, _cabi
from vag_nmt_b200.optim import ClipAdam

cfg = synthetic.DE
torch.manual_seed(1234)
model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], 256, 256, 512, 512, 0.99, tied_emb=True).cuda()
opt = ClipAdam(model, lr=4e-4)
w = torch.ones(cfg["tgt_size"], device="cuda"); w[0] = 0
crit = torch.nn.NLLLoss(weight=w, reduction="none")
cv = vag.PairwiseRankingLoss(margin=0.1)
bt = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100)
src, tgt, im = bt.src.cuda(), bt.tgt.cuda(), bt.im.cuda()
print("Ts", src.shape[1], "Tt", tgt.shape[1])
lib = _cabi.lib()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    model.train(); opt.zero_grad()
    n0 = lib.vag_launch_count(); t0 = sync()
    loss, _, _ = model(src, bt.src_lengths, tgt, im, 1.0, criterion_mt=crit, criterion_vse=cv)
    n1 = lib.vag_launch_count(); t1 = sync()
    loss.backward()
    n2 = lib.vag_launch_count(); t2 = sync()
    opt.step()
    n3 = lib.vag_launch_count(); t3 = sync()
    if it >= 3:
        print(f"fwd {1e3*(t1-t0):6.2f} ms ({n1-n0} launches)  bwd {1e3*(t2-t1):6.2f} ms ({n2-n1})  opt {1e3*(t3-t2):6.2f} ms ({n3-n2})")
