"""Experiment: capture forward+backward of one training batch in a CUDA graph; replay time = pure GPU time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vag_nmt_b200 as vag
from vag_nmt_b200 import synthetic, _cabi
from vag_nmt_b200.optim import ClipAdam

cfg = synthetic.DE
torch.manual_seed(1234)
model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], 256, 256, 512, 512, 0.99, tied_emb=True).cuda()
if len(sys.argv) > 1: model.precision = sys.argv[1]
opt = ClipAdam(model, lr=4e-4)
w = torch.ones(cfg["tgt_size"], device="cuda"); w[0] = 0
crit = torch.nn.NLLLoss(weight=w, reduction="none")
cv = vag.PairwiseRankingLoss(margin=0.1)
bt = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100)
src, tgt, im = bt.src.cuda(), bt.tgt.cuda(), bt.im.cuda()
print("Ts", src.shape[1], "Tt", tgt.shape[1])
model.train()
def fb():
    opt.zero_grad()
    loss, lm, lv = model(src, bt.src_lengths, tgt, im, 1.0, criterion_mt=crit, criterion_vse=cv)
    with model.precision_scope():
        loss.backward()
    return loss
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): fb()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = fb()
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): g.replay()
b.record(); torch.cuda.synchronize()
print("graph replay fwd+bwd: %.3f ms, loss %.6f" % (a.elapsed_time(b) / 10, float(loss)))
t0 = time.perf_counter()
for _ in range(10): fb()
torch.cuda.synchronize()
print("eager fwd+bwd: %.3f ms" % ((time.perf_counter() - t0) * 100))
