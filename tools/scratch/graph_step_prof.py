"""A few graphed training steps on ONE batch shape (for an ncu launch list: the last step's kernels are the tail of the list)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vag_nmt_b200 as vag
from vag_nmt_b200 import synthetic, _cabi
from vag_nmt_b200.optim import ClipAdam
from vag_nmt_b200.train import GraphedTrainStep
cfg = synthetic.DE
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(1234)
model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], 256, 256, 512, 512, 0.99, tied_emb=True).cuda()
model.precision = prec
opt = ClipAdam(model, lr=4e-4)
w = torch.ones(cfg["tgt_size"], device="cuda"); w[0] = 0
crit = torch.nn.NLLLoss(weight=w, reduction="none")
cv = vag.PairwiseRankingLoss(margin=0.1)
bt = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100)
print("Ts", bt.src.shape[1], "Tt", bt.tgt.shape[1])
stepper = GraphedTrainStep(model, opt, crit, cv)
lib = _cabi.lib()
for i in range(n):
    n0 = lib.vag_launch_count()
    loss = stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im, 1.0)[0]
    torch.cuda.synchronize()
    print("step", i, "host launches", lib.vag_launch_count() - n0, "loss", float(loss))
