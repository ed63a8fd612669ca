import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from vag_nmt_b200 import synthetic
from vag_nmt_b200.optim import ClipAdam
from vag_nmt_b200.train import DistributedPairwiseRankingLoss, train_imagine_beam
cfg = synthetic.DE
model = bench.build_cpu_params().cuda()
model.precision = os.environ.get("PREC", "fp32")
opt = ClipAdam(model, lr=4e-4)
w = torch.ones(cfg["tgt_size"], device="cuda"); w[0] = 0
crit_mt = torch.nn.NLLLoss(weight=w, reduction="none")
crit_vse = DistributedPairwiseRankingLoss(margin=0.1)
bt = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100)
src, tgt, im = bt.src.cuda(), bt.tgt.cuda(), bt.im.cuda()
for i in range(int(os.environ.get("STEPS", "6"))):
    train_imagine_beam(src, tgt, im, bt.src_lengths, model, opt, crit_mt, crit_vse, 0.99, 1.0, sync=False)
torch.cuda.synchronize()
print("done")
