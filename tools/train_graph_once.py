"""One CUDA-graph-replayed training step for profilers (bf16 mode, EN→DE, batch 32): python tools/train_graph_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vag_nmt_b200 as vag  # noqa: E402
from vag_nmt_b200 import synthetic  # noqa: E402
from vag_nmt_b200.optim import ClipAdam  # noqa: E402
from vag_nmt_b200.train import GraphedTrainStep  # noqa: E402

cfg = synthetic.DE
dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
model.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
opt = ClipAdam(model, lr=4e-4)
w = torch.ones(cfg["tgt_size"], device=dev)
w[0] = 0
stepper = GraphedTrainStep(model, opt, torch.nn.NLLLoss(weight=w, reduction="none"), vag.PairwiseRankingLoss(margin=0.1))
bt = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100)
for _ in range(4):
    loss = stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im, 1.0)[0]
torch.cuda.synchronize()
print("Ts", bt.src.shape[1], "Tt", bt.tgt.shape[1], "loss", float(loss))
if "time" in sys.argv:      # replay timing (CUDA events, 3 x 40 steps, best): python tools/train_graph_once.py bf16 time
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            stepper.step(bt.src, bt.src_lengths, bt.tgt, bt.im, 1.0)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    print(f"ms_per_step {best:.4f}  PDL_MASK={os.environ.get('VAG_PDL_MASK', 'default')}")
