"""Eager step driver vs GraphedTrainStep on the bench's training batches."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vag_nmt_b200 as vag
from vag_nmt_b200 import synthetic
from vag_nmt_b200.optim import ClipAdam
from vag_nmt_b200.train import GraphedTrainStep, train_imagine_beam
cfg = synthetic.DE
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for graphed in (False, True):
    torch.manual_seed(1234)
    model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], 256, 256, 512, 512, 0.99, tied_emb=True).cuda()
    model.precision = prec
    opt = ClipAdam(model, lr=4e-4)
    w = torch.ones(cfg["tgt_size"], device="cuda"); w[0] = 0
    crit = torch.nn.NLLLoss(weight=w, reduction="none")
    cv = vag.PairwiseRankingLoss(margin=0.1)
    batches = [synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100 + 8 * i) for i in range(8)]
    pinned = [(bt.src.pin_memory(), bt.tgt.pin_memory(), bt.im.pin_memory(), bt.src_lengths) for bt in batches]
    toks = [int((bt.tgt != 0).sum()) for bt in batches]
    stepper = GraphedTrainStep(model, opt, crit, cv)
    def step(i):
        src, tgt, im, lens = pinned[i % 8]
        if graphed:
            return stepper.step(src, lens, tgt, im, 1.0)[0]
        return train_imagine_beam(src.cuda(non_blocking=True), tgt.cuda(non_blocking=True), im.cuda(non_blocking=True), lens, model, opt, crit, cv, 0.99, 1.0, sync=False)[0]
    for i in range(8): loss = step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 32
    a.record()
    for i in range(n): loss = step(i)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"{prec} graphed={graphed}: {ms:.3f} ms/step, {sum(toks) * (n // 8) / (ms * n / 1e3):.0f} tok/s, loss {float(loss):.5f}")
