"""Small workload for compute-sanitizer: the tcgen05 / TMA / mbarrier / cluster kernels of the fused decode step (CTA-pair
contraction, vocabulary top-2 kernel, summary selection, attention, reorder) and one training step (≤ 32-row kernels, GRU-fused
epilogues, tensor-core backward contractions).  python tools/sanitize_run.py [decode|train]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("VAG_DECODE_GRAPH", "0")
import bench  # noqa: E402
import vag_nmt_b200 as vag  # noqa: E402
from vag_nmt_b200 import synthetic  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "decode"
dev = torch.device("cuda", 0)
cfg = synthetic.DE
model = bench.build_cpu_params().to(dev)
if what == "decode":
    b = synthetic.make_batch(160, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=21)      # 1920 rows: the fused tensor-core step
    out = model.beamsearch_decode(b.src, b.src_lengths, b.im, beam_size=12, max_length=4)
    model.precision = "bf16"
    out2 = model.beamsearch_decode(b.src, b.src_lengths, b.im, beam_size=12, max_length=3)
    print("decoded", len(out), len(out2))
else:
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import train_imagine_beam
    b = synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=7, max_len=12)
    w = torch.ones(cfg["tgt_size"], device=dev)
    w[0] = 0
    for prec in ("fp32", "bf16"):
        model.precision = prec
        loss = train_imagine_beam(b.src, b.tgt, b.im, b.src_lengths, model, ClipAdam(model, lr=4e-4), torch.nn.NLLLoss(weight=w, reduction="none"),
                                  vag.PairwiseRankingLoss(margin=0.1), 0.99, 1.0)
        print("train step", prec, loss)
torch.cuda.synchronize()
