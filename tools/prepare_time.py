"""Time the per-call preparation of a decode (encoder, visual pooling, decoder init, key projection) on the bench corpus:
python tools/prepare_time.py [sentences]   (VAG_ENC_FUSED=0 selects the three-kernel encoder steps)"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from vag_nmt_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
cfg = synthetic.DE
sents, im = synthetic.make_corpus(n, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, _ = synthetic.pad_and_sort(sents, im)
src_d, im_d = src.to(dev), im_s.to(dev)
for _ in range(3):
    model._prepare(src_d, lens, im_d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    model._prepare(src_d, lens, im_d)
e1.record()
torch.cuda.synchronize()
print(f"prepare {n} sentences: {e0.elapsed_time(e1) / 20:.3f} ms per call, encoder steps:",
      "three-kernel" if os.environ.get("VAG_ENC_FUSED") == "0" else "fused cell")
