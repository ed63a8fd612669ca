"""One beam-12 decode for profilers: python tools/decode_once.py [sentences] [max_length] [precision] (no graph replay, no polling)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("VAG_DECODE_GRAPH", "0")
import bench  # noqa: E402
from vag_nmt_b200 import ops, synthetic, _cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
model.precision = prec
cfg = synthetic.DE
sents, im = synthetic.make_corpus(n, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
src_d, im_d = src.to(dev), im_s.to(dev)
with _cabi.precision_scope(prec):
    w, ctx, mask, keys, h0, _, _ = model._prepare(src_d, lens, im_d)
    for _ in range(2):
        ops.beam_decode(w, h0, keys, ctx, mask, 12, L, early_stop=False)
torch.cuda.synchronize()
print("decoded", n, "sentences,", L, "steps,", prec)
