"""Where the end-to-end decode call spends its time beyond the device loop: python tools/e2e_phases.py
(host clock with a synchronise after every phase — the phases do not overlap in this script, in the real call some do)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from vag_nmt_b200 import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
cfg = synthetic.DE
sents, im = synthetic.make_corpus(1000, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, _ = synthetic.pad_and_sort(sents, im)
src_pin, im_pin = src.pin_memory(), im_s.pin_memory()
for _ in range(2):
    model.beamsearch_decode(src_pin.to(dev), lens, im_pin.to(dev), beam_size=12, max_length=80)
acc = {}


def lap(name, t0):
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


N = 5
for _ in range(N):
    torch.cuda.synchronize()
    t = time.perf_counter()
    s = src_pin.to(dev, non_blocking=True)
    i = im_pin.to(dev, non_blocking=True)
    t = lap("h2d", t)
    with model.precision_scope():
        w, ctx, mask, keys, h0, _, _ = model._prepare(s, lens, i)
        t = lap("prepare (encoder, pooling, keys, h0)", t)
        hyp, hyp_len = model._beam_decode(w, h0, keys, ctx, mask, 12, 80)
        t = lap("beam loop", t)
    rows = hyp.cpu().numpy()
    ln = hyp_len.cpu().tolist()
    t = lap("d2h", t)
    out = [rows[b, :ln[b]].tolist() for b in range(rows.shape[0])]
    t = lap("python token lists", t)
    t0 = time.perf_counter()
    torch.cuda.synchronize()
    full = model.beamsearch_decode(src_pin.to(dev, non_blocking=True), lens, im_pin.to(dev, non_blocking=True), beam_size=12, max_length=80)
    lap("whole call (for comparison)", t0)
for k, v in acc.items():
    print(f"{v / N:8.3f} ms  {k}")
