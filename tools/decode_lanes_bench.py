"""Decode lanes A/B on one GPU: n sentences (what one rank of an N-GPU run holds) with 1, 2, 4, 8 lanes, and the reference's
eval batching with 1-8 batches in flight.   python tools/decode_lanes_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vag_nmt_b200 import synthetic  # noqa: E402
from vag_nmt_b200.translate import decode_corpus, decode_corpus_pipelined  # noqa: E402

cfg = synthetic.DE
dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
K, L = 12, 80
sents, im = synthetic.make_corpus(1000, cfg["src_size"], cfg["im_feats_size"], seed=7)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = {}
for n in (125, 250, 500, 1000):
    order = sorted(range(1000), key=lambda i: -len(sents[i]))[:: 1000 // n][:n]
    src, lens, im_s, _ = synthetic.pad_and_sort([sents[i] for i in order], im[order])
    src, im_s = src.to(dev), im_s.to(dev)
    ref = None
    for lanes in ("1", "2", "3", "4", "6", "8"):
        os.environ["VAG_DECODE_LANES"] = lanes
        plan = model._lane_plan(n, K)
        if lanes != "1" and plan is None:
            continue
        hyp, hl = model.decode_device(src, lens, im_s, K, L)
        if ref is None:
            ref = (hyp.clone(), hl.clone())
        same = bool(torch.equal(hl, ref[1])) and all(torch.equal(hyp[b, :int(hl[b])], ref[0][b, :int(hl[b])]) for b in range(0, n, 7))
        ms = timed(lambda: model.decode_device(src, lens, im_s, K, L), 3)
        print(f"n={n:5d} lanes={lanes:>4s} ({'-' if plan is None else len(plan)}): {ms:8.2f} ms  {n / ms * 1e3:9.0f} sent/s  same_tokens={same}", flush=True)
os.environ["VAG_DECODE_LANES"] = "1"
fn = lambda s_, l_, i_, K_, L_: model.beamsearch_decode(s_, l_, i_, beam_size=K_, max_length=L_)
seq = decode_corpus(fn, sents, im, K, L, batch_size=16)
ms = timed(lambda: decode_corpus(fn, sents, im, K, L, batch_size=16), 1)
print(f"reference batching sequential: {ms:8.1f} ms  {1000 / ms * 1e3:8.0f} sent/s", flush=True)
for lanes in (2, 4, 8, 16):
    got = decode_corpus_pipelined(model, sents, im, K, L, 16, lanes=lanes)
    ms = timed(lambda: decode_corpus_pipelined(model, sents, im, K, L, 16, lanes=lanes), 1)
    print(f"reference batching, {lanes} in flight: {ms:8.1f} ms  {1000 / ms * 1e3:8.0f} sent/s  same_tokens={sum(int(a == b) for a, b in zip(got, seq))}/1000", flush=True)
