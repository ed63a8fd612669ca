"""Where the time of one reference-sized decode call (16 sentences, beam 12, max_length 80) goes: encoder + pooling + hoists
(eager launches) against the graph-replayed 80-step loop.  python tools/ref_batching_profile.py"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from vag_nmt_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
cfg = synthetic.DE
sents, im = synthetic.make_corpus(64, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, order = synthetic.pad_and_sort(sents[:16], im[:16])
src_d, im_d = src.to(dev), im_s.to(dev)


def timeit(fn, n=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


from vag_nmt_b200 import _cabi
with _cabi.precision_scope("fp32"):
    t_prep = timeit(lambda: model._prepare(src_d, lens, im_d))
    w, ctx, mask, keys, h0, _, _ = model._prepare(src_d, lens, im_d)
    t_graph = timeit(lambda: model._beam_decode(w, h0, keys, ctx, mask, 12, 80))
    import os
    os.environ["VAG_DECODE_GRAPH"] = "0"
    t_eager_poll = timeit(lambda: ops.beam_decode(w, h0, keys, ctx, mask, 12, 80, early_stop=True))
    t_eager = timeit(lambda: ops.beam_decode(w, h0, keys, ctx, mask, 12, 80, early_stop=False))
    os.environ["VAG_DECODE_GRAPH"] = "1"
t_full = timeit(lambda: model.beamsearch_decode(src, lens, im_s, beam_size=12, max_length=80))
print(f"B=16 T={src.shape[1]}: prepare {t_prep:.2f} ms | loop graph {t_graph:.2f} ms | loop eager {t_eager:.2f} ms (polled {t_eager_poll:.2f}) | "
      f"beamsearch_decode end to end {t_full:.2f} ms")
