"""Stream-enqueued (polled) decode loop against CUDA-graph replay (chunks of steps, VAG_DECODE_CHUNK; 0 = the whole loop as one
graph) at several batch sizes:  python tools/graph_rows_ab.py [fp32|bf16] 125 250 500 1000"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from vag_nmt_b200 import synthetic, _cabi  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
sizes = [int(a) for a in sys.argv[2:]] or [125, 250, 500, 1000]
dev = torch.device("cuda", 0)
model = bench.build_cpu_params().to(dev)
model.precision = prec
cfg = synthetic.DE
for n in sizes:
    sents, im = synthetic.make_corpus(n, cfg["src_size"], cfg["im_feats_size"], seed=7)
    src, lens, im_s, _ = synthetic.pad_and_sort(sents, im)
    src_d, im_d = src.to(dev), im_s.to(dev)
    res = {}
    for graph in (False, True, False, True):
        type(model)._GRAPH_ROWS_MAX = 10 ** 9 if graph else 0
        ts = []
        with _cabi.precision_scope(prec):
            w, ctx, mask, keys, h0, _, _ = model._prepare(src_d, lens, im_d)
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                hyp, hl = model._beam_decode(w, h0, keys, ctx, mask, 12, 80)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
        ts.sort()
        res.setdefault(graph, []).append(ts[len(ts) // 2])
    print(f"{prec} n={n:5d}: stream {min(res[False]):8.3f} ms   graph {min(res[True]):8.3f} ms")
