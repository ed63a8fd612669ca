import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from vag_nmt_b200 import synthetic
cfg = synthetic.DE
model = bench.build_cpu_params().cuda()
sents, im = synthetic.make_corpus(1000, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
src, im_s = src.cuda(), im_s.cuda()
def run(): return model.decode_device(src, lens, im_s, 12, 80)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
print("eager ms/step", e0.elapsed_time(e1) / 3)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    run()
torch.cuda.current_stream().wait_stream(s)
try:
    with torch.cuda.graph(g):
        out = run()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("graph ms/step", e0.elapsed_time(e1) / 3)
    ref = run()
    torch.cuda.synchronize()
    print("same tokens", bool((ref[0] == out[0]).all()))
except Exception as ex:
    print("capture failed:", repr(ex)[:300])
