import os, sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from vag_nmt_b200 import synthetic
cfg = synthetic.DE
model = bench.build_cpu_params().cuda()
sents, im = synthetic.make_corpus(1000, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
src_pin, im_pin = src.pin_memory(), im_s.pin_memory()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0 = sync()
    s = src_pin.to("cuda", non_blocking=True); i = im_pin.to("cuda", non_blocking=True)
    t1 = sync()
    hyp, hl = model.decode_device(s, lens, i, 12, 80)
    t2 = time.perf_counter()   # enqueue done
    t3 = sync()
    rows = hyp.cpu().tolist(); ls = hl.cpu().tolist()
    t4 = time.perf_counter()
    out = [rows[b][:ls[b]] for b in range(len(ls))]
    t5 = time.perf_counter()
    print(f"h2d {1e3*(t1-t0):.2f}  enqueue {1e3*(t2-t1):.2f}  gpu-wait {1e3*(t3-t2):.2f}  d2h+tolist {1e3*(t4-t3):.2f}  lists {1e3*(t5-t4):.2f}  total {1e3*(t5-t0):.2f}")
t0 = sync()
for _ in range(3):
    s = src_pin.to("cuda", non_blocking=True); i = im_pin.to("cuda", non_blocking=True)
    model.beamsearch_decode(s, lens, i, beam_size=12, max_length=80)
t1 = sync()
print("beamsearch_decode e2e ms", 1e3*(t1-t0)/3)
