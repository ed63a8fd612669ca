import os, sys, torch
sys.path.insert(0, '/root/repo')
import bench
from vag_nmt_b200 import _cabi, synthetic
lib = _cabi.lib()
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.vag_tc_set_debug(dbg.data_ptr())
cfg = synthetic.DE
model = bench.build_cpu_params().cuda()
sents, im = synthetic.make_corpus(1000, cfg["src_size"], cfg["im_feats_size"], seed=7)
src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
model.decode_device(src.cuda(), lens, im_s.cuda(), 12, 80)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
print("refills", d[21], "recomputes", d[20], "fraction", d[20] / max(d[21], 1))
