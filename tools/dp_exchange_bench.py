"""Peer-memory gradient exchange against NCCL on the 64 MB flat gradient buffer: torchrun --nproc-per-node N tools/dp_exchange_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vag_nmt_b200.optim import PeerGradientExchange  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 15_944_624
ex = PeerGradientExchange(n, dev)
ss = torch.zeros(1, device=dev)
flat2 = torch.randn(n, device=dev)
ex.flat[:n].copy_(flat2)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


t_p2p = timed(lambda: ex.allreduce(ss))
t_nccl = timed(lambda: dist.all_reduce(flat2, op=dist.ReduceOp.AVG))
if rank == 0:
    print(f"world {dist.get_world_size()}: 63.8 MB gradient exchange  peer-memory kernels (+ norm) {t_p2p * 1e3:.1f} us   NCCL all-reduce alone {t_nccl * 1e3:.1f} us")
dist.destroy_process_group()
