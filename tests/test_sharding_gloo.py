"""CPU, world_size 2, gloo: sentence-sharded decoding returns exactly the single-process result on every rank."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from conftest import build_mm, cpu_params
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.translate import decode_corpus, decode_corpus_sharded
    cfg = synthetic.TINY
    p = cpu_params(build_mm(cfg, 3))
    sents, im = synthetic.make_corpus(9, cfg["src_size"], cfg["im_feats_size"], seed=4, max_len=8, min_len=1, mean=4, std=2)
    fn = lambda src, lens, im_b, K, L: O.multimodal_beamsearch_decode(p, src, lens, im_b, K, L)
    sharded = decode_corpus_sharded(fn, sents, im, 3, 9)
    single = decode_corpus(fn, sents, im, 3, 9)
    torch.save(dict(ok=sharded == single, n=len(sharded)), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_decode_equals_single(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        res = torch.load(tmp_path / f"r{r}.pt")
        assert res["ok"] and res["n"] == 9


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vag_nmt_b200.optim import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.Linear(5, 3)
    x = torch.arange(20, dtype=torch.float32).reshape(4, 5) / 10 + rank     # a different shard per rank
    net(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    allreduce_gradients(list(net.parameters()))
    torch.save(dict(local=local, avg=[p.grad.clone() for p in net.parameters()]), os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_is_the_mean(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    for a0, a1, l0, l1 in zip(r0["avg"], r1["avg"], r0["local"], r1["local"]):
        assert torch.equal(a0, a1)                                   # every rank ends with the same gradient
        assert torch.allclose(a0, (l0 + l1) / 2, atol=1e-6)          # … the mean of the per-rank gradients
