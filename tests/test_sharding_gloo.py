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
    contiguous = decode_corpus_sharded(fn, sents, im, 3, 9, balance=False)
    single = decode_corpus(fn, sents, im, 3, 9)
    # retrieval evaluation: sharded embedding + all-gather + replicated recall == single process
    from vag_nmt_b200.translate import embed_corpus, embed_corpus_sharded, retrieval_eval_sharded
    efn = lambda src, lens, im_b: O.embed_sent_im(p, src, lens, im_b)
    lim, ltxt = embed_corpus_sharded(efn, sents, im, batch_size=4)
    lim1, ltxt1 = embed_corpus(efn, sents, im, batch_size=4)
    recall = retrieval_eval_sharded(efn, sents, im, batch_size=4, rank_fn=O.t2i)
    emb_ok = bool(torch.allclose(lim, lim1, atol=1e-6) and torch.allclose(ltxt, ltxt1, atol=1e-6))
    torch.save(dict(ok=sharded == single and contiguous == single, n=len(sharded), emb_ok=emb_ok, recall=[float(x) for x in recall],
                    recall_single=[float(x) for x in O.t2i(lim1, ltxt1)]), os.path.join(out_dir, f"r{rank}.pt"))
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
        assert res["emb_ok"] and res["recall"] == res["recall_single"]


def test_shard_indices_balanced_and_complete():
    from vag_nmt_b200.translate import shard_indices
    lengths = [5, 9, 3, 9, 7, 1, 4, 8, 2, 6, 5]
    for world in (1, 2, 3, 4, 8):
        for balance in (True, False):
            parts = [shard_indices(lengths, world, r, balance) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(len(lengths)))      # a partition of the corpus
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        loads = [sum(lengths[i] for i in shard_indices(lengths, world, r)) for r in range(world)]
        assert max(loads) - min(loads) <= max(lengths)                                   # round-robin over the sorted corpus


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
