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
