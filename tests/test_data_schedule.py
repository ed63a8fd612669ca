"""CPU: the host-side rows of SURVEY.md section 8f — bucketing sampler against golden schedules generated from the
reference's own BucketBatchSampler (oracle/make_golden_sampler.py), batch builders, plateau scheduler against torch's,
early stopping, checkpoint round trip."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bucket_sampler.json")


def test_bucket_sampler_reproduces_reference_schedules():
    from vag_nmt_b200.data import BucketBatchSampler
    for case in json.load(open(GOLD)):
        s = BucketBatchSampler(case["lengths"], case["batch_size"], case["max_len"])
        assert len(s) == case["len"]
        np.random.seed(case["seed"])
        for want in case["epochs"]:
            got = [[int(i) for i in b] for b in s]
            assert got == want
            for b in got:      # the property the training loop relies on: one target length per batch
                assert len({case["lengths"][i] for i in b}) == 1


def test_bucket_sampler_data_parallel_equal_slices_of_global_batches():
    """Data-parallel mode: every rank runs the same number of steps with the SAME local batch size at every step (what the
    global-batch ranking loss and the averaging gradient all-reduce need), one target length per global batch, no sample twice,
    and at most world-1 samples of each bucket dropped per epoch."""
    from vag_nmt_b200.data import BucketBatchSampler
    lengths = [int(x) for x in np.random.RandomState(5).randint(4, 20, size=333)]
    world, bs = 4, 16
    per_rank = [[[int(i) for i in b] for b in BucketBatchSampler(lengths, bs, world_size=world, rank=r, seed=9)] for r in range(world)]
    assert len({len(p) for p in per_rank}) == 1 and len(per_rank[0]) == len(BucketBatchSampler(lengths, bs, world_size=world, rank=0, seed=9))
    for step in range(len(per_rank[0])):
        sizes = {len(per_rank[r][step]) for r in range(world)}
        assert len(sizes) == 1 and 1 <= next(iter(sizes)) <= bs               # equal local batch on every rank
        assert len({lengths[i] for r in range(world) for i in per_rank[r][step]}) == 1
    seen = [i for p in per_rank for b in p for i in b]
    assert len(seen) == len(set(seen))
    n_buckets = len(set(lengths))
    n_batches = sum(-(-lengths.count(n) // (bs * world)) for n in set(lengths))
    assert len(lengths) - len(seen) <= (world - 1) * n_batches and n_buckets > 0
    # the private RandomState advances per epoch and never touches the global numpy RNG the reference sampler uses
    np.random.seed(1)
    before = np.random.get_state()[1].copy()
    s0 = BucketBatchSampler(lengths, bs, world_size=world, rank=0, seed=9)
    e1, e2 = [list(map(int, b)) for b in s0], [list(map(int, b)) for b in s0]
    assert e1 != e2 and (np.random.get_state()[1] == before).all()


def test_train_generator_pads_sorts_and_keeps_rows_together():
    from vag_nmt_b200.data import data_generator_tl_mtv
    rng = np.random.RandomState(0)
    pairs = []
    for i in range(50):
        lx, ly = int(rng.randint(3, 12)), int(rng.randint(3, 6))
        pairs.append(([int(v) for v in rng.randint(4, 90, size=lx - 1)] + [3], [1000 + i] * (ly - 1) + [3]))
    im = rng.rand(50, 6).astype(np.float32)
    im[:, 0] = np.arange(50)
    np.random.seed(3)
    n_rows = 0
    for bx, by, bim, xl, yl in data_generator_tl_mtv(pairs, im, 8):
        assert xl == sorted(xl, reverse=True) and bx.shape == (len(xl), max(xl)) and len(set(yl)) == 1
        for r in range(len(xl)):
            i = int(by[r, 0]) - 1000                       # which sample this row came from
            assert int(bim[r, 0]) == i
            assert bx[r, :xl[r]].tolist() == pairs[i][0] and (bx[r, xl[r]:] == 0).all()
        n_rows += len(xl)
    assert n_rows == 50


def test_eval_generator_and_reorder_round_trip():
    from vag_nmt_b200.data import data_generator_mtv, translation_reorder, translation_reorder_BPE
    rng = np.random.RandomState(1)
    pairs = [([int(v) for v in rng.randint(4, 50, size=int(rng.randint(2, 9)))] + [3], [5, 3]) for _ in range(21)]
    out = []
    for bx, by, bim, xl, yl, order in data_generator_mtv(pairs, None, 8):
        sorted_rows = [bx[r, :xl[r]].tolist() for r in range(len(xl))]
        out.extend(translation_reorder(sorted_rows, order))
    assert out == [p[0] for p in pairs]
    words = {4: "ein@@", 5: "mal", 6: "hund"}
    assert translation_reorder_BPE([[6], [4, 5, 99]], [1, 0], words) == [["einmal", "<unk>"], ["hund"]]


class _FakeOpt:
    def __init__(self, lrs):
        self.param_groups = [{"lr": lr, "params": []} for lr in lrs]


def test_plateau_scheduler_matches_torch():
    from vag_nmt_b200.schedule import ReduceLROnPlateau
    rng = np.random.RandomState(2)
    for trial in range(5):
        p = torch.nn.Parameter(torch.zeros(1))
        ref_opt = torch.optim.Adam([{"params": [p], "lr": 4e-4}, {"params": [torch.nn.Parameter(torch.zeros(1))], "lr": 2e-4}])
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, factor=0.2, patience=3 + trial)
        mine_opt = _FakeOpt([4e-4, 2e-4])
        mine = ReduceLROnPlateau(mine_opt, factor=0.2, patience=3 + trial)
        vals = np.abs(np.cumsum(rng.randn(80)) * 0.05 + 3.0) * np.linspace(1.0, 0.9, 80)
        for v in vals:
            ref.step(float(v))
            mine.step(float(v))
            assert [g["lr"] for g in mine_opt.param_groups] == [g["lr"] for g in ref_opt.param_groups]


def test_early_stopping_counter():
    from vag_nmt_b200.schedule import EarlyStopping
    es = EarlyStopping(3)
    assert es.update(10.0) and es.counter == 3
    assert not es.update(9.0) and es.counter == 2
    assert es.update(11.0) and es.counter == 3
    for _ in range(3):
        assert not es.should_stop
        es.update(1.0)
    assert es.should_stop


def test_checkpoint_round_trip(tmp_path):
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.schedule import ReduceLROnPlateau, load_checkpoint, save_checkpoint
    cfg = dict(synthetic.TINY)
    def build(seed):
        torch.manual_seed(seed)
        return vag.NMT_Seq2Seq_Beam_V2(cfg["src_size"], cfg["tgt_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
                                       cfg["hidden_size"], tied_emb=True)
    a, b = build(1), build(2)
    oa, ob = ClipAdam(a, lr=1e-3), ClipAdam(b, lr=5e-3)
    oa.step_count = 7
    for p in list(a.parameters())[:3]:
        oa.state[p] = (torch.full_like(p, 0.5), torch.full_like(p, 0.25))
    sa, sb = ReduceLROnPlateau(oa, factor=0.2, patience=2), ReduceLROnPlateau(ob, factor=0.2, patience=2)
    for v in (3.0, 3.1, 3.2, 3.3):
        sa.step(v)
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, a, oa, sa, extra={"iter": 42})
    extra = load_checkpoint(path, b, ob, sb)
    assert extra == {"iter": 42} and ob.step_count == 7
    for (k, v), (_, u) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, u), k
    assert [g["lr"] for g in ob.param_groups] == [g["lr"] for g in oa.param_groups] and oa.param_groups[0]["lr"] < 1e-3
    assert sb.best == sa.best and sb.num_bad_epochs == sa.num_bad_epochs
    assert len(ob.state) == 3 and all(float(m.mean()) == 0.5 for m, _ in ob.state.values())
