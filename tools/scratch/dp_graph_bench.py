"""torchrun: data-parallel training step, eager vs GraphedTrainStep with the collectives (all-gather of the embeddings for
the global ranking loss) captured inside the graph; gradient all-reduce stays eager in ClipAdam.step."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vag_nmt_b200 as vag
from vag_nmt_b200 import synthetic
from vag_nmt_b200.optim import ClipAdam
from vag_nmt_b200.train import GraphedTrainStep, DistributedPairwiseRankingLoss
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = synthetic.DE
for graphed in (False, True):
    torch.manual_seed(1234)
    model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], 256, 256, 512, 512, 0.99, tied_emb=True).cuda()
    model.precision = "bf16"
    opt = ClipAdam(model, lr=4e-4)
    w = torch.ones(cfg["tgt_size"], device="cuda"); w[0] = 0
    crit = torch.nn.NLLLoss(weight=w, reduction="none")
    cv = DistributedPairwiseRankingLoss(margin=0.1)
    batches = [synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100 + 8 * i + rank) for i in range(8)]
    pinned = [(bt.src.pin_memory(), bt.tgt.pin_memory(), bt.im.pin_memory(), bt.src_lengths) for bt in batches]
    toks = [int((bt.tgt != 0).sum()) for bt in batches]
    stepper = GraphedTrainStep(model, opt, crit, cv, enabled=graphed)
    for i in range(8): loss = stepper.step(*[pinned[i % 8][j] for j in (0, 3, 1, 2)], 1.0)[0]
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 32
    a.record()
    for i in range(n): loss = stepper.step(*[pinned[i % 8][j] for j in (0, 3, 1, 2)], 1.0)[0]
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / n], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]; dist.all_gather(allc, chk)
    if rank == 0:
        print(f"world {world} graphed={graphed}: {float(ms):.3f} ms/step, ~{sum(toks) * (n // 8) * world / (float(ms) * n / 1e3):.0f} tok/s, loss {float(loss):.5f}, replicas in sync: {all(float(c) == float(allc[0]) for c in allc)}")
dist.barrier(); torch.cuda.synchronize()
if rank == 0: print("clean exit", flush=True)
dist.destroy_process_group()
