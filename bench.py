#!/usr/bin/env python
"""Headline benchmark: beam-12 decoding of a 1000-sentence Multi30K-shaped test set (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: every rank decodes ITS OWN 1000-sentence
shard (weak scaling, no data-path collective — sentences are independent) with
``model.beamsearch_decode`` at beam 12, max_length 80 (the reference's MAX_LENGTH), EN→DE shapes, FP32
token-exact mode.  Prints ONE JSON line on rank 0 (contract in the task statement):

  value     decoded sentences/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public API with HOST (pinned) inputs and host token lists out
  roofline  the dominant kernel (vocabulary-projection contraction) timed alone with CUDA events
  cpu_baseline  the CPU oracle (a port of the reference's algorithm, reference batching 16) on a bounded sample

``--impl reference`` times that CPU port alone (rank 0 only), each step = one reference batch of 16 sentences.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "decoded sents/sec (beam=12)"
UNIT = "sentences/s"
BEAM, MAX_LEN, N_SENT, REF_BATCH = 12, 80, 1000, 16
P_STEP_DE = 6_663_936          # weights touched per decoder-step row


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sentences", type=int, default=N_SENT, help="sentences per rank and step")
    ap.add_argument("--beam", type=int, default=BEAM)
    ap.add_argument("--max-length", type=int, default=MAX_LEN)
    ap.add_argument("--cpu-sample", type=int, default=32, help="sentences of the cpu_baseline sample (0 = skip)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------ CPU reference arm
def oracle_decode_time(params, sents, im, beam, max_len, batch, threads):
    """Decode `sents` with the CPU oracle the way the reference's driver does (eval batches of 16, per-batch
    length sort, attn_e recomputed every step as in NMT_Decoder.py:47).  Returns seconds."""
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    torch.set_num_threads(threads)
    out = [None] * len(sents)
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(0, len(sents), batch):
            src, lens, im_s, order = synthetic.pad_and_sort(sents[i:i + batch], im[i:i + batch])
            toks = O.multimodal_beamsearch_decode(params, src, lens, im_s, beam, max_len, hoist_keys=False)
            for r, c in enumerate(order):   # un-sort like translation_reorder_BPE (preprocessing.py:475-486)
                out[i + c] = toks[r]
    return time.perf_counter() - t0, out


def build_cpu_params(seed=1234):
    """Random-init EN→DE weights with the reference's init (host tensors for the oracle)."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    cfg = synthetic.DE
    torch.manual_seed(seed)
    m = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"],
                                                  cfg["src_embedding_size"], cfg["tgt_embedding_size"], cfg["hidden_size"],
                                                  cfg["shared_embedding_size"], 0.99, tied_emb=True).eval()
    return m


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vag_nmt_b200 import synthetic
    threads = os.cpu_count() or 1
    model = build_cpu_params()
    params = {k: v.detach() for k, v in model.state_dict().items()}
    sents, im = synthetic.make_corpus(REF_BATCH * (args.steps + args.warmup), synthetic.DE["src_size"],
                                      synthetic.DE["im_feats_size"], seed=7)
    times = []
    for s in range(args.warmup + args.steps):
        lo = s * REF_BATCH
        dt, _ = oracle_decode_time(params, sents[lo:lo + REF_BATCH], im[lo:lo + REF_BATCH], args.beam, args.max_length, REF_BATCH, threads)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = REF_BATCH * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, per_step=REF_BATCH),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{len(times)} reference eval batches of {REF_BATCH} sentences, beam {args.beam}, max_length {args.max_length}, "
                                       "oracle/vag_oracle.py (torch CPU fp32, all host threads)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, per_step):
    return {"workload": "VAG-NMT EN->DE beam-12 decoding of a 1000-sentence test-set shape (BASELINE configs[2])",
            "model": "NMT_AttentionImagine_Seq2Seq_Beam_V11 E256 H512 S512 I2048 V9391 random-init (seed 1234)",
            "sentences_per_rank_per_step": per_step, "beam": args.beam, "max_length": args.max_length,
            "src_len": "clip(round(N(14,4.5)),4,40)", "parallelism": f"dp{args.gpus} (sentence-sharded, no collective)",
            "l2": "one bench step = 80 decoder steps over 12000 rows; the activations + summaries + weights each decoder step touches (~0.5 GB) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------ training step (second metric)
def oracle_train_step_time(model_cpu, batches, threads, lr=4e-4):
    """One optimisation step the reference's way (train.py:36-51) through the CPU oracle + torch autograd."""
    from oracle import vag_oracle as O
    torch.set_num_threads(threads)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in model_cpu.state_dict().items()}
    p["decoder.out.weight"] = p["decoder.embedding.weight"]
    params = {k: v for k, v in p.items() if k != "decoder.out.weight"}
    opt = torch.optim.Adam([{"params": [v for k, v in params.items() if "bias" not in k], "weight_decay": 1e-5},
                            {"params": [v for k, v in params.items() if "bias" in k]}], lr=lr)
    w = torch.ones(p["decoder.embedding.weight"].shape[0])
    w[0] = 0
    times, toks = [], 0
    for i, bt in enumerate(batches):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _, _ = O.multimodal_forward(p, bt.src, bt.src_lengths, bt.tgt, bt.im, True, w, "pairwise", 0.1)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        opt.step()
        dt = time.perf_counter() - t0
        if i > 0:           # first step warms the thread pool / allocator
            times.append(dt)
            toks += int((bt.tgt != 0).sum())
    return toks / sum(times), len(times)


def measure_train(args, dev, world, rank, timed, precision="bf16", french=False):
    """BASELINE configs[1]/[3] shape: EN->DE multimodal training step, batch 32 per GPU, teacher forced, dropout 0
    (parity configuration), gradients all-reduced over ranks.  precision "bf16": every tensor-core / FFMA contraction of the
    forward AND backward pass rounds its operands to bfloat16 and accumulates in FP32; state, soft-max, losses, Adam FP32."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import DistributedPairwiseRankingLoss, GraphedTrainStep
    cfg = synthetic.FR if french else synthetic.DE
    if french:   # BASELINE configs[3]: EN->FR defaults of nmt_multimodal_beam_FR.py:55-67 (dropout 0.2 / 0.4 / 0.4, lr 1e-3)
        torch.manual_seed(1234)
        model = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"],
                                                          cfg["tgt_embedding_size"], cfg["hidden_size"], cfg["shared_embedding_size"], 0.99,
                                                          dropout_emb=0.2, dropout_ctx=0.4, dropout_out=0.4, dropout_im_emb=0.2,
                                                          tied_emb=True).to(dev)
    else:
        model = build_cpu_params().to(dev)
    model.precision = precision
    opt = ClipAdam(model, lr=1e-3 if french else 4e-4)
    w = torch.ones(cfg["tgt_size"], device=dev)
    w[0] = 0
    crit_mt = torch.nn.NLLLoss(weight=w, reduction="none")
    crit_vse = DistributedPairwiseRankingLoss(margin=0.1)
    B = 32
    batches = [synthetic.make_batch(B, cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], seed=100 + 8 * i + rank) for i in range(8)]
    pinned = [(bt.src.pin_memory(), bt.tgt.pin_memory(), bt.im.pin_memory(), bt.src_lengths) for bt in batches]
    tokens = [int((bt.tgt != 0).sum()) for bt in batches]
    state = {"i": 0, "loss": None}

    # the step driver of train.py:36-51 with zero_grad/forward/backward replayed from one CUDA graph per batch shape
    # (single process; under data parallelism the collectives keep the step eager)
    stepper = GraphedTrainStep(model, opt, crit_mt, crit_vse, clip=1.0)

    def step():
        src, tgt, im, lens = pinned[state["i"] % len(pinned)]
        state["i"] += 1
        out = stepper.step(src, lens, tgt, im, 1.0)       # pinned host batch → device copies inside the step
        state["loss"] = out[0]

    for _ in range(len(pinned)):      # one pass over the batch shapes: every shape's graph is captured before the timed region
        step()
    state["i"] = 0
    steps = max(args.steps, 8)
    ms = timed(step, steps)
    final_loss = float(state["loss"])
    tok = sum(tokens[i % len(tokens)] for i in range(steps)) * world
    res = {"metric": "train tgt tokens/sec", "value": tok / (ms / 1e3), "unit": "tokens/s", "ms_per_step": ms / steps, "steps": steps,
           "batch_per_gpu": B, "global_batch": B * world, "dtype": "bf16" if precision == "bf16" else "f32", "loss_after": final_loss,
           "note": ("EN->FR multimodal, teacher forcing 1.0, dropout emb 0.2 / ctx 0.4 / out 0.4 (masks drawn on the device inside the step), "
                    "pairwise ranking loss over the global batch, clip 1.0 + Adam(lr 1e-3, wd 1e-5 on non-bias); " if french else
                    "EN->DE multimodal, teacher forcing 1.0, dropout 0, pairwise ranking loss over the global batch, "
                    "clip 1.0 + Adam(lr 4e-4, wd 1e-5 on non-bias); ") + "host batches (pinned) copied in the timed region; "
                   + ("forward+backward replayed from a CUDA graph per batch shape" if stepper.enabled else "eager launches (collectives in the step)")}
    if rank == 0 and world == 1 and args.cpu_sample > 0 and precision == "bf16" and not french:
        threads = os.cpu_count() or 1
        cpu_model = build_cpu_params()
        v, n = oracle_train_step_time(cpu_model, batches[:4], threads)
        res["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port",
                               "sample": f"{n} optimisation steps, batch {B}, oracle forward + torch autograd + clip + Adam on CPU fp32"}
    return res


def measure_text_only(args, dev, world, rank, timed, src_dev, lens):
    """BASELINE configs[4]: the text-only baseline (NMT_Seq2Seq_Beam_V2, no visual grounding) — beam-12 decoding of the same
    1000-sentence shard and the training step at batch 32 (bf16 mode), same timing rules as the multimodal numbers."""
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.optim import ClipAdam
    from vag_nmt_b200.train import GraphedTrainStep
    cfg = synthetic.DE
    torch.manual_seed(1234)
    model = vag.NMT_Seq2Seq_Beam_V2(cfg["src_size"], cfg["tgt_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
                                    cfg["hidden_size"], tied_emb=True).to(dev).eval()
    K, L = args.beam, args.max_length
    for _ in range(3):
        model.decode_device(src_dev, lens, None, K, L)
    ms = timed(lambda: model.decode_device(src_dev, lens, None, K, L), args.steps)
    out = {"decode": {"metric": METRIC, "value": args.sentences * world * args.steps / (ms / 1e3), "unit": UNIT,
                      "ms_per_step": ms / args.steps, "dtype": "f32"}}
    model.precision = "bf16"
    opt = ClipAdam(model, lr=4e-4)
    w = torch.ones(cfg["tgt_size"], device=dev)
    w[0] = 0
    crit = torch.nn.NLLLoss(weight=w, reduction="none")
    batches = [synthetic.make_batch(32, cfg["src_size"], cfg["tgt_size"], None, seed=100 + 8 * i + rank) for i in range(8)]
    pinned = [(bt.src.pin_memory(), bt.tgt.pin_memory(), bt.src_lengths) for bt in batches]
    tokens = [int((bt.tgt != 0).sum()) for bt in batches]
    stepper = GraphedTrainStep(model, opt, crit, None, clip=1.0)
    state = {"i": 0}

    def step():
        src, tgt, ls = pinned[state["i"] % len(pinned)]
        state["i"] += 1
        state["loss"] = stepper.step(src, ls, tgt, None, 1.0)[0]

    for _ in range(len(pinned)):
        step()
    state["i"] = 0
    steps = max(args.steps, 8)
    ms = timed(step, steps)
    tok = sum(tokens[i % len(tokens)] for i in range(steps)) * world
    out["train"] = {"metric": "train tgt tokens/sec", "value": tok / (ms / 1e3), "unit": "tokens/s", "ms_per_step": ms / steps,
                    "batch_per_gpu": 32, "dtype": "bf16", "loss_after": float(state["loss"])}
    return out


# ------------------------------------------------------------------------------------------ B200 arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import vag_nmt_b200 as vag
    from vag_nmt_b200 import _cabi, ops, synthetic
    lib = _cabi.lib()
    cfg = synthetic.DE
    model = build_cpu_params().to(dev)
    sents, im = synthetic.make_corpus(args.sentences, cfg["src_size"], cfg["im_feats_size"], seed=7 + rank)
    src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
    src_pin, im_pin = src.pin_memory(), im_s.pin_memory()
    src_dev, im_dev = src.to(dev), im_s.to(dev)
    K, L = args.beam, args.max_length

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_device():
        model.decode_device(src_dev, lens, im_dev, K, L)

    result = {}

    def step_e2e():
        s = src_pin.to(dev, non_blocking=True)
        i = im_pin.to(dev, non_blocking=True)
        result["tokens"] = model.beamsearch_decode(s, lens, i, beam_size=K, max_length=L)

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    n0 = lib.vag_launch_count()
    ms = timed(step_device, args.steps)
    launches = lib.vag_launch_count() - n0
    clocks = sampler.finish()
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # bf16 mode of the same workload (north_star's second arithmetic mode; not token-exact, see tests/test_gpu_bf16.py)
    model.precision = "bf16"
    for _ in range(2):
        step_device()
    ms_bf16 = timed(step_device, args.steps)
    model.precision = "fp32"

    total_sent = args.sentences * world
    value = total_sent * args.steps / (ms / 1e3)
    e2e = total_sent * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: the vocabulary projection of one decoder step, [B*K, E] x [E, V] reduced in its
    #      epilogue to per-slice top-2 / soft-max summaries (vocab_top2_pair_kernel — tcgen05 cta_group::2, FP16-split
    #      operands, exactly the launch the beam loop makes), timed alone with CUDA events on the launching stream.
    pk = peaks()
    N = args.sentences * K
    E, V = cfg["tgt_embedding_size"], cfg["tgt_size"]
    x = torch.tanh(torch.randn(N, E, device=dev))
    wgt, bias = model.decoder.out.weight.detach(), model.decoder.out.bias.detach()
    xs, wsplit = ops.tc_split(x), ops.tc_split(wgt)
    for _ in range(3):
        ops.tc_gemm_top2(xs, wsplit, N, E, V, bias)
    reps = 20
    ms_k = timed(lambda: ops.tc_gemm_top2(xs, wsplit, N, E, V, bias), reps) / reps
    flops = 2.0 * N * E * V
    achieved = flops / (ms_k / 1e3) / 1e12
    peak_tf = pk["bf16"] / 3.0
    traffic = None
    tpath = ROOT / "profiles" / "dominant_kernel_traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text()).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "vocab_top2_pair_kernel (vag_tc_gemm_top2_f32) vocabulary projection + top-2/soft-max "
                                             "summaries, rows=%d K=%d N=%d" % (N, E, V),
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "peak_note": f"{pk['src']} bf16 burst {pk['bf16']} TF/s ÷ 3: FP32-exact mode issues 3 FP16 tensor products "
                             "(hi·hi, hi·lo, lo·hi) per algorithmic MAC; algorithmic FLOPs = 2·rows·K·N",
                "ms_per_launch": ms_k, "flops_per_launch": flops}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, per_step=args.sentences),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(src_pin.numel() * 8 + im_pin.numel() * 4),
                    "d2h_bytes_per_step": int(args.sentences * (L * 8 + 4)), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "decode_bf16": {"value": total_sent * args.steps / (ms_bf16 / 1e3), "unit": UNIT, "ms_per_step": ms_bf16 / args.steps,
                            "note": "same workload with model.precision = 'bf16' (bfloat16 operands, one tensor product, FP32 accumulation)"},
            "algorithmic": {"gflop_per_sentence": 2.0 * (1 + (L - 1) * K) * P_STEP_DE / 1e9,
                            "achieved_tflops_whole_job": value * 2.0 * (1 + (L - 1) * K) * P_STEP_DE / 1e12 / world}}

    if rank == 0 and world == 1 and args.cpu_sample > 0:
        threads = os.cpu_count() or 1
        params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        n = args.cpu_sample
        dt, cpu_tokens = oracle_decode_time(params, sents[:n], im[:n], K, L, REF_BATCH, threads)
        # the sample doubles as a parity spot check: the CPU port must produce the tokens the GPU produced
        inv = {c: r for r, c in enumerate(order)}
        agree = sum(int(cpu_tokens[i] == result["tokens"][inv[i]]) for i in range(n))
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first {n} sentences of the shard in reference eval batches of {REF_BATCH}, beam {K}, "
                                          f"max_length {L}, oracle/vag_oracle.py (torch CPU fp32)",
                                "token_exact_sentences": f"{agree}/{n}"}
    del model
    torch.cuda.empty_cache()
    line["text_only"] = measure_text_only(args, dev, world, rank, timed, src_dev, lens)   # BASELINE configs[4]
    torch.cuda.empty_cache()
    line["train"] = measure_train(args, dev, world, rank, timed, "bf16")       # BASELINE configs[1]: training step bf16
    line["train_f32"] = measure_train(args, dev, world, rank, timed, "fp32")
    try:
        line["train_fr"] = measure_train(args, dev, world, rank, timed, "bf16", french=True)   # BASELINE configs[3] shape (32 per GPU)
    except Exception as exc:                     # additional line item only: never take the headline down with it
        line["train_fr"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
