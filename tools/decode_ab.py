"""A/B timing of the beam-12 decode loop under different environment switches, one fresh process per configuration.

    python tools/decode_ab.py [--n 1000] [--L 80] [--prec fp32] "base" "VAG_SELECT_NT=256" "VAG_PDL_MASK=29,VAG_SELECT_NT=256"

Every configuration (comma-separated KEY=VALUE pairs; the literal `base` = no switch) decodes the same synthetic corpus
(seed 7, like bench.py) `reps` times after two warm-up decodes and prints the median / minimum of the device-timed loop plus a
checksum of the tokens, so that a switch that changes the translation is visible at once.
"""
import argparse
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def worker(n, L, prec, reps):
    import torch
    sys.path.insert(0, str(ROOT))
    os.environ.setdefault("VAG_DECODE_GRAPH", "0")
    import bench
    from vag_nmt_b200 import ops, synthetic, _cabi
    dev = torch.device("cuda", 0)
    model = bench.build_cpu_params().to(dev)
    model.precision = prec
    cfg = synthetic.DE
    sents, im = synthetic.make_corpus(n, cfg["src_size"], cfg["im_feats_size"], seed=7)
    src, lens, im_s, _ = synthetic.pad_and_sort(sents, im)
    src_d, im_d = src.to(dev), im_s.to(dev)
    times = []
    with _cabi.precision_scope(prec):
        w, ctx, mask, keys, h0, _, _ = model._prepare(src_d, lens, im_d)
        for i in range(2 + reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            hyp, hyp_len = ops.beam_decode(w, h0, keys, ctx, mask, 12, L, early_stop=os.environ.get("AB_EARLY_STOP") == "1")
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1))
    times.sort()
    chk = int((hyp.to(torch.int64) * torch.arange(1, hyp.numel() + 1, device=dev).view_as(hyp) % 1000003).sum().item())
    print(f"RESULT median {times[len(times) // 2]:.3f} ms  min {times[0]:.3f} ms  ({n / times[len(times) // 2]:.2f} k sent/s)  tokens#{chk}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--L", type=int, default=80)
    ap.add_argument("--prec", default="fp32")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("configs", nargs="*")
    a = ap.parse_args()
    if a.worker:
        worker(a.n, a.L, a.prec, a.reps)
        return
    for cfg in a.configs or ["base"]:
        env = dict(os.environ)
        if cfg != "base":
            for kv in cfg.split(","):
                k, v = kv.split("=", 1)
                env[k] = v
        r = subprocess.run([sys.executable, __file__, "--worker", "--n", str(a.n), "--L", str(a.L), "--prec", a.prec, "--reps", str(a.reps)],
                           env=env, capture_output=True, text=True, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
        print(f"{cfg:48s} {a.prec} n={a.n}: {line[0][7:] if line else 'FAILED: ' + (r.stderr.strip().splitlines() or ['?'])[-1]}", flush=True)


if __name__ == "__main__":
    main()
