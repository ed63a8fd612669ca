"""Executed instructions / stall samples per CUDA source line of one kernel instance in an .ncu-rep captured with
--import-source on:  python tools/ncu_source_lines.py report.ncu-rep <kernel regex> [launch-skip] [top]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""
data = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue     # SASS rows (no line number) are already summed into their source line's row
    try:
        n, s = int(r[hdr.index("Instructions Executed")].replace(",", "")), int(r[hdr.index("# Samples")].replace(",", ""))
    except ValueError:
        continue
    data.append((n, s, fname, r[0], r[1].strip()[:120]))
tot, tots = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total warp instructions {tot}, samples {tots}")
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{d[0] / max(tot, 1) * 100:5.1f}% instr {d[1] / max(tots, 1) * 100:5.1f}% samples  {d[2]}:{d[3]:>5s}  {d[4]}")
