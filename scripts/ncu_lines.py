#!/usr/bin/env python
"""Per-source-line totals (samples, executed warp instructions) from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py <report.ncu-rep> [top N]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
samples, execd, text = defaultdict(int), defaultdict(int), {}
cur = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        cur = int(r[0])
        # the source text may contain commas: everything up to the last 7 columns
        text[cur] = ",".join(r[1:-7]).strip()
        continue
    if r[0] == "" and cur is not None and len(r) >= 10 and r[2].startswith("0x"):
        try:
            samples[cur] += int(r[6])
            execd[cur] += int(r[7])
        except ValueError:
            pass
ts, te = sum(samples.values()), sum(execd.values())
print(f"total samples {ts}, warp instructions executed {te}")
for ln in sorted(samples, key=lambda k: -samples[k])[:top]:
    print(f"{ln:5d} samples {samples[ln]:6d} {100 * samples[ln] / max(ts, 1):5.1f}%  exec {execd[ln]:10d} {100 * execd[ln] / max(te, 1):5.1f}%  {text.get(ln, '')[:110]}")
