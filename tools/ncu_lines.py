#!/usr/bin/env python
"""Joins an `ncu --page source --print-source sass --csv` export with `nvdisasm -g` line info of the same cubin
and prints executed warp instructions / stall samples per source line.
  python tools/ncu_lines.py <ncu_sass.csv> <nvdisasm_function.dis> [source.cu] [top]
The ncu export may hold several launches of the kernel back to back: the LAST one is used."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
# split into launches: a header row starts with "Address"
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[starts[-1]]
body = [r for r in rows[starts[-1] + 1:] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
dis = open(sys.argv[2]).read().splitlines()
lines = []   # per instruction: (file, line)
cur = ("?", 0)
for l in dis:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        lines.append(cur)
assert len(lines) == len(body), (len(lines), len(body))
src = open(sys.argv[3]).read().splitlines() if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
inst = defaultdict(int)
samp = defaultdict(int)
for (f, ln), r in zip(lines, body):
    inst[(f, ln)] += int(r[col["Instructions Executed"]])
    samp[(f, ln)] += int(r[col["# Samples"]])
ti = sum(inst.values())
ts = sum(samp.values())
print("total warp instructions %d, samples %d" % (ti, ts))
for k in sorted(inst, key=lambda k: -samp[k])[:top]:
    text = src[k[1] - 1].strip()[:100] if src and k[0].endswith(".cu") and k[1] <= len(src) else ""
    print("%-28s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0][:28], k[1], 100.0 * inst[k] / ti, 100.0 * samp[k] / ts, text))
