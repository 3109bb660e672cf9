#!/usr/bin/env python
"""Aggregate `ncu --page source --csv` per-line counters of one source file into phases delimited by the
`// ---- X` comments of that file: share of executed warp-instructions and of stall samples per phase.
usage: phase_inst.py source_page.csv path/to/file.cuh"""
import csv
import re
import sys

page, src = sys.argv[1], sys.argv[2]
fname = src.split("/")[-1]
lines = open(src).read().split("\n")
marks = [(1, "prologue")]
for i, l in enumerate(lines, 1):
    m = re.match(r"\s*// ---- (.*)", l)
    if m:
        marks.append((i, m.group(1)[:60]))
agg, other, cur, hdr = {}, {}, None, None
for r in csv.reader(open(page)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        i, s = int(r[ii]), int(r[si])
    except ValueError:
        continue
    if cur == fname:
        name = [n for (b, n) in marks if b <= int(r[0])][-1]
        a = agg.setdefault(name, [0, 0])
    else:
        a = other.setdefault(cur, [0, 0])
    a[0] += i
    a[1] += s
ti = sum(v[0] for v in agg.values()) + sum(v[0] for v in other.values())
ts = sum(v[1] for v in agg.values()) + sum(v[1] for v in other.values())
print(f"total warp-instructions {ti}, stall samples {ts}")
for b, n in marks:
    if n in agg:
        print(f"{100 * agg[n][0] / ti:5.1f}% inst {100 * agg[n][1] / ts:5.1f}% samples  {fname}:{b}  {n}")
for f, v in sorted(other.items(), key=lambda kv: -kv[1][0]):
    if v[0]:
        print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / ts:5.1f}% samples  (inlined from {f})")
