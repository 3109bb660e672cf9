#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by source line:
warp-instructions executed and stall samples, top N lines."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] not in ("", "-") and r[0].isdigit():
        try:
            agg.append((cur_file, int(r[0]), r[1].strip()[:110], int(r[i_inst]), int(r[i_samp])))
        except ValueError:
            pass
tot_i = sum(a[3] for a in agg) or 1
tot_s = sum(a[4] for a in agg) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("-- by instructions")
for f, l, s, i, sm in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{100 * i / tot_i:5.1f}% inst {100 * sm / tot_s:5.1f}% samp  {f}:{l}  {s}")
if len(sys.argv) > 3:
    print("-- by stall samples")
    for f, l, s, i, sm in sorted(agg, key=lambda a: -a[4])[:top]:
        print(f"{100 * sm / tot_s:5.1f}% samp {100 * i / tot_i:5.1f}% inst  {f}:{l}  {s}")
