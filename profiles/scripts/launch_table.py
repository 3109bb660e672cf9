"""stdin: `ncu --metrics gpu__time_duration.sum --csv` output -> per-kernel count, last duration and total of the LAST `calls` invocations"""
import csv, sys
rows = [r for r in csv.reader(sys.stdin) if len(r) > 10 and r[0].isdigit()]
agg = {}
for r in rows:
    name = r[4].split("(")[0][-46:]
    agg.setdefault(name, []).append(float(r[-1]) / 1000)
for k, v in agg.items():
    print(f"{k:48s} n={len(v):4d} last={v[-1]:8.1f} us  mean={sum(v)/len(v):8.1f} us  total={sum(v):9.1f} us")
