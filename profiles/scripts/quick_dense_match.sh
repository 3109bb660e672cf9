#!/bin/bash
# per-kernel times of the dense-label grid matcher (ncu launch list of prof_train_grid.py)
python profiles/scripts/prof_train_grid.py 1024 3 > /dev/null 2>&1 || echo "plain run failed"
ncu --metrics gpu__time_duration.sum --clock-control none --csv python profiles/scripts/prof_train_grid.py 1024 3 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
last={}
for r in rows:
    name=r[4].split('(')[0][-40:]; last.setdefault(name,[]).append(float(r[-1]))
for k,v in last.items():
    if 'match' in k or 'subsample' in k: print(k, [round(x/1000,1) for x in v[-3:]])"
