#!/bin/bash
# Round-2 ncu evidence (run under gpurun from the repo root): usage capture_r02.sh <tag> <kernel-regex> <driver.py> [args...]
# A plain run of the same command must exit 0 first; numbers printed under ncu are never bench values.
TAG=$1; REGEX=$2; shift 2
OUT=gpurun_out/ncu; mkdir -p $OUT
python "$@" > $OUT/$TAG.plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/$TAG.plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv python "$@" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -c 8 -f -o $OUT/$TAG python "$@" > $OUT/$TAG.ncu.log 2>&1
ls -la $OUT/$TAG.ncu-rep
