#!/bin/bash
# compute-sanitizer over one small invocation of every kernel family (profiles/scripts/sanitize_driver.py):
# memcheck (out-of-bounds / misaligned), racecheck (shared-memory hazards), synccheck (divergent barriers), initcheck
# (reads of uninitialised global memory).  Run under gpurun from the repo root; the summaries land in gpurun_out/sanitize/.
# Every tool first needs a plain run of the same command that exits 0.
OUT=gpurun_out/sanitize
mkdir -p $OUT
DRV="python profiles/scripts/sanitize_driver.py"
$DRV > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  for sec in yolo nms dense train misc; do
    log=$OUT/${tool}_${sec}.log
    extra=""
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
    timeout 900 compute-sanitizer --tool $tool $extra --print-limit 30 $DRV $sec > $log 2>&1
    rc=$?
    echo "== $tool $sec rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|section ok' $log | tr '\n' ' ')"
  done
done
grep -h -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/*.log | sort | uniq -c > $OUT/summary.txt
cat $OUT/summary.txt
