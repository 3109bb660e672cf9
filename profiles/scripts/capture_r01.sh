# Round-1 evidence capture (run under gpurun from the repo root).  Every ncu pass follows a plain run of the same
# command that exited 0; numbers printed under ncu are never bench values.
set -x
python bench.py --no-extras --steps 2000 > gpurun_out/b_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --no-extras --steps 20 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
python profiles/scripts/prof_yolo.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:yolo_fast -s 3 -c 1 -f -o gpurun_out/r01_yolo_fast python profiles/scripts/prof_yolo.py > /dev/null 2>&1
python profiles/scripts/prof_dense.py 256 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_decode -s 6 -c 1 -f -o gpurun_out/r01_dense_flat_n256 python profiles/scripts/prof_dense.py 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_decode -s 6 -c 1 -f -o gpurun_out/r01_dense_flat_n32 python profiles/scripts/prof_dense.py 32 > /dev/null 2>&1
python profiles/scripts/prof_train.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:rpn_loss|match_pass1|match_pass2|subsample" -s 4 -c 4 -f -o gpurun_out/r01_train python profiles/scripts/prof_train.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python profiles/scripts/prof_train.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
