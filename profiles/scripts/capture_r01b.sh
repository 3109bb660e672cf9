# Round-1 (session 3) evidence capture for det_dense_detect (run under gpurun from the repo root).  Every ncu pass
# follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err || tail -5 gpurun_out/bench_full.err
python profiles/scripts/prof_dense_detect.py > gpurun_out/dd_prof.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dense_detect.csv python profiles/scripts/prof_dense_detect.py 32 256 --once > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:dense_decode_flat|dense_detect_nms" -c 4 -f -o gpurun_out/r01_dense_detect_n256 python profiles/scripts/prof_dense_detect.py 256 --once > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:dense_decode_flat|dense_detect_nms" -c 4 -f -o gpurun_out/r01_dense_detect_n32 python profiles/scripts/prof_dense_detect.py 32 --once > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
