mkdir -p gpurun_out/ncu
ncu --set full --clock-control none --import-source on -k "regex:rpn_select" -s 2 -c 1 -f -o gpurun_out/ncu/r02_rpn_select python profiles/scripts/prof_proposals.py 64 2000 1000 > /dev/null 2>&1
ls -la gpurun_out/ncu/r02_rpn_select.ncu-rep
