#!/bin/bash
# one ncu capture (full set + source) of the batch-8 icp_kernel launch at the dense stretch
mkdir -p gpurun_out
T=${1:-r2_batch8}
LS_BATCH=${LS_BATCH:-8} LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 900 ncu --set full --import-source on --clock-control none -k regex:icp_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/$T python tools/prof_one.py 1 30 > gpurun_out/${T}_ncu.log 2>&1
tail -5 gpurun_out/${T}_ncu.log
ls -la gpurun_out/$T.ncu-rep
