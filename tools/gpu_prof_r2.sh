#!/bin/bash
# round-2 profile artefacts: launch list of the bench command, full ncu captures of the dominant kernel (config 2 with 32
# registrations per launch, config 5 with 8), written to gpurun_out/ (summaries are copied to profiles/ afterwards)
mkdir -p gpurun_out
LS_BENCH_NO_HOST_ARM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_launches_bench.log 2>&1; echo "launch list rc $?"
LS_BATCH=32 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 900 ncu --set full --import-source on --clock-control none -k regex:icp_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_cfg2_batch32 python tools/prof_one.py 1 30 > gpurun_out/r2_cfg2_batch32_ncu.log 2>&1; echo "cfg2 capture rc $?"
LS_BATCH=8 LS_PROF_SENSOR=1 LS_PROF_K=8 timeout 900 ncu --set full --import-source on --clock-control none -k regex:icp_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_cfg5_batch8 python tools/prof_one.py 1 50 > gpurun_out/r2_cfg5_batch8_ncu.log 2>&1; echo "cfg5 capture rc $?"
LS_PROF_SENSOR=1 LS_PROF_K=8 timeout 900 ncu --set full --clock-control none -k regex:icp_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_cfg5_single python tools/prof_one.py 2 50 > gpurun_out/r2_cfg5_single_ncu.log 2>&1; echo "cfg5 single capture rc $?"
LS_BATCH=8 LS_PROF_SENSOR=1 LS_PROF_K=8 timeout 300 python tools/prof_one.py 2 50 2>&1 | grep -v "^\[ls\]" | tail -4
ls -la gpurun_out/*.ncu-rep | tail -5
