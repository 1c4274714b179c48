#!/bin/bash
# quick perf check: parity of the ICP tests only, then phase timing single + batch 8 (dense stretch)
mkdir -p gpurun_out
T=${1:-r2b}
timeout 900 python -m pytest tests/test_gpu_icp.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
LS_PHASE_TIMING=1 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 2 30 > gpurun_out/${T}_phase_single.log 2>&1
LS_BATCH=8 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 2 30 > gpurun_out/${T}_batch8.log 2>&1
LS_BATCH=16 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 2 30 > gpurun_out/${T}_batch16.log 2>&1
LS_BATCH=8 LS_PHASE_TIMING=1 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 1 30 > gpurun_out/${T}_phase_batch8.log 2>&1
grep -v "^\[ls\]" gpurun_out/${T}_phase_single.log | tail -2
grep -v "^\[ls\]" gpurun_out/${T}_batch8.log | tail -2
grep -v "^\[ls\]" gpurun_out/${T}_batch16.log | tail -2
