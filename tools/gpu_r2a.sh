#!/bin/bash
# round-2 GPU check: parity suite, phase timing (single + batch 8, dense stretch), short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
LS_PHASE_TIMING=1 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 2 30 > gpurun_out/r2a_phase_single.log 2>&1
LS_BATCH=8 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 2 30 > gpurun_out/r2a_batch8.log 2>&1
LS_BATCH=8 LS_PHASE_TIMING=1 LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 300 python tools/prof_one.py 1 30 > gpurun_out/r2a_phase_batch8.log 2>&1
grep -v "^\[ls\]" gpurun_out/r2a_batch8.log | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
cat gpurun_out/r2a_bench.json | cut -c1-600
