#!/bin/bash
# A/B of library builds: LS_B200_LIB=<path> selects the build; batched config-2 registrations at the dense stretch
mkdir -p gpurun_out
T=${1:-ab}
for v in "" _vA _vB _vC _vD; do
  lib=laser_slam_b200/_build$v/libls_b200.so
  [ -f $lib ] || continue
  for B in ${BATCHES:-8 16}; do
    LS_B200_LIB=$PWD/$lib LS_BATCH=$B LS_PROF_SCAN=14 LS_PROF_Y=-20 timeout 200 python tools/prof_one.py 2 30 > gpurun_out/${T}${v}_b$B.log 2>&1
    echo "build '$v' batch $B: $(grep -v '^\[ls\]' gpurun_out/${T}${v}_b$B.log | grep 'batch' | tail -1 | cut -c1-75)"
  done
done
