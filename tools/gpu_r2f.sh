#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2f_pytest.log; tail -4 gpurun_out/r2f_pytest.log
TRACKS="8 32" bash tools/gpu_bench.sh r2f
