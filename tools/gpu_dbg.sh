#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py -m gpu -x -q > gpurun_out/dbg_pytest_default.log 2>&1; tail -40 gpurun_out/dbg_pytest_default.log | cut -c1-180
