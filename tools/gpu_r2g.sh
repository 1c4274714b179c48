#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_posegraph.py tests/test_host_layer.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2g_pytest.log; tail -15 gpurun_out/r2g_pytest.log | cut -c1-200
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/r2g_cfg4.json 2> gpurun_out/r2g_cfg4.err; echo "cfg4 rc $?"; tail -3 gpurun_out/r2g_cfg4.err; cut -c1-1500 gpurun_out/r2g_cfg4.json
timeout 900 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2g_cfg5.json 2> gpurun_out/r2g_cfg5.err; echo "cfg5 rc $?"; tail -3 gpurun_out/r2g_cfg5.err; cut -c1-700 gpurun_out/r2g_cfg5.json
