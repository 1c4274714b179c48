#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_posegraph.py tests/test_host_layer.py -m gpu -x -q > gpurun_out/pg_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/pg_pytest.log; tail -12 gpurun_out/pg_pytest.log | cut -c1-200
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/pg_cfg4.json 2> gpurun_out/pg_cfg4.err; echo "cfg4 rc $?"; tail -3 gpurun_out/pg_cfg4.err
python - <<PY
import json
d = json.load(open("gpurun_out/pg_cfg4.json"))
print("ms/estimate", d["ms_per_step"], "e2e solves/s", d["e2e"]["value"], "conv", d["config"]["to_convergence"], "marg", d["config"]["marginals"], "launches", d["gpu_launches"])
PY
