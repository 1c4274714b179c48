#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_layer.py tests/test_normals.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2h_pytest.log; tail -12 gpurun_out/r2h_pytest.log | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2h_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "host", d["e2e"].get("host_layer"), "parity", d["config"]["parity_check"])
PY
timeout 900 python bench.py --config 3 --steps 60 > gpurun_out/r2h_cfg3.json 2> gpurun_out/r2h_cfg3.err; echo "cfg3 rc $?"; tail -3 gpurun_out/r2h_cfg3.err; cut -c1-1800 gpurun_out/r2h_cfg3.json
