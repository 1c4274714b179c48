#!/bin/bash
# final validation of a round: every -m gpu test, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log; tail -8 gpurun_out/final_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $?"; tail -2 gpurun_out/final_bench.err | cut -c1-250
python - <<PY
import json
d = json.load(open("gpurun_out/final_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "host", d["e2e"].get("host_layer", {}).get("value"), "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"])
PY
