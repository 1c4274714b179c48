#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2i_pytest.log; tail -6 gpurun_out/r2i_pytest.log | cut -c1-220
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2i_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2i_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "host", d["e2e"].get("host_layer", {}).get("value"), "cpu", d["cpu_baseline"]["value"])
PY
