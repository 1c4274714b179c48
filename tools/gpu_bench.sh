#!/bin/bash
mkdir -p gpurun_out
T=${1:-bench}
for B in ${TRACKS:-8 16 32}; do
  timeout 900 python bench.py --steps 20 --warmup 3 --tracks $B > gpurun_out/${T}_t$B.json 2> gpurun_out/${T}_t$B.err; echo "tracks $B rc $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_t$B.json"))
    print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4), "single", round(d["config"]["single_stream_ms_per_registration"], 3), "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("  parse failed", e)
PY
done
