#!/bin/bash
# usage (on the GPU box, repo root): tools/gpu_check.sh <label> [pytest -k expression]
# runs the -m gpu suite and a short bench, leaves gpurun_out/t_<label>.txt and gpurun_out/b_<label>.json
label=$1; kexpr=$2
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --timeout-method thread -p no:cacheprovider -k "$kexpr" 2>&1 | tail -40 > gpurun_out/t_$label.txt
else
  timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --timeout-method thread -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_$label.txt
fi
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/b_$label.json 2> gpurun_out/b_$label.err
tail -6 gpurun_out/t_$label.txt
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b_$label.json"))
    print("$label", "ms/step", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "update", {k: round(v, 1) for k, v in d.get("update_latency", {}).items() if k.startswith("p50")})
except Exception as e:
    print("bench failed:", e)
PY
tail -3 gpurun_out/b_$label.err
