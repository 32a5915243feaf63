#!/bin/bash
# usage (GPU box, repo root): tools/multi_bench.sh <N> [workloads...]   -- torchrun bench lines into gpurun_out/b_<workload>_<N>gpu.json
N=$1; shift
mkdir -p gpurun_out
port=29520
for wl in "$@"; do
  port=$((port+1))
  steps=5; [ "$wl" = "config4" ] && steps=2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps $steps --warmup 3 --workload $wl > gpurun_out/b_${wl}_${N}gpu.json 2> gpurun_out/b_${wl}_${N}gpu.err
  python - <<PY
import json
for ln in open("gpurun_out/b_${wl}_${N}gpu.json").read().splitlines():
    if ln.startswith("{"):
        d = json.loads(ln)
        print("$wl", d["n_gpus"], "gpus", round(d["ms_per_step"], 2), "ms", round(d["value"] / 1e6, 1), "Mevals/s  e2e", round(d["e2e"]["ms_per_step"], 2), "ms", round(d["e2e"]["value"] / 1e6, 1), "gather", d.get("gather_check"), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  grep -v "^\*\*\*\|OMP_NUM" gpurun_out/b_${wl}_${N}gpu.err | tail -3
done
