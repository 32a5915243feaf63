#!/bin/bash
# usage: tools/quick_bench.sh <label> [precisions...]   -- appends concise bench lines to gpurun_out/quick.txt
label=$1; shift
mkdir -p gpurun_out
for p in "$@"; do
  python bench.py --precision $p --no-cpu-baseline --no-update-latency --no-torch-baseline --steps 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', d['config']['precision'], round(d['ms_per_step'],2), 'ms', round(d['value']/1e6,1), 'Mevals/s')" >> gpurun_out/quick.txt
done
