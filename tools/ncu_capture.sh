#!/bin/bash
# Round-2 ncu artefacts (run on the GPU box from the repo root, AFTER the same commands exited 0 without ncu):
#   launch lists (gpu__time_duration.sum) of the default bench and of the config-5 bench, and one `--set full` capture of
#   each dominant kernel: k_separate_tc3 (both launches of a config-2 pass), k_cif_wide (config 5), k_update (1000 agents).
mkdir -p gpurun_out
B="--no-cpu-baseline --no-update-latency --no-torch-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_f16x3.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_config5.csv python bench.py --workload config5 --steps 2 --warmup 3 $B > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_separate_tc3 --launch-skip 6 --launch-count 2 -o gpurun_out/r02_tc3 -f python bench.py --steps 1 --warmup 3 $B > gpurun_out/ncu_tc3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cif_wide --launch-skip 3 --launch-count 1 -o gpurun_out/r02_tcw -f python bench.py --workload config5 --steps 1 --warmup 3 $B > gpurun_out/ncu_tcw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update --launch-skip 4 --launch-count 2 -o gpurun_out/r02_update -f python tools/bench_update.py 1000 > gpurun_out/ncu_upd.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_*.csv
