#!/bin/bash
# Profiling pass for BASELINE config 4 (brandubh 4096 games x 200 sims, 64ch x 4 net) on ONE B200 (run under gpurun):
# launch list of two bench steps + one `ncu --set full` capture of the fused tree kernel; outputs in gpurun_out/brandubh/.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/brandubh
B="python bench.py --game brandubh --games 4096 --sims 200 --net brandubh_train --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --sustain-seconds 0 --no-select-events --preroll 24"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 900 --csv --log-file gpurun_out/brandubh/launches.csv $B > gpurun_out/brandubh/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand_select -s 300 -c 2 -f -o gpurun_out/brandubh/prof_expand_select $B --no-round-graph > gpurun_out/brandubh/prof_expand_select.log 2>&1
ls -la gpurun_out/brandubh
