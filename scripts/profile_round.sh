#!/bin/bash
# Round profiling pass on ONE B200 (run under gpurun): the launch list of a bench step and one `ncu --set full`
# capture per hot kernel, written to gpurun_out/final/; scripts/summarize_profiles.py <tag> turns them into the
# tracked summaries under profiles/.  Numbers printed by a run under ncu are never bench values.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/final
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --sustain-seconds 0 --no-select-events --preroll 24"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/final/launches.csv $B > gpurun_out/final/launches.log 2>&1
for k in select:k_select expand_select:k_expand_select trunk:k_trunk_tc head:k_head_tc play:k_play_moves; do
  tag=${k%%:*}; pat=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 300 -c 2 -f -o gpurun_out/final/prof_$tag $B --no-round-graph > gpurun_out/final/prof_$tag.log 2>&1
done
ls -la gpurun_out/final
