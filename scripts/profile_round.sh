#!/bin/bash
# Round profiling pass on ONE B200 (run under gpurun): the launch list of a bench step and one `ncu --set full`
# capture per hot kernel, written to gpurun_out/final/; scripts/summarize_profiles.py <tag> turns them into the
# tracked summaries under profiles/.  Numbers printed by a run under ncu are never bench values.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/final
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --sustain-seconds 0 --no-select-events --preroll 24"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/final/launches.csv $B > gpurun_out/final/launches.log 2>&1
for k in select:k_select expand_select:k_expand_select; do
  tag=${k%%:*}; pat=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 300 -c 2 -f -o gpurun_out/final/prof_$tag $B --no-round-graph > gpurun_out/final/prof_$tag.log 2>&1
done
# the evaluator's kernels on the bench's batch shape (6960 non-terminal leaves of 8192 Connect4 games), default precision
ncu --set full --clock-control none --import-source on -k regex:k_trunk_tc -s 2 -c 2 -f -o gpurun_out/final/prof_trunk python scripts/nn_once.py connect4 bf16x2 6960 3 > gpurun_out/final/prof_trunk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_head_tc -s 2 -c 2 -f -o gpurun_out/final/prof_head python scripts/nn_once.py connect4 bf16x2 6960 3 > gpurun_out/final/prof_head.log 2>&1
ls -la gpurun_out/final
