#!/bin/bash
# A/B of the whole step with the trunk's two tile plans (same box, alternating)
X="--no-cpu-baseline --no-e2e --no-alt --steps 40 --sustain-seconds 3"
for rep in 1 2; do
for p in 0 1; do
  AZB_NNG_PERSIST=$p timeout 200 python bench.py $X 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('persist=$p c4', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', 'nn p50', d['roofline_nn']['launch_us_p10_p50_p90'][1], 'sustained', round(d['sustained']['value']/1e6,2), d['sustained']['clocks']['sm_mhz'])"
done
done
for p in 0 1; do
  AZB_NNG_PERSIST=$p timeout 300 python bench.py $X --steps 10 --game brandubh --games 4096 --sims 200 --net brandubh_train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('persist=$p brandubh', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', 'nn p50', d['roofline_nn']['launch_us_p10_p50_p90'][1], 'sustained', round(d['sustained']['value']/1e6,2), d['sustained']['clocks']['sm_mhz'])"
done
