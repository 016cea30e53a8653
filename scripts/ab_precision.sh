X="--no-cpu-baseline --no-e2e --no-alt --steps 40 --sustain-seconds 3"
for rep in 1 2; do
for p in bf16x2 fp16x2; do
  timeout 200 python bench.py $X --precision $p 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$p c4', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', 'nn p50', d['roofline_nn']['launch_us_p10_p50_p90'][1], 'sustained', round(d['sustained']['value']/1e6,2), d['sustained']['clocks']['sm_mhz'], 'err', d['nn_error']['max_abs_error_vs_f32_module'])"
done
done
