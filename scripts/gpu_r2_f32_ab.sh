#!/bin/bash
# f32 chain: launch-shape A/B under the power cap (burst and >= 1 s sustained), one bench.py --no-configs run per setting
mkdir -p gpurun_out
for cfg in "CB_MIN_BLOCKS=5" "CB_MIN_BLOCKS=4" "CB_MIN_BLOCKS=3" "CB_MIN_BLOCKS=6" "CB_MIN_BLOCKS=4 CB_UNROLL=2" "CB_MIN_BLOCKS=8 CB_UNROLL=2" "CB_THREADS=128 CB_MIN_BLOCKS=8" "CB_THREADS=512 CB_MIN_BLOCKS=2" "CB_WAVES=4" "CB_WAVES=64"; do
  env $cfg python bench.py --no-configs --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$cfg', '| burst', round(d['value'],1), '| sustained', round(r['achieved_sustained'],1), round(r['frac_sustained'],4), '@', r['sustained_clocks']['sm_mhz'], 'MHz', r['sustained_clocks'].get('power_w_max'))"
done 2>&1 | tee gpurun_out/f32_shape_ab.log
