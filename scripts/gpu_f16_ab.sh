#!/bin/bash
mkdir -p gpurun_out
for cfg in "CB_PAIR=1" "CB_PAIR=0" "CB_PAIR=1 CB_UNROLL=2" "CB_PAIR=1 CB_MIN_BLOCKS=3" "CB_PAIR=1 CB_UNROLL=2 CB_MIN_BLOCKS=6"; do
  echo "== $cfg"; env $cfg python scripts/bench_configs.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(v['GB/s'],v['frac_of_measured_peak']) for k,v in d.items() if 'f16' in k or k in ('chain8_fwd_f32','cheap8_fwd_f32')})"
done 2>&1 | tee gpurun_out/f16_ab.log
