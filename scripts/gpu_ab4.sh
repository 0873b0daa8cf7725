#!/bin/bash
mkdir -p gpurun_out
for cfg in "CB_UNROLL=4 CB_MIN_BLOCKS=4" "CB_UNROLL=2 CB_MIN_BLOCKS=6" "CB_UNROLL=4 CB_MIN_BLOCKS=5" "CB_UNROLL=4 CB_MIN_BLOCKS=6" "CB_UNROLL=2 CB_MIN_BLOCKS=8" "CB_UNROLL=3 CB_MIN_BLOCKS=6" "CB_UNROLL=2 CB_MIN_BLOCKS=6 CB_WAVES=32" "CB_UNROLL=1 CB_MIN_BLOCKS=8"; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --warmup 5 --extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'sustained',round(d['sustained']['value'],1),d['sustained']['clocks']['sm_mhz'])
print({k:round(v['GB/s'],0) for k,v in d['extra'].items() if k in ('cheap8_f32','chain8_f32','chain8_f16','unary_grad_cos_f32')})"
done 2>&1 | tee gpurun_out/perf_ab4.log
