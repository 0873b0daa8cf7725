#!/bin/bash
mkdir -p gpurun_out
for cfg in "CB_WAVES=1" "CB_WAVES=2" "CB_WAVES=4" "CB_WAVES=8" "CB_WAVES=16" "CB_WAVES=64" "CB_WAVES=4096" "CB_WAVES=16 CB_UNROLL=2" "CB_WAVES=16 CB_UNROLL=8 CB_MIN_BLOCKS=3" "CB_WAVES=8 CB_THREADS=512 CB_MIN_BLOCKS=2" "CB_WAVES=16 CB_THREADS=128 CB_MIN_BLOCKS=8 CB_BLOCKS_PER_SM=16"; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --warmup 5 --extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'clk',d['clocks']['sm_mhz'],d['clocks']['reasons'],'pw',d['clocks']['power_w_max'])
print({k:round(v['GB/s'],0) for k,v in d['extra'].items()})"
done 2>&1 | tee gpurun_out/perf_ab2.log
