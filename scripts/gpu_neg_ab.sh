#!/bin/bash
mkdir -p gpurun_out
for cfg in "CB_NEG_XOR=1" "CB_NEG_XOR=0" "CB_NEG_XOR=1" "CB_NEG_XOR=0"; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('burst', round(d['value'],1), 'GB/s', d['ms_per_step'], 'ms | sustained', round(d['sustained']['value'],1), 'GB/s @', d['sustained']['clocks']['sm_mhz'], 'MHz', d['sustained']['clocks'].get('power_w_max'), 'W')"
done 2>&1 | tee gpurun_out/neg_ab.log
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3 | tee -a gpurun_out/neg_ab.log
