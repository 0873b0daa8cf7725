#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -s -k "transcendental or chain or config1 or unary_grad or exact or reference_kats" > gpurun_out/pytest_math.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_math.log
grep -E "ulp|passed|failed|rc=|Error" gpurun_out/pytest_math.log | tail -20
for cfg in "CB_WAVES=16" "CB_WAVES=1" "CB_WAVES=16 CB_UNROLL=2" "CB_WAVES=16 CB_THREADS=128 CB_MIN_BLOCKS=8 CB_BLOCKS_PER_SM=16" "CB_WAVES=16 CB_THREADS=512 CB_MIN_BLOCKS=2" "CB_WAVES=64" ; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --warmup 5 --extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'clk',d['clocks']['sm_mhz'],d['clocks']['reasons'],'pw',d['clocks']['power_w_max'])
print({k:round(v['GB/s'],0) for k,v in d['extra'].items()})"
done 2>&1 | tee gpurun_out/perf_ab3.log
