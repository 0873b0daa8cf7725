#!/bin/bash
# new dtypes (bf16, i8, i16, u16, u64, bool) + native f16 arithmetic: parity tests, then f16/bf16 chain throughput A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dtypes.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_dtypes.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dtypes.log
tail -15 gpurun_out/pytest_dtypes.log
for cfg in "CB_H_NATIVE=1" "CB_H_NATIVE=0"; do
  echo "== $cfg"; env $cfg python scripts/bench_configs.py 2>gpurun_out/bench_configs.err | tee gpurun_out/configs_${cfg}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(v['GB/s'],v['frac_of_measured_peak']) for k,v in d.items() if 'f16' in k or k in ('chain8_fwd_f32','cheap8_fwd_f32')})"
done 2>&1 | tee gpurun_out/dtypes_ab.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 600 gpurun_out/bench.json
