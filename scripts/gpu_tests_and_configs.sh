#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python scripts/bench_configs.py > gpurun_out/configs_n1.json 2> gpurun_out/configs_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/configs_n1.json'))
for k,v in d.items():
    if k=='unary_grad_each': print(k,{kk:vv['GB/s'] for kk,vv in v.items()})
    elif 'GB/s' in v: print(k, v['GB/s'], v.get('frac_of_measured_peak'))
    else: print(k,v)
PY
