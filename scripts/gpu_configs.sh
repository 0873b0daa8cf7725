#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_configs.py > gpurun_out/configs_n1.json 2> gpurun_out/configs_n1.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/configs_n1.json'))
for k,v in d.items():
    if k=='unary_grad_each': print(k,{kk:vv['GB/s'] for kk,vv in v.items()})
    else: print(k,v)
PY
tail -3 gpurun_out/configs_n1.err
