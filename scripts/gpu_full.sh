#!/bin/bash
# Full GPU pass (run through gpurun): smoke, every GPU test, the bench line (+extra rows), reference arm.
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -2 gpurun_out/smoke.log; tail -6 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('sustained',d['sustained'])
print('e2e',d['e2e']); print('cpu',d['cpu_baseline']['value'])
print({k:round(v['GB/s']) for k,v in d['extra'].items()})
print(open('gpurun_out/bench_ref.json').read()[:400])
PY
