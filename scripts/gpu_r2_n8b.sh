#!/bin/bash
# round 2 (8 GPUs), second pass: the sharded sum with programmatic dependent launch (A/B), the 8-rank tests, the bench line
mkdir -p gpurun_out
for pdl in 1 0; do
  echo "== CB_SUM_PDL=$pdl N=8"
  CB_SUM_PDL=$pdl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$pdl scripts/bench_configs.py --sum-only 2>/dev/null | tail -1
  echo "== CB_SUM_PDL=$pdl CB_COMM_P2P=0 N=8"
  CB_COMM_P2P=0 CB_SUM_PDL=$pdl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$pdl scripts/bench_configs.py --sum-only 2>/dev/null | tail -1
done 2>&1 | tee gpurun_out/sum_pdl_ab_n8.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8 or misses" > gpurun_out/pytest_multi_n8b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_n8b.log; tail -4 gpurun_out/pytest_multi_n8b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/bench_n8b.err | tail -1 > gpurun_out/bench_n8b.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n8b.json')); m=d.get('multi_gpu',{})
print('N=8 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
print(json.dumps(m)[:1600])"
