#!/bin/bash
# round 2, call 3 (2 GPUs): full GPU suite incl. the multi-GPU tests at world 2, sharded sum at N=1,2, e2e pipeline A/B at N=2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log; tail -12 gpurun_out/pytest_gpu3.log
python scripts/bench_configs.py --sum-only > gpurun_out/sum_n1.json 2>gpurun_out/sum_n1.err; cat gpurun_out/sum_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_configs.py --sum-only > gpurun_out/sum_n2.json 2>gpurun_out/sum_n2.err; cat gpurun_out/sum_n2.json
for cfg in "CB_PIPE_SLOTS=3 CB_PIPE_CHUNK_MIB=16" "CB_PIPE_SLOTS=6 CB_PIPE_CHUNK_MIB=16" "CB_PIPE_SLOTS=4 CB_PIPE_CHUNK_MIB=64" "CB_PIPE_SLOTS=8 CB_PIPE_CHUNK_MIB=8"; do
  echo "== N=2 $cfg"
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('value', round(d['value'],1), 'e2e', round(e['value'],1), 'host_link', {k:(round(v,1) if isinstance(v,float) else v) for k,v in e['host_link'].items() if k!='how'})"
done 2>&1 | tee gpurun_out/e2e_pipe_ab_n2.log
for cfg in "CB_PIPE_SLOTS=3 CB_PIPE_CHUNK_MIB=16" "CB_PIPE_SLOTS=6 CB_PIPE_CHUNK_MIB=16" "CB_PIPE_SLOTS=4 CB_PIPE_CHUNK_MIB=64"; do
  echo "== N=1 $cfg"
  env $cfg python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('value', round(d['value'],1), 'e2e', round(e['value'],1), 'host_link', {k:(round(v,1) if isinstance(v,float) else v) for k,v in e['host_link'].items() if k!='how'})"
done 2>&1 | tee gpurun_out/e2e_pipe_ab_n1.log
nvidia-smi topo -m > gpurun_out/r2_topology_n2.txt 2>&1
