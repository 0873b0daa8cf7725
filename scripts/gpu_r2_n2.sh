#!/bin/bash
# round 2 (2 GPUs): GPU suite incl. world-2 tests, the bench line at N=2 (strong scaling, collective, e2e + concurrent
# host-link ceiling) and the end-to-end pipeline A/B under contention
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n2.log; tail -6 gpurun_out/pytest_gpu_n2.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 ${@:2} 2>gpurun_out/bench_n2.err | tail -1; }
run2 29521 > gpurun_out/bench_n2.json; head -c 300 gpurun_out/bench_n2.json; echo
for cfg in "CB_PIPE_SLOTS=3 CB_PIPE_CHUNK_MIB=16" "CB_PIPE_SLOTS=6 CB_PIPE_CHUNK_MIB=16" "CB_PIPE_SLOTS=4 CB_PIPE_CHUNK_MIB=64" "CB_PIPE_SLOTS=3 CB_PIPE_CHUNK_MIB=128" "CB_PIPE_SLOTS=8 CB_PIPE_CHUNK_MIB=4" "CB_BENCH_WC=1"; do
  echo "== N=2 $cfg"
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('value', round(d['value'],1), 'e2e', round(e['value'],1), 'host_link', {k:(round(v,2) if isinstance(v,float) else v) for k,v in e['host_link'].items() if k!='how'})"
done 2>&1 | tee gpurun_out/e2e_pipe_ab_n2.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; head -c 200 gpurun_out/bench_n1.json; echo
