#!/bin/bash
# round 2 (8 GPUs): the multi-GPU tests at world 2 / 4 / 8, the bench line at N = 8 and 4 (strong scaling, the timed
# collective with in-run parity, e2e against the concurrent host-link ceiling), topology
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topology_n8.txt 2>&1; { nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; } >> gpurun_out/r2_topology_n8.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi_n8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_n8.log; tail -5 gpurun_out/pytest_multi_n8.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/bench_n$n.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_n$n.json')); m=d.get('multi_gpu',{})
print('N=$n value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e']['host_link'].items() if k!='how'})
print(json.dumps(m)[:1500])"
done
