#!/bin/bash
# run with: gpurun --gpus N -- bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
    python -c "
import json,sys
d=json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1])
print($n,'gpus value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'sustained',round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/scale_$n.err
  fi
done
# sharded sum/mean over 2^30 f32 (BASELINE configs[3]): contiguous slices + NCCL combine of the partials
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then python scripts/bench_configs.py --sum-only > gpurun_out/sum_$n.json 2> gpurun_out/sum_$n.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 scripts/bench_configs.py --sum-only > gpurun_out/sum_$n.json 2> gpurun_out/sum_$n.err; fi
    tail -1 gpurun_out/sum_$n.json || tail -3 gpurun_out/sum_$n.err
  fi
done
