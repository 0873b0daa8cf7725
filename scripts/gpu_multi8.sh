#!/bin/bash
# final 8-GPU pass: multi-GPU tests, weak-scaling bench at 1/2/4/8, sharded sum with both exchange paths
bash scripts/gpu_multi.sh 8 2>&1 | grep -v "sum_f32" 
for mode in 1 0; do for n in 2 4 8; do
  CB_COMM_P2P=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 scripts/bench_configs.py --sum-only > gpurun_out/sum_p2p${mode}_$n.json 2> gpurun_out/sum_p2p${mode}_$n.err
  tail -1 gpurun_out/sum_p2p${mode}_$n.json || tail -3 gpurun_out/sum_p2p${mode}_$n.err
done; done
