#!/bin/bash
# run with: gpurun --gpus N -- bash scripts/gpu_multi_sum.sh N   (sharded sum: test + timing, both exchange paths)
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
for mode in 1 0; do for n in 2 4 8; do
  if [ $n -le $N ]; then
    CB_COMM_P2P=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 scripts/bench_configs.py --sum-only > gpurun_out/sum_p2p${mode}_$n.json 2> gpurun_out/sum_p2p${mode}_$n.err
    tail -1 gpurun_out/sum_p2p${mode}_$n.json || tail -3 gpurun_out/sum_p2p${mode}_$n.err
  fi
done; done
