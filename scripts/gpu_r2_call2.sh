#!/bin/bash
# round 2, call 2: GPU suite on the new tree, the fused backward bench, ncu of the chain-grad kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log; tail -15 gpurun_out/pytest_gpu2.log
python scripts/bench_bwd.py > gpurun_out/bench_bwd.json 2> gpurun_out/bench_bwd.err; cat gpurun_out/bench_bwd.json; tail -3 gpurun_out/bench_bwd.err
ncu --set full --clock-control none --import-source on -k regex:cb_chain_grad_vec -s 3 -c 1 -o gpurun_out/prof_chain_grad -f python scripts/bench_bwd.py $((1<<26)) > gpurun_out/ncu_chain_grad.log 2>&1
tail -2 gpurun_out/ncu_chain_grad.log
