#!/bin/bash
# Runs on the GPU box (via gpurun): GPU test suite, the bench line, and one ncu --set full
# capture of the dominant kernel.  Everything it writes goes to gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
ncu --set full --clock-control none --import-source on -k regex:cb_apply_vec -s 5 -c 2 -o gpurun_out/prof_chain8 -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_chain8.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; head -c 700 gpurun_out/bench2.json
