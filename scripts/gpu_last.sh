#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
ncu --set full --clock-control none -k regex:cb_apply_vec -c 4 -o gpurun_out/prof_half -f python scripts/profile_half.py > gpurun_out/ncu_half.log 2>&1; tail -2 gpurun_out/ncu_half.log
