#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --launch-skip 40 -c 40 -o gpurun_out/prof_all_kernels -f python scripts/profile_kernels.py > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log; ls -la gpurun_out/prof_all_kernels.ncu-rep
