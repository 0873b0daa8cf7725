#!/bin/bash
# Profiling evidence for profiles/ (run through gpurun):
#  1. launch list of the bench command (gpu__time_duration per launch, cold-cache, serialised)
#  2. one ncu --set full capture of the dominant kernel (fused chain8, f32)
#  3. ncu --set full of the other hot kernels through the --extra rows (one launch each)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cb_apply_vec -s 5 -c 1 -o gpurun_out/prof_chain8_f32 -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none -k regex:"binary_vec_kernel|sum_pass1_kernel|cb_unary_grad_vec|fill16_kernel|copy16_kernel" -c 12 -o gpurun_out/prof_others -f python bench.py --steps 3 --warmup 3 --extra > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/launches.csv
