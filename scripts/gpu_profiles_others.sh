#!/bin/bash
# ncu --set full of the other hot kernels (one launch each) through bench.py --extra
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:"binary_vec_kernel|sum_pass1_kernel|sum_pass2_kernel|cb_unary_grad_vec|fill16_kernel|copy16_kernel" -c 14 -o gpurun_out/prof_others -f python bench.py --steps 3 --warmup 3 --extra > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log; ls -la gpurun_out/prof_others.ncu-rep
