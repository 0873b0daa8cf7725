#!/bin/bash
# one ncu --set full capture of the fused chain kernel (run through gpurun); extra env passes tunables
mkdir -p gpurun_out
NAME=${1:-prof_chain8}
ncu --set full --clock-control none --import-source on -k regex:cb_apply_vec -s 5 -c 1 -o gpurun_out/$NAME -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_$NAME.log 2>&1
tail -3 gpurun_out/ncu_$NAME.log
