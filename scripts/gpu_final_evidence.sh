#!/bin/bash
# final single-GPU evidence pass: tests, bench line (+extra), reference arm, launch list, ncu of the shipped chain kernel
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cb_apply_vec -s 5 -c 1 -o gpurun_out/prof_chain8_final -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/smoke.log; tail -4 gpurun_out/pytest_gpu.log; head -c 400 gpurun_out/bench.json; echo; head -c 300 gpurun_out/bench_ref.json; echo; wc -l gpurun_out/launches.csv
