#!/bin/bash
# round 2 evidence on one B200: the GPU suite, compute-sanitizer over every kernel, the launch list of the bench command,
# ncu --set full of the three headline kernels inside the bench command, the bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -5 gpurun_out/pytest_gpu_final.log
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_kernels.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_kernels|launches" gpurun_out/r2_sanitize_$tool.log | tail -4
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; head -c 300 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cb_apply_vec -s 5 -c 1 -o gpurun_out/r2_prof_chain8 -f python bench.py --steps 3 --warmup 3 --no-configs > gpurun_out/ncu_r2_chain8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cb_chain_grad_vec -s 3 -c 1 -o gpurun_out/r2_prof_chain_grad -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_r2_chain_grad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lut16_kernel -s 2 -c 1 -o gpurun_out/r2_prof_lut16 -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_r2_lut16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sum_kernel -s 3 -c 1 -o gpurun_out/r2_prof_sum -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_r2_sum.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
