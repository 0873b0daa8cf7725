#!/bin/bash
# First GPU call of the next round: re-capture the evidence with the scale-and-shift fusion on (the committed ncu
# captures predate it), and measure the one tunable left open (CB_NEG_XOR on top of the fusion).
#   gpurun --timeout 1500 -- bash scripts/gpu_round2_evidence.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 --extra > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 300 gpurun_out/bench.json; echo
for cfg in "CB_NEG_XOR=0" "CB_NEG_XOR=1" "CB_FUSE_SCALE_ADD=0"; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('burst', round(d['value'],1), 'GB/s | sustained', round(d['sustained']['value'],1), 'GB/s @', d['sustained']['clocks']['sm_mhz'], 'MHz')"
done 2>&1 | tee gpurun_out/round2_ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cb_apply_vec -s 5 -c 1 -o gpurun_out/prof_chain8_fused -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_fused.log 2>&1
python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
wc -l gpurun_out/launches.csv; ls -la gpurun_out/*.ncu-rep
# host-side ASan/UBSan WITH a device: the module layer (handles, ids, aliasing, fusing) under the module tests and fuzz
python -m custos_b200.build --sanitize > gpurun_out/asan_build.log 2>&1
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  CUSTOS_B200_LIB=custos_b200/lib/libcustos_b200_asan.so \
  python -m pytest tests/test_gpu_modules.py tests/test_gpu_fuzz_modules.py tests/test_gpu_reference_suite.py tests/test_gpu_untyped.py -m gpu -q -p no:cacheprovider > gpurun_out/asan_gpu.log 2>&1
tail -3 gpurun_out/asan_gpu.log
# host topology of the box (for the multi-GPU end-to-end analysis)
{ nvidia-smi topo -m; echo; numactl -H 2>/dev/null || echo "no numactl"; echo; lscpu | head -30; echo; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c; nproc; free -g; } > gpurun_out/r2_topology_n1.txt 2>&1
