#!/bin/bash
# exact scale-add fusion (CB_FUSE_SCALE_ADD=1): parity tests with it on, then the bench line
mkdir -p gpurun_out
export CB_FUSE_SCALE_ADD=1
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz_expr.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/fuse_ab.log
python bench.py --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('FUSE=1 burst', round(d['value'],1), 'GB/s', d['ms_per_step'], 'ms | sustained', round(d['sustained']['value'],1), 'GB/s @', d['sustained']['clocks']['sm_mhz'], 'MHz | check', d['config']['sampled_check_max_abs_err'])" | tee -a gpurun_out/fuse_ab.log
