#!/bin/bash
# A/B of the tile-level fast-path decision (CB_TILE_REDO) and launch shape around it; then parity tests with the default
mkdir -p gpurun_out
for cfg in "CB_TILE_REDO=1" "CB_TILE_REDO=0" "CB_TILE_REDO=1 CB_MIN_BLOCKS=4" "CB_TILE_REDO=1 CB_MIN_BLOCKS=6" "CB_TILE_REDO=1 CB_UNROLL=2 CB_MIN_BLOCKS=6" "CB_TILE_REDO=1 CB_UNROLL=2 CB_MIN_BLOCKS=8"; do
  echo "== $cfg"; env $cfg python bench.py --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('burst', round(d['value'],1), 'GB/s', d['ms_per_step'], 'ms | sustained', round(d['sustained']['value'],1), 'GB/s @', d['sustained']['clocks']['sm_mhz'], 'MHz', d['sustained']['clocks'].get('power_w_max'), 'W')"
done 2>&1 | tee gpurun_out/tile_ab.log
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dtypes.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -5 | tee -a gpurun_out/tile_ab.log
