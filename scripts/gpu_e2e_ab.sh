#!/bin/bash
mkdir -p gpurun_out
for mib in 4 8 16 32 64 128; do
  echo "== CB_PIPE_CHUNK_MIB=$mib"; CB_PIPE_CHUNK_MIB=$mib python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2))"
done 2>&1 | tee gpurun_out/e2e_ab.log
nvidia-smi topo -m 2>/dev/null | head -5; nvidia-smi -q -d PCIE 2>/dev/null | grep -iE "Link Width|Link Gen|Generation" | head -8
