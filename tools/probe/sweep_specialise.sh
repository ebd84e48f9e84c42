#!/bin/bash
# persistent PVQ kernel: generic vs size-class specialised band code, with / without the no-reference prepass
mkdir -p gpurun_out
for sp in 0 1; do
  DAALA_B200_NVCC_FLAGS="-DDAALA_PERSIST_SPECIALISE=$sp" python -c "
import os
from daala_b200 import build
os.utime('daala_b200/csrc/kf_engine.cu')
build.build()"
  for pp in 0 1; do
    timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --prepass $pp 2>> gpurun_out/r2n_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('specialise', $sp, 'prepass', $pp, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['kernels_ms']['pvq_luma(gather+k_pvq_persist<intra>+finish)'])"
  done
  if [ $sp = 1 ]; then timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --prepass 1 --frames 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frames 1 specialise 1 prepass 1', d['ms_per_step'])"; fi
done
