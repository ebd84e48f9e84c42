#!/bin/bash
# persistent PVQ kernel: warps per SM (register cap = 65536 / 32 / warps) with one-warp CTAs
mkdir -p gpurun_out
for mb in 4 5 6 8; do
  DAALA_B200_NVCC_FLAGS="-DDAALA_PERSIST_MIN_CTAS=$mb" python -c "
import os
from daala_b200 import build
os.utime('daala_b200/csrc/kf_engine.cu')
build.build()"
  timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>> gpurun_out/r2q_sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps/SM', 4*$mb, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['kernels_ms']['pvq_luma(gather+k_pvq_persist<intra>+finish)'])"
done
