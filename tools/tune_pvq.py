#!/usr/bin/env python
"""Times the PVQ band kernels per size class and launch geometry on the bench
workload (dev tool; run on the GPU box)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from daala_b200 import _native, pvq  # noqa: E402
from daala_b200.frame import Geometry  # noqa: E402
from daala_b200.pipeline import HotPath  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
geom = Geometry(bench.PIC_W, bench.PIC_H)
frames = bench.make_host_frames(geom, F)
hp = HotPath(geom, nframes=F, q0=bench.Q0, pvq_qm_q4=np.full((3, 30), bench.PVQ_QM_Q4, np.uint8))
for f, (planes, bsize) in enumerate(frames):
    hp.fb.upload(planes, bsize, frame=f)
hp.set_block_sizes([fr[1] for fr in frames])
hp.fb.forward()
b = hp.batch
b.gather()
L = pvq._bind()
p = ctypes.byref(b.params)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ref = None
for mode in (2, 0, 11, 12, 13):
    line = "mode %2d:" % mode
    for nmax in (16, 32, 128):
        lst = b.lists[nmax]
        for _ in range(2):
            _native.check(L.daala_b200_pvq_encode_bands_mode(p, lst.data_ptr(), lst.numel(), nmax, mode, s), "x")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            _native.check(L.daala_b200_pvq_encode_bands_mode(p, lst.data_ptr(), lst.numel(), nmax, mode, s), "x")
        e1.record()
        torch.cuda.synchronize()
        line += "  n<=%3d: %8.3f ms (%d bands)" % (nmax, e0.elapsed_time(e1) / 3, lst.numel())
    sig = (int(b.res_k.sum().item()), int(b.out.sum().item()), int(b.y.abs().sum().item()))
    if ref is None:
        ref = sig
    line += "  checksum %s" % ("ok" if sig == ref else "MISMATCH %s vs %s" % (sig, ref))
    print(line, flush=True)
