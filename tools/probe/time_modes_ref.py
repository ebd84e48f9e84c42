"""Probe: throughput of every band-kernel variant on with-reference data (bench workload):
luma intra waves 0/1 and the chroma (CfL) batch.  usage: python tools/probe/time_modes_ref.py [frames]"""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from daala_b200 import _native, pvq
from daala_b200.frame import Geometry
from daala_b200.pipeline import HotPath
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
geom = Geometry(bench.PIC_W, bench.PIC_H)
frames = bench.make_host_frames(geom, F)
hp = HotPath(geom, nframes=F, q0=bench.Q0, pvq_qm_q4=np.full((3, 30), bench.PVQ_QM_Q4, np.uint8), keyframe_prediction=True)
for f, (planes, bsize) in enumerate(frames):
    hp.fb.upload(planes, bsize, frame=f)
hp.set_block_sizes([fr[1] for fr in frames])
hp.run(); torch.cuda.synchronize()     # leaves in/ref of both batches populated with real data
L = pvq._bind()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def t(fn, reps=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

MODES = {16: [2, 20, 21, 3, 30, 31], 32: [2, 20, 21, 3, 30, 31], 128: [2, 3, 30, 31, 11]}
b = hp.batch_luma
p = ctypes.byref(b.params)
for k in (128, 32, 16):
    for w in (0, 1, 3):
        a, c = b.chain_slices[k][w]
        ptr = b.chain_lists[k].data_ptr() + 4 * a
        row = ["%d:%.3f" % (m, t(lambda: _native.check(L.daala_b200_pvq_encode_bands_mode(p, ptr, c, k, m, s), "b")))
               for m in MODES[k]]
        print("luma class %3d wave %d (%7d entries)  ms by mode  %s" % (k, w, c, "  ".join(row)), flush=True)
    lst = b.bulk_lists[k]
    if lst.numel():
        row = ["%d:%.3f" % (m, t(lambda: _native.check(L.daala_b200_pvq_encode_bands_mode(p, lst.data_ptr(), lst.numel(), k, m, s), "b")))
               for m in MODES[k]]
        print("luma class %3d bulk   (%7d entries)  ms by mode  %s" % (k, lst.numel(), "  ".join(row)), flush=True)
b = hp.batch_chroma
p = ctypes.byref(b.params)
for k in (128, 32, 16):
    lst = b.lists[k]
    row = ["%d:%.3f" % (m, t(lambda: _native.check(L.daala_b200_pvq_encode_bands_mode(p, lst.data_ptr(), lst.numel(), k, m, s), "b")))
           for m in MODES[k]]
    print("chroma class %3d (%7d entries)  ms by mode  %s" % (k, lst.numel(), "  ".join(row)), flush=True)
