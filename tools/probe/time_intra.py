import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from daala_b200 import _native, pvq
from daala_b200.frame import Geometry
from daala_b200.pipeline import HotPath
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
geom = Geometry(bench.PIC_W, bench.PIC_H)
frames = bench.make_host_frames(geom, F)
hp = HotPath(geom, nframes=F, q0=bench.Q0, pvq_qm_q4=np.full((3, 30), bench.PVQ_QM_Q4, np.uint8), keyframe_prediction=True)
for f, (planes, bsize) in enumerate(frames):
    hp.fb.upload(planes, bsize, frame=f)
hp.set_block_sizes([fr[1] for fr in frames])
b = hp.batch_luma
L = pvq._bind()
p = ctypes.byref(b.params)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn, reps=2):
    hp.fb.forward(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0
    for _ in range(reps):
        hp.fb.forward(); b.epoch += 1; torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps
for bs in range(5):
    ids = b.class_ids[bs]
    def run(bs=bs, ids=ids):
        deps = (b.dep_top.data_ptr(), b.dep_left.data_ptr(), b.done.data_ptr(), b.epoch)
        if bs == 0:
            _native.check(L.daala_b200_pvq_luma_intra_ids(p, ids.data_ptr(), ids.numel(), *deps, s), "x")
        else:
            _native.check(L.daala_b200_pvq_luma_intra_class(p, ids.data_ptr(), ids.numel(), bs, *deps, s), "x")
    print("class bs=%d: %d blocks, %.3f ms" % (bs, ids.numel(), t(run)), flush=True)
# no-dependency variant: all deps = -1
b.dep_top.fill_(-1); b.dep_left.fill_(-1)
for bs in range(5):
    ids = b.class_ids[bs]
    def run(bs=bs, ids=ids):
        deps = (b.dep_top.data_ptr(), b.dep_left.data_ptr(), b.done.data_ptr(), b.epoch)
        if bs == 0:
            _native.check(L.daala_b200_pvq_luma_intra_ids(p, ids.data_ptr(), ids.numel(), *deps, s), "x")
        else:
            _native.check(L.daala_b200_pvq_luma_intra_class(p, ids.data_ptr(), ids.numel(), bs, *deps, s), "x")
    print("NO DEPS class bs=%d: %.3f ms" % (bs, t(run)), flush=True)
