"""Probe: where the band-granular luma intra wavefront spends its time (bench workload).
usage: python tools/probe/time_bandwaves.py [frames]"""
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
b = hp.batch_luma
L = pvq._bind()
p = ctypes.byref(b.params)
cur = torch.cuda.current_stream()
s = ctypes.c_void_p(cur.cuda_stream)
top, left = b.dep_top.data_ptr(), b.dep_left.data_ptr()

def t(fn, reps=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        hp.fb.forward(); torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

for mode in ("bands", "waves"):
    b.intra_mode = mode
    print("run_luma_intra mode=%s: %.3f ms" % (mode, t(lambda: b.run_luma_intra())), flush=True)
b.intra_mode = "bands"
_native.check(L.daala_b200_coding_order_gather(p, b.nblocks, 0, s), "g")

def chain(k, mode, small=0, small_mode=3):
    for w, (a, c) in enumerate(b.chain_slices[k]):
        ptr = b.chain_lists[k].data_ptr() + 4 * a
        if w > 0:
            _native.check(L.daala_b200_pvq_intra_band_ref(p, top, left, ptr, c, s), "r")
        _native.check(L.daala_b200_pvq_encode_bands_mode(p, ptr, c, k, small_mode if c < small else mode, s), "b")

for k in (128, 32, 16):
    sl = b.chain_slices[k]
    print("class %d: %d waves, counts %s" % (k, len(sl), [c for _, c in sl]), flush=True)
    print("  chain alone (mode 0): %.3f ms" % t(lambda: chain(k, 0)), flush=True)
    lst = b.bulk_lists[k]
    if lst.numel():
        print("  bulk (%d entries): %.3f ms" % (lst.numel(), t(lambda: _native.check(
            L.daala_b200_pvq_encode_bands_mode(p, lst.data_ptr(), lst.numel(), k, 0, s), "b"))), flush=True)
    # latency of one small wave (64 entries of the last big-enough wave) per kernel variant
    a, c = next((a, c) for a, c in reversed(sl) if c >= 64)
    ptr = b.chain_lists[k].data_ptr() + 4 * a
    modes = [2, 3, 11, 12] + ([13] if k == 128 else [])
    for m in modes:
        for cnt in (1, 64, 2048):
            if cnt > c:
                continue
            ms = t(lambda: _native.check(L.daala_b200_pvq_encode_bands_mode(p, ptr, cnt, k, m, s), "b"))
            print("  mode %2d, %4d entries: %.3f ms" % (m, cnt, ms), flush=True)
