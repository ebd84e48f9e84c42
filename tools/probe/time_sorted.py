"""Probe: how much do the band kernels gain when the entries of a launch are ordered by work
(so that the lanes of a warp run similar trip counts)?  Keys tried: the actual K of a previous
run (upper bound) and the band energy of the input (what a pre-pass could compute)."""
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
hp.run(); torch.cuda.synchronize()
L = pvq._bind()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
EDGES = torch.tensor(pvq.BAND_EDGES, device="cuda")

def t(fn, reps=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

def keys(b, lst):
    e = lst.to(torch.int64) & 0xffffffff
    blk, band = e >> 4, e & 15
    k_actual = b.res_k[blk * 9 + band].to(torch.float64)
    noref = (b.res_theta[blk * 9 + band] < 0).to(torch.float64)
    off = torch.from_numpy(b.blocks_np["coef_off"].astype(np.int64)).cuda()[blk]
    cs = torch.cumsum(b.in_.to(torch.float64) ** 2, 0)
    cs = torch.cat([torch.zeros(1, dtype=torch.float64, device="cuda"), cs])
    energy = cs[off + EDGES[band + 1]] - cs[off + EDGES[band]]
    return {"actual_k": k_actual * 4 + (1 - noref), "energy": energy, "n_then_energy": (EDGES[band + 1] - EDGES[band]).double() * 1e12 + energy}

def study(name, b, lst, k, modes):
    p = ctypes.byref(b.params)
    base = ["%d:%.3f" % (m, t(lambda: _native.check(L.daala_b200_pvq_encode_bands_mode(p, lst.data_ptr(), lst.numel(), k, m, s), "b"))) for m in modes]
    print("%s (%d entries) unsorted  %s" % (name, lst.numel(), "  ".join(base)), flush=True)
    for kn, kv in keys(b, lst).items():
        order = torch.argsort(kv, descending=True)
        sl = lst[order].contiguous()
        row = ["%d:%.3f" % (m, t(lambda: _native.check(L.daala_b200_pvq_encode_bands_mode(p, sl.data_ptr(), sl.numel(), k, m, s), "b"))) for m in modes]
        print("   sorted by %-14s %s" % (kn, "  ".join(row)), flush=True)

bc = hp.batch_chroma
for k in (128, 32, 16):
    study("chroma class %d" % k, bc, bc.lists[k], k, [2, 3, 11])
bl = hp.batch_luma
for k in (128, 32, 16):
    for w in (0, 1):
        a, c = bl.chain_slices[k][w]
        study("luma class %d wave %d" % (k, w), bl, bl.chain_lists[k][a:a + c].contiguous(), k, [2, 3, 11])
    if bl.bulk_lists[k].numel():
        study("luma class %d bulk" % k, bl, bl.bulk_lists[k], k, [2, 3, 11])
