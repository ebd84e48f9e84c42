"""ncu target: one quantise() of the CfL chroma batch (all with-reference searches) on the bench
workload.  usage: ncu --profile-from-start off ... python tools/probe/prof_chroma.py [frames]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from daala_b200.frame import Geometry
from daala_b200.pipeline import HotPath
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
geom = Geometry(bench.PIC_W, bench.PIC_H)
frames = bench.make_host_frames(geom, F)
hp = HotPath(geom, nframes=F, q0=bench.Q0, pvq_qm_q4=np.full((3, 30), bench.PVQ_QM_Q4, np.uint8), keyframe_prediction=True)
for f, (planes, bsize) in enumerate(frames):
    hp.fb.upload(planes, bsize, frame=f)
hp.set_block_sizes([fr[1] for fr in frames])
hp.run(); torch.cuda.synchronize()
hp.fb.forward(); hp.batch_luma.run_luma_intra(); hp.batch_chroma.cfl_pred(hp.cfl_plane)
hp.batch_chroma.gather(); torch.cuda.synchronize()
torch.cuda.profiler.start()
hp.batch_chroma.quantise()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
