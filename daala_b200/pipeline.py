"""The device-resident hot path of one batch of frames:

    forward lapped transform -> PVQ band quantisation -> inverse transform

mirroring the order of od_encode_coefficients (reference src/encode.c:2539):
od_ref_plane_to_coeff + od_apply_prefilter_frame_sbs + od_compute_dcts, then
per leaf block the quantisation half of od_block_encode (:1311-1389), then
idct_2d + od_postfilter_split + od_apply_postfilter_frame_sbs +
od_coeff_to_ref_plane.  The serial remainder of the reference (range coder,
RDO on coder state, DC Haar coding) is host work and not part of this path.
"""
import numpy as np
import torch

from . import pvq
from .frame import FrameBuffers


class HotPath:
    def __init__(self, geom, nframes=1, device="cuda:0", q0=38, is_keyframe=1, use_masking=1,
                 lam=pvq.PVQ_LAMBDA, pvq_qm_q4=None, sb_row0=0, sb_rows=None, keyframe_prediction=False,
                 pvq_groups=1):
        self.geom = geom
        self.nframes = nframes
        self.device = torch.device(device)
        self.fb = FrameBuffers(geom, device, nframes)
        self.fb.sb_row0 = sb_row0
        self.fb.sb_rows = geom.nvsb - sb_row0 if sb_rows is None else sb_rows
        self.fb.haar_dc = 1 if is_keyframe else 0
        self.pred = None
        self.q0, self.is_keyframe, self.use_masking, self.lam = q0, is_keyframe, use_masking, lam
        self.pvq_qm_q4 = pvq_qm_q4 if pvq_qm_q4 is not None else np.full((3, 30), 16, np.uint8)
        self.batch = None
        # keyframes with the reference's predictors: luma H/V intra (wavefront kernel), chroma CfL
        self.keyframe_prediction = bool(keyframe_prediction) and bool(is_keyframe)
        self.batch_luma = self.batch_chroma = self.cfl_plane = None
        # keyframe prediction: the frames of the batch are split into `pvq_groups` groups, each with its
        # own luma / chroma PVQ batches on its own stream, so that the latency-bound tail of one group's
        # luma wavefront overlaps the throughput-bound chroma / early-wave work of another
        self.pvq_groups = max(1, min(int(pvq_groups), nframes))
        self.groups = []

    def set_block_sizes(self, bsizes, device_lists=False):
        """bsizes: one map per frame (host numpy).  Builds the block / band
        lists of this rank's superblock rows and uploads them.  device_lists (keyframe prediction only):
        derive every list from the uploaded maps with tensor ops on the device (daala_b200/lists_torch.py,
        identical arrays) instead of numpy on the host."""
        for f, b in enumerate(bsizes):
            self.fb.bsize[f].copy_(torch.from_numpy(np.ascontiguousarray(b)))
        if device_lists:
            assert self.keyframe_prediction and self.pvq_groups == 1
            assert self.fb.sb_row0 == 0 and self.fb.sb_rows == self.geom.nvsb
            from . import lists_torch
            kw = dict(q0=self.q0, is_keyframe=1, use_masking=self.use_masking, lam=self.lam,
                      pvq_qm_q4=self.pvq_qm_q4, device=self.device)
            L = lists_torch.keyframe_lists(self.fb.bsize, self.geom.nhsb, self.geom.nvsb)
            self.cfl_plane = torch.zeros_like(self.fb.coeffs[1])
            bl = pvq.PvqBatch(dict(records=L["luma"], total=L["luma_total"], lists=L["chain"]), self.fb.coeffs, None,
                              **kw)
            bl.setup_intra_device(L)
            bc = pvq.PvqBatch(dict(records=L["chroma"], total=L["chroma_total"], lists=L["chroma_lists"]),
                              self.fb.coeffs, [self.cfl_plane] * 3, **kw)
            self.groups = [(bl, bc, None)]
            self.batch_luma, self.batch_chroma, self.batch = bl, bc, bc
            return
        lists = []
        for f, b in enumerate(bsizes):
            lists.append(pvq.block_list(b, self.geom, frame=f, sb_row0=self.fb.sb_row0, sb_rows=self.fb.sb_rows))
        blocks = np.concatenate(lists)
        if self.keyframe_prediction:
            assert self.fb.sb_row0 == 0 and self.fb.sb_rows == self.geom.nvsb, \
                "intra prediction chains cross superblock rows: keyframe_prediction needs whole frames per rank"
            kw = dict(q0=self.q0, is_keyframe=1, use_masking=self.use_masking, lam=self.lam,
                      pvq_qm_q4=self.pvq_qm_q4, device=self.device)
            self.cfl_plane = torch.zeros_like(self.fb.coeffs[1])
            self.groups = []
            bounds = np.linspace(0, self.nframes, self.pvq_groups + 1).astype(int)
            for g in range(self.pvq_groups):
                sel = blocks[(blocks["frame"] >= bounds[g]) & (blocks["frame"] < bounds[g + 1])]
                luma, top, left, depth = pvq.sort_by_depth(pvq.raster_order(sel[sel["pli"] == 0]), list(bsizes),
                                                           self.geom)
                bl = pvq.PvqBatch(luma, self.fb.coeffs, None, **kw)
                bl.setup_intra(top, left, depth)
                chroma = pvq.mark_luma4x4(sel[sel["pli"] != 0], list(bsizes))
                chroma = chroma[np.argsort(chroma["bs"], kind="stable")]
                bc = pvq.PvqBatch(chroma, self.fb.coeffs, [self.cfl_plane] * 3, **kw)
                stream = torch.cuda.Stream(device=self.device) if self.pvq_groups > 1 else None
                self.groups.append((bl, bc, stream))
            self.batch_luma, self.batch_chroma = self.groups[0][0], self.groups[0][1]
            self.batch = self.batch_chroma
            return
        blocks = blocks[np.argsort(blocks["bs"], kind="stable")]
        self.batch = pvq.PvqBatch(blocks, self.fb.coeffs, self.pred.coeffs if self.pred else None, q0=self.q0,
                                  is_keyframe=self.is_keyframe, use_masking=self.use_masking, lam=self.lam,
                                  pvq_qm_q4=self.pvq_qm_q4, device=self.device)

    def pvq_batches(self):
        """Every PVQ batch of this pass (what produces symbols for the host entropy coder)."""
        if self.groups:
            return [b for bl, bc, _ in self.groups for b in (bl, bc)]
        return [self.batch]

    def use_prediction(self, pred_fb):
        """Inter frames: `pred_fb.coeffs` hold the transformed motion-compensated
        prediction (the reference's mdtmp planes)."""
        self.pred = pred_fb

    def capture(self):
        """Record one whole pass (all streams, ~10^2-10^3 launches, many of them latency-bound wave
        kernels) into a CUDA graph; `replay()` then costs one launch on the host.  Pointers, lists and
        tensor maps are baked in: call again after set_block_sizes()."""
        torch.cuda.synchronize(self.device)
        self.run()                       # loads modules, creates streams outside the capture
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.graph_launches = self.run()
        self.graph = graph
        return graph

    def replay(self):
        self.graph.replay()
        return self.graph_launches

    def run_pvq(self):
        """The PVQ stage of one pass (after forward()); returns the number of kernel launches."""
        if not self.keyframe_prediction:
            return self.batch.run()
        n = 0
        main = torch.cuda.current_stream(self.device)
        for bl, bc, st in self.groups:
            if st is not None:
                st.wait_stream(main)
            n += bl.run_luma_intra(st)
            n += bc.cfl_pred(self.cfl_plane, st)
            n += bc.run(st)
        for _, _, st in self.groups:
            if st is not None:
                main.wait_stream(st)
        return n

    def run(self, exchange=None):
        """One pass; returns the number of kernel launches.  `exchange` (multi-GPU)
        is called between the two halves of the inverse to trade lapped border rows."""
        self.fb.forward()
        n = 1 + self.run_pvq()
        self.fb.inverse(lapped_only=True)
        if exchange is not None:
            exchange()
        self.fb.sb_postfilter_store()
        return n + 2
