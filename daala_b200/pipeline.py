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
                 lam=pvq.PVQ_LAMBDA, pvq_qm_q4=None, sb_row0=0, sb_rows=None):
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

    def set_block_sizes(self, bsizes):
        """bsizes: one map per frame (host numpy).  Builds the block / band
        lists of this rank's superblock rows and uploads them."""
        lists = []
        for f, b in enumerate(bsizes):
            self.fb.bsize[f].copy_(torch.from_numpy(np.ascontiguousarray(b)))
            lists.append(pvq.block_list(b, self.geom, frame=f, sb_row0=self.fb.sb_row0, sb_rows=self.fb.sb_rows))
        blocks = np.concatenate(lists)
        blocks = blocks[np.argsort(blocks["bs"], kind="stable")]
        self.batch = pvq.PvqBatch(blocks, self.fb.coeffs, self.pred.coeffs if self.pred else None, q0=self.q0,
                                  is_keyframe=self.is_keyframe, use_masking=self.use_masking, lam=self.lam,
                                  pvq_qm_q4=self.pvq_qm_q4, device=self.device)

    def use_prediction(self, pred_fb):
        """Inter frames: `pred_fb.coeffs` hold the transformed motion-compensated
        prediction (the reference's mdtmp planes)."""
        self.pred = pred_fb

    def run(self, exchange=None):
        """One pass; returns the number of kernel launches.  `exchange` (multi-GPU)
        is called between the two halves of the inverse to trade lapped border rows."""
        self.fb.forward()
        n = 1 + self.batch.run()
        self.fb.inverse(lapped_only=True)
        if exchange is not None:
            exchange()
        self.fb.sb_postfilter_store()
        return n + 2
