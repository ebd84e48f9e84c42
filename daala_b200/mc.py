"""Host side of the motion-compensation / block-matching entry points
(include/daala_b200.h, "Motion compensation" section): job array layouts and
launch wrappers.  Reference planes are torch uint8 tensors WITH padding; the
kernels take a pointer to pixel (0, 0) and reach into the padding for displaced
windows, like the reference's ref_imgs with OD_BUFFER_PADDING (src/state.h:101)."""
import ctypes

import numpy as np
import torch

from . import _native

MC_BLOCK_DTYPE = np.dtype([("mvx", "<i4", 4), ("mvy", "<i4", 4), ("x0", "<u2"), ("y0", "<u2"),
                           ("log_xblk", "u1"), ("log_yblk", "u1"), ("oc", "u1"), ("s", "u1")])
MATCH_JOB_DTYPE = np.dtype([("mvx", "<i4"), ("mvy", "<i4"), ("x0", "<u2"), ("y0", "<u2"), ("log_blk", "u1"),
                            ("pad_", "u1", 3)])
OD_BUFFER_PADDING = 96

BMA_JOB_DTYPE = np.dtype([("bx", "<i4"), ("by", "<i4"), ("mvx", "<i4"), ("mvy", "<i4"), ("log_mvb_sz", "<i4")])

assert MC_BLOCK_DTYPE.itemsize == 40 and MATCH_JOB_DTYPE.itemsize == 16 and BMA_JOB_DTYPE.itemsize == 20


def _bind():
    L = _native.lib()
    if getattr(L, "_mc_bound", False):
        return L
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.daala_b200_mc_predict_blocks.argtypes = [vp, ci, vp, ci, vp, ci, vp]
    L.daala_b200_mc_match_candidates.argtypes = [vp, ci, vp, ci, vp, ci, ci, vp, vp]
    L.daala_b200_mc_predict1fmv_batch.argtypes = [vp, ci, vp, ci, vp, ci, ci, vp]
    p3, i3 = ctypes.c_void_p * 3, ctypes.c_int * 3
    L.daala_b200_mv_bma_sad.argtypes = [p3, i3, p3, i3, ci, ci, ci, vp, ci, vp, vp]
    L.daala_b200_mv_est_sad.argtypes = [p3, i3, p3, i3, ci, ci, ci, vp, ci, vp, vp]
    L._mc_bound = True
    return L


class PaddedPlane:
    """A u8 plane with `pad` pixels of edge-replicated border on every side
    (od_img_edge_ext, src/state.c:1102)."""

    def __init__(self, plane, pad=OD_BUFFER_PADDING, device="cuda:0"):
        a = np.pad(np.ascontiguousarray(plane), pad, mode="edge")
        self.pad = pad
        self.h, self.w = plane.shape
        self.buf = torch.from_numpy(a).to(device)
        self.stride = self.buf.stride(0)

    @property
    def origin_ptr(self):
        return self.buf.data_ptr() + self.pad * self.stride + self.pad


def to_device(arr, device="cuda:0"):
    return torch.from_numpy(arr.view(np.uint8).reshape(-1)).to(device)


def _stream(stream, device):
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return ctypes.c_void_p(s.cuda_stream)


def predict_blocks(ref, dst, blocks_dev, count, stream=None):
    """OBMC-predict `count` blocks of `ref` (PaddedPlane) into `dst` (2-D uint8 tensor)."""
    L = _bind()
    _native.check(L.daala_b200_mc_predict_blocks(ref.origin_ptr, ref.stride, dst.data_ptr(), dst.stride(0),
                                                 blocks_dev.data_ptr(), count, _stream(stream, dst.device)),
                  "mc_predict_blocks")


def match_candidates(cur, ref, jobs_dev, count, use_satd=False, out=None, stream=None):
    """SAD/SATD of `count` candidate jobs; cur: 2-D uint8 tensor, ref: PaddedPlane."""
    L = _bind()
    if out is None:
        out = torch.empty(count, dtype=torch.int32, device=cur.device)
    _native.check(L.daala_b200_mc_match_candidates(cur.data_ptr(), cur.stride(0), ref.origin_ptr, ref.stride,
                                                   jobs_dev.data_ptr(), count, int(use_satd), out.data_ptr(),
                                                   _stream(stream, cur.device)), "mc_match_candidates")
    return out


def _planes(cur, ref):
    p3, i3 = ctypes.c_void_p * 3, ctypes.c_int * 3
    return (p3(*[c.data_ptr() for c in cur]), i3(*[c.stride(0) for c in cur]), p3(*[r.origin_ptr for r in ref]),
            i3(*[r.stride for r in ref]))


def bma_sad(cur, ref, pic_w, pic_h, jobs_dev, count, out=None, use_chroma=True, stream=None):
    """od_mv_est_bma_sad of `count` half-pel BMA candidates (BMA_JOB_DTYPE); cur: three 2-D uint8 tensors,
    ref: three PaddedPlane (chroma padded by OD_BUFFER_PADDING >> 1 at least)."""
    L = _bind()
    if out is None:
        out = torch.empty(count, dtype=torch.int32, device=cur[0].device)
    _native.check(L.daala_b200_mv_bma_sad(*_planes(cur, ref), pic_w, pic_h, 3 if use_chroma else 1, jobs_dev.data_ptr(),
                                          count, out.data_ptr(), _stream(stream, cur[0].device)), "mv_bma_sad")
    return out


def est_sad(cur, ref, pic_w, pic_h, blocks_dev, count, out=None, use_chroma=True, stream=None):
    """od_mv_est_sad of `count` MV-grid blocks; blocks_dev: [count][3] MC_BLOCK_DTYPE records (one per plane)."""
    L = _bind()
    if out is None:
        out = torch.empty(count, dtype=torch.int32, device=cur[0].device)
    _native.check(L.daala_b200_mv_est_sad(*_planes(cur, ref), pic_w, pic_h, 3 if use_chroma else 1, blocks_dev.data_ptr(),
                                          count, out.data_ptr(), _stream(stream, cur[0].device)), "mv_est_sad")
    return out
