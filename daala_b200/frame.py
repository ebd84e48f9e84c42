"""Frame geometry and the device-resident transform pipeline (host side).

Mirrors the geometry rules of the reference's od_state_init_impl
(src/state.c:376-379: frame padded to whole 64x64 superblocks) and drives the
batch entry points of include/daala_b200.h.  torch supplies device memory and
the stream; every kernel is in libdaala_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _native

OD_BSIZE_MAX = 64


class Geometry:
    """Padded frame geometry of a 4:2:0 (or 4:4:4) picture."""

    def __init__(self, pic_w, pic_h, xdec=1, nplanes=3):
        self.pic_w, self.pic_h = pic_w, pic_h
        self.nhsb = (pic_w + OD_BSIZE_MAX - 1) // OD_BSIZE_MAX
        self.nvsb = (pic_h + OD_BSIZE_MAX - 1) // OD_BSIZE_MAX
        self.frame_w = self.nhsb * OD_BSIZE_MAX
        self.frame_h = self.nvsb * OD_BSIZE_MAX
        self.nplanes = nplanes
        self.xdec = [0] + [xdec] * (nplanes - 1)

    def plane_shape(self, pli):
        return self.frame_h >> self.xdec[pli], self.frame_w >> self.xdec[pli]

    @property
    def bsize_shape(self):
        return self.nvsb * 8, self.nhsb * 8

    @property
    def luma_pixels(self):
        return self.pic_w * self.pic_h

    @property
    def padded_samples(self):
        return sum(h * w for h, w in (self.plane_shape(p) for p in range(self.nplanes)))

    def shard_rows(self, rank, world):
        """Contiguous superblock rows of `rank` (34 rows over 8
        ranks -> 5,5,4,4,4,4,4,4)."""
        base, extra = divmod(self.nvsb, world)
        n = base + (1 if rank < extra else 0)
        r0 = rank * base + min(rank, extra)
        return r0, n


class FrameBuffers:
    """Device buffers of a batch of `nframes` frames of one geometry: u8 planes
    in, int32 coefficient (`d`) planes, int32 lapped (`c`) planes, u8 planes
    out, and one block-size map per frame.  Tensors are [nframes, h, w]."""

    def __init__(self, geom, device="cuda:0", nframes=1):
        self.geom = geom
        self.nframes = nframes
        self.device = torch.device(device)
        g = geom

        def alloc(dtype, p):
            return torch.zeros((nframes,) + g.plane_shape(p), dtype=dtype, device=self.device)

        self.pixels = [alloc(torch.uint8, p) for p in range(g.nplanes)]
        self.coeffs = [alloc(torch.int32, p) for p in range(g.nplanes)]
        self.lapped = [alloc(torch.int32, p) for p in range(g.nplanes)]
        self.pixels_out = [alloc(torch.uint8, p) for p in range(g.nplanes)]
        self.bsize = torch.zeros((nframes,) + g.bsize_shape, dtype=torch.uint8, device=self.device)
        self.haar_dc = 1
        self.sb_row0, self.sb_rows = 0, g.nvsb

    def descriptor(self):
        g = self.geom
        f = _native.Frame()
        for p in range(g.nplanes):
            pl = f.plane[p]
            pl.pixels = self.pixels[p].data_ptr()
            pl.coeffs = self.coeffs[p].data_ptr()
            pl.lapped = self.lapped[p].data_ptr()
            pl.pixels_out = self.pixels_out[p].data_ptr()
            pl.pixel_stride = self.pixels[p].stride(1)
            pl.coeff_stride = self.coeffs[p].stride(1)
            pl.lapped_stride = self.lapped[p].stride(1)
            pl.pixel_out_stride = self.pixels_out[p].stride(1)
            pl.pixel_frame_pitch = self.pixels[p].stride(0)
            pl.coeff_frame_pitch = self.coeffs[p].stride(0)
            pl.lapped_frame_pitch = self.lapped[p].stride(0)
            pl.pixel_out_frame_pitch = self.pixels_out[p].stride(0)
            pl.xdec = g.xdec[p]
        f.bsize = self.bsize.data_ptr()
        f.bstride = self.bsize.stride(1)
        f.bsize_frame_pitch = self.bsize.stride(0)
        f.nhsb, f.nvsb = g.nhsb, g.nvsb
        f.pic_w, f.pic_h = g.pic_w, g.pic_h
        f.haar_dc = int(self.haar_dc)
        f.nframes = self.nframes
        f.sb_row0, f.sb_rows = self.sb_row0, self.sb_rows
        return f

    # --- host <-> device -------------------------------------------------
    def upload(self, planes, bsize=None, frame=0):
        for p, a in enumerate(planes):
            self.pixels[p][frame].copy_(torch.from_numpy(np.ascontiguousarray(a)), non_blocking=True)
        if bsize is not None:
            self.bsize[frame].copy_(torch.from_numpy(np.ascontiguousarray(bsize)), non_blocking=True)

    # --- kernels ---------------------------------------------------------
    def _call(self, name, stream):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        f = self.descriptor()
        fn = getattr(_native.lib(), name)
        _native.check(fn(ctypes.byref(f), self.geom.nplanes, ctypes.c_void_p(s.cuda_stream)), name)

    def forward(self, stream=None, tma=True):
        self._call("daala_b200_forward_frame" if tma else "daala_b200_forward_frame_no_tma", stream)

    def inverse(self, stream=None, lapped_only=False):
        self._call("daala_b200_inverse_frame_lapped" if lapped_only else "daala_b200_inverse_frame", stream)

    def sb_postfilter_store(self, stream=None):
        self._call("daala_b200_sb_postfilter_store_frame", stream)
