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


class FrameBuffers:
    """Device buffers of one frame: u8 planes in, int32 d planes, int32 lapped
    planes, u8 planes out and the block-size map.  `batch` frames are stacked
    along the superblock-row axis (independent frames simply extend nvsb for
    the all-intra transform path... NOT used: each frame keeps its own edges),
    so a batch is a list of FrameBuffers instead."""

    def __init__(self, geom, device="cuda:0"):
        self.geom = geom
        self.device = torch.device(device)
        g = geom
        self.pixels = [torch.zeros(g.plane_shape(p), dtype=torch.uint8, device=self.device) for p in range(g.nplanes)]
        self.coeffs = [torch.zeros(g.plane_shape(p), dtype=torch.int32, device=self.device) for p in range(g.nplanes)]
        self.lapped = [torch.zeros(g.plane_shape(p), dtype=torch.int32, device=self.device) for p in range(g.nplanes)]
        self.pixels_out = [torch.zeros(g.plane_shape(p), dtype=torch.uint8, device=self.device) for p in range(g.nplanes)]
        self.bsize = torch.zeros(g.bsize_shape, dtype=torch.uint8, device=self.device)
        self.haar_dc = 1
        self._desc = None

    def descriptor(self):
        g = self.geom
        f = _native.Frame()
        for p in range(g.nplanes):
            pl = f.plane[p]
            pl.pixels = self.pixels[p].data_ptr()
            pl.coeffs = self.coeffs[p].data_ptr()
            pl.lapped = self.lapped[p].data_ptr()
            pl.pixels_out = self.pixels_out[p].data_ptr()
            pl.pixel_stride = self.pixels[p].stride(0)
            pl.coeff_stride = self.coeffs[p].stride(0)
            pl.lapped_stride = self.lapped[p].stride(0)
            pl.pixel_out_stride = self.pixels_out[p].stride(0)
            pl.xdec = g.xdec[p]
        f.bsize = self.bsize.data_ptr()
        f.bstride = self.bsize.stride(0)
        f.nhsb, f.nvsb = g.nhsb, g.nvsb
        f.pic_w, f.pic_h = g.pic_w, g.pic_h
        f.haar_dc = int(self.haar_dc)
        return f

    # --- host <-> device -------------------------------------------------
    def upload(self, planes, bsize=None):
        for p, a in enumerate(planes):
            self.pixels[p].copy_(torch.from_numpy(np.ascontiguousarray(a)), non_blocking=True)
        if bsize is not None:
            self.bsize.copy_(torch.from_numpy(np.ascontiguousarray(bsize)), non_blocking=True)

    # --- kernels ---------------------------------------------------------
    def forward(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        f = self.descriptor()
        _native.check(_native.lib().daala_b200_forward_frame(ctypes.byref(f), self.geom.nplanes,
                                                             ctypes.c_void_p(s.cuda_stream)), "forward_frame")

    def inverse(self, stream=None, lapped_only=False):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        f = self.descriptor()
        fn = _native.lib().daala_b200_inverse_frame_lapped if lapped_only else _native.lib().daala_b200_inverse_frame
        _native.check(fn(ctypes.byref(f), self.geom.nplanes, ctypes.c_void_p(s.cuda_stream)), "inverse_frame")
