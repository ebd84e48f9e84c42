"""GPU parity of the deringing kernel (csrc/dering_kernels.cu) against the pinned CPU oracle
(oracle/port_dering.c, itself checked against od_dering by tests/test_oracle_dering.py): bit-exact
int16 planes and direction maps, luma and 4:2:0 chroma, every superblock position (frame edges,
interior), skip flags with and without the lapping ring."""
import ctypes

import numpy as np
import pytest

from tests import oracle_lib
from tests.oracle_lib import addr

pytestmark = pytest.mark.gpu


class DeringParams(ctypes.Structure):
    _fields_ = [("y", ctypes.c_void_p), ("x", ctypes.c_void_p), ("dir", ctypes.c_void_p), ("bskip", ctypes.c_void_p),
                ("sb_threshold", ctypes.c_void_p),
                ("ystride", ctypes.c_int), ("xstride", ctypes.c_int), ("dir_stride", ctypes.c_int),
                ("skip_stride", ctypes.c_int), ("nhsb", ctypes.c_int), ("nvsb", ctypes.c_int), ("xdec", ctypes.c_int),
                ("pli", ctypes.c_int), ("threshold", ctypes.c_int), ("overlap", ctypes.c_int),
                ("coeff_shift", ctypes.c_int), ("dir_format", ctypes.c_int)]


@pytest.mark.parametrize("xdec", [0, 1])
@pytest.mark.parametrize("threshold,overlap", [(19, 0), (64, 1), (200, 1)])
def test_dering_plane_matches_oracle(xdec, threshold, overlap):
    import torch
    from daala_b200 import _native
    from tests.golden.make_golden import dering_image
    L = _native.lib()
    L.daala_b200_dering_plane.argtypes = [ctypes.POINTER(DeringParams), ctypes.c_void_p]
    port = oracle_lib.load_port()
    nhsb, nvsb = 3, 2
    sb = 64 >> xdec
    img = dering_image(nvsb, nhsb, sb, 31 + xdec)
    h, w = img.shape
    rng = np.random.default_rng(99)
    units = 16 >> xdec
    skip_stride = nhsb * units + 3
    bskip = (rng.random((nvsb * units, skip_stride)) < 0.4).astype(np.uint8)
    dirs = rng.integers(0, 8, size=(nvsb * 8, nhsb * 8)).astype(np.int32)
    # oracle, superblock by superblock
    want = np.zeros_like(img)
    want_dir = dirs.copy()
    Dir = (ctypes.c_int * 8) * 8
    for sby in range(nvsb):
        for sbx in range(nhsb):
            d = Dir()
            for r in range(8):
                for c in range(8):
                    d[r][c] = int(dirs[sby * 8 + r, sbx * 8 + c])
            y = np.zeros((sb, sb), np.int16)
            port.port_dering(addr(y), sb, addr(img, sby * sb * w + sbx * sb), w, 8, 8, sbx, sby, nhsb, nvsb, xdec, d,
                             1 if xdec else 0, addr(bskip, (sby * units) * skip_stride + sbx * units), skip_stride,
                             threshold, overlap, 4)
            want[sby * sb:(sby + 1) * sb, sbx * sb:(sbx + 1) * sb] = y
            want_dir[sby * 8:(sby + 1) * 8, sbx * 8:(sbx + 1) * 8] = np.array([list(r) for r in d])
    x_dev = torch.from_numpy(img).cuda()
    y_dev = torch.zeros_like(x_dev)
    dir_dev = torch.from_numpy(dirs.copy()).cuda()
    skip_dev = torch.from_numpy(bskip).cuda()
    p = DeringParams(y=y_dev.data_ptr(), x=x_dev.data_ptr(), dir=dir_dev.data_ptr(), bskip=skip_dev.data_ptr(),
                     sb_threshold=None, ystride=w, xstride=w, dir_stride=nhsb * 8, skip_stride=skip_stride,
                     nhsb=nhsb, nvsb=nvsb, xdec=xdec, pli=1 if xdec else 0, threshold=threshold, overlap=overlap,
                     coeff_shift=4, dir_format=0)
    assert L.daala_b200_dering_plane(ctypes.byref(p), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    torch.cuda.synchronize()
    assert np.array_equal(dir_dev.cpu().numpy(), want_dir)
    assert np.array_equal(y_dev.cpu().numpy(), want)
