"""ctypes binding of libdaala_b200.so (the C ABI of include/daala_b200.h).

Loading fails loudly when the library has not been built: there is no CPU
fallback anywhere in this package.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdaala_b200.so")

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_void_p = ctypes.c_void_p


class Plane(ctypes.Structure):
    """struct daala_b200_plane"""
    _fields_ = [
        ("pixels", c_void_p), ("coeffs", c_void_p), ("lapped", c_void_p), ("pixels_out", c_void_p),
        ("pixel_stride", c_int), ("coeff_stride", c_int), ("lapped_stride", c_int),
        ("pixel_out_stride", c_int), ("xdec", c_int), ("pad_", c_int),
        ("pixel_frame_pitch", c_ll), ("coeff_frame_pitch", c_ll), ("lapped_frame_pitch", c_ll),
        ("pixel_out_frame_pitch", c_ll),
    ]


class Frame(ctypes.Structure):
    """struct daala_b200_frame"""
    _fields_ = [
        ("plane", Plane * 3), ("bsize", c_void_p), ("bstride", c_int), ("nhsb", c_int),
        ("nvsb", c_int), ("pic_w", c_int), ("pic_h", c_int), ("haar_dc", c_int),
        ("nframes", c_int), ("sb_row0", c_int), ("sb_rows", c_int), ("pad_", c_int),
        ("bsize_frame_pitch", c_ll), ("post16", ctypes.c_void_p * 3),
    ]


_lib = None


def lib():
    """The loaded library; raises if it is missing (build with daala_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libdaala_b200.so is not built (run `python -m daala_b200.build` or "
                "__graft_entry__.build()); daala_b200 has no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        fp = ctypes.POINTER(Frame)
        for name in ("daala_b200_forward_frame", "daala_b200_forward_frame_no_tma", "daala_b200_inverse_frame",
                     "daala_b200_inverse_frame_lapped", "daala_b200_sb_postfilter_store_frame"):
            fn = getattr(L, name)
            fn.argtypes = [fp, c_int, c_void_p]
            fn.restype = c_int
        L.daala_b200_plane_sb_filter.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.daala_b200_plane_sb_filter.restype = c_int
        L.daala_b200_block_transform.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p]
        L.daala_b200_block_transform.restype = c_int
        L.daala_b200_device_count.restype = c_int
        L.daala_b200_version.restype = ctypes.c_char_p
        _lib = L
    return _lib


class CudaError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise CudaError("%s failed with cudaError %d" % (what, rc))
