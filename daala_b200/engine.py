"""Host side of the keyframe engine (include/daala_b200.h, "Keyframe engine"; csrc/kf_engine.cu):
ctypes binding + numpy marshalling.  The engine is the batched equivalent of od_encode_coefficients
(reference src/encode.c:2539) for keyframes without the entropy coder: u8 planes + block-size maps in,
reconstruction + PVQ symbols out, everything in between on the GPU (work lists included).

No torch here: device memory, streams and the CUDA graph belong to the engine."""
import ctypes

import numpy as np

from . import _native, pvq
from .frame import Geometry

c_int, c_ll, c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p

PH_LISTS, PH_FORWARD, PH_PVQ_LUMA, PH_INVERSE, PH_PVQ_CHROMA = 1, 2, 4, 8, 16
PH_PVQ, PH_ALL, PH_SEARCH_ONLY = 20, 31, 64
CNT = dict(n_luma=0, n_chroma=1, luma_coefs=2, chroma_coefs=3, items_l=4, items_c=7,
           total_hi=14, n_heads=15, n_heads0=16, error=17)


class Config(ctypes.Structure):
    _fields_ = [("pic_w", c_int), ("pic_h", c_int), ("nframes", c_int), ("q0", c_int), ("use_masking", c_int),
                ("qm_stride", c_int), ("pvq_norm_lambda", ctypes.c_double), ("pvq_qm_q4", (ctypes.c_ubyte * 32) * 3),
                ("qm", c_void_p), ("qm_inv", c_void_p), ("sb_row0", c_int), ("sb_rows", c_int),
                ("max_blocks_div", c_int), ("persist_ctas_per_sm", c_int), ("split_free", c_int), ("dering", c_int), ("noref_prepass", c_int), ("level_chains", c_int), ("stream", c_void_p),
                ("coded_quantizer", c_int), ("qm_is_flat", c_int), ("dering_lambda", ctypes.c_double)]


class Totals(ctypes.Structure):
    _fields_ = [("n_luma", c_ll), ("luma_coefs", c_ll), ("n_chroma", c_ll), ("chroma_coefs", c_ll)]


class IO(ctypes.Structure):
    _fields_ = [("pixels", c_void_p * 3), ("bsize", c_void_p), ("dering_level", c_void_p), ("totals", ctypes.POINTER(Totals)),
                ("pixels_out", c_void_p * 3), ("luma_blocks", c_void_p), ("chroma_blocks", c_void_p),
                ("luma_res", c_void_p), ("chroma_res", c_void_p), ("luma_y16", c_void_p), ("chroma_y16", c_void_p),
                ("luma_skip_diff", c_void_p), ("chroma_skip_diff", c_void_p), ("chroma_flip", c_void_p),
                ("counts", c_void_p), ("dering_level_out", c_void_p)]


class Buffers(ctypes.Structure):
    _fields_ = [("pixels", c_void_p * 3), ("coeffs", c_void_p * 3), ("lapped", c_void_p * 3),
                ("pixels_out", c_void_p * 3), ("plane_w", c_int * 3), ("plane_h", c_int * 3), ("bsize", c_void_p),
                ("counts", c_void_p), ("luma_blocks", c_void_p), ("chroma_blocks", c_void_p), ("dep_top", c_void_p),
                ("dep_left", c_void_p), ("succ_bottom", c_void_p), ("succ_right", c_void_p), ("luma_items", c_void_p * 3),
                ("luma_heads", c_void_p), ("luma_heads0", c_void_p), ("chroma_items", c_void_p * 3),
                ("luma_res", c_void_p), ("chroma_res", c_void_p), ("luma_y16", c_void_p), ("chroma_y16", c_void_p),
                ("luma_skip_diff", c_void_p), ("chroma_skip_diff", c_void_p), ("chroma_flip", c_void_p),
                ("max_luma_blocks", c_int), ("max_chroma_blocks", c_int), ("stream", c_void_p),
                ("bytes_allocated", c_ll)]


def _bind():
    L = _native.lib()
    if getattr(L, "_kf_bound", False):
        return L
    L.daala_b200_kf_create.argtypes = [ctypes.POINTER(Config)]
    L.daala_b200_kf_create.restype = c_void_p
    L.daala_b200_kf_destroy.argtypes = [c_void_p]
    L.daala_b200_kf_destroy.restype = None
    L.daala_b200_kf_error.argtypes = [c_void_p]
    L.daala_b200_kf_error.restype = ctypes.c_char_p
    L.daala_b200_kf_device_buffers.argtypes = [c_void_p, ctypes.POINTER(Buffers)]
    L.daala_b200_kf_launches_per_step.argtypes = [c_void_p]
    L.daala_b200_kf_run_device.argtypes = [c_void_p, c_int, c_int]
    L.daala_b200_kf_time_device.argtypes = [c_void_p, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_float)]
    L.daala_b200_kf_count_blocks.argtypes = [c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_int,
                                             ctypes.POINTER(Totals)]
    for name in ("daala_b200_kf_submit", "daala_b200_kf_encode"):
        getattr(L, name).argtypes = [c_void_p, ctypes.POINTER(IO)]
    L.daala_b200_kf_wait.argtypes = [c_void_p]
    L.daala_b200_device_copy.argtypes = [c_void_p, c_void_p, ctypes.c_size_t, c_int]
    L.daala_b200_host_alloc.argtypes = [ctypes.c_size_t]
    L.daala_b200_host_alloc.restype = c_void_p
    L.daala_b200_host_free.argtypes = [c_void_p]
    L.daala_b200_host_free.restype = None
    L._kf_bound = True
    return L


class Pinned:
    """A numpy array over page-locked host memory (daala_b200_host_alloc)."""

    def __init__(self, shape, dtype):
        self.L = _bind()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.daala_b200_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("daala_b200_host_alloc(%d) failed" % self.nbytes)
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.daala_b200_host_free(self.ptr)
            self.ptr = None


class KeyframeEngine:
    """One engine = one set of device buffers + one CUDA graph for batches of `nframes` keyframes."""

    def __init__(self, geom, nframes=1, q0=38, use_masking=1, lam=pvq.PVQ_LAMBDA, pvq_qm_q4=None, qm=None,
                 qm_inv=None, sb_row0=0, sb_rows=0, max_blocks_div=0, persist_ctas_per_sm=0, split_free=0, level_chains=0, noref_prepass=0, dering=0, coded_quantizer=0,
                 qm_is_flat=0, dering_lambda=None, pinned=True):
        self.L = _bind()
        self.geom, self.F = geom, nframes
        if qm is None:
            qm, qm_inv = pvq.default_qm(True)
        self._qm = np.ascontiguousarray(qm, np.int16)
        self._qm_inv = np.ascontiguousarray(qm_inv, np.int16)
        cfg = Config()
        cfg.pic_w, cfg.pic_h, cfg.nframes, cfg.q0, cfg.use_masking = geom.pic_w, geom.pic_h, nframes, int(q0), int(use_masking)
        cfg.qm_stride = pvq.OD_QM_STRIDE
        cfg.pvq_norm_lambda = float(lam)
        q4 = pvq_qm_q4 if pvq_qm_q4 is not None else np.full((3, 30), 16, np.uint8)
        for p in range(3):
            for i in range(30):
                cfg.pvq_qm_q4[p][i] = int(q4[p][i])
        cfg.qm, cfg.qm_inv = self._qm.ctypes.data, self._qm_inv.ctypes.data
        cfg.sb_row0, cfg.sb_rows = int(sb_row0), int(sb_rows)
        cfg.max_blocks_div, cfg.persist_ctas_per_sm = int(max_blocks_div), int(persist_ctas_per_sm)
        cfg.split_free = int(split_free)
        cfg.level_chains = int(level_chains)
        cfg.noref_prepass = int(noref_prepass)
        cfg.dering = int(dering)
        # dering == 2 (level search): scale of od_compute_dist and enc->dering_lambda = 0.67 * OD_PVQ_LAMBDA * q^2
        # (src/rate.c:1086; the target quantizer is this engine's q0)
        cfg.coded_quantizer = int(coded_quantizer)
        cfg.qm_is_flat = int(qm_is_flat)
        cfg.dering_lambda = float(0.67 * pvq.PVQ_LAMBDA * q0 * q0 if dering_lambda is None else dering_lambda)
        self.dering_lambda = cfg.dering_lambda
        self.dering = int(dering)
        self.kf = self.L.daala_b200_kf_create(ctypes.byref(cfg))
        if not self.kf:
            raise RuntimeError("daala_b200_kf_create failed (no CUDA device, or out of memory)")
        self.buf = Buffers()
        self._check(self.L.daala_b200_kf_device_buffers(self.kf, ctypes.byref(self.buf)), "device_buffers")
        self.sb_row0 = sb_row0
        self.sb_rows = sb_rows if sb_rows > 0 else geom.nvsb
        self.pinned = pinned
        self._host = {}
        self._io = None
        self._out = None
        self.totals = None

    def launches_per_step(self):
        return int(self.L.daala_b200_kf_launches_per_step(self.kf))

    def close(self):
        if getattr(self, "kf", None):
            self.L.daala_b200_kf_destroy(self.kf)
            self.kf = None
        for v in self._host.values():
            if isinstance(v, Pinned):
                v.free()
        self._host = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise _native.CudaError("%s failed with cudaError %d (%s)" % (
                what, rc, self.L.daala_b200_kf_error(self.kf).decode() if self.kf else ""))

    # --- host buffers --------------------------------------------------------------------------
    def _arr(self, name, shape, dtype):
        """(Re)usable host array `name`; pinned when the engine was created with pinned=True."""
        cur = self._host.get(name)
        n = int(np.prod(shape))
        if cur is not None:
            a = cur.array if isinstance(cur, Pinned) else cur
            if a.dtype == np.dtype(dtype) and a.size >= n:
                return a.reshape(-1)[:n].reshape(shape)
            if isinstance(cur, Pinned):
                cur.free()
        if self.pinned:
            cur = Pinned((max(n, 1),), dtype)
            self._host[name] = cur
            return cur.array[:n].reshape(shape)
        a = np.zeros(max(n, 1), dtype)
        self._host[name] = a
        return a[:n].reshape(shape)

    def count_blocks(self, bsize):
        t = Totals()
        g = self.geom
        b = np.ascontiguousarray(bsize, np.uint8)
        assert b.shape == (self.F,) + tuple(g.bsize_shape)
        self._check(self.L.daala_b200_kf_count_blocks(b.ctypes.data, self.F, b.strides[0], b.strides[1], g.nhsb, g.nvsb,
                                                      self.sb_row0, self.sb_rows if self.sb_rows != g.nvsb else 0,
                                                      ctypes.byref(t)), "count_blocks")
        return t

    def stage_dering_levels(self, levels):
        """levels: [F, nvsb, nhsb] uint8 (0..5), staged next to the other inputs (engines created with dering=1)."""
        g = self.geom
        a = self._arr("dlev", (self.F, g.nvsb, g.nhsb), np.uint8)
        a[...] = levels

    def stage_inputs(self, planes, bsize):
        """Copies one batch into the engine's (pinned) host input buffers.  planes: per plane an array
        [F, h, w] u8 (padded geometry); bsize: [F, nvsb*8, nhsb*8]."""
        g = self.geom
        for p in range(3):
            a = self._arr("in%d" % p, (self.F,) + g.plane_shape(p), np.uint8)
            a[...] = planes[p]
        b = self._arr("bsize", (self.F,) + tuple(g.bsize_shape), np.uint8)
        b[...] = bsize
        self.totals = self.count_blocks(b)

    def prepare_io(self, symbols=True, recon=True):
        """Builds the daala_b200_kf_io record over the staged inputs and result buffers sized for them."""
        g, t = self.geom, self.totals
        io = IO()
        for p in range(3):
            io.pixels[p] = self._arr("in%d" % p, (self.F,) + g.plane_shape(p), np.uint8).ctypes.data
        io.bsize = self._arr("bsize", (self.F,) + tuple(g.bsize_shape), np.uint8).ctypes.data
        if self.dering == 1:
            io.dering_level = self._arr("dlev", (self.F, g.nvsb, g.nhsb), np.uint8).ctypes.data
        io.totals = ctypes.pointer(t)
        out = {}
        if recon:
            for p in range(3):
                out["recon%d" % p] = self._arr("out%d" % p, (self.F,) + g.plane_shape(p), np.uint8)
                io.pixels_out[p] = out["recon%d" % p].ctypes.data
        if symbols:
            nl, nc = int(t.n_luma), int(t.n_chroma)
            out["luma_blocks"] = self._arr("lb", (nl,), pvq.BLOCK_DTYPE)
            out["chroma_blocks"] = self._arr("cb", (nc,), pvq.BLOCK_DTYPE)
            out["luma_res"] = self._arr("lr", (nl, 9, 4), np.int16)
            out["chroma_res"] = self._arr("cr", (nc, 9, 4), np.int16)
            out["luma_y16"] = self._arr("ly", (int(t.luma_coefs),), np.int16)
            out["chroma_y16"] = self._arr("cy", (int(t.chroma_coefs),), np.int16)
            out["luma_skip_diff"] = self._arr("ls", (nl,), np.float64)
            out["chroma_skip_diff"] = self._arr("cs", (nc,), np.float64)
            out["chroma_flip"] = self._arr("cf", (nc,), np.int32)
            for k in ("luma_blocks", "chroma_blocks", "luma_res", "chroma_res", "luma_y16", "chroma_y16",
                      "luma_skip_diff", "chroma_skip_diff", "chroma_flip"):
                setattr(io, k, out[k].ctypes.data)
        out["counts"] = self._arr("cnt", (32,), np.int32)
        io.counts = out["counts"].ctypes.data
        if self.dering:
            out["dering_levels"] = self._arr("dlev_out", (self.F, g.nvsb, g.nhsb), np.uint8)
            io.dering_level_out = out["dering_levels"].ctypes.data
        self._io, self._out = io, out
        self.h2d_bytes = sum(int(np.prod(g.plane_shape(p))) for p in range(3)) * self.F + int(np.prod(g.bsize_shape)) * self.F
        self.d2h_bytes = sum(v.nbytes for v in out.values())
        return out

    def submit(self):
        self._check(self.L.daala_b200_kf_submit(self.kf, ctypes.byref(self._io)), "kf_submit")

    def wait(self):
        self._check(self.L.daala_b200_kf_wait(self.kf), "kf_wait")
        return self._out

    def encode(self, planes, bsize, symbols=True, recon=True, dering_levels=None):
        """One batch end to end through the C ABI with host buffers; returns the result arrays (views of
        the engine's host buffers: copy what must survive the next call)."""
        self.stage_inputs(planes, bsize)
        if self.dering == 1:
            self.stage_dering_levels(dering_levels)
        self.prepare_io(symbols, recon)
        self.submit()
        out = self.wait()
        if int(out["counts"][CNT["error"]]):
            raise RuntimeError("keyframe engine: block capacity exceeded (max_blocks_div too large)")
        return out

    # --- device-resident use -------------------------------------------------------------------
    def run_device(self, phases=PH_ALL, graph=True):
        self._check(self.L.daala_b200_kf_run_device(self.kf, phases, 1 if graph else 0), "kf_run_device")

    def time_device(self, phases=PH_ALL, graph=True, reps=1):
        """Milliseconds (CUDA events on the engine's stream) for `reps` repetitions of the phases."""
        ms = ctypes.c_float()
        self._check(self.L.daala_b200_kf_time_device(self.kf, phases, 1 if graph else 0, reps, ctypes.byref(ms)),
                    "kf_time_device")
        return float(ms.value)

    def upload(self, planes, bsize):
        g = self.geom
        for p in range(3):
            a = np.ascontiguousarray(planes[p], np.uint8)
            assert a.shape == (self.F,) + g.plane_shape(p)
            self._check(self.L.daala_b200_device_copy(self.buf.pixels[p], a.ctypes.data, a.nbytes, 0), "upload")
        b = np.ascontiguousarray(bsize, np.uint8)
        assert b.shape == (self.F,) + tuple(g.bsize_shape)
        self._check(self.L.daala_b200_device_copy(self.buf.bsize, b.ctypes.data, b.nbytes, 0), "upload")

    def download(self, ptr, shape, dtype):
        self.wait()
        a = np.zeros(shape, dtype)
        if a.nbytes:
            self._check(self.L.daala_b200_device_copy(a.ctypes.data, ptr, a.nbytes, 1), "download")
        return a

    def counts(self):
        return self.download(self.buf.counts, (32,), np.int32)

    def coeff_plane(self, p):
        return self.download(self.buf.coeffs[p], (self.F,) + self.geom.plane_shape(p), np.int32)

    def recon_plane(self, p):
        return self.download(self.buf.pixels_out[p], (self.F,) + self.geom.plane_shape(p), np.uint8)


def band_records(blocks, res, geom, pli, frame):
    """[h/4, w/4, 9, 4] int16 array of one plane of one frame with each block's band records stored at
    its origin (the layout the oracle's recording hook uses); unwritten entries are -32768."""
    h, w = geom.plane_shape(pli)
    out = np.full((h // 4, w // 4, 9, 4), -32768, np.int16)
    sel = (blocks["pli"] == pli) & (blocks["frame"] == frame)
    b = blocks[sel]
    r = res[sel]
    nb = np.array([1, 4, 7, 9, 9])[b["bs"]]
    for band in range(9):
        m = nb > band
        out[b["y0"][m] >> 2, b["x0"][m] >> 2, band] = r[m, band]
    return out
