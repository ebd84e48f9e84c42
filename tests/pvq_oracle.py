"""Block-level PVQ oracle (test infrastructure): the band loop and CfL flip of
od_pvq_encode (src/pvq_encoder.c:847-880) restated in Python around the
oracle's pvq_theta (the reference's static function through
oracle/ref_hooks_pvq.c, or the plain-C port)."""
import numpy as np

from tests import pvq_cases
from tests.oracle_lib import addr

BAND_EDGES = [1, 16, 24, 32, 64, 96, 128, 256, 384, 512]
NBANDS = {0: 1, 1: 4, 2: 7, 3: 9, 4: 9}
OD_QM_STRIDE = 5456


def qm_offset(bs, xdec):
    return xdec * OD_QM_STRIDE + (((1 << (2 * bs)) - 1) << 4) // 3  # od_qm_offset, src/pvq.c:306


def coding_order(lib, prefix, plane, x0, y0, n):
    """od_raster_to_coding_order of the n x n block at (x0, y0) of `plane`."""
    dst = np.zeros(n * n, np.int32)
    stride = plane.shape[1]
    fn = lib.od_raster_to_coding_order if prefix == "ref" else lib.port_raster_to_coding_order
    fn(addr(dst), n, addr(plane, y0 * stride + x0), stride)
    return dst


def from_coding_order(lib, prefix, plane, x0, y0, n, vec):
    stride = plane.shape[1]
    fn = lib.od_coding_order_to_raster if prefix == "ref" else lib.port_coding_order_to_raster
    fn(addr(plane, y0 * stride + x0), stride, addr(np.ascontiguousarray(vec)), n)


def block(lib, prefix, dvec, pvec, bs, pli, xdec, q0, is_keyframe, use_masking, lam, qm, qm_inv, pvq_qm_q4):
    n = 4 << bs
    ref = pvec.copy()
    flip = 0
    off = qm_offset(bs, xdec)
    if pli != 0 and is_keyframe:
        xy = 0
        for i in range(1, 16):
            rq = int(ref[i]) * int(qm[off + i])
            inq = int(dvec[i]) * int(qm[off + i])
            xy += (rq * inq) >> 30
        if xy < 0:
            flip = 1
            ref[1:BAND_EDGES[NBANDS[bs]]] *= -1
    out = np.zeros(n * n, np.int32)
    y = np.zeros(n * n, np.int32)
    bands = []
    skip_diff = 0.0
    for band in range(NBANDS[bs]):
        a, b = BAND_EDGES[band], BAND_EDGES[band + 1]
        idx = bs * (bs + 1) + (band + 1) - (band + 1) // 3
        q = max(1, (q0 * int(pvq_qm_q4[pli][idx])) >> 4)
        beta = 6144 if (use_masking and pli == 0 and bs > 0) else 4096
        c = dict(n=b - a, x0=dvec[a:b], r0=ref[a:b], q0=q, beta=beta, is_keyframe=is_keyframe, pli=pli,
                 qm_off=off + a, lam=lam)
        r = pvq_cases.run_theta(lib, prefix, c, qm, qm_inv)
        out[a:b] = r["out"]
        y[a:a + len(r["y"])] = r["y"]
        skip_diff += r["skip_diff"]
        bands.append(r)
    return dict(flip=flip, ref=ref, out=out, y=y, bands=bands, skip_diff=skip_diff)
