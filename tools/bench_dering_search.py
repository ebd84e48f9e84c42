"""Times the deringing level search of one 4K luma plane: daala_b200_dering_search (device filters and
distortions + host decision, scratch allocation included) next to the reference's loop on one host core
(oracle/ref_hooks_encode.c::oracle_ref_dering_search), and checks the levels agree.  Prints one JSON line.
Not part of bench.py: the search is host-driven and outside the engine's graph."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import oracle_lib, test_dering_search as T   # noqa: E402
from tests.oracle_lib import addr                        # noqa: E402


def main():
    ref = oracle_lib.load_ref()
    L = T.lib_decide()
    nhsb, nvsb, q = 60, 34, 72
    rng = np.random.default_rng(1)
    src, ctmp = T.synth_pair(rng, nhsb, nvsb, 1.0)
    lam = T.DERING_LAMBDA_SCALE * q * q
    cdf0 = np.zeros((11, 6), np.uint16)
    L.daala_b200_dering_cdf_init(addr(cdf0), None)
    cdf_ref = cdf0.copy()
    t0 = time.perf_counter()
    lv_ref, _ = T.ref_search(ref, src, ctmp, nhsb, nvsb, q, 1, 1, lam, None, cdf_ref)
    t_ref = time.perf_counter() - t0

    class P(ctypes.Structure):
        _fields_ = [("etmp", ctypes.c_void_p), ("src", ctypes.c_void_p), ("bskip", ctypes.c_void_p),
                    ("etmp_stride", ctypes.c_int), ("src_stride", ctypes.c_int), ("skip_stride", ctypes.c_int),
                    ("nhsb", ctypes.c_int), ("nvsb", ctypes.c_int), ("quantizer", ctypes.c_int),
                    ("coded_quantizer", ctypes.c_int), ("qm_is_flat", ctypes.c_int), ("use_activity_masking", ctypes.c_int),
                    ("is_keyframe", ctypes.c_int), ("dering_lambda", ctypes.c_double)]
    L.daala_b200_dering_search.argtypes = [ctypes.POINTER(P), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]
    d_etmp = torch.from_numpy(ctmp.astype(np.int16)).cuda()
    d_src = torch.from_numpy(src).cuda()
    prm = P(d_etmp.data_ptr(), d_src.data_ptr(), None, nhsb * 64, nhsb * 64, 0, nhsb, nvsb, q, q, 0, 1, 1, lam)
    lv = np.zeros(nhsb * nvsb, np.uint8)
    times = []
    for it in range(6):
        cdf = cdf0.copy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = L.daala_b200_dering_search(ctypes.byref(prm), addr(cdf), 128, addr(lv), None, None)
        times.append(time.perf_counter() - t0)
        assert r == 0
    print(json.dumps({"what": "dering level search, one 3840x2176 luma plane, q=72", "superblocks": nhsb * nvsb,
                      "ours_ms": round(1e3 * min(times[1:]), 3), "ours_ms_all": [round(1e3 * t, 3) for t in times],
                      "reference_one_core_ms": round(1e3 * t_ref, 1), "levels_equal": bool(np.array_equal(lv, lv_ref)),
                      "level_histogram": np.bincount(lv, minlength=6).tolist()}))


if __name__ == "__main__":
    main()
