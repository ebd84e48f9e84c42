"""The plain-C port (oracle/port_*.c) against tests/golden/reference_vectors.npz -- outputs of the real
reference build recorded by tests/golden/make_golden.py.  Unlike tests/test_oracle_*.py these need no
oracle/_ref: the pin travels with the repository."""
import ctypes
import os

import numpy as np
import pytest

from tests import pvq_cases
from tests.golden import make_golden
from tests.oracle_lib import addr

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(PATH)


def test_transforms_and_filters_match_recorded_reference_outputs(port, gold):
    for ln in (2, 3, 4, 5, 6):
        n = 1 << ln
        for x, y in zip(gold["dct_x_%d" % ln], gold["dct_y_%d" % ln]):
            got = np.zeros((n, n), np.int32)
            port.port_bin_fdct2d(ln, addr(got), n, addr(np.ascontiguousarray(x)), n)
            assert np.array_equal(got, y)
            back = np.zeros((n, n), np.int32)
            port.port_bin_idct2d(ln, addr(back), n, addr(got), n)
            assert np.array_equal(back, x)
    for ln in (1, 2, 3, 4, 5, 6):
        n = 1 << ln
        got = np.zeros((n, n), np.int32)
        x = np.ascontiguousarray(gold["haar_x_%d" % ln])   # keep alive across the call
        port.port_haar(addr(got), n, addr(x), n, ln)
        assert np.array_equal(got, gold["haar_y_%d" % ln])
    for n in (4, 8, 16, 32):
        x = gold["filt_x_%d" % n]
        for i in range(len(x)):
            pre = np.zeros(n, np.int32)
            post = np.zeros(n, np.int32)
            if n == 4:
                port.port_pre_filter4(addr(pre), addr(np.ascontiguousarray(x[i])))
                port.port_post_filter4(addr(post), addr(np.ascontiguousarray(x[i])))
            else:
                port.port_pre_filter_n(n, addr(pre), addr(np.ascontiguousarray(x[i])))
                port.port_post_filter_n(n, addr(post), addr(np.ascontiguousarray(x[i])))
            assert np.array_equal(pre, gold["filt_pre_%d" % n][i])
            assert np.array_equal(post, gold["filt_post_%d" % n][i])


def test_pvq_theta_matches_recorded_reference_outputs(port, gold):
    qm, qm_inv = gold["qm"], gold["qm_inv"]
    rows = gold["theta_rows"]
    xo = yo = oo = 0
    coded = 0
    for i, (n, key, pli, beta, q0, qm_off, gain, itheta, max_theta, k, ny) in enumerate(rows.tolist()):
        c = dict(n=n, is_keyframe=key, pli=pli, beta=beta, q0=q0, qm_off=qm_off, lam=float(gold["theta_lam"][i]),
                 x0=gold["theta_x"][xo:xo + n], r0=gold["theta_r"][xo:xo + n])
        b = pvq_cases.run_theta(port, "port", c, qm, qm_inv)
        assert (b["gain"], b["itheta"], b["max_theta"], b["k"]) == (gain, itheta, max_theta, k), i
        assert np.array_equal(b["y"], gold["theta_y"][yo:yo + ny]), i
        assert np.array_equal(b["out"], gold["theta_out"][oo:oo + n]), i
        assert b["skip_diff"] == gold["theta_skip_diff"][i], i
        xo += n
        yo += ny
        oo += n
        coded += k > 0
    assert len(rows) > 250 and coded > 80


def test_motion_compensation_matches_recorded_reference_outputs(port, gold):
    img = make_golden.mc_image()
    w = img.shape[1]
    for (lx, ly, mvx, mvy, x0, y0), want in zip(gold["mc1_rows"].tolist(), gold["mc1_crc"].tolist()):
        a = np.zeros((1 << lx) * (1 << ly), np.uint8)
        port.port_mc_predict1fmv8(addr(a), addr(img, y0 * w + x0), w, mvx, mvy, lx, ly)
        assert make_golden.crc(a) == want, (lx, ly, mvx, mvy)
    I4 = ctypes.c_int32 * 4
    for row, want in zip(gold["obmc_rows"].tolist(), gold["obmc_crc"].tolist()):
        ln, oc, s = row[:3]
        n = 1 << ln
        a = np.zeros((n, n), np.uint8)
        port.port_mc_predict(addr(a), n, addr(img, 45 * w + 50), w, I4(*row[3:7]), I4(*row[7:11]), oc, s, ln, ln)
        assert make_golden.crc(a) == want, row
    for ln, sad, satd in gold["sad_rows"].tolist():
        n = 1 << ln
        a, b = np.ascontiguousarray(gold["sad_a_%d" % ln]), np.ascontiguousarray(gold["sad_b_%d" % ln])
        assert port.port_mc_compute_sad8(addr(a), n, addr(b), n, n, n) == sad
        assert port.port_mc_compute_satd8(ln, addr(a), n, addr(b), n) == satd


def test_frame_drivers_match_recorded_reference_plane_checksums(port, gold):
    got = make_golden.frame_crcs(port, "port", gold["qm"], gold["qm_inv"])
    want = dict(zip(gold["frame_keys"].tolist(), gold["frame_crc"].tolist()))
    assert got == want


def test_dering_port_matches_recorded_reference_outputs(port, gold):
    rows, crcs = make_golden.dering_cases(port.port_dering)
    assert rows == gold["dering_rows"].tolist()
    assert crcs == gold["dering_crc"].tolist()


def test_frame_drivers_with_the_reference_encoders_block_sizes(port, gold):
    """Same picture, block sizes decided by the whole reference encoder (recorded in the fixture)."""
    got = make_golden.real_map_crcs(port, "port", gold["qm"], gold["qm_inv"], gold["real_bsize"])
    assert got == dict(zip(gold["real_keys"].tolist(), gold["real_crc"].tolist()))
    assert gold["real_bsize"].max() <= 4 and gold["real_packet"][0] > 500
