"""dsdneo_b200_sps_fir_design (host C) against the UNMODIFIED design_sps_fir() of the reference (src/dsp/dsd_filters.c:94-170,
opened by oracle/ref_shim_filters.c -> oracle/_ref/libdsdneo_ref_filt.so) and against the committed golden designs
(tests/golden/sps_fir_design.npz, made from the same library): tap counts and every tap bit for bit, for the five matched
filters at every samples-per-symbol from 2 to 210, and the taps the symbolizer tests use (the reference filters' impulse
responses) are these designs."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import _harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _pkg():
    import __graft_entry__ as g
    return g.load_package()


def _ref():
    path = os.path.join(H.REF_DIR, "libdsdneo_ref_filt.so")
    return C.CDLL(path) if os.path.exists(path) else None


def _golden():
    return np.load(os.path.join(H.GOLDEN_DIR, "sps_fir_design.npz"))


def _descriptor(z, which):
    bsps, kind, alpha = z["f%d_desc" % which]
    return z["f%d_base" % which], int(bsps), int(kind), float(np.float32(alpha))


def test_design_equals_committed_golden():
    b200, z = _pkg(), _golden()
    n = 0
    for which in range(5):
        base, bsps, kind, alpha = _descriptor(z, which)
        for key in z.files:
            if key.startswith("f%d_sps" % which):
                sps = int(key.split("sps")[1])
                ours = b200.sps_fir_design(kind, base, bsps, alpha, sps)
                assert ours.size == z[key].size, (which, sps)
                assert np.array_equal(ours.view(np.uint32), z[key].view(np.uint32)), (which, sps)
                n += 1
    assert n == 80


def test_design_equals_unmodified_reference_at_every_sps():
    R = _ref()
    if R is None:
        pytest.skip("oracle/_ref/libdsdneo_ref_filt.so not built (reference tree absent)")
    b200 = _pkg()
    for which in range(5):
        base = np.zeros(1024, np.float32)
        bsps, kind, alpha = C.c_int(), C.c_int(), C.c_float()
        n = R.ref_filt_descriptor(which, base.ctypes.data_as(C.c_void_p), 1024, C.byref(bsps), C.byref(kind), C.byref(alpha))
        base = base[:n]
        for sps in range(2, 211):
            t = np.zeros(1024, np.float32)
            m = R.ref_filt_design(which, sps, t.ctypes.data_as(C.c_void_p), 1024)
            ours = b200.sps_fir_design(kind.value, base, bsps.value, alpha.value, sps)
            assert ours.size == m, (which, sps)
            assert np.array_equal(ours.view(np.uint32), t[:m].view(np.uint32)), (which, sps)
    # what the reference refuses, and what does not fit
    assert R.ref_filt_design(0, 1, t.ctypes.data_as(C.c_void_p), 1024) == 0
    with pytest.raises(Exception):
        b200.sps_fir_design(0, base, 10, 0.0, 1)


def test_designs_are_the_filters_the_symbolizer_tests_use():
    """The taps the symbolizer / receive-bank tests pass in (impulse responses of the reference's p25_filter / dmr_filter ...)
    are these designs (an impulse response drops trailing zero taps and cannot tell -0 from +0)."""
    b200, z = _pkg(), _golden()
    for which, sps in ((0, 10), (1, 10), (2, 20), (3, 20), (4, 10), (0, 8), (1, 5)):
        base, bsps, kind, alpha = _descriptor(z, which)
        ours = b200.sps_fir_design(kind, base, bsps, alpha, sps)
        used = H.sps_fir_taps(which, sps) if (H.ref_sym() is not None or (which, sps) in ((0, 10), (1, 10))) else None
        if used is None:
            continue
        k = used.size
        assert k <= ours.size and not np.any(ours[: ours.size - k]), (which, sps)
        assert np.array_equal(ours[ours.size - k:], used), (which, sps)


def test_rejects_and_limits():
    b200 = _pkg()
    base = np.ones(91, np.float32)
    for bad in (dict(sps=1), dict(sps=0), dict(sps=-3)):
        with pytest.raises(Exception):
            b200.sps_fir_design(0, base, 10, 0.0, bad["sps"])
    with pytest.raises(Exception):
        b200.sps_fir_design(2, base, 10, 0.0, 10)  # unknown design kind
    with pytest.raises(Exception):
        b200.sps_fir_design(0, base, 0, 0.0, 10)  # base sps 0
    # longest design: the reference caps at 1023 taps
    t = b200.sps_fir_design(1, base, 10, 0.7, 400)
    assert t.size == 1023 and abs(float(t.astype(np.float64).sum()) - 1.0) < 1e-5
