"""Writes tests/golden/sps_fir_design.npz from the UNMODIFIED reference (oracle/_ref/libdsdneo_ref_filt.so: src/dsp/dsd_filters.c
compiled in place with its static design_sps_fir() opened by oracle/ref_shim_filters.c): for each of the five matched filters
its descriptor (coefficient table, base sps, design kind, roll-off) and the designed taps at a spread of samples per symbol.
Run in the dev container: python tests/golden/make_sps_fir_design_golden.py"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SPS = [2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 25, 40, 64, 100, 160, 200]


def main():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_filt.so"))
    out = {}
    for which in range(5):
        base = np.zeros(1024, np.float32)
        bsps, kind, alpha = C.c_int(), C.c_int(), C.c_float()
        n = R.ref_filt_descriptor(which, base.ctypes.data_as(C.c_void_p), 1024, C.byref(bsps), C.byref(kind), C.byref(alpha))
        assert n > 0
        out["f%d_base" % which] = base[:n].copy()
        out["f%d_desc" % which] = np.array([bsps.value, kind.value, alpha.value], np.float64)
        for sps in SPS:
            t = np.zeros(1024, np.float32)
            m = R.ref_filt_design(which, sps, t.ctypes.data_as(C.c_void_p), 1024)
            assert m > 0
            out["f%d_sps%d" % (which, sps)] = t[:m].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sps_fir_design.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
