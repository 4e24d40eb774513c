"""Regenerates the committed golden fixtures from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libdsdneo_ref.so, built from /root/reference).  Run in the dev container:  python tests/golden/make_golden.py

  sps_fir_taps.npz   normalised matched-filter taps of the reference (impulse responses of p25/dmr/nxdn/dpmr/m17_filter)
  full_demod.npz     seeded 4FSK IQ -> reference full_demod() discriminator output (parity + avx2 builds), 3 blocks of 1024 pairs
  symbols.npz        seeded discriminator stream -> reference getDibitSoft() dibits / symbols / soft metrics (P25p1 +, DMR BS data)
  halfband.npz       seeded cf32 -> the reference's half-band cascade (simd_hb_decim2_complex chained as
                     full_demod_apply_halfband_decimation does), 3 passes, 3 blocks of 512 pairs, parity + avx2 builds
                     (python tests/golden/make_golden.py halfband  regenerates only this file)
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _harness as H  # noqa: E402


def ref_hb_cascade(variant, x, block_pairs, n_blocks, passes):
    """Chains the reference's simd_hb_decim2_complex exactly like demod_pipeline.cpp:983-1001 (31 taps, then 15)."""
    R = C.CDLL(H._ref_path(variant))
    R.simd_hb_decim2_complex.restype = C.c_int
    hb15 = (C.c_float * 15).in_dll(R, "hb_q15_taps")
    hb31 = (C.c_float * 31).in_dll(R, "hb31_q15_taps")
    hist = np.zeros((passes, 2, 30), np.float32)
    outs = []
    for b in range(n_blocks):
        cur = np.ascontiguousarray(x[b * block_pairs:(b + 1) * block_pairs]).reshape(-1).copy()
        for i in range(passes):
            dst = np.zeros(cur.size // 2 + 2, np.float32)
            n = R.simd_hb_decim2_complex(H._ptr(cur), cur.size, H._ptr(dst), H._ptr(hist[i, 0]), H._ptr(hist[i, 1]),
                                         hb31 if i == 0 else hb15, 31 if i == 0 else 15)
            cur = dst[:n].copy()
        outs.append(cur.reshape(-1, 2))
    return np.concatenate(outs)


def halfband():
    rng = np.random.default_rng(20261018)
    bp, nb, passes = 512, 3, 3
    x = rng.standard_normal((bp * nb, 2)).astype(np.float32)
    out = {"x": x, "block_pairs": bp, "n_blocks": nb, "passes": passes, "ref_par": ref_hb_cascade("par", x, bp, nb, passes)}
    if H.ref_available("avx2") and H.ref("avx2").simd_fir_get_impl_name() == b"avx2":
        out["ref_avx2"] = ref_hb_cascade("avx2", x, bp, nb, passes)
    np.savez_compressed(os.path.join(HERE, "halfband.npz"), **out)
    print("halfband.npz written")


def main():
    assert H.ref_available("par"), "build oracle/_ref first (make -C oracle ref)"
    if len(sys.argv) > 1 and sys.argv[1] == "halfband":
        return halfband()
    R = H.ref_sym()
    taps = {}
    for which in range(5):
        for sps in (5, 8, 10, 20):
            buf = np.zeros(1024, np.float32)
            n = R.ref_sps_fir_taps(which, sps, H._ptr(buf), 1024)
            taps["f%d_sps%d" % (which, sps)] = buf[:n].copy()
    np.savez_compressed(os.path.join(HERE, "sps_fir_taps.npz"), **taps)

    rng = np.random.default_rng(20261017)
    bp, nb = 1024, 3
    iq = H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=14.0)[: bp * nb]
    out = {"iq": iq, "block_pairs": bp, "n_blocks": nb}
    out["ref_par"] = H.RefDemod("par").run(iq, bp, nb)
    if H.ref_available("avx2") and H.ref("avx2").simd_fir_get_impl_name() == b"avx2":
        out["ref_avx2"] = H.RefDemod("avx2").run(iq, bp, nb)
    np.savez_compressed(os.path.join(HERE, "full_demod.npz"), **out)

    sym = {}
    for name, sync in (("p25p1_pos", H.SYNC_P25P1_POS), ("dmr_bs_data", H.SYNC_DMR_BS_DATA_POS)):
        x, _ = H.synth_disc(rng, 1500, 10, 9000.0, 1200.0, drift=800.0)
        h = R.ref_sym_create(48000, 4800, sync, sync, 1, 128, 1024)
        R.ref_sym_feed(h, H._ptr(x), x.size)
        n_max = 1500
        d = np.zeros(n_max, np.uint8); r = np.zeros(n_max, np.uint8); l = np.zeros(2 * n_max, np.int16); s = np.zeros(n_max, np.float32)
        n = R.ref_sym_get_dibits(h, n_max, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
        R.ref_sym_destroy(h)
        sym[name + "_x"] = x
        sym[name + "_dibits"] = d[:n]
        sym[name + "_rel"] = r[:n]
        sym[name + "_llr"] = l[:2 * n]
        sym[name + "_symbols"] = s[:n]
    np.savez_compressed(os.path.join(HERE, "symbols.npz"), **sym)
    halfband()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
