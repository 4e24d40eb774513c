"""Generates tests/golden/c1_p25p1_c4fm_{cc,vc}.npz: configuration C1 of BASELINE.json -- the reference's own IQ-replay
fixtures tests/fixtures/iq/p25p1_c4fm_cc.iq (cu8, 48 kS/s, 2 s of a P25 Phase 1 C4FM control channel; the reference's CLI test
DECODE_IQ_P25P1_C4FM_CC expects "NAC/CC: 140" from it, tests/CMakeLists.txt:8888-8893) and p25p1_c4fm_vc.iq (3 s of a voice
channel, NAC 293) pushed through the UNMODIFIED reference block side (full_demod), sample side (getSymbol hunt, then
getDibitSoft through the hook seam) and frame handlers (dsd_dispatch_handle_p25p1 replaying those dibits, FEC leaves
recorded through ld --wrap: oracle/ref_shim_p25.c), all compiled into oracle/_ref.  Run in the dev container:

    python tests/golden/make_c1_golden.py

Stored: the cu8 capture itself (input), CRC32s of the reference's discriminator output and hunt symbols, the reference's
dibits / reliabilities / LLRs / symbols, and one reference frame record per frame sync found in the dibit stream."""
import ctypes as C
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _harness as H  # noqa: E402

FIXTURES = {"c1_p25p1_c4fm_cc": ("/root/reference/tests/fixtures/iq/p25p1_c4fm_cc.iq", 0x140),
            "c1_p25p1_c4fm_vc": ("/root/reference/tests/fixtures/iq/p25p1_c4fm_vc.iq", 0x293)}
BP, N_HUNT = 8000, 4000  # full_demod block size; samples given to the hunt-mode launch


def widen(u8):
    """widen_u8_to_f32_bias127 (src/dsp/simd_widen.cpp:139-147)"""
    return ((u8.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).reshape(-1, 2)


def main():
    for name, (fixture, nac) in FIXTURES.items():
        one(name, fixture, nac)


def one(name, fixture, expected_nac):
    u = np.fromfile(fixture, dtype=np.uint8)
    x = widen(u)
    nb = x.shape[0] // BP
    disc = H.RefDemod("par", rate=48000, symrate=4800, profile=4).run(x, BP, nb)
    # launch structure of the GPU test, mirrored with the (pinned) oracle to learn how many symbols the hunt launch yields
    taps = {0: H.sps_fir_taps(0, 10)}
    ch = H.OracleSymChan()
    O = H.oracle_sym()
    O.oracle_sym_init(C.byref(ch), 48000, 4800, 1, 2, 1, 0, H._ptr(taps[0]), taps[0].size, 128, 1024)
    hunt = np.zeros(N_HUNT // 8, np.float32)
    cons = C.c_long(0)
    k_a = O.oracle_sym_run_symbols(C.byref(ch), 0, H._ptr(disc), N_HUNT, 12, H._ptr(hunt), hunt.size, C.byref(cons))
    # the unmodified reference: same number of hunt symbols, then getDibitSoft until fewer than 600 samples are left
    R = H.ref_sym()
    h = R.ref_sym_create(48000, 4800, H.SYNC_P25P1_POS, H.SYNC_P25P1_POS, 1, 128, 1024)
    R.ref_sym_feed(h, H._ptr(disc), disc.size)
    ref_hunt = np.zeros(k_a, np.float32)
    assert R.ref_sym_get_symbols(h, 0, k_a, 600, H._ptr(ref_hunt)) == k_a
    assert R.ref_sym_consumed(h) == cons.value and H.bits_equal(ref_hunt, hunt[:k_a])
    n = disc.size // 9
    d, r, l, s = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(2 * n, np.int16), np.zeros(n, np.float32)
    k_b = R.ref_sym_get_dibits(h, n, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
    R.ref_sym_destroy(h)
    crc = lambda a: np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))
    # frame level: the unmodified handlers on the unmodified slicer's dibits, one record per frame sync in the stream
    dd, rr, ll = d[:k_b], r[:k_b], l[:2 * k_b].reshape(-1, 2)
    sync = np.array(H.P25P1_SYNC_DIBITS, np.uint8)
    win = np.lib.stride_tricks.sliding_window_view(dd, 24)
    pos, recs = [], []
    for i in np.nonzero((win == sync).all(axis=1))[0]:
        rec = H.ref_p25_decode(dd, rr, ll, int(i) + 23, 0)
        if not rec["overrun"]:
            pos.append(int(i) + 23)
            recs.append(rec)
    frame_ref = np.array(recs, H.REF_P25_DTYPE).view(np.uint8).reshape(len(recs), -1)
    print(name, "frames", [(int(x["duid"]), int(x["nid_status"]), hex(int(x["nac"]))) for x in recs])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), iq_cu8=u, frame_pos=np.array(pos, np.int32), frame_ref=frame_ref, block_pairs=np.int32(BP), n_hunt_samples=np.int32(N_HUNT),
                        disc_crc=crc(disc), disc_head=disc[:64], hunt_count=np.int32(k_a), hunt_consumed=np.int32(cons.value),
                        hunt_crc=crc(ref_hunt), dibits=d[:k_b], reliab=r[:k_b], llr=l[:2 * k_b].reshape(-1, 2), symbols_crc=crc(s[:k_b]),
                        expected_nac=np.int32(expected_nac))
    print("hunt symbols", k_a, "consumed", cons.value, "dibits", k_b, "file", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    main()
