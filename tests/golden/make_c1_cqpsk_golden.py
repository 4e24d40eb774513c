"""Generates tests/golden/c1_p25p1_cqpsk_cc.npz: the reference's own IQ-replay fixture tests/fixtures/iq/p25p1_cqpsk_cc.iq (cu8,
48 kS/s, 2 s of a P25 Phase 1 CQPSK / LSM control channel; CLI test DECODE_IQ_P25P1_CQPSK_CC, tests/CMakeLists.txt:8900-8905)
through the UNMODIFIED reference CQPSK block side (full_demod with output_kind SYMBOL_CQPSK, ted_sps 10) and its symbol-rate
sample side (getDibitSoft with output kind 2, rf_mod 1) compiled into oracle/_ref.  Run in the dev container:

    python tests/golden/make_c1_cqpsk_golden.py
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _harness as H  # noqa: E402

FIXTURE = "/root/reference/tests/fixtures/iq/p25p1_cqpsk_cc.iq"
BP = 4800


def main():
    u = np.fromfile(FIXTURE, dtype=np.uint8)
    x = ((u.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).reshape(-1, 2)
    nb = x.shape[0] // BP
    ref = H.RefCqpsk("par", rate=48000, symrate=4800, sps=10)
    sym, counts = ref.run(x, BP, nb)
    R = H.ref_sym()
    R.ref_sym_create_cqpsk.restype = C.c_void_p
    R.ref_sym_create_cqpsk.argtypes = [C.c_int] * 7 + [C.c_double]
    h = R.ref_sym_create_cqpsk(4800, H.SYNC_P25P1_POS, H.SYNC_P25P1_POS, 128, 1024, 0, 1, -100.0)
    R.ref_sym_feed(h, H._ptr(sym), sym.size)
    n = sym.size
    d, r, l, s = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(2 * n, np.int16), np.zeros(n, np.float32)
    k = R.ref_sym_get_dibits(h, n, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
    R.ref_sym_destroy(h)
    crc = lambda a: np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))
    np.savez_compressed(os.path.join(HERE, "c1_p25p1_cqpsk_cc.npz"), iq_cu8=u, block_pairs=np.int32(BP), counts=counts, symbols_crc=crc(sym),
                        symbols_head=sym[:64], dibits=d[:k], reliab=r[:k], llr=l[:2 * k].reshape(-1, 2), expected_nac=np.int32(0xD6))
    print("symbols", sym.size, "dibits", k, "file", os.path.getsize(os.path.join(HERE, "c1_p25p1_cqpsk_cc.npz")))


if __name__ == "__main__":
    main()
