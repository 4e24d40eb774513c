"""CPU tests: pin the FEC oracle (oracle/oracle_fec.c) against the reference's known-answer vectors
(tests/fec/test_fec_block_codes.c, test_fec_bptc_rs.c, tests/protocol/p25/test_p25p1_soft_rs.cpp,
test_p25_12_list.c) and against the compiled reference on exhaustive / seeded random inputs."""
import ctypes as C
import itertools

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_available("par"), reason="oracle/_ref not built (no /root/reference)")

HAM = {0: (7, 4, "Hamming_7_4_decode"), 1: (12, 8, "Hamming_12_8_decode"), 2: (13, 9, "Hamming_13_9_decode"),
       3: (15, 11, "Hamming_15_11_decode"), 4: (16, 11, "Hamming_16_11_4_decode")}

# DMR BPTC(196,96) reference codeword from the reference's own test (tests/fec/test_fec_bptc_rs.c:19-27), packed
# MSB-first into hex (bit 0 is the reserved bit R(3)); payload rule: payload[i] = (17 i + i/5) & 1, R = {1,0,1}.
BPTC_KAT_HEX = "552294a452954a552944a532948a52a94ab2821ed9d2c9f70"


def bptc_kat_bits():
    v = int(BPTC_KAT_HEX, 16)  # 196 bits, MSB first
    bits = np.array([(v >> (195 - i)) & 1 for i in range(196)], dtype=np.uint8)
    return bits


def test_bptc_reference_kat():
    """The reference's fixed codeword decodes to its payload rule with 0 irrecoverable lines; 1 flipped bit is
    corrected; a 5x5 block of flips is reported (tests/fec/test_fec_bptc_rs.c:30-71)."""
    O = H.oracle_fec()
    cw = bptc_kat_bits()
    payload = np.array([((i * 17) + (i // 5)) & 1 for i in range(96)], dtype=np.uint8)
    out, r, und = np.zeros(96, np.uint8), np.zeros(3, np.uint8), C.c_int(0)
    assert O.oracle_bptc_196x96_extract(H._ptr(cw, H.u8p), H._ptr(out, H.u8p), H._ptr(r, H.u8p), C.byref(und)) == 0
    assert np.array_equal(out, payload) and list(r) == [1, 0, 1]
    one = cw.copy()
    one[1 + 4 * 15 + 5] ^= 1
    assert O.oracle_bptc_196x96_extract(H._ptr(one, H.u8p), H._ptr(out, H.u8p), H._ptr(r, H.u8p), C.byref(und)) == 0
    assert np.array_equal(out, payload)
    bad = cw.copy()
    for row in range(5):
        for col in range(5):
            bad[1 + row * 15 + col] ^= 1
    assert O.oracle_bptc_196x96_extract(H._ptr(bad, H.u8p), H._ptr(out, H.u8p), H._ptr(r, H.u8p), C.byref(und)) > 0


@needs_ref
def test_bptc_kat_constant_matches_reference_test_vector():
    """Guard the packed KAT against transcription errors: the reference decodes it to the same payload."""
    R = H.ref_fec()
    cw = bptc_kat_bits()
    out, r = np.zeros(96, np.uint8), np.zeros(3, np.uint8)
    assert R.BPTC_196x96_Extract_Data(H._ptr(cw.copy(), H.u8p), H._ptr(out, H.u8p), H._ptr(r, H.u8p)) == 0
    assert np.array_equal(out, np.array([((i * 17) + (i // 5)) & 1 for i in range(96)], dtype=np.uint8))


@needs_ref
@pytest.mark.parametrize("code", [0, 1, 2, 3, 4])
def test_hamming_exhaustive_vs_reference(code):
    """Every n-bit word (2^n <= 65536 of them) through oracle and reference: same corrected bits, decoded bits, bool."""
    O, R = H.oracle_fec(), H.ref_fec()
    n, k, name = HAM[code]
    fn = getattr(R, name)
    for w in range(1 << n):
        bits = np.array([(w >> (n - 1 - i)) & 1 for i in range(n)], dtype=np.uint8)
        a, b = bits.copy(), bits.copy()
        da, db = np.full(k, 7, np.uint8), np.full(k, 7, np.uint8)
        if code == 0:
            ra = O.oracle_hamming_decode(code, H._ptr(a, H.u8p), None)
            rb = fn(H._ptr(b, H.u8p))
        else:
            ra = O.oracle_hamming_decode(code, H._ptr(a, H.u8p), H._ptr(da, H.u8p))
            rb = fn(H._ptr(b, H.u8p), H._ptr(db, H.u8p), 1)
        assert bool(ra) == bool(rb), (w, ra, rb)
        assert np.array_equal(a, b) and np.array_equal(da, db), w


@needs_ref
def test_golay_qr_tables_vs_reference_all_syndromes():
    """All error patterns of weight <= 4 on the zero codeword plus random codewords: the table-driven decoders agree
    with the reference bit for bit (this exercises every reachable syndrome and the enumeration-order tie-breaks)."""
    O, R = H.oracle_fec(), H.ref_fec()
    for n, fo, fr in ((24, O.oracle_golay_24_12_decode, R.Golay_24_12_decode),
                      (20, O.oracle_golay_20_8_decode, R.Golay_20_8_decode),
                      (16, O.oracle_qr_16_7_6_decode, R.QR_16_7_6_decode)):
        for w in range(0, 5):
            for pos in itertools.combinations(range(n), w):
                bits = np.zeros(n, np.uint8)
                bits[list(pos)] = 1
                a, b = bits.copy(), bits.copy()
                ra, rb = fo(H._ptr(a, H.u8p)), fr(H._ptr(b, H.u8p))
                assert bool(ra) == bool(rb) and np.array_equal(a, b), (n, pos)
    rng = np.random.default_rng(5)
    for _ in range(3000):
        data = rng.integers(0, 2, 12).astype(np.uint8)
        ea, eb = np.zeros(24, np.uint8), np.zeros(24, np.uint8)
        O.oracle_golay_24_12_encode(H._ptr(data, H.u8p), H._ptr(ea, H.u8p))
        R.Golay_24_12_encode(H._ptr(data, H.u8p), H._ptr(eb, H.u8p))
        assert np.array_equal(ea, eb)
        flips = rng.choice(24, size=rng.integers(0, 5), replace=False)
        ea[flips] ^= 1
        eb[flips] ^= 1
        ra, rb = O.oracle_golay_24_12_decode(H._ptr(ea, H.u8p)), R.Golay_24_12_decode(H._ptr(eb, H.u8p))
        assert bool(ra) == bool(rb) and np.array_equal(ea, eb)
        if len(flips) <= 3:
            assert ra and np.array_equal(ea[:12], data)


def test_golay_reference_kats():
    """Properties asserted by the reference's own test (tests/fec/test_fec_block_codes.c): encode->decode round trip,
    up to 3 errors corrected for (24,12), up to 2 for (20,8) and QR(16,7,6)."""
    O = H.oracle_fec()
    data = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1, 1, 0], np.uint8)
    cw = np.zeros(24, np.uint8)
    O.oracle_golay_24_12_encode(H._ptr(data, H.u8p), H._ptr(cw, H.u8p))
    for pos in itertools.combinations(range(24), 3):
        x = cw.copy()
        x[list(pos)] ^= 1
        assert O.oracle_golay_24_12_decode(H._ptr(x, H.u8p)) and np.array_equal(x, cw)


@needs_ref
def test_bptc_random_vs_reference():
    O, R = H.oracle_fec(), H.ref_fec()
    rng = np.random.default_rng(6)
    cw = bptc_kat_bits()
    n_checked = 0
    for trial in range(4000):
        x = cw.copy()
        nflip = int(rng.integers(0, 14))
        x[rng.choice(196, size=nflip, replace=False)] ^= 1
        if trial % 3 == 0:
            x = rng.integers(0, 2, 196).astype(np.uint8)  # garbage bursts
        # interleave/deinterleave round trip
        inter = np.zeros(196, np.uint8)
        inter[(np.arange(196) * 181) % 196] = x  # transmit order: standard's 181 is the inverse of 13 mod 196
        da, db = np.zeros(196, np.uint8), np.zeros(196, np.uint8)
        O.oracle_bptc_deinterleave(H._ptr(inter, H.u8p), H._ptr(da, H.u8p))
        R.BPTCDeInterleaveDMRData(H._ptr(inter, H.u8p), H._ptr(db, H.u8p))
        assert np.array_equal(da, db) and np.array_equal(da, x)
        oa, ra, und = np.zeros(96, np.uint8), np.zeros(3, np.uint8), C.c_int(0)
        ob, rb = np.zeros(96, np.uint8), np.zeros(3, np.uint8)
        ea = O.oracle_bptc_196x96_extract(H._ptr(x.copy(), H.u8p), H._ptr(oa, H.u8p), H._ptr(ra, H.u8p), C.byref(und))
        eb = R.BPTC_196x96_Extract_Data(H._ptr(x.copy(), H.u8p), H._ptr(ob, H.u8p), H._ptr(rb, H.u8p))
        if und.value:
            continue  # the reference reads an uninitialised buffer here (first column of a pass uncorrectable)
        n_checked += 1
        assert ea == eb, (trial, ea, eb)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb), trial
    assert n_checked > 2500


@needs_ref
def test_p25_12_vs_reference_and_roundtrip():
    O, R = H.oracle_fec(), H.ref_fec()
    rng = np.random.default_rng(7)
    i16p, u32p = C.POINTER(C.c_int16), C.POINTER(C.c_uint32)
    for trial in range(400):
        dib, tx = H.p25_trellis_encode(rng)
        noise = [0.0, 60.0, 150.0, 260.0][trial % 4]
        llr = H.dibits_to_llr(tx, 200, rng, noise)
        if trial % 10 == 9:
            llr = rng.integers(-300, 300, 196).astype(np.int16)  # pure noise incl. ties
        if trial % 25 == 0:
            llr[:] = 0  # all ties: lowest predecessor / lowest final state must win
        a, b = np.zeros(12, np.uint8), np.zeros(12, np.uint8)
        ma = O.oracle_p25_12_soft_llr(llr.ctypes.data_as(i16p), H._ptr(a, H.u8p))
        mb = R.p25_12_soft_llr(None, llr.ctypes.data_as(i16p), H._ptr(b, H.u8p))
        assert ma == mb and np.array_equal(a, b), trial
        if noise == 0.0 and trial % 10 != 9 and trial % 25 != 0:
            want = np.array([(dib[4 * i] << 6) | (dib[4 * i + 1] << 4) | (dib[4 * i + 2] << 2) | dib[4 * i + 3] for i in range(12)], np.uint8)
            assert np.array_equal(a, want) and ma == 0
        # list decoder
        for maxc in (8, 3, 1):
            cb = np.zeros((8, 12), np.uint8)
            cm = np.zeros(8, np.uint32)
            na = O.oracle_p25_12_soft_llr_list(llr.ctypes.data_as(i16p), H._ptr(cb, H.u8p), cm.ctypes.data_as(u32p), maxc)
            cands = (H.P25Candidate * 8)()
            nb = R.p25_12_soft_llr_list(None, llr.ctypes.data_as(i16p), cands, maxc)
            assert na == nb, (trial, na, nb)
            for c in range(na):
                assert bytes(cb[c]) == bytes(cands[c].bytes) and int(cm[c]) == cands[c].metric, (trial, c)


# RS parity vectors pinned by the reference (tests/protocol/p25/test_p25p1_soft_rs.cpp:44-59): data symbol i = (i*7+3)&63 etc.
def _rs_words(symbols):
    return np.array([[(s >> (5 - b)) & 1 for b in range(6)] for s in symbols], dtype=np.uint8).reshape(-1)


@needs_ref
@pytest.mark.parametrize("n,k", [(36, 20), (24, 12), (24, 16)])
def test_rs63_vs_reference(n, k):
    O, R = H.oracle_fec(), H.ref_fec()
    tt = (n - k) // 2
    fn = {(36, 20): R.check_and_fix_redsolomon_36_20_17, (24, 12): R.check_and_fix_reedsolomon_24_12_13,
          (24, 16): R.check_and_fix_reedsolomon_24_16_9}[(n, k)]
    rng = np.random.default_rng(8 + n + k)
    for trial in range(1500):
        data = np.zeros(63 - 2 * tt, np.int32)
        data[:k] = rng.integers(0, 64, k)
        cw = np.zeros(63, np.int32)
        O.oracle_rs63_encode(tt, data.ctypes.data_as(H.i32p), cw.ctypes.data_as(H.i32p))
        # a valid codeword has zero syndromes for the reference decoder
        chk = np.zeros(63, np.int32)
        assert R.ref_rs63_decode(tt, cw.ctypes.data_as(H.i32p), chk.ctypes.data_as(H.i32p)) == 0 and np.array_equal(chk, cw)
        nerr = int(rng.integers(0, tt + 4))
        rx = cw.copy()
        pos = rng.choice(n, size=min(nerr, n), replace=False)  # errors inside the shortened length
        rx[pos] ^= rng.integers(1, 64, pos.size).astype(np.int32)
        par, dat = _rs_words(rx[:2 * tt]), _rs_words(rx[2 * tt:2 * tt + k])
        da, db = dat.copy(), dat.copy()
        ra = O.oracle_p25_rs_decode(n, k, H._ptr(da, H.u8p), H._ptr(par, H.u8p))
        rb = fn(H._ptr(db, H.u8p), H._ptr(par, H.u8p))
        assert ra == rb, (trial, nerr, ra, rb)
        assert np.array_equal(da, db), (trial, nerr)
        if nerr <= tt:
            assert ra == 0 and np.array_equal(da, _rs_words(cw[2 * tt:2 * tt + k]))
        # raw 63-symbol decoder incl. errors in the zero padding
        rx2 = cw.copy()
        pos2 = rng.choice(63, size=int(rng.integers(0, tt + 3)), replace=False)
        rx2[pos2] ^= rng.integers(1, 64, pos2.size).astype(np.int32)
        oa, ob = np.zeros(63, np.int32), np.zeros(63, np.int32)
        r1 = O.oracle_rs63_decode(tt, rx2.ctypes.data_as(H.i32p), oa.ctypes.data_as(H.i32p))
        r2 = R.ref_rs63_decode(tt, rx2.ctypes.data_as(H.i32p), ob.ctypes.data_as(H.i32p))
        assert r1 == r2 and np.array_equal(oa, ob), (trial, pos2)


def _bind_conv(R):
    u16p = C.POINTER(C.c_uint16)
    R.viterbi_decode.restype = C.c_uint32
    R.viterbi_decode.argtypes = [H.u8p, u16p, C.c_uint16]
    R.viterbi_decode_punctured.restype = C.c_uint32
    R.viterbi_decode_punctured.argtypes = [H.u8p, u16p, H.u8p, C.c_uint16, C.c_uint16]
    R.CNXDNConvolution_decode.argtypes = [C.c_uint8, C.c_uint8]
    R.CNXDNConvolution_decode_soft.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint8]
    R.CNXDNConvolution_chainback.argtypes = [H.u8p, C.c_uint]


def _bind_conv_oracle(O):
    u16p = C.POINTER(C.c_uint16)
    O.oracle_viterbi_k5_decode.restype = C.c_uint32
    O.oracle_viterbi_k5_decode.argtypes = [H.u8p, u16p, C.c_int]
    O.oracle_viterbi_k5_decode_punctured.restype = C.c_uint32
    O.oracle_viterbi_k5_decode_punctured.argtypes = [H.u8p, u16p, H.u8p, C.c_int, C.c_int]
    O.oracle_nxdn_conv_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, u16p, H.u8p]


def viterbi_cases(rng, n_cases):
    """(cost array uint16, len) test inputs: clean / noisy encodings of terminated messages and pure noise."""
    cases = []
    for t in range(n_cases):
        nbits = [240, 96, 40, 176][t % 4]
        msg = np.concatenate([rng.integers(0, 2, nbits), np.zeros(4, np.int64)])
        enc = H.conv_k5_encode(msg).astype(np.float64)
        soft = enc * 65535.0 + rng.standard_normal(enc.size) * [0.0, 9000.0, 20000.0, 30000.0][(t // 4) % 4]
        cost = np.clip(np.rint(soft), 0, 65535).astype(np.uint16)
        if t % 11 == 10:
            cost = rng.integers(0, 65536, cost.size).astype(np.uint16)
        if t % 13 == 12:
            cost[:] = 0x7FFF  # all erasures: every ACS is a tie
        cases.append((cost, msg))
    return cases


@needs_ref
def test_viterbi_k5_vs_reference():
    O, R = H.oracle_fec(), H.ref_fec()
    _bind_conv(R)
    _bind_conv_oracle(O)
    rng = np.random.default_rng(21)
    u16p = C.POINTER(C.c_uint16)
    for t, (cost, msg) in enumerate(viterbi_cases(rng, 200)):
        n = cost.size
        a, b = np.full(64, 0x55, np.uint8), np.full(64, 0x55, np.uint8)
        ma = O.oracle_viterbi_k5_decode(H._ptr(a, H.u8p), cost.ctypes.data_as(u16p), n)
        mb = R.viterbi_decode(H._ptr(b, H.u8p), cost.ctypes.data_as(u16p), n)
        assert ma == mb and np.array_equal(a, b), t
        if t % 16 < 4 and t % 11 != 10 and t % 13 != 12:  # clean: decodes back to the message (first message bit lands at bit 8)
            nb = n // 2
            bits = np.unpackbits(a)[8:8 + nb]
            assert np.array_equal(bits[: msg.size - 4], msg[: msg.size - 4]) and ma == 0
    # punctured: M17-style 2-of-... pattern and "no puncture"
    for pattern in ([1, 1, 1, 0], [1] * 8, [1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1]):
        punct = np.array(pattern, np.uint8)
        for t in range(40):
            in_len = int(rng.integers(40, 300)) & ~1
            cost = rng.integers(0, 65536, in_len).astype(np.uint16)
            a, b = np.zeros(80, np.uint8), np.zeros(80, np.uint8)
            ma = O.oracle_viterbi_k5_decode_punctured(H._ptr(a, H.u8p), cost.ctypes.data_as(u16p), H._ptr(punct, H.u8p), in_len, punct.size)
            mb = R.viterbi_decode_punctured(H._ptr(b, H.u8p), cost.ctypes.data_as(u16p), H._ptr(punct, H.u8p), in_len, punct.size)
            assert ma == mb and np.array_equal(a, b), (pattern, t)


@needs_ref
def test_nxdn_convolution_vs_reference():
    O, R = H.oracle_fec(), H.ref_fec()
    _bind_conv(R)
    _bind_conv_oracle(O)
    rng = np.random.default_rng(22)
    u16p = C.POINTER(C.c_uint16)
    R.CNXDNConvolution_init()
    metrics = np.zeros(32, np.uint16)  # oracle-side carried ping-pong metric arrays; the reference carries its own statics
    for t in range(120):
        n_steps = int(rng.integers(20, 300))
        n_out = int(rng.integers(1, n_steps + 1)) if t % 3 else n_steps - 4
        n_out = max(1, n_out)
        msg = rng.integers(0, 2, n_steps)
        enc = H.conv_k5_encode(msg)
        sym = (enc * 2).astype(np.uint8)
        flip = rng.random(sym.size) < [0.0, 0.03, 0.1, 0.5][t % 4]
        sym[flip] = 2 - sym[flip]
        if t % 7 == 6:
            sym[rng.random(sym.size) < 0.1] = 1  # erasure value used by the NXDN depuncturer
        soft = (t % 2 == 1)
        rel = rng.integers(0, 256, sym.size).astype(np.uint8)
        R.CNXDNConvolution_start()
        for i in range(n_steps):
            if soft:
                R.CNXDNConvolution_decode_soft(int(sym[2 * i]), int(sym[2 * i + 1]), int(rel[2 * i]), int(rel[2 * i + 1]))
            else:
                R.CNXDNConvolution_decode(int(sym[2 * i]), int(sym[2 * i + 1]))
        b = np.full(40, 0xA5, np.uint8)
        R.CNXDNConvolution_chainback(H._ptr(b, H.u8p), n_out)
        a = np.full(40, 0xA5, np.uint8)
        O.oracle_nxdn_conv_decode(H._ptr(sym, H.u8p), H._ptr(rel, H.u8p) if soft else None, n_steps, n_out,
                                  metrics.ctypes.data_as(u16p), H._ptr(a, H.u8p))
        assert np.array_equal(a, b), t
