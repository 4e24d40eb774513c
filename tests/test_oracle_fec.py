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


def _near_valid_128x77(rng, n_flips):
    """A matrix whose 7 Hamming rows are valid codewords and whose parity row is consistent, with n_flips bit errors."""
    O = H.oracle_fec()
    m = rng.integers(0, 2, 128).astype(np.uint8)
    for i in range(7):
        line, dec = m[16 * i:16 * i + 16].copy(), np.zeros(11, np.uint8)
        while not O.oracle_hamming_decode(4, H._ptr(line, H.u8p), H._ptr(dec, H.u8p)):
            line = rng.integers(0, 2, 16).astype(np.uint8)
        m[16 * i:16 * i + 16] = line
    m[112:128] = m[:112].reshape(7, 16).sum(axis=0) % 2
    for e in rng.integers(0, 128, n_flips):
        m[e] ^= 1
    return m


@needs_ref
def test_bptc_128x77_and_16x2_match_reference():
    """oracle == BPTC_128x77_Extract_Data / BPTC_16x2_Extract_Data (src/fec/bptc.c:167-333) on clean, lightly corrupted
    and random inputs with dirty upper bits; inputs on which the reference reads an uninitialised buffer are excluded."""
    O, R = H.oracle_fec(), H.ref("par")
    R.InitAllFecFunction()
    R.BPTC_128x77_Extract_Data.restype = C.c_uint32
    R.BPTC_16x2_Extract_Data.restype = C.c_uint32
    O.oracle_bptc_128x77_extract.restype = C.c_uint
    O.oracle_bptc_16x2_extract.restype = C.c_uint
    rng = np.random.default_rng(5)
    compared = 0
    for t in range(3000):
        x = _near_valid_128x77(rng, int(rng.integers(0, 5))) if t % 3 else rng.integers(0, 2, 128).astype(np.uint8)
        x = x | (rng.integers(0, 128, 128) * 2).astype(np.uint8)
        a, b, oa, ob, und = x.copy(), x.copy(), np.zeros(77, np.uint8), np.zeros(77, np.uint8), C.c_int(0)
        rb = O.oracle_bptc_128x77_extract(H._ptr(b, H.u8p), H._ptr(ob, H.u8p), C.byref(und))
        if und.value:
            continue
        ra = R.BPTC_128x77_Extract_Data(H._ptr(a, H.u8p), H._ptr(oa, H.u8p))
        assert ra == rb and np.array_equal(oa, ob)
        compared += 1
    assert compared > 1500
    compared = 0
    for t in range(3000):
        x = rng.integers(0, 256, 32).astype(np.uint8)
        for odd in (0, 1):
            a, b, oa, ob, und = x.copy(), x.copy(), np.zeros(32, np.uint8), np.zeros(32, np.uint8), C.c_int(0)
            rb = O.oracle_bptc_16x2_extract(H._ptr(b, H.u8p), H._ptr(ob, H.u8p), C.c_uint(odd), C.byref(und))
            if und.value:
                continue
            ra = R.BPTC_16x2_Extract_Data(H._ptr(a, H.u8p), H._ptr(oa, H.u8p), C.c_uint32(odd))
            assert ra == rb and np.array_equal(oa, ob)
            compared += 1
    assert compared > 1500


def test_bptc_128x77_reference_kat():
    """The reference's own known-answer matrix (tests/fec/test_fec_bptc_rs.c:74-125): zero errors, payload = 0xA5 pattern
    LSB first, CRC bits zero; a single flipped bit is corrected."""
    O = H.oracle_fec()
    O.oracle_bptc_128x77_extract.restype = C.c_uint
    ref = np.array([[1, 0, 1, 0, 0, 1, 0, 1, 1, 0, 1, 1, 0, 0, 0, 1], [0, 0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1],
                    [0, 1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0], [1, 0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 1, 0, 1, 1, 0],
                    [1, 0, 0, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0], [0, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0, 1, 1, 1, 0],
                    [0, 1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0], [1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 0, 1, 0, 0]], np.uint8)
    want = np.array([(0xA5 >> (i % 8)) & 1 for i in range(72)] + [0] * 5, np.uint8)
    for flip in (None, (1, 3)):
        m = ref.copy()
        if flip:
            m[flip] ^= 1
        out = np.zeros(77, np.uint8)
        assert O.oracle_bptc_128x77_extract(H._ptr(np.ascontiguousarray(m.reshape(-1)), H.u8p), H._ptr(out, H.u8p), None) == 0
        assert np.array_equal(out, want)


RS_SHAPES = {(36, 20): 0, (24, 12): 1, (24, 16): 2}


def make_rs_soft_cases(rng, n, k, count, dirty=True):
    """Random shortened RS words with 0..2t+2 symbol errors and per-symbol reliabilities that mostly flag the errors."""
    O = H.oracle_fec()
    tt = (n - k) // 2
    dat, par, rel_d, rel_p, ers, n_ers, truth = [], [], [], [], [], [], []
    for trial in range(count):
        data = np.zeros(63 - 2 * tt, np.int32)
        data[:k] = rng.integers(0, 64, k)
        cw = np.zeros(63, np.int32)
        O.oracle_rs63_encode(tt, data.ctypes.data_as(H.i32p), cw.ctypes.data_as(H.i32p))
        rx = cw.copy()
        pos = rng.choice(n, int(rng.integers(0, 2 * tt + 3)), replace=False)
        rx[pos] ^= rng.integers(1, 64, pos.size).astype(np.int32)
        rel = np.full(n, 255, np.uint8)
        for q in pos:
            if rng.random() < 0.8:
                rel[q] = rng.integers(0, 120)
        for q in rng.choice(n, int(rng.integers(0, 4)), replace=False):
            rel[q] = rng.integers(0, 255)
        d = _rs_words(rx[2 * tt:n])
        if dirty and trial % 5 == 0:
            d = (d * rng.integers(1, 120, d.size)).astype(np.uint8)  # callers pass any non-zero byte as a 1
        kk = int(rng.integers(1, 2 * tt + 1))
        e = (list(pos[:kk]) + [q for q in rng.permutation(n) if q not in pos])[:kk]
        dat.append(d); par.append(_rs_words(rx[:2 * tt])); rel_d.append(rel[2 * tt:]); rel_p.append(rel[:2 * tt])
        ers.append(np.array(e + [0] * (16 - kk), np.int32)); n_ers.append(kk); truth.append(_rs_words(cw[2 * tt:n]))
    return (np.array(dat), np.array(par), np.array(rel_d), np.array(rel_p), np.array(ers), np.array(n_ers, np.int32), np.array(truth))


def oracle_rs_erasures(n, k, d, p, ers, n_er):
    """check_and_fix_*_soft semantics from the oracle parts: hard decode (writes back), then one erasure decode."""
    O = H.oracle_fec()
    tt = (n - k) // 2
    d = d.copy()
    sym = lambda bits: [int("".join("1" if b else "0" for b in bits[6 * i:6 * i + 6]), 2) for i in range(len(bits) // 6)]
    inn = np.zeros(63, np.int32)
    inn[:2 * tt] = sym(p)
    inn[2 * tt:n] = sym(d)
    rc = O.oracle_p25_rs_decode(n, k, H._ptr(d, H.u8p), H._ptr(p, H.u8p))
    if rc != 0:
        out = np.zeros(63, np.int32)
        e = np.ascontiguousarray(ers[:n_er], np.int32)
        rc = O.oracle_rs63_decode_with_erasures(tt, inn.ctypes.data_as(H.i32p), out.ctypes.data_as(H.i32p), e.ctypes.data_as(H.i32p), int(n_er))
        if rc == 0:
            d = _rs_words(out[2 * tt:n])
    return d, rc


@needs_ref
@pytest.mark.parametrize("n,k", [(36, 20), (24, 12), (24, 16)])
def test_rs_soft_vs_reference(n, k):
    """Erasure decoding and the ranked-erasure soft wrappers == the reference (check_and_fix_*_soft,
    p25p1_rs_*_soft_reliability, p25p1_build_rs_ranked_erasures), return codes and corrected bits."""
    O, R = H.oracle_fec(), H.ref_fec()
    f_ers = {(36, 20): R.check_and_fix_redsolomon_36_20_17_soft, (24, 12): R.check_and_fix_reedsolomon_24_12_13_soft,
             (24, 16): R.check_and_fix_reedsolomon_24_16_9_soft}[(n, k)]
    f_soft = {(36, 20): R.p25p1_rs_36_20_17_soft_reliability, (24, 16): R.p25p1_rs_24_16_9_soft_reliability}.get((n, k))
    rng = np.random.default_rng(70 + n + k)
    dat, par, rel_d, rel_p, ers, n_ers, truth = make_rs_soft_cases(rng, n, k, 1200)
    wins = 0
    for i in range(dat.shape[0]):
        a = dat[i].copy()
        e = np.ascontiguousarray(ers[i, :n_ers[i]])
        ra = f_ers(H._ptr(a, H.u8p), H._ptr(par[i], H.u8p), e.ctypes.data_as(H.i32p), int(n_ers[i]))
        b, rb = oracle_rs_erasures(n, k, dat[i], par[i], ers[i], n_ers[i])
        assert ra == rb and np.array_equal(a, b), i
        if f_soft is not None:
            a, b = dat[i].copy(), dat[i].copy()
            ra = f_soft(H._ptr(a, H.u8p), H._ptr(par[i], H.u8p), H._ptr(rel_d[i].copy(), H.u8p), H._ptr(rel_p[i].copy(), H.u8p))
            rb = O.oracle_p25_rs_soft_reliability(n, k, H._ptr(b, H.u8p), H._ptr(par[i], H.u8p), H._ptr(rel_d[i].copy(), H.u8p),
                                                  H._ptr(rel_p[i].copy(), H.u8p), 64)
            assert ra == rb and np.array_equal(a, b), i
            wins += ra == 0 and np.array_equal(a, truth[i])
        # ranking helper
        want, got = np.zeros(16, np.int32), np.zeros(16, np.int32)
        tt = (n - k) // 2
        nw = R.p25p1_build_rs_ranked_erasures(H._ptr(rel_d[i].copy(), H.u8p), k, H._ptr(rel_p[i].copy(), H.u8p), n - k, tt,
                                              want.ctypes.data_as(H.i32p), 2 * tt)
        ng = O.oracle_p25_rs_ranked_erasures(H._ptr(rel_d[i].copy(), H.u8p), k, H._ptr(rel_p[i].copy(), H.u8p), n - k, tt, 64,
                                             got.ctypes.data_as(H.i32p), 2 * tt)
        assert nw == ng and np.array_equal(want[:nw], got[:ng])
    if f_soft is not None:
        assert wins > 300


def test_rs_soft_reference_kats():
    """The reference's pinned vectors (tests/protocol/p25/test_p25p1_soft_rs.cpp:44-59,83-160): parity of fill_data(seed)
    words, ten erased data symbols corrected by RS(36,20,17) where the hard decoder fails, ranked-erasure mapping."""
    O = H.oracle_fec()
    fill = lambda count, seed: np.array([(seed + i * 7) & 0x3F for i in range(count)], np.int32)
    kats = [(8, 20, 3, [0x1F, 0x01, 0x38, 0x24, 0x0C, 0x2B, 0x29, 0x35, 0x1C, 0x11, 0x2D, 0x0A, 0x11, 0x3D, 0x12, 0x32]),
            (8, 20, 29, [0x19, 0x3C, 0x0A, 0x2A, 0x33, 0x2F, 0x23, 0x23, 0x08, 0x0C, 0x0F, 0x16, 0x12, 0x04, 0x17, 0x01]),
            (6, 12, 11, [0x05, 0x08, 0x00, 0x2B, 0x10, 0x32, 0x1F, 0x0D, 0x03, 0x2F, 0x27, 0x22]),
            (4, 16, 23, [0x25, 0x2A, 0x2C, 0x09, 0x0A, 0x0F, 0x06, 0x29])]
    for tt, k, seed, parity in kats:
        data = np.zeros(63 - 2 * tt, np.int32)
        data[:k] = fill(k, seed)
        cw = np.zeros(63, np.int32)
        O.oracle_rs63_encode(tt, data.ctypes.data_as(H.i32p), cw.ctypes.data_as(H.i32p))
        assert list(cw[:2 * tt]) == parity
    # HDU: 10 corrupted data symbols, all flagged as erasures (positions 16..25)
    data = fill(20, 3)
    bad = data.copy()
    for i in range(10):
        bad[i] ^= (0x21 + i) & 0x3F
    par = _rs_words(np.array(kats[0][3]))
    hard = _rs_words(bad)
    assert O.oracle_p25_rs_decode(36, 20, H._ptr(hard, H.u8p), H._ptr(par, H.u8p)) == 1
    fixed, rc = oracle_rs_erasures(36, 20, _rs_words(bad), par, np.arange(16, 26, dtype=np.int32), 10)
    assert rc == 0 and np.array_equal(fixed, _rs_words(data))
    er = np.zeros(16, np.int32)
    d, p = np.full(20, 255, np.uint8), np.full(16, 255, np.uint8)
    p[3], d[4] = 10, 20
    assert O.oracle_p25_rs_ranked_erasures(H._ptr(d, H.u8p), 20, H._ptr(p, H.u8p), 16, 2, 64, er.ctypes.data_as(H.i32p), 16) == 2
    assert list(er[:2]) == [3, 20]


# ---------------------------------------------------------------- P25 word codes: Golay(24,6/12), Hamming(10,6,3), BCH(63,16,11)

def bch_63_16_encode(data16):
    """Systematic BCH(63,16,11) codeword for 16 data bits (MSB first): input bit i is coefficient 62 - i, the generator is the
    LCM of the minimal polynomials of alpha^1..alpha^22 over GF(64) (x^6 + x + 1)."""
    exp, v = [], 1
    for _ in range(63):
        exp.append(v)
        v <<= 1
        if v & 0x40:
            v ^= 0x43
    log = {e: i for i, e in enumerate(exp)}
    mul = lambda a, b: 0 if a == 0 or b == 0 else exp[(log[a] + log[b]) % 63]
    roots, seen = [], set()
    for i in range(1, 23):
        c = i
        while c not in seen:
            seen.add(c)
            roots.append(c)
            c = (2 * c) % 63
    g = [1]
    for r in roots:  # g(x) *= (x + alpha^r), coefficients in GF(64), ends up binary
        ng = [0] * (len(g) + 1)
        for k, co in enumerate(g):
            ng[k + 1] ^= co
            ng[k] ^= mul(co, exp[r])
        g = ng
    assert len(g) == 48 and all(c in (0, 1) for c in g)
    msg = [0] * 63
    for i, b in enumerate(data16):
        msg[62 - i] = int(b)
    rem = msg[:]
    for k in range(62, 46, -1):
        if rem[k]:
            for j, co in enumerate(g):
                rem[k - 47 + j] ^= co
    cw = [msg[k] if k >= 47 else rem[k] for k in range(63)]
    return np.array([cw[62 - i] for i in range(63)], np.uint8)


def make_p25_word_cases(rng, code, n):
    """(data, parity) pairs: random words, about 2 % with non-binary bytes."""
    db, pb = {0: (6, 12), 1: (12, 12), 2: (6, 4)}[code]
    d = rng.integers(0, 2, (n, db)).astype(np.uint8)
    p = rng.integers(0, 2, (n, pb)).astype(np.uint8)
    d[::53, 0] = 2
    p[7::61, 1] = 3
    return d, p


def oracle_p25_word(code, d, p):
    O = H.oracle_fec()
    d = d.copy()
    if code == 2:
        return d, O.oracle_hamming_10_6_3_decode(H._ptr(d, H.u8p), H._ptr(p, H.u8p)), None
    fx = C.c_int(0)
    rc = O.oracle_p25_golay24_decode(6 if code == 0 else 12, H._ptr(d, H.u8p), H._ptr(p, H.u8p), C.byref(fx))
    return d, rc, fx.value


@needs_ref
def test_p25_word_codes_match_reference():
    """Golay(24,6)/(24,12) (check_and_fix_golay_24_6/_12), Hamming(10,6,3) (exhaustive) and the NID BCH(63,16,11) decoder ==
    the compiled reference: status, corrected bits, reported error counts, behaviour on non-binary input."""
    O, R = H.oracle_fec(), H.ref_fec()
    rng = np.random.default_rng(40)
    for code, fn in ((0, R.check_and_fix_golay_24_6), (1, R.check_and_fix_golay_24_12)):
        d, p = make_p25_word_cases(rng, code, 6000)
        for i in range(d.shape[0]):
            a, fa = d[i].copy(), C.c_int(-7)
            ra = fn(H._ptr(a, H.u8p), H._ptr(p[i], H.u8p), C.byref(fa))
            b, rb, fb = oracle_p25_word(code, d[i], p[i])
            assert ra == rb and fa.value == fb and np.array_equal(a, b), (code, i)
    for v in range(1024):
        d = np.array([(v >> (9 - i)) & 1 for i in range(6)], np.uint8)
        p = np.array([(v >> (3 - i)) & 1 for i in range(4)], np.uint8)
        a = d.copy()
        ra = R.hamming_10_6_3_decode(H._ptr(a, H.u8p), H._ptr(p, H.u8p))
        b, rb, _ = oracle_p25_word(2, d, p)
        assert ra == rb and np.array_equal(a, b), v
    wins = 0
    for t in range(4000):
        cw = bch_63_16_encode(rng.integers(0, 2, 16))
        x = cw.copy()
        for e in rng.choice(63, int(rng.integers(0, 15)), replace=False):
            x[e] ^= 1
        if t % 7 == 0:
            x = rng.integers(0, 2, 63).astype(np.uint8)
        oa, ob, ea, eb = np.full(16, 9, np.uint8), np.full(16, 9, np.uint8), C.c_int(-1), C.c_int(-1)
        ra = R.ref_bch_63_16_decode(H._ptr(x, H.u8p), H._ptr(oa, H.u8p), C.byref(ea))
        rb = O.oracle_bch_63_16_decode(H._ptr(x, H.u8p), H._ptr(ob, H.u8p), C.byref(eb))
        assert ra == rb and ea.value == eb.value and np.array_equal(oa, ob), t
        wins += ra
    assert wins > 2500


def test_bch_63_16_corrects_up_to_eleven_errors():
    """Property pin without the reference tree: every systematic codeword decodes to its data with error_count = number of
    flipped bits for 0..11 errors."""
    O = H.oracle_fec()
    rng = np.random.default_rng(41)
    for t in range(300):
        data = rng.integers(0, 2, 16).astype(np.uint8)
        cw = bch_63_16_encode(data)
        ne = t % 12
        x = cw.copy()
        for e in rng.choice(63, ne, replace=False):
            x[e] ^= 1
        out, ec = np.zeros(16, np.uint8), C.c_int(-1)
        assert O.oracle_bch_63_16_decode(H._ptr(x, H.u8p), H._ptr(out, H.u8p), C.byref(ec)) == 1
        assert ec.value == ne and np.array_equal(out, data)


# ---------------------------------------------------------------------------- P25p1 NID decode (hard + NAC retry + Chase)

VALID_DUIDS = [0x0, 0x3, 0x5, 0x7, 0xA, 0xC, 0xF]


def make_nid(rng, nac=None, duid=None):
    """63 BCH bits + parity bit of one NID (TIA-102.BAAA-A: parity 1 for LDU1 / LDU2)."""
    nac = int(rng.integers(1, 0xFFF)) if nac is None else nac
    duid = int(rng.choice(VALID_DUIDS)) if duid is None else duid
    info = np.array([(nac >> (11 - i)) & 1 for i in range(12)] + [(duid >> (3 - i)) & 1 for i in range(4)], np.uint8)
    return bch_63_16_encode(info).astype(np.uint8), (1 if duid in (0x5, 0xA) else 0), nac, duid


def nid_cases(rng, n):
    """Noisy NIDs around the BCH radius with reliabilities that mostly (not always) mark the flipped bits as weak."""
    cases = []
    for k in range(n):
        cw, par, nac, duid = make_nid(rng)
        n_err = int(rng.choice([0, 3, 9, 11, 12, 13, 14, 15, 18, 25]))
        pos = rng.choice(63, n_err, replace=False)
        x = cw.copy()
        x[pos] ^= 1
        rel = rng.integers(60, 256, 63).astype(np.uint8)
        weak = pos[rng.random(n_err) < 0.8]
        rel[weak] = rng.integers(0, 90, weak.size)
        if k % 5 == 0:
            rel[rng.choice(63, 10, replace=False)] = rng.integers(0, 64, 10)  # ties and extra weak positions
        if k % 7 == 0:
            rel[:] = rng.integers(0, 3, 63) * 100                                # heavy ties
        parity = par ^ int(rng.random() < 0.2)
        observed = [0, nac, nac, int(rng.integers(1, 0xFFF)), 0xFFF][k % 5]
        cases.append((x, rel, observed, parity, int(rng.integers(0, 256))))
    return cases


@needs_ref
def test_p25p1_nid_decode_matches_reference():
    """oracle_p25p1_nid_decode == the reference's p25p1_nid_decode (status, NAC, DUID, correction count) with and without
    reliabilities, with / without a known NAC, around and beyond the BCH radius."""
    O, R = H.oracle_fec(), H.ref_fec()
    O.oracle_p25p1_nid_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    R.ref_p25p1_nid_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    rng = np.random.default_rng(63)
    seen = set()
    for x, rel, observed, parity, prel in nid_cases(rng, 1500):
        for use_rel in (True, False):
            a, b = [C.c_int() for _ in range(3)], [C.c_int() for _ in range(3)]
            rp = H._ptr(rel, H.u8p) if use_rel else None
            sa = R.ref_p25p1_nid_decode(H._ptr(x, H.u8p), rp, observed, parity, prel, *[C.byref(v) for v in a])
            sb = O.oracle_p25p1_nid_decode(H._ptr(x, H.u8p), rp, observed, parity, prel, 64, *[C.byref(v) for v in b])
            assert sa == sb, (sa, sb)
            if sa > 0:
                assert [v.value for v in a] == [v.value for v in b]
            else:
                assert a[2].value == b[2].value == 0
            seen.add((sa, use_rel))
    assert {(0, True), (1, True), (2, True), (1, False), (0, False)} <= seen


def _oracle_cut(dib, llr, pos_last_sync, n_payload):
    O = H.oracle_fec()
    O.oracle_p25p1_frame_cut.argtypes = [H.u8p, C.POINTER(C.c_int16), C.c_int, C.c_int, C.c_int, H.u8p, H.u8p, H.u8p, H.u8p,
                                         H.u8p, C.POINTER(C.c_int16)]
    dib = np.ascontiguousarray(dib, np.uint8)
    llr = np.ascontiguousarray(llr, np.int16)
    code, rel = np.zeros(63, np.uint8), np.zeros(63, np.uint8)
    par, prel = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    pd, pl = np.zeros(max(n_payload, 1), np.uint8), np.zeros((max(n_payload, 1), 2), np.int16)
    flags = O.oracle_p25p1_frame_cut(H._ptr(dib, H.u8p), llr.ctypes.data_as(C.POINTER(C.c_int16)), dib.size, pos_last_sync,
                                     n_payload, H._ptr(code, H.u8p), H._ptr(rel, H.u8p), H._ptr(par, H.u8p), H._ptr(prel, H.u8p),
                                     H._ptr(pd, H.u8p), pl.ctypes.data_as(C.POINTER(C.c_int16)))
    return flags, code, rel, int(par[0]), int(prel[0]), pd[:n_payload], pl[:n_payload]


def test_p25p1_frame_cut_round_trip():
    """A TSDU built with status symbols on the air-interface grid comes back through the sequential cutter + the pinned
    decoders: NAC / DUID from the NID, the TSBK dibits from the three trellis blocks; truncated streams are flagged."""
    O = H.oracle_fec()
    O.oracle_p25p1_nid_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    rng = np.random.default_rng(8)
    for trial in range(20):
        nac = int(rng.integers(1, 0xFFF))
        frame, payloads = H.p25p1_build_tsdu(rng, nac, 3, bch_63_16_encode)
        assert frame.size == 24 + 33 + 3 * 98 + 8 + 1  # 9 status symbols up to frame offset 359
        pre = int(rng.integers(0, 50))
        dib = np.concatenate([rng.integers(0, 4, pre), frame, rng.integers(0, 4, 7)])
        llr = np.stack([np.where(dib & 2, 200, -200), np.where(dib & 1, 300, -300)], axis=1).astype(np.int16)  # positive = bit 1
        flags, code, rel, par, prel, pd, pl = _oracle_cut(dib, llr, pre + 23, 3 * 98)
        assert flags == 3 and par == 0 and prel == 255 and set(rel.tolist()) <= {200, 255}
        v = [C.c_int() for _ in range(3)]
        st = O.oracle_p25p1_nid_decode(H._ptr(code, H.u8p), H._ptr(rel, H.u8p), 0, par, prel, 64, *[C.byref(x) for x in v])
        assert st == 1 and v[0].value == nac and v[1].value == 7 and v[2].value == 0
        for b in range(3):
            blk = pd[98 * b:98 * (b + 1)]
            out12 = np.zeros(12, np.uint8)
            O.oracle_p25_12_soft_llr(pl[98 * b:98 * (b + 1)].reshape(-1).ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(out12, H.u8p))
            want = np.zeros(12, np.uint8)
            for j in range(48):
                want[j // 4] |= int(payloads[b][j]) << (6 - 2 * (j % 4))
            assert np.array_equal(out12, want), (trial, b)
            assert np.array_equal(np.where(pl[98 * b:98 * (b + 1), 0] > 0, 2, 0) + np.where(pl[98 * b:98 * (b + 1), 1] > 0, 1, 0), blk)
        # truncated: payload incomplete, then NID incomplete
        assert _oracle_cut(dib[:pre + frame.size - 1], llr[:pre + frame.size - 1], pre + 23, 3 * 98)[0] == 3  # only the trailing status symbol is missing
        assert _oracle_cut(dib[:pre + frame.size - 2], llr[:pre + frame.size - 2], pre + 23, 3 * 98)[0] == 1
        assert _oracle_cut(dib[:pre + 40], llr[:pre + 40], pre + 23, 3 * 98)[0] == 0


def hamming_soft_cases(rng, n):
    """10-bit words around Hamming(10,6,3) codewords with 0..3 flips and reliabilities that mostly mark the flips as weak."""
    bits, rel = np.zeros((n, 10), np.uint8), np.zeros((n, 10), np.int32)
    for k in range(n):
        d = rng.integers(0, 2, 6).astype(np.uint8)
        cw = np.concatenate([d, [d[0] ^ d[1] ^ d[2] ^ d[5], d[0] ^ d[1] ^ d[3] ^ d[5], d[0] ^ d[2] ^ d[3] ^ d[4],
                                 d[1] ^ d[2] ^ d[3] ^ d[4]]]).astype(np.uint8)
        pos = rng.choice(10, int(rng.integers(0, 4)), replace=False)
        cw[pos] ^= 1
        r = rng.integers(40, 300, 10)
        weak = pos[rng.random(pos.size) < 0.75]
        r[weak] = rng.integers(-5, 80, weak.size)
        if k % 6 == 0:
            r[:] = rng.integers(0, 3, 10) * 64  # ties on the threshold
        bits[k], rel[k] = cw, r
    return bits, rel


@needs_ref
def test_hamming_10_6_3_soft_matches_reference():
    O, R = H.oracle_fec(), H.ref_fec()
    O.oracle_hamming_10_6_3_soft.argtypes = [H.u8p, H.i32p, C.c_int, C.c_int, H.u8p]
    R.hamming_10_6_3_soft.argtypes = [H.u8p, H.i32p, H.u8p]
    rng = np.random.default_rng(1063)
    bits, rel = hamming_soft_cases(rng, 6000)
    seen = set()
    for k in range(bits.shape[0]):
        oa, ob = np.zeros(10, np.uint8), np.zeros(10, np.uint8)
        ra = R.hamming_10_6_3_soft(H._ptr(bits[k], H.u8p), rel[k].ctypes.data_as(H.i32p), H._ptr(oa, H.u8p))
        rb = O.oracle_hamming_10_6_3_soft(H._ptr(bits[k], H.u8p), rel[k].ctypes.data_as(H.i32p), 1, 64, H._ptr(ob, H.u8p))
        assert ra == rb and np.array_equal(oa, ob), (k, ra, rb, bits[k], rel[k], oa, ob)
        seen.add(ra)
    assert seen == {0, 1, 2}


def golay_soft_cases(rng, length, n):
    """Noisy Golay(24,6) / (24,12) words (the zero codeword plus 0..6 flips: the code is linear) with reliabilities that mostly
    mark the flips as weak; returns data [n, length], parity [n, 12], reliab [n, length + 12] int32."""
    w = length + 12
    bits = np.zeros((n, w), np.uint8)
    rel = rng.integers(40, 300, (n, w)).astype(np.int32)
    for k in range(n):
        pos = rng.choice(w, int(rng.integers(0, 7)), replace=False)
        bits[k, pos] ^= 1
        weak = pos[rng.random(pos.size) < 0.8]
        rel[k, weak] = rng.integers(-3, 90, weak.size)
        if k % 6 == 0:
            rel[k] = rng.integers(0, 3, w) * 64
    return np.ascontiguousarray(bits[:, :length]), np.ascontiguousarray(bits[:, length:]), rel


@needs_ref
@pytest.mark.parametrize("length", [6, 12])
def test_p25_golay24_soft_matches_reference(length):
    O, R = H.oracle_fec(), H.ref_fec()
    O.oracle_p25_golay24_soft.argtypes = [C.c_int, H.u8p, H.u8p, H.i32p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    fn = R.check_and_fix_golay_24_6_soft if length == 6 else R.check_and_fix_golay_24_12_soft
    fn.argtypes = [H.u8p, H.u8p, H.i32p, C.POINTER(C.c_int)]
    rng = np.random.default_rng(2400 + length)
    data, par, rel = golay_soft_cases(rng, length, 2500)
    seen = set()
    for k in range(data.shape[0]):
        da, db = data[k].copy(), data[k].copy()
        fa, fb = C.c_int(0), C.c_int(0)
        ra = fn(H._ptr(da, H.u8p), H._ptr(par[k], H.u8p), rel[k].ctypes.data_as(H.i32p), C.byref(fa))
        rb = O.oracle_p25_golay24_soft(length, H._ptr(db, H.u8p), H._ptr(par[k], H.u8p), rel[k].ctypes.data_as(H.i32p), 1, 64, C.byref(fb))
        assert ra == rb and np.array_equal(da, db) and fa.value == fb.value, (k, ra, rb, fa.value, fb.value)
        seen.add(ra)
    assert 0 in seen


def test_dmr_burst_cut_round_trip():
    """A DMR BS data burst built from BPTC(196,96) / Golay(20,8) / Hamming(7,4) codewords comes back through the sequential
    cutter + the pinned decoders; the inverted-DMR switch and truncated streams behave as in dmr_data_sync."""
    O = H.oracle_fec()
    O.oracle_dmr_burst_cut.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, H.u8p, H.u8p, H.u8p, H.u8p]
    rng = np.random.default_rng(18)
    for trial in range(6):
        payload = rng.integers(0, 2, 96).astype(np.uint8)
        cc, dt = int(rng.integers(0, 16)), int(rng.integers(0, 11))
        burst, sent = H.dmr_build_bs_data_burst(rng, payload, cc, dt)
        pre = int(rng.integers(0, 40))
        dib = np.concatenate([rng.integers(0, 4, pre), burst, rng.integers(0, 4, 5)]).astype(np.uint8)
        rel = rng.integers(0, 256, dib.size).astype(np.uint8)
        pos = pre + 12 + 49 + 5 + 23
        cach, info, rel98, slot = np.zeros(24, np.uint8), np.zeros(196, np.uint8), np.zeros(98, np.uint8), np.zeros(20, np.uint8)
        args = [H._ptr(a, H.u8p) for a in (cach, info, rel98, slot)]
        assert O.oracle_dmr_burst_cut(H._ptr(dib, H.u8p), H._ptr(rel, H.u8p), dib.size, pos, 0, *args) == 1
        assert np.array_equal(cach, sent["cach"]) and np.array_equal(info, sent["info"]) and np.array_equal(slot, sent["slot"])
        assert np.array_equal(rel98[:49], rel[pre + 12:pre + 61]) and np.array_equal(rel98[49:], rel[pos + 6:pos + 55])
        dei, out, r3, und = np.zeros(196, np.uint8), np.zeros(96, np.uint8), np.zeros(3, np.uint8), C.c_int(0)
        O.oracle_bptc_deinterleave(H._ptr(info, H.u8p), H._ptr(dei, H.u8p))
        assert O.oracle_bptc_196x96_extract(H._ptr(dei, H.u8p), H._ptr(out, H.u8p), H._ptr(r3, H.u8p), C.byref(und)) == 0
        assert np.array_equal(out, payload)
        s2 = slot.copy()
        assert O.oracle_golay_20_8_decode(H._ptr(s2, H.u8p))
        assert int("".join(map(str, s2[:4])), 2) == cc and int("".join(map(str, s2[4:8])), 2) == dt
        # inverted DMR: the part up to the sync's end is XORed with 2
        dinv = dib.copy()
        dinv[:pos + 1] ^= 2
        assert O.oracle_dmr_burst_cut(H._ptr(dinv, H.u8p), H._ptr(rel, H.u8p), dib.size, pos, 1, *args) == 1
        assert np.array_equal(info, sent["info"]) and np.array_equal(slot, sent["slot"])
        assert O.oracle_dmr_burst_cut(H._ptr(dib, H.u8p), H._ptr(rel, H.u8p), pos + 54, pos, 0, *args) == 0


# ---- DMR rate 3/4 trellis and RS(12,9) ----------------------------------------------------------------------------------------

# the reference's own known-answer vectors: tests/protocol/dmr/dmr_r34_reference_vectors.h (payload, transmitted dibits)
R34_VECTORS = [
    ("02550B350F9F838235DA49FB52ACE4645BA8",
     "02222110323133223303302102023310133332322133102133130022101023022011033231112022112331223122222003"),
    ("9032A5943D763939B97FE808AB2783BE51F8",
     "13211222200320022021300302221311310212113322220010333303332301203103133220011101322110133223011123"),
    ("3DD640810B98BC52429A00252F64BA72E596",
     "31111320203033023023212313312133330001120220120303313232102332021322301123300222122032331132023303"),
]
RS129_CODEWORD = [3, 20, 37, 54, 71, 88, 105, 122, 139, 208, 63, 250]  # tests/fec/test_fec_bptc_rs.c:194


def oracle_r34(dibits, reliab=None):
    O = H.oracle_fec()
    O.oracle_dmr_r34_decode.argtypes = [H.u8p, H.u8p, H.u8p]
    out = np.zeros(18, np.uint8)
    d = np.ascontiguousarray(dibits, np.uint8)
    r = None if reliab is None else np.ascontiguousarray(reliab, np.uint8)
    assert O.oracle_dmr_r34_decode(H._ptr(d, H.u8p), None if r is None else H._ptr(r, H.u8p), H._ptr(out, H.u8p)) == 0
    return out


def oracle_rs129(cw):
    O = H.oracle_fec()
    O.oracle_rs_12_9_decode.argtypes = [H.u8p, H.u8p, H.u8p]
    c = np.ascontiguousarray(cw, np.uint8).copy()
    syn, ef = np.zeros(3, np.uint8), np.zeros(1, np.uint8)
    res = O.oracle_rs_12_9_decode(H._ptr(c, H.u8p), H._ptr(syn, H.u8p), H._ptr(ef, H.u8p))
    return res, c, syn, int(ef[0])


def make_r34_cases(seed, n):
    """Valid codewords (the reference vectors) with 0..8 random dibit errors and random reliabilities, plus pure noise."""
    rng = np.random.default_rng(seed)
    base = [np.array([int(ch) for ch in d], np.uint8) for _, d in R34_VECTORS]
    dib, rel = np.zeros((n, 98), np.uint8), rng.integers(0, 256, (n, 98)).astype(np.uint8)
    for i in range(n):
        if i % 5 == 4:
            dib[i] = rng.integers(0, 4, 98)
            continue
        d = base[i % 3].copy()
        pos = rng.choice(98, int(rng.integers(0, 9)), replace=False)
        d[pos] = rng.integers(0, 4, pos.size)
        rel[i, pos] = rng.integers(0, 64, pos.size)
        dib[i] = d
    return dib, rel


def make_rs129_cases(seed, n):
    rng = np.random.default_rng(seed)
    cw = np.tile(np.array(RS129_CODEWORD, np.uint8), (n, 1))
    for i in range(n):
        k = i % 5  # 0 clean, 1..3 corrupted bytes, 4 random words
        if k == 4:
            cw[i] = rng.integers(0, 256, 12)
        else:
            pos = rng.choice(12, k, replace=False)
            cw[i, pos] ^= rng.integers(1, 256, k).astype(np.uint8)
    return cw


def test_oracle_r34_reference_vectors():
    for payload, dibits in R34_VECTORS:
        d = np.array([int(ch) for ch in dibits], np.uint8)
        assert oracle_r34(d).tobytes().hex().upper() == payload
        assert oracle_r34(d, np.full(98, 200, np.uint8)).tobytes().hex().upper() == payload


def test_oracle_rs129_reference_codeword():
    res, c, syn, ef = oracle_rs129(RS129_CODEWORD)
    assert res == 0 and not syn.any()
    bad = np.array(RS129_CODEWORD, np.uint8)
    bad[4] ^= 0x5A
    res, c, syn, ef = oracle_rs129(bad)
    assert res == 2 and ef == 1 and list(c) == RS129_CODEWORD and syn.any()


@needs_ref
def test_oracle_r34_matches_compiled_reference():
    R = H.ref_fec("par")
    R.dmr_r34_viterbi_decode.argtypes = [H.u8p, H.u8p]
    R.dmr_r34_viterbi_decode_soft.argtypes = [H.u8p, H.u8p, H.u8p]
    dib, rel = make_r34_cases(5, 400)
    for i in range(dib.shape[0]):
        out = np.zeros(18, np.uint8)
        assert R.dmr_r34_viterbi_decode(H._ptr(dib[i], H.u8p), H._ptr(out, H.u8p)) == 0
        assert np.array_equal(out, oracle_r34(dib[i])), i
        assert R.dmr_r34_viterbi_decode_soft(H._ptr(dib[i], H.u8p), H._ptr(rel[i], H.u8p), H._ptr(out, H.u8p)) == 0
        assert np.array_equal(out, oracle_r34(dib[i], rel[i])), i


@needs_ref
def test_oracle_rs129_matches_compiled_reference():
    R = H.ref_fec("par")
    R.rs_12_9_calc_syndrome.argtypes = [H.u8p, H.u8p]
    R.rs_12_9_check_syndrome.argtypes = [H.u8p]
    R.rs_12_9_check_syndrome.restype = C.c_uint8
    R.rs_12_9_correct_errors.argtypes = [H.u8p, H.u8p, H.u8p]
    R.rs_12_9_correct_errors.restype = C.c_uint8
    cases = make_rs129_cases(6, 1500)
    seen = set()
    for i in range(cases.shape[0]):
        c = cases[i].copy()
        syn, ef = np.zeros(6, np.uint8), np.zeros(1, np.uint8)
        R.rs_12_9_calc_syndrome(H._ptr(c, H.u8p), H._ptr(syn, H.u8p))
        want = 0
        if R.rs_12_9_check_syndrome(H._ptr(syn, H.u8p)):
            want = 1 + R.rs_12_9_correct_errors(H._ptr(c, H.u8p), H._ptr(syn, H.u8p), H._ptr(ef, H.u8p))
        res, oc, osyn, oef = oracle_rs129(cases[i])
        assert (res, oef) == (want, int(ef[0])) and np.array_equal(oc, c) and np.array_equal(osyn, syn[:3]), i
        seen.add(res)
    assert seen == {0, 1, 2, 3}
