"""GPU parity tests for the batched FEC kernels: bit-exact against the oracle (pinned to the reference by
tests/test_oracle_fec.py) on exhaustive / seeded random batches, through the host-buffer C-ABI."""
import ctypes as C
import itertools

import numpy as np
import pytest

import _harness as H
from test_oracle_fec import (HAM, bptc_kat_bits, _rs_words, make_rs_soft_cases, oracle_rs_erasures, bch_63_16_encode,
                             make_p25_word_cases, oracle_p25_word)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("code", [0, 1, 2, 3, 4])
def test_hamming_exhaustive(gpu, code):
    O = H.oracle_fec()
    n, k, _ = HAM[code]
    words = np.arange(1 << n, dtype=np.uint32)
    bits = ((words[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1).astype(np.uint8)
    bits[::7] |= 0xA4  # dirty upper bits must be ignored on input and preserved in place (fec.c sums raw bytes mod 2)
    want_bits = bits.copy()
    want_dec = np.full((bits.shape[0], k), 9, np.uint8)
    want_ok = np.zeros(bits.shape[0], np.uint8)
    for i in range(bits.shape[0]):
        want_ok[i] = O.oracle_hamming_decode(code, H._ptr(want_bits[i], H.u8p), None if code == 0 else H._ptr(want_dec[i], H.u8p))
    got_bits = bits.copy()
    got_dec = np.full((bits.shape[0], k), 9, np.uint8)
    got_ok = gpu.fec_block_decode(code, got_bits, None if code == 0 else got_dec)
    assert np.array_equal(got_ok, want_ok)
    assert np.array_equal(got_bits, want_bits)
    assert np.array_equal(got_dec, want_dec)


@pytest.mark.parametrize("code,n,fn", [(5, 20, "oracle_golay_20_8_decode"), (6, 24, "oracle_golay_24_12_decode"),
                                        (7, 16, "oracle_qr_16_7_6_decode")])
def test_golay_qr_all_light_patterns(gpu, code, n, fn):
    O = H.oracle_fec()
    pats = [p for w in range(0, 5) for p in itertools.combinations(range(n), w)]
    bits = np.zeros((len(pats), n), np.uint8)
    for i, p in enumerate(pats):
        bits[i, list(p)] = 1
    rng = np.random.default_rng(code)
    rnd = rng.integers(0, 2, (5000, n)).astype(np.uint8)
    bits = np.concatenate([bits, rnd])
    want = bits.copy()
    want_ok = np.array([getattr(O, fn)(H._ptr(want[i], H.u8p)) for i in range(want.shape[0])], np.uint8)
    got = bits.copy()
    got_ok = gpu.fec_block_decode(code, got)
    assert np.array_equal(got_ok != 0, want_ok != 0)
    assert np.array_equal(got, want)


def test_golay_encode(gpu):
    import torch

    O = H.oracle_fec()
    data = ((np.arange(4096)[:, None] >> (11 - np.arange(12))[None, :]) & 1).astype(np.uint8)
    d = torch.from_numpy(data).cuda()
    out = torch.empty((4096, 24), dtype=torch.uint8, device="cuda")
    gpu.check(gpu.lib().dsdneo_b200_fec_golay_24_12_encode_batch(d.data_ptr(), out.data_ptr(), 4096, None))
    got = out.cpu().numpy()
    for i in range(0, 4096, 37):
        want = np.zeros(24, np.uint8)
        O.oracle_golay_24_12_encode(H._ptr(data[i], H.u8p), H._ptr(want, H.u8p))
        assert np.array_equal(got[i], want)


def test_bptc_batch(gpu):
    O = H.oracle_fec()
    rng = np.random.default_rng(61)
    cw = bptc_kat_bits()
    n = 6000
    bursts = np.tile(cw, (n, 1))
    for i in range(n):
        if i % 3 == 0:
            bursts[i] = rng.integers(0, 2, 196)
        else:
            bursts[i, rng.choice(196, size=int(rng.integers(0, 14)), replace=False)] ^= 1
    want_out, want_r, want_e, undef = np.zeros((n, 96), np.uint8), np.zeros((n, 3), np.uint8), np.zeros(n, np.uint32), np.zeros(n, bool)
    for i in range(n):
        u = C.c_int(0)
        want_e[i] = O.oracle_bptc_196x96_extract(H._ptr(bursts[i], H.u8p), H._ptr(want_out[i], H.u8p), H._ptr(want_r[i], H.u8p), C.byref(u))
        undef[i] = bool(u.value)
    out, r3, errs = gpu.bptc_196x96(bursts, interleaved=False)
    assert np.array_equal(errs, want_e) and np.array_equal(out, want_out) and np.array_equal(r3, want_r)  # incl. the documented choice
    # fused de-interleave: transmit order in, same answers out
    inter = np.zeros_like(bursts)
    inter[:, (np.arange(196) * 181) % 196] = bursts
    out2, r32, errs2 = gpu.bptc_196x96(inter, interleaved=True)
    assert np.array_equal(errs2, want_e) and np.array_equal(out2, want_out) and np.array_equal(r32, want_r)
    assert (~undef).sum() > 3000


def test_p25_12_batch(gpu):
    O = H.oracle_fec()
    rng = np.random.default_rng(71)
    i16p, u32p = C.POINTER(C.c_int16), C.POINTER(C.c_uint32)
    n = 512
    llrs = np.zeros((n, 196), np.int16)
    for t in range(n):
        dib, tx = H.p25_trellis_encode(rng)
        llrs[t] = H.dibits_to_llr(tx, 200, rng, [0.0, 60.0, 150.0, 260.0][t % 4])
        if t % 10 == 9:
            llrs[t] = rng.integers(-300, 300, 196)
        if t % 25 == 0:
            llrs[t] = 0
        if t % 50 == 1:
            llrs[t] = rng.integers(-32768, 32768, 196)  # full-range LLRs incl. -32768
    out, met = gpu.p25_12_soft_llr(llrs)
    for t in range(n):
        w = np.zeros(12, np.uint8)
        m = O.oracle_p25_12_soft_llr(llrs[t].ctypes.data_as(i16p), H._ptr(w, H.u8p))
        assert m == met[t] and np.array_equal(out[t], w), t
    for maxc in (8, 3):
        cands, cnt = gpu.p25_12_soft_llr_list(llrs, maxc)
        for t in range(n):
            cb, cm = np.zeros((8, 12), np.uint8), np.zeros(8, np.uint32)
            na = O.oracle_p25_12_soft_llr_list(llrs[t].ctypes.data_as(i16p), H._ptr(cb, H.u8p), cm.ctypes.data_as(u32p), maxc)
            assert na == cnt[t], (t, na, cnt[t])
            for c in range(na):
                assert bytes(cb[c]) == bytes(cands[8 * t + c].bytes) and int(cm[c]) == cands[8 * t + c].metric, (t, c)


@pytest.mark.parametrize("variant,n,k", [(0, 36, 20), (1, 24, 12), (2, 24, 16)])
def test_p25_rs_batch(gpu, variant, n, k):
    O = H.oracle_fec()
    tt = (n - k) // 2
    rng = np.random.default_rng(81 + variant)
    nw = 3000
    data_bits = np.zeros((nw, k * 6), np.uint8)
    par_bits = np.zeros((nw, 2 * tt * 6), np.uint8)
    for t in range(nw):
        data = np.zeros(63 - 2 * tt, np.int32)
        data[:k] = rng.integers(0, 64, k)
        cw = np.zeros(63, np.int32)
        O.oracle_rs63_encode(tt, data.ctypes.data_as(H.i32p), cw.ctypes.data_as(H.i32p))
        nerr = int(rng.integers(0, tt + 5))
        pos = rng.choice(n, size=nerr, replace=False)
        cw[pos] ^= rng.integers(1, 64, pos.size).astype(np.int32)
        par_bits[t] = _rs_words(cw[:2 * tt])
        data_bits[t] = _rs_words(cw[2 * tt:2 * tt + k])
    want = data_bits.copy()
    want_st = np.array([O.oracle_p25_rs_decode(n, k, H._ptr(want[t], H.u8p), H._ptr(par_bits[t], H.u8p)) for t in range(nw)], np.uint8)
    got = data_bits.copy()
    st = gpu.p25_rs_decode(variant, got, par_bits)
    assert np.array_equal(st, want_st)
    assert np.array_equal(got, want)
    assert (want_st == 0).sum() > nw // 3 and (want_st == 1).sum() > 50


def test_viterbi_k5_batch(gpu):
    from test_oracle_fec import _bind_conv_oracle, viterbi_cases

    O = H.oracle_fec()
    _bind_conv_oracle(O)
    rng = np.random.default_rng(91)
    u16p = C.POINTER(C.c_uint16)
    cases = viterbi_cases(rng, 208)
    for nbits in (240, 96, 40, 176):
        grp = [c for c, m in cases if c.size == 2 * (nbits + 4)]
        cost = np.stack(grp)
        init = np.full((cost.shape[0], 64), 0x55, np.uint8)
        out, met = gpu.viterbi_k5_decode(cost, cost.shape[1], out_pitch=64, out_init=init)
        for i in range(cost.shape[0]):
            w = np.full(64, 0x55, np.uint8)
            m = O.oracle_viterbi_k5_decode(H._ptr(w, H.u8p), cost[i].ctypes.data_as(u16p), cost.shape[1])
            assert m == met[i] and np.array_equal(out[i], w), (nbits, i)
    for pattern in ([1, 1, 1, 0], [1] * 8, [1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1]):
        punct = np.array(pattern, np.uint8)
        in_len = 250
        cost = rng.integers(0, 65536, (64, in_len)).astype(np.uint16)
        out, met = gpu.viterbi_k5_decode(cost, in_len, punct=punct, out_pitch=80)
        for i in range(64):
            w = np.zeros(80, np.uint8)
            m = O.oracle_viterbi_k5_decode_punctured(H._ptr(w, H.u8p), cost[i].ctypes.data_as(u16p), H._ptr(punct, H.u8p), in_len, punct.size)
            assert m == met[i] and np.array_equal(out[i], w), (pattern, i)


def test_nxdn_conv_batch_with_carried_metrics(gpu):
    from test_oracle_fec import _bind_conv_oracle

    O = H.oracle_fec()
    _bind_conv_oracle(O)
    rng = np.random.default_rng(92)
    u16p = C.POINTER(C.c_uint16)
    n = 96
    g_metrics = np.zeros((n, 32), np.uint16)
    o_metrics = np.zeros((n, 32), np.uint16)
    for frame in range(6):  # consecutive frames per channel: the metric arrays carry over like the reference's statics
        n_steps = [191, 300, 37, 96, 255, 20][frame]
        n_out = [187, 296, 37, 92, 1, 16][frame]
        soft = frame % 2 == 1
        sym = (rng.integers(0, 2, (n, 2 * n_steps)) * 2).astype(np.uint8)
        sym[rng.random(sym.shape) < 0.05] = 1
        rel = rng.integers(0, 256, sym.shape).astype(np.uint8)
        init = np.full((n, 40), 0xA5, np.uint8)
        out = gpu.nxdn_conv_decode(sym, rel if soft else None, n_steps, n_out, g_metrics, out_pitch=40, out_init=init)
        for i in range(n):
            w = np.full(40, 0xA5, np.uint8)
            O.oracle_nxdn_conv_decode(H._ptr(sym[i], H.u8p), H._ptr(rel[i], H.u8p) if soft else None, n_steps, n_out,
                                      o_metrics[i].ctypes.data_as(u16p), H._ptr(w, H.u8p))
            assert np.array_equal(out[i], w), (frame, i)
        assert np.array_equal(g_metrics, o_metrics), frame


def test_bptc_128x77_and_16x2_bit_exact(gpu):
    """Batched BPTC 128x77 / 16x2 == oracle (== reference, tests/test_oracle_fec.py) on valid, corrupted and random
    inputs; items on which the reference itself is undefined (uninitialised buffer) are excluded."""
    O = H.oracle_fec()
    O.oracle_bptc_128x77_extract.restype = C.c_uint
    O.oracle_bptc_16x2_extract.restype = C.c_uint
    rng = np.random.default_rng(91)
    n = 4096
    mats = rng.integers(0, 2, (n, 128)).astype(np.uint8)
    # make three quarters of them near-valid: valid Hamming rows + consistent parity row + 0..4 flips
    for i in range(n):
        if i % 4:
            for r in range(7):
                line, dec = mats[i, 16 * r:16 * r + 16].copy(), np.zeros(11, np.uint8)
                O.oracle_hamming_decode(4, H._ptr(line, H.u8p), H._ptr(dec, H.u8p))
                mats[i, 16 * r:16 * r + 16] = line
            mats[i, 112:] = mats[i, :112].reshape(7, 16).sum(axis=0) % 2
            for e in rng.integers(0, 128, int(rng.integers(0, 5))):
                mats[i, e] ^= 1
    mats |= (rng.integers(0, 128, (n, 128)) * 2).astype(np.uint8)  # dirty upper bits, as callers pass them
    out, errs = gpu.bptc_128x77(mats)
    checked = 0
    for i in range(n):
        w, und = np.zeros(77, np.uint8), C.c_int(0)
        e = O.oracle_bptc_128x77_extract(H._ptr(mats[i].copy(), H.u8p), H._ptr(w, H.u8p), C.byref(und))
        if und.value:
            continue
        assert e == errs[i] and np.array_equal(out[i], w), i
        checked += 1
    assert checked > n // 2
    words = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    for odd in (False, True):
        out, errs = gpu.bptc_16x2(words, odd)
        checked = 0
        for i in range(n):
            w, und = np.zeros(32, np.uint8), C.c_int(0)
            e = O.oracle_bptc_16x2_extract(H._ptr(words[i].copy(), H.u8p), H._ptr(w, H.u8p), C.c_uint(1 if odd else 0), C.byref(und))
            if und.value:
                continue
            assert e == errs[i] and np.array_equal(out[i], w), i
            checked += 1
        assert checked > n // 4


@pytest.mark.parametrize("n,k,variant", [(36, 20, 0), (24, 12, 1), (24, 16, 2)])
def test_rs_soft_decoders_bit_exact(gpu, n, k, variant):
    """Batched errors-and-erasures RS decode and the ranked-erasure soft wrapper == oracle (== reference,
    tests/test_oracle_fec.py::test_rs_soft_vs_reference): status and every data bit, incl. untouched dirty bytes on failure."""
    O = H.oracle_fec()
    rng = np.random.default_rng(300 + variant)
    dat, par, rel_d, rel_p, ers, n_ers, truth = make_rs_soft_cases(rng, n, k, 1500)
    got, st = gpu.p25_rs_decode_erasures(variant, dat, par, ers, n_ers)
    for i in range(dat.shape[0]):
        want, rc = oracle_rs_erasures(n, k, dat[i], par[i], ers[i], n_ers[i])
        assert rc == st[i] and np.array_equal(got[i], want), i
    got, st = gpu.p25_rs_soft_reliability(variant, dat, par, rel_d, rel_p, 64)
    solved = 0
    for i in range(dat.shape[0]):
        b = dat[i].copy()
        rc = O.oracle_p25_rs_soft_reliability(n, k, H._ptr(b, H.u8p), H._ptr(par[i], H.u8p), H._ptr(rel_d[i].copy(), H.u8p),
                                              H._ptr(rel_p[i].copy(), H.u8p), 64)
        assert rc == st[i] and np.array_equal(got[i], b), i
        solved += rc == 0 and np.array_equal(b, truth[i])
    assert solved > 400
    # a different threshold changes the ranked count, not the order
    got2, st2 = gpu.p25_rs_soft_reliability(variant, dat[:200], par[:200], rel_d[:200], rel_p[:200], 200)
    for i in range(200):
        b = dat[i].copy()
        rc = O.oracle_p25_rs_soft_reliability(n, k, H._ptr(b, H.u8p), H._ptr(par[i], H.u8p), H._ptr(rel_d[i].copy(), H.u8p),
                                              H._ptr(rel_p[i].copy(), H.u8p), 200)
        assert rc == st2[i] and np.array_equal(got2[i], b), i


def test_p25_word_codes_and_nid_bch_bit_exact(gpu):
    """Batched Golay(24,6)/(24,12), Hamming(10,6,3) and BCH(63,16,11) == oracle (== reference, tests/test_oracle_fec.py)."""
    O = H.oracle_fec()
    rng = np.random.default_rng(410)
    for code in (0, 1, 2):
        d, p = make_p25_word_cases(rng, code, 5000)
        got, st, fx = gpu.p25_word_decode(code, d, p)
        for i in range(d.shape[0]):
            w, rc, f = oracle_p25_word(code, d[i], p[i])
            assert rc == st[i] and np.array_equal(got[i], w), (code, i)
            if code != 2:
                assert f == fx[i], (code, i)
    words = []
    for t in range(3000):
        x = bch_63_16_encode(rng.integers(0, 2, 16))
        for e in rng.choice(63, int(rng.integers(0, 15)), replace=False):
            x[e] ^= 1
        words.append(rng.integers(0, 2, 63).astype(np.uint8) if t % 7 == 0 else x)
    words = np.array(words)
    init = np.full((words.shape[0], 16), 9, np.uint8)
    out, ok, ec = gpu.bch_63_16_decode(words, init)
    for i in range(words.shape[0]):
        w, e = np.full(16, 9, np.uint8), C.c_int(-1)
        r = O.oracle_bch_63_16_decode(H._ptr(words[i], H.u8p), H._ptr(w, H.u8p), C.byref(e))
        assert r == ok[i] and e.value == ec[i] and np.array_equal(out[i], w), i
    assert ok.sum() > 1800


def test_p25p1_nid_decode_bit_exact_vs_oracle(gpu):
    """Batched p25p1_nid_decode (hard, NAC retry, Chase search) == oracle (pinned to the reference) on noisy NIDs around and
    beyond the BCH radius: status, NAC, DUID, correction count; with and without reliabilities / known NAC."""
    from test_oracle_fec import nid_cases

    O = H.oracle_fec()
    O.oracle_p25p1_nid_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    rng = np.random.default_rng(631)
    cases = nid_cases(rng, 3000)
    code = np.stack([c[0] for c in cases]).astype(np.uint8)
    rel = np.stack([c[1] for c in cases]).astype(np.uint8)
    obs = np.array([c[2] for c in cases], np.int32)
    par = np.array([c[3] for c in cases], np.uint8)
    prel = np.array([c[4] for c in cases], np.uint8)
    for use_rel, use_obs in [(True, True), (False, True), (True, False)]:
        st, nac, duid, errs = gpu.p25p1_nid_decode(code, rel if use_rel else None, obs if use_obs else None, par,
                                                   prel if use_rel else None, 64)
        seen = set()
        for i in range(len(cases)):
            v = [C.c_int() for _ in range(3)]
            want = O.oracle_p25p1_nid_decode(H._ptr(code[i], H.u8p), H._ptr(rel[i], H.u8p) if use_rel else None,
                                             int(obs[i]) if use_obs else 0, int(par[i]), int(prel[i]) if use_rel else 0, 64,
                                             *[C.byref(x) for x in v])
            assert (st[i], nac[i], duid[i], errs[i]) == (want, v[0].value, v[1].value, v[2].value), (i, use_rel, use_obs)
            seen.add(want)
        assert seen == {0, 1, 2}


def test_hamming_10_6_3_soft_bit_exact_vs_oracle(gpu):
    from test_oracle_fec import hamming_soft_cases

    O = H.oracle_fec()
    O.oracle_hamming_10_6_3_soft.argtypes = [H.u8p, H.i32p, C.c_int, C.c_int, H.u8p]
    rng = np.random.default_rng(1064)
    bits, rel = hamming_soft_cases(rng, 8000)
    for override in (1, 0):
        out, st = np.zeros_like(bits), np.zeros(bits.shape[0], np.uint8)
        gpu.check(gpu.lib().dsdneo_b200_hamming_10_6_3_soft_batch_host(bits.ctypes.data, rel.ctypes.data, override, 64,
                                                                      out.ctypes.data, st.ctypes.data, bits.shape[0]))
        for k in range(bits.shape[0]):
            want = np.zeros(10, np.uint8)
            rc = O.oracle_hamming_10_6_3_soft(H._ptr(bits[k], H.u8p), rel[k].ctypes.data_as(H.i32p), override, 64, H._ptr(want, H.u8p))
            assert rc == st[k] and np.array_equal(out[k], want), (k, override)
        assert set(st.tolist()) == {0, 1, 2}


@pytest.mark.parametrize("length", [6, 12])
def test_p25_golay24_soft_bit_exact_vs_oracle(gpu, length):
    from test_oracle_fec import golay_soft_cases

    O = H.oracle_fec()
    O.oracle_p25_golay24_soft.argtypes = [C.c_int, H.u8p, H.u8p, H.i32p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    rng = np.random.default_rng(2450 + length)
    data, par, rel = golay_soft_cases(rng, length, 4000)
    code = gpu.P25_WORD_GOLAY_24_6 if length == 6 else gpu.P25_WORD_GOLAY_24_12
    for override in (1, 0):
        got = data.copy()
        st, fx = np.zeros(data.shape[0], np.uint8), np.zeros(data.shape[0], np.int32)
        gpu.check(gpu.lib().dsdneo_b200_p25_golay_soft_batch_host(code, got.ctypes.data, par.ctypes.data, rel.ctypes.data, override, 64,
                                                                 st.ctypes.data, fx.ctypes.data, data.shape[0]))
        for k in range(data.shape[0]):
            want = data[k].copy()
            f = C.c_int(0)
            rc = O.oracle_p25_golay24_soft(length, H._ptr(want, H.u8p), H._ptr(par[k], H.u8p), rel[k].ctypes.data_as(H.i32p), override, 64,
                                           C.byref(f))
            assert rc == st[k] and np.array_equal(got[k], want) and f.value == fx[k], (k, override, rc, st[k], f.value, fx[k])


def test_dmr_r34_hard_and_soft(gpu):
    from test_oracle_fec import R34_VECTORS, make_r34_cases, oracle_r34

    dib, rel = make_r34_cases(11, 3000)
    for payload, dibits in R34_VECTORS:
        d = np.array([[int(ch) for ch in dibits]], np.uint8)
        out = np.zeros((1, 18), np.uint8)
        gpu.check(gpu.lib().dsdneo_b200_dmr_r34_decode_batch_host(d.ctypes.data, None, out.ctypes.data, 1), "r34")
        assert out.tobytes().hex().upper() == payload
    n = dib.shape[0]
    hard, soft = np.zeros((n, 18), np.uint8), np.zeros((n, 18), np.uint8)
    gpu.check(gpu.lib().dsdneo_b200_dmr_r34_decode_batch_host(dib.ctypes.data, None, hard.ctypes.data, n), "r34")
    gpu.check(gpu.lib().dsdneo_b200_dmr_r34_decode_batch_host(dib.ctypes.data, rel.ctypes.data, soft.ctypes.data, n), "r34 soft")
    for i in range(n):
        assert np.array_equal(hard[i], oracle_r34(dib[i])), i
        assert np.array_equal(soft[i], oracle_r34(dib[i], rel[i])), i


def test_rs_12_9(gpu):
    from test_oracle_fec import make_rs129_cases, oracle_rs129

    cw = make_rs129_cases(12, 5000)
    n = cw.shape[0]
    got = cw.copy()
    syn, res, ef = np.zeros((n, 3), np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    gpu.check(gpu.lib().dsdneo_b200_rs_12_9_decode_batch_host(got.ctypes.data, syn.ctypes.data, res.ctypes.data, ef.ctypes.data, n), "rs129")
    for i in range(n):
        r, c, s, e = oracle_rs129(cw[i])
        assert (int(res[i]), int(ef[i])) == (r, e) and np.array_equal(got[i], c) and np.array_equal(syn[i], s), i
    assert set(res.tolist()) == {0, 1, 2, 3}
