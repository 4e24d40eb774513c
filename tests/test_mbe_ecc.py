"""Vocoder frame ECC (ambe_fr / imbe_fr -> ambe_d / imbe_d) and the DMR BS voice burst cutter.

PARITY UNPINNED for the ECC: mbelib-neo is not in the reference tree (oracle/oracle_mbe.c header).  What can be checked without
it is checked here: the two block codes have the distance / correction radius of the [23,12] Golay and [15,11] Hamming codes
(both perfect, so every syndrome decoder gives the same word), encode -> modulate -> corrupt -> decode round trips recover the
parameter bits with the right error counts, and the CUDA kernels equal the CPU restatement bit for bit.  The burst cutter's
de-interleave schedule is compared with the reference's own table when the reference tree is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _harness as H

u8p = C.POINTER(C.c_uint8)
REF_MAP = "/root/reference/include/dsd-neo/core/ambe_interleave.h"


def _o():
    L = H.oracle()
    L.oracle_mbe_golay2312_encode.restype = C.c_uint
    L.oracle_mbe_hamming1511_encode.restype = C.c_uint
    return L


def _bits(word, n):
    return np.array([(word >> i) & 1 for i in range(n)], np.uint8)


def _pn(seed12, n):
    pr, out = 16 * seed12, []
    for _ in range(n):
        pr = (173 * pr + 13849) & 0xFFFF
        out.append(pr >> 15)
    return out


def ambe_encode(L, d49):
    """ambe_d[49] -> ambe_fr[4][24] (the transmit side of TIA-102.BABA-style AMBE+2 3600x2450: test helper)."""
    fr = np.zeros((4, 24), np.uint8)
    u0 = int("".join(map(str, d49[:12])), 2)
    u1 = int("".join(map(str, d49[12:24])), 2)
    c0 = L.oracle_mbe_golay2312_encode(u0)
    fr[0, 1:24] = _bits(c0, 23)
    fr[0, 0] = bin(c0).count("1") & 1
    c1 = _bits(L.oracle_mbe_golay2312_encode(u1), 23)
    pn = _pn(u0, 23)
    for k, j in enumerate(range(22, -1, -1)):
        c1[j] ^= pn[k]
    fr[1, :23] = c1
    fr[2, :11] = d49[24:35][::-1]
    fr[3, :14] = d49[35:49][::-1]
    return fr


def imbe_encode(L, d88):
    fr = np.zeros((8, 23), np.uint8)
    u = [int("".join(map(str, d88[12 * i:12 * i + 12])), 2) for i in range(4)]
    v = [int("".join(map(str, d88[48 + 11 * i:59 + 11 * i])), 2) for i in range(3)]
    pn = _pn(u[0], 114)
    k = 0
    fr[0] = _bits(L.oracle_mbe_golay2312_encode(u[0]), 23)
    for i in range(1, 4):
        w = _bits(L.oracle_mbe_golay2312_encode(u[i]), 23)
        for j in range(22, -1, -1):
            w[j] ^= pn[k]
            k += 1
        fr[i] = w
    for i in range(3):
        w = _bits(L.oracle_mbe_hamming1511_encode(v[i]), 15)
        for j in range(14, -1, -1):
            w[j] ^= pn[k]
            k += 1
        fr[4 + i, :15] = w
    fr[7, :7] = d88[81:88][::-1]
    return fr


def _oracle_ambe(L, fr):
    d = np.zeros(49, np.uint8)
    a, b = C.c_int(), C.c_int()
    L.oracle_ambe3600x2450_decode(np.ascontiguousarray(fr).ctypes.data_as(u8p), d.ctypes.data_as(u8p), C.byref(a), C.byref(b))
    return d, a.value, b.value


def _oracle_imbe(L, fr):
    d = np.zeros(88, np.uint8)
    a, b = C.c_int(), C.c_int()
    L.oracle_imbe7200x4400_decode(np.ascontiguousarray(fr).ctypes.data_as(u8p), d.ctypes.data_as(u8p), C.byref(a), C.byref(b))
    return d, a.value, b.value


def test_codes_are_the_perfect_golay_and_hamming_codes():
    L = _o()
    cw = [L.oracle_mbe_golay2312_encode(d) for d in range(4096)]
    assert min(bin(c).count("1") for c in cw[1:]) == 7
    assert len({c ^ a ^ b for c in cw[:64] for a in cw[:8] for b in cw[:8]} - set(cw)) == 0  # linear
    hw = [L.oracle_mbe_hamming1511_encode(d) for d in range(2048)]
    assert min(bin(c).count("1") for c in hw[1:]) == 3
    rng = np.random.default_rng(1)
    out = np.zeros(23, np.uint8)
    for _ in range(4000):
        d = int(rng.integers(0, 4096))
        pos = rng.choice(23, size=int(rng.integers(0, 4)), replace=False)
        r = cw[d]
        for p in pos:
            r ^= 1 << int(p)
        errs = L.oracle_mbe_golay2312(_bits(r, 23).ctypes.data_as(u8p), out.ctypes.data_as(u8p))
        assert int("".join(map(str, out[11:][::-1])), 2) == d
        assert errs == sum(1 for p in pos if p >= 11)  # the count covers the data bits only, as in mbe_golay2312
    out = np.zeros(15, np.uint8)
    for d in range(2048):
        for p in range(-1, 15):
            r = hw[d] ^ ((1 << p) if p >= 0 else 0)
            errs = L.oracle_mbe_hamming1511(_bits(r, 15).ctypes.data_as(u8p), out.ctypes.data_as(u8p))
            assert int("".join(map(str, out[4:][::-1])), 2) == d and errs == (p >= 0)


def test_oracle_round_trips():
    L = _o()
    rng = np.random.default_rng(2)
    for _ in range(300):
        d = rng.integers(0, 2, 49).astype(np.uint8)
        fr = ambe_encode(L, d)
        e0 = rng.choice(np.arange(1, 24), size=int(rng.integers(0, 4)), replace=False)
        e1 = rng.choice(23, size=int(rng.integers(0, 4)), replace=False)
        fr[0, e0] ^= 1
        fr[1, e1] ^= 1
        fr[0, 0] ^= int(rng.integers(0, 2))  # the overall parity bit is not used
        got, c0, tot = _oracle_ambe(L, fr)
        assert np.array_equal(got, d)
        assert c0 == int((e0 >= 12).sum()) and tot == c0 + int((e1 >= 11).sum())
    for _ in range(300):
        d = rng.integers(0, 2, 88).astype(np.uint8)
        fr = imbe_encode(L, d)
        want = 0
        for i in range(4):
            e = rng.choice(23, size=int(rng.integers(0, 4)), replace=False)
            fr[i, e] ^= 1
            want += int((e >= 11).sum())
            if i == 0:
                want0 = want
        for i in range(4, 7):
            if rng.integers(0, 2):
                fr[i, int(rng.integers(0, 15))] ^= 1
                want += 1
        got, c0, tot = _oracle_imbe(L, fr)
        assert np.array_equal(got, d) and c0 == want0 and tot == want


def _voice_burst(rng, inverted=False):
    dib = rng.integers(0, 4, 144).astype(np.uint8)
    return dib


def test_voice_cut_oracle_inverts_the_interleave():
    """Cutting is a permutation: every vocoder dibit lands in exactly one cell, the unreached cells stay 0, the CACH rule is the
    data cutter's, and the schedule equals the reference's table (when the tree is present)."""
    L = _o()
    rng = np.random.default_rng(3)
    dib = _voice_burst(rng)
    cach, fr, sync = np.zeros(24, np.uint8), np.zeros((3, 4, 24), np.uint8), np.zeros(48, np.uint8)
    L.oracle_dmr_voice_cut(dib.ctypes.data_as(u8p), 0, cach.ctypes.data_as(u8p), fr.ctypes.data_as(u8p), sync.ctypes.data_as(u8p))
    total_bits = int(((dib[12:66] >> 1) & 1).sum() + (dib[12:66] & 1).sum() + ((dib[90:] >> 1) & 1).sum() + (dib[90:] & 1).sum())
    assert int(fr.sum()) == total_bits
    assert fr[:, 1, 23].sum() == 0 and fr[:, 2, 11:].sum() == 0 and fr[:, 3, 14:].sum() == 0
    assert np.array_equal(sync[0::2], (dib[66:90] >> 1) & 1) and np.array_equal(sync[1::2], dib[66:90] & 1)
    inv = np.zeros_like(fr)
    c2, s2 = np.zeros(24, np.uint8), np.zeros(48, np.uint8)
    L.oracle_dmr_voice_cut(dib.ctypes.data_as(u8p), 1, c2.ctypes.data_as(u8p), inv.ctypes.data_as(u8p), s2.ctypes.data_as(u8p))
    d2 = dib.copy()
    d2[:90] ^= 2
    chk = np.zeros_like(fr)
    L.oracle_dmr_voice_cut(d2.ctypes.data_as(u8p), 0, c2.ctypes.data_as(u8p), chk.ctypes.data_as(u8p), s2.ctypes.data_as(u8p))
    assert np.array_equal(inv, chk)
    if os.path.exists(REF_MAP):
        txt = open(REF_MAP).read()
        body = txt[txt.index("dsd_ambe_2450_dibit_map[DSD_AMBE_2450_DIBITS] = {"):]
        ref = np.array([[int(x) for x in m] for m in re.findall(r"\{(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\}", body)][:36], np.uint8)
        # frame 0 of a burst whose only set bit is dibit 12 + i (high) tells where map entry i sends the high bit
        for i in range(36):
            one = np.zeros(144, np.uint8)
            one[12 + i] = 2
            f = np.zeros((3, 4, 24), np.uint8)
            L.oracle_dmr_voice_cut(one.ctypes.data_as(u8p), 0, c2.ctypes.data_as(u8p), f.ctypes.data_as(u8p), s2.ctypes.data_as(u8p))
            assert f[0, ref[i, 0], ref[i, 1]] == 1 and f.sum() == 1


@pytest.mark.gpu
def test_gpu_ecc_equals_oracle(gpu):
    L = _o()
    rng = np.random.default_rng(4)
    n = 3000
    afr = rng.integers(0, 2, (n, 4, 24)).astype(np.uint8)  # arbitrary words: beyond the correction radius too
    ifr = rng.integers(0, 2, (n, 8, 23)).astype(np.uint8)
    for k in range(0, n, 3):  # a third are valid frames with a few errors
        afr[k] = ambe_encode(L, rng.integers(0, 2, 49).astype(np.uint8))
        afr[k, rng.integers(0, 2), rng.integers(1, 23)] ^= 1
        ifr[k] = imbe_encode(L, rng.integers(0, 2, 88).astype(np.uint8))
        ifr[k, rng.integers(0, 7), rng.integers(0, 15)] ^= 1
    d, c0, tot = gpu.ambe3600x2450_decode(afr)
    for k in range(n):
        wd, w0, wt = _oracle_ambe(L, afr[k])
        assert np.array_equal(d[k], wd) and c0[k] == w0 and tot[k] == wt, k
    d, c0, tot = gpu.imbe7200x4400_decode(ifr)
    for k in range(n):
        wd, w0, wt = _oracle_imbe(L, ifr[k])
        assert np.array_equal(d[k], wd) and c0[k] == w0 and tot[k] == wt, k
    assert gpu.ambe3600x2450_decode(afr[:0])[0].shape == (0, 49)


@pytest.mark.gpu
def test_gpu_voice_cut_equals_oracle(gpu):
    import torch

    L = _o()
    rng = np.random.default_rng(5)
    n_ch, max_hits, n_bursts, pitch = 5, 4, 3, 2000
    dib = rng.integers(0, 4, (n_ch, pitch)).astype(np.uint8)
    counts = np.array([2000, 1500, 700, 100, 2000], np.int32)
    hits = np.zeros((n_ch, max_hits, 2), np.int32)
    n_hits = np.array([4, 2, 3, 1, 0], np.int32)
    hits[0, :, 0] = [89, 400, 1000, 1900]   # first: exactly enough history; last: runs past the end
    hits[1, :2, 0] = [50, 1100]             # first: not enough history
    hits[2, :3, 0] = [100, 300, 500]
    hits[3, :1, 0] = [95]
    for inverted in (False, True):
        cach, fr, sync, valid = gpu.dmr_voice_cut(torch.from_numpy(dib).cuda(), torch.from_numpy(counts).cuda(), torch.from_numpy(hits).cuda(),
                                                  torch.from_numpy(n_hits).cuda(), max_hits, n_bursts, inverted)
        cach, fr, sync, valid = (t.cpu().numpy() for t in (cach, fr, sync, valid))
        n_valid = 0
        for ch in range(n_ch):
            for h in range(max_hits):
                for j in range(n_bursts):
                    r = (ch * max_hits + h) * n_bursts + j
                    start = hits[ch, h, 0] + 1 - 90 + 144 * j
                    ok = h < n_hits[ch] and start >= 0 and start + 144 <= counts[ch]
                    assert valid[r] == ok, (ch, h, j)
                    if not ok:
                        continue
                    n_valid += 1
                    wc, wf, ws = np.zeros(24, np.uint8), np.zeros((3, 4, 24), np.uint8), np.zeros(48, np.uint8)
                    b = np.ascontiguousarray(dib[ch, start:start + 144])
                    L.oracle_dmr_voice_cut(b.ctypes.data_as(u8p), int(inverted and j == 0), wc.ctypes.data_as(u8p), wf.ctypes.data_as(u8p),
                                           ws.ctypes.data_as(u8p))
                    assert np.array_equal(cach[r], wc) and np.array_equal(fr[r], wf) and np.array_equal(sync[r], ws), (ch, h, j)
        assert n_valid >= 15
    m = gpu.ambe_2450_dibit_map()
    if os.path.exists(REF_MAP):
        txt = open(REF_MAP).read()
        body = txt[txt.index("dsd_ambe_2450_dibit_map[DSD_AMBE_2450_DIBITS] = {"):]
        ref = np.array([[int(x) for x in mm] for mm in re.findall(r"\{(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\}", body)][:36], np.uint8)
        assert np.array_equal(m, ref)


@pytest.mark.gpu
def test_gpu_imbe_decode_from_voice_records(gpu):
    """The bank's packed voice records (nine IMBE frames per LDU) decode like the unpacked imbe_fr arrays."""
    import torch

    L = _o()
    rng = np.random.default_rng(6)
    n = 40
    rec = np.zeros(n, dtype=gpu.P25_VOICE_DTYPE)
    fr = rng.integers(0, 2, (n, 9, 8, 23)).astype(np.uint8)
    for k in range(0, n, 2):
        for v in range(9):
            fr[k, v] = imbe_encode(L, rng.integers(0, 2, 88).astype(np.uint8))
            fr[k, v, rng.integers(0, 7), rng.integers(0, 15)] ^= 1
    fr[:, :, 4:7, 15:] = 0
    fr[:, :, 7, 7:] = 0
    rec["bits"] = (fr.astype(np.uint32) << np.arange(23, dtype=np.uint32)).sum(axis=-1)
    d, c0, tot = gpu.p25p1_voice_imbe_decode(torch.from_numpy(rec.view(np.uint8).reshape(n, -1)).cuda(), n)
    d, c0, tot = d.cpu().numpy(), c0.cpu().numpy(), tot.cpu().numpy()
    for k in range(n):
        for v in range(9):
            wd, w0, wt = _oracle_imbe(L, fr[k, v])
            assert np.array_equal(d[k, v], wd) and c0[k, v] == w0 and tot[k, v] == wt, (k, v)


# ---- the voice burst cutter against the UNMODIFIED dmrBSBootstrap() / dmrBS() (oracle/ref_shim_dmr.c) ----

def _ref_dmr_voice():
    import _harness as HH

    if not os.path.exists(os.path.join(HH.REF_DIR, "libdsdneo_ref_dmr.so")):
        return None
    R = HH.ref_dmr()
    if not hasattr(R, "ref_dmr_voice_run"):
        return None
    R.ref_dmr_voice_run.argtypes = [u8p, u8p, C.c_long, C.c_long, u8p, C.POINTER(C.c_long), C.c_int, C.POINTER(C.c_long)]
    R.ref_dmr_voice_run.restype = C.c_int
    return R


def _qr_16_7_6_encode(data7):
    O = H.oracle_fec()
    for par in range(512):
        w = np.array(list(data7) + [(par >> (8 - i)) & 1 for i in range(9)], np.uint8)
        t = w.copy()
        if O.oracle_qr_16_7_6_decode(t.ctypes.data_as(u8p)) and np.array_equal(t, w):
            return w
    raise AssertionError("no QR(16,7,6) codeword")


def _bs_voice_stream(rng, n_super=3, cc=5):
    """BS stream in dibits: bursts alternate slot 0 / slot 1, both slots carry voice superframes A..F: A has the voice sync,
    B..F a valid EMB (colour code, PI, LCSS under QR(16,7,6)) around 32 arbitrary embedded-signalling bits."""
    L = _o()
    m = np.array([[int(x) for x in e] for e in re.findall(r"\{(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\}",
                  open(REF_MAP).read().split("dsd_ambe_2450_dibit_map[DSD_AMBE_2450_DIBITS] = {")[1])][:36])
    emb = [_qr_16_7_6_encode([(cc >> 3) & 1, (cc >> 2) & 1, (cc >> 1) & 1, cc & 1, 0, (l >> 1) & 1, l & 1]) for l in range(4)]
    parts, sent = [rng.integers(0, 4, 30)], []
    for b in range(12 * n_super):
        slot, vc = b % 2, (b // 2) % 6
        cach = np.zeros(24, np.uint8)
        cach[:7] = H.hamming_7_4_encode_bruteforce((1, slot, 0, 0))
        cach[7:] = rng.integers(0, 2, 17)
        tx = np.array([cach[H.DMR_CACH_INTERLEAVE[i]] for i in range(24)], np.uint8)
        dib = np.zeros(144, np.int64)
        dib[:12] = (tx[0::2] << 1) | tx[1::2]
        frames = rng.integers(0, 2, (3, 49)).astype(np.uint8)
        frs = [ambe_encode(L, d) for d in frames]
        for f, off, cnt, m0 in [(0, 12, 36, 0), (1, 48, 18, 0), (1, 90, 18, 18), (2, 108, 36, 0)]:
            for i in range(cnt):
                hr, hc, lr, lc = m[m0 + i]
                dib[off + i] = (int(frs[f][hr, hc]) << 1) | int(frs[f][lr, lc])
        if vc == 0:
            dib[66:90] = [int(c) for c in "131111333113313313113313"]
        else:
            e = emb[int(rng.integers(0, 4))]
            bits = np.concatenate([e[:8], rng.integers(0, 2, 32).astype(np.uint8), e[8:]])
            dib[66:90] = (bits[0::2] << 1) | bits[1::2]
        parts.append(dib)
        sent.append(frs)
    return np.concatenate(parts).astype(np.uint8), sent


@pytest.mark.skipif(not os.path.exists(REF_MAP), reason="reference tree not present")
def test_voice_cutter_pinned_to_unmodified_dmrbs():
    """Every ambe_fr[4][24] the unmodified dmrBSBootstrap() + dmrBS() loop hands to processMbeFrame on a replayed BS voice stream
    (its colour-code confidence gate opens after a few valid EMBs) equals the oracle cutter's frame for the same burst."""
    R = _ref_dmr_voice()
    if R is None:
        pytest.skip("oracle/_ref/libdsdneo_ref_dmr.so not built")
    L = _o()
    rng = np.random.default_rng(8)
    for inverted in (0, 1):
        dib, sent = _bs_voice_stream(rng)
        stream = dib.copy()
        sync_end = 30 + 89
        if inverted:
            stream[:sync_end + 1] ^= 2  # the hunt's buffer holds the raw dibits; the reference un-inverts the 90 it takes from it
        rel = np.full(stream.size, 200, np.uint8)
        R.ref_dmr_reset(inverted)
        frames = np.zeros((400, 96), np.uint8)
        at = (C.c_long * 400)()
        consumed = C.c_long(0)
        n_calls = R.ref_dmr_voice_run(stream.ctypes.data_as(u8p), rel.ctypes.data_as(u8p), stream.size, sync_end,
                                      frames.ctypes.data_as(u8p), at, 400, C.byref(consumed))
        assert n_calls >= 3 * 12, n_calls
        seen = set()
        for k in range(0, min(n_calls, 400), 3):
            assert at[k] == at[k + 1] == at[k + 2] and (at[k] - 30) % 144 == 0
            j = (at[k] - 30) // 144 - 1  # the burst that ended at this stream position
            cach, fr, sync = np.zeros(24, np.uint8), np.zeros((3, 4, 24), np.uint8), np.zeros(48, np.uint8)
            start = 30 + 144 * j
            L.oracle_dmr_voice_cut(np.ascontiguousarray(stream[start:start + 144]).ctypes.data_as(u8p), int(inverted and j == 0),
                                   cach.ctypes.data_as(u8p), fr.ctypes.data_as(u8p), sync.ctypes.data_as(u8p))
            assert np.array_equal(frames[k:k + 3].reshape(3, 4, 24), fr), (inverted, j, k)
            assert np.array_equal(fr, np.stack(sent[j])), (inverted, j)  # and it is what was transmitted
            seen.add(int(j))
        assert len(seen) >= 12 and {0, 1} & seen or len(seen) >= 12
