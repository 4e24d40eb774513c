"""The oracle's restatement of the P25 Phase 1 frame handlers (oracle/oracle_p25p1_frame.c: NID, TSBK, HDU, LDU1, LDU2 incl.
status-symbol stripping, IMBE de-interleave, word-level hard + soft FEC, Reed-Solomon hard + ranked erasures, LSD) against
the UNMODIFIED reference handlers replaying the same dibit stream (oracle/_ref/libdsdneo_ref_p25.so = dispatch_p25p1.c +
p25p1_{hdu,ldu,ldu1,ldu2,tsbk}.c, FEC leaves recorded through ld --wrap) and against the committed golden records."""
import os

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_p25_available(), reason="oracle/_ref/libdsdneo_ref_p25.so not built")
SYNC = np.array(H.P25P1_SYNC_DIBITS, np.uint8)


def _find_syncs(dibits):
    n = dibits.size
    win = np.lib.stride_tricks.sliding_window_view(dibits, 24)
    return [int(i) + 23 for i in np.nonzero((win == SYNC).all(axis=1))[0] if i + 24 < n]


def _soft_from_dibits(rng, dibits, flip=0.0, weak=0.15, coarse=False):
    """A plausible slicer output for a dibit stream: LLR sign = bit, magnitude high, a fraction of weak / flipped bits.
    coarse: magnitudes 1..3 only, so equal path / penalty metrics (tie-break rules) are exercised everywhere."""
    n = dibits.size
    bits = np.stack([(dibits >> 1) & 1, dibits & 1], axis=1).astype(np.int64)
    mag = rng.integers(120, 256, (n, 2)) if not coarse else rng.integers(1, 4, (n, 2))
    if not coarse:
        mag = np.where(rng.random((n, 2)) < weak, rng.integers(0, 60, (n, 2)), mag)
    err = rng.random((n, 2)) < flip
    if not coarse:
        mag = np.where(err, rng.integers(0, 90, (n, 2)), mag)  # wrong bits tend to be weak
    rx = bits ^ err
    llr = np.where(rx == 1, mag, -mag).astype(np.int16)
    d = (rx[:, 0] * 2 + rx[:, 1]).astype(np.uint8)
    rel = np.minimum(np.abs(llr).min(axis=1), 255).astype(np.uint8)
    return d, rel, llr


def _compare_stream(d, rel, llr, observed_nac=0, min_frames=1, positions=None):
    n_ok, kinds = 0, {}
    for p in (positions if positions is not None else _find_syncs(d)):
        ref = H.ref_p25_decode(d, rel, llr, p, observed_nac)
        n, f, v = H.oracle_p25_decode(d, llr, p, observed_nac)
        if ref["overrun"]:
            assert n == -1 or True
            continue
        bad = H.p25_frames_agree(ref, f, v)
        assert not bad, (p, bad, int(ref["duid"]), int(ref["nid_status"]))
        if ref["nid_status"] > 0 and int(ref["duid"]) in (0, 5, 7, 10, 15):
            assert n == int(ref["consumed"]), (p, n, int(ref["consumed"]))
        kinds[int(f["duid"])] = kinds.get(int(f["duid"]), 0) + 1
        n_ok += 1
    assert n_ok >= min_frames
    return kinds


@needs_ref
@pytest.mark.parametrize("flip,coarse", [(0.0, False), (0.01, False), (0.03, False), (0.06, False), (0.02, True), (0.05, True)])
def test_synthetic_frames_oracle_equals_reference_handlers(flip, coarse):
    rng = np.random.default_rng(int(flip * 1000) + 5 + 100 * coarse)
    parts = [rng.integers(0, 4, 100)]
    positions, at = [], 100
    for rep in range(3):
        nac = int(rng.integers(1, 0xFFE))
        for build in (lambda: H.p25p1_build_hdu(rng, nac)[0], lambda: H.p25p1_build_ldu(rng, nac, False)[0],
                      lambda: H.p25p1_build_ldu(rng, nac, True)[0], lambda: H.p25p1_build_tdulc(rng, nac)[0],
                      lambda: H.p25p1_build_tsdu(rng, nac, int(rng.integers(1, 4)), H._bch_nid_encoder(), valid_crc=rep != 1)[0]):
            frame, gap = build(), rng.integers(0, 4, int(rng.integers(0, 20)))
            positions.append(at + 23)
            at += frame.size + gap.size
            parts += [frame, gap]
    parts.append(rng.integers(0, 4, 900))
    tx = np.concatenate(parts).astype(np.uint8)
    d, rel, llr = _soft_from_dibits(rng, tx, flip=flip, coarse=coarse)
    kinds = _compare_stream(d, rel, llr, min_frames=15, positions=positions)
    if flip == 0.0:
        assert kinds.get(0, 0) == 3 and kinds.get(5, 0) == 3 and kinds.get(10, 0) == 3 and kinds.get(7, 0) == 3, kinds
        assert kinds.get(15, 0) == 3, kinds


def test_clean_frames_decode_to_the_transmitted_payloads():
    rng = np.random.default_rng(99)
    for build in (lambda: H.p25p1_build_hdu(rng, 0x293), lambda: H.p25p1_build_ldu(rng, 0x293, False),
                  lambda: H.p25p1_build_ldu(rng, 0x293, True), lambda: H.p25p1_build_tdulc(rng, 0x293)):
        frame, truth = build()
        tx = np.concatenate([rng.integers(0, 4, 40), frame, rng.integers(0, 4, 40)]).astype(np.uint8)
        d, rel, llr = _soft_from_dibits(rng, tx, flip=0.0, weak=0.0)
        n, f, v = H.oracle_p25_decode(d, llr, 40 + 23)
        assert n > 0 and f["nid_status"] == 1 and f["nac"] == 0x293 and f["duid"] == truth["duid"]
        k = truth["rs_data"].size
        assert f["rs_status"] == 0 and np.array_equal(f["rs_data"][:k], truth["rs_data"])
        if truth["duid"] in (5, 10):
            bits = ((v["bits"][:, :, None] >> np.arange(23, dtype=np.uint32)) & 1).reshape(9, 184)
            assert np.array_equal(bits, truth["voice"]) and f["lsd_ok"] == 3 and np.array_equal(f["lsd"], truth["lsd"])


@pytest.mark.parametrize("name", ["c1_p25p1_c4fm_cc", "c1_p25p1_c4fm_vc"])
def test_fixture_frames_oracle_equals_golden_reference_records(name):
    """The reference's own captures: every frame the unmodified handlers decoded from the unmodified slicer's dibits (golden,
    tests/golden/make_c1_golden.py) is reproduced by the oracle from the same dibits."""
    path = os.path.join(H.GOLDEN_DIR, name + ".npz")
    g = np.load(path)
    d, llr = g["dibits"], g["llr"]
    pos, recs = g["frame_pos"], g["frame_ref"].view(H.REF_P25_DTYPE).reshape(-1)
    assert pos.size >= 8
    for p, ref in zip(pos, recs):
        n, f, v = H.oracle_p25_decode(d, llr, int(p), 0)
        bad = H.p25_frames_agree(ref, f, v)
        assert not bad, (name, int(p), bad)
    nacs = {int(r["nac"]) for r in recs if r["nid_status"] == 1}
    assert int(g["expected_nac"]) in nacs


def _plain_final_state(llr196):
    """The best final state of p25_12_soft_llr (lowest state wins ties), by a direct restatement of its recursion."""
    tbl = []
    for g in range(4):
        for j in range(2 * g, 98, 8):
            tbl += [j, j + 1]
    dei = np.zeros(196, np.int64)
    for i in range(98):
        dei[2 * tbl[i]], dei[2 * tbl[i] + 1] = llr196[2 * i], llr196[2 * i + 1]
    dtm = [2, 12, 1, 15, 14, 0, 13, 3, 9, 7, 10, 4, 5, 11, 6, 8]
    bc = lambda l, bit: (-l if l < 0 else 0) if bit else (l if l > 0 else 0)
    pm = [0, 256, 256, 256]
    for i in range(49):
        l = [int(v) for v in dei[4 * i:4 * i + 4]]
        pm = [min(pm[pv] + sum(bc(l[k], (dtm[pv * 4 + nx] >> (3 - k)) & 1) for k in range(4)) for pv in range(4)) for nx in range(4)]
    return min(range(4), key=lambda s: (pm[s], s))


def test_list_decoder_candidate_zero_is_the_plain_path_when_it_ends_in_state_zero():
    """What the device TSBK shortcut relies on: when p25_12_soft_llr's best path ends in state 0 (every valid block does: its
    49th dibit is the flush), candidate 0 of p25_12_soft_llr_list is that path, ties included (coarse LLRs make equal path
    metrics common).  For other final states the list's de-duplication can shadow it -- also exercised here."""
    import ctypes as C

    O = H.oracle_fec()
    rng = np.random.default_rng(12)
    n_zero = n_shadowed = 0
    for trial in range(1200):
        scale = [1, 2, 3, 40, 300][trial % 5]
        if trial % 2 == 0:  # a valid block under noise: mostly ends in state 0
            tx = H.p25_trellis_encode(rng)[1]
            bits = np.stack([(tx >> 1) & 1, tx & 1], axis=1).reshape(-1)
            llr = (np.where(bits == 1, scale, -scale) + rng.integers(-scale, scale + 1, 196)).astype(np.int16)
        else:
            llr = rng.integers(-scale, scale + 1, 196).astype(np.int16)
        if trial % 7 == 0:
            llr[rng.integers(0, 196, 60)] = 0
        plain = np.zeros(12, np.uint8)
        O.oracle_p25_12_soft_llr(llr.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(plain, H.u8p))
        cand = np.zeros(8 * 12, np.uint8)
        met = np.zeros(8, np.uint32)
        n = O.oracle_p25_12_soft_llr_list(llr.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(cand, H.u8p),
                                          met.ctypes.data_as(C.POINTER(C.c_uint32)), 8)
        assert n >= 1
        if _plain_final_state(llr) == 0:
            assert np.array_equal(cand[:12], plain), trial
            n_zero += 1
        else:
            n_shadowed += not np.array_equal(cand[:12], plain)
    assert n_zero > 300 and n_shadowed > 0, (n_zero, n_shadowed)
