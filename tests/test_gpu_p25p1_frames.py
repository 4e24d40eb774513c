"""The device P25 Phase 1 frame decoder (dsdneo_b200_p25p1_frames_decode_batch: TSBK / HDU / LDU1 / LDU2 incl. status-symbol
stripping, IMBE de-interleave, word-level hard + soft FEC, Reed-Solomon hard + ranked erasures, LSD) against the oracle's
restatement of the reference handlers on synthetic multi-channel streams, and against the golden records of the UNMODIFIED
reference handlers on the reference's own captures (tests/golden/c1_p25p1_c4fm_{cc,vc}.npz)."""
import os

import numpy as np
import pytest

import _harness as H
from test_oracle_p25p1_frame import _soft_from_dibits

pytestmark = pytest.mark.gpu


def _run(gpu, streams, positions, max_hits=32, observed_nac=None):
    """streams: list of (dibits, llr) per channel; positions: list of lists of last-sync-dibit indices."""
    import torch

    n_ch = len(streams)
    pitch = max(d.size for d, _ in streams) + 8
    dib = np.zeros((n_ch, pitch), np.uint8)
    llr = np.zeros((n_ch, pitch, 2), np.int16)
    cnt = np.zeros(n_ch, np.int32)
    hits = np.zeros((n_ch, max_hits, 2), np.int32)
    n_hits = np.zeros(n_ch, np.int32)
    for c, (d, l) in enumerate(streams):
        dib[c, :d.size], llr[c, :d.size], cnt[c] = d, l, d.size
        n_hits[c] = len(positions[c])
        hits[c, :min(len(positions[c]), max_hits), 0] = positions[c][:max_hits]
    t = lambda a: torch.from_numpy(a).cuda()
    return gpu.p25p1_frames_decode(t(dib), t(llr), t(cnt), t(hits), t(n_hits), observed_nac=observed_nac)


def _voice_of(frames, voices, i):
    vi = int(frames[i]["voice_index"])
    return voices[vi] if vi >= 0 else np.zeros(1, H.P25_VOICE_DTYPE)[0]


@pytest.mark.parametrize("flip,coarse", [(0.0, False), (0.02, False), (0.05, False), (0.02, True), (0.06, True)])
def test_frames_equal_the_oracle_on_synthetic_channels(gpu, flip, coarse):
    rng = np.random.default_rng(int(flip * 100) + 11 + 50 * coarse)
    n_ch = 24
    streams, positions = [], []
    for c in range(n_ch):
        parts, pos, at = [rng.integers(0, 4, 30 + c)], [], 30 + c
        nac = int(rng.integers(1, 0xFFE))
        builders = [lambda: H.p25p1_build_hdu(rng, nac)[0], lambda: H.p25p1_build_ldu(rng, nac, False)[0],
                    lambda: H.p25p1_build_ldu(rng, nac, True)[0], lambda: H.p25p1_build_tdulc(rng, nac)[0],
                    lambda: H.p25p1_build_tsdu(rng, nac, int(rng.integers(1, 4)), H._bch_nid_encoder(), valid_crc=c % 3 != 0)[0]]
        for k in rng.permutation(10):
            frame, gap = builders[k % 5](), rng.integers(0, 4, int(rng.integers(0, 12)))
            pos.append(at + 23)
            at += frame.size + gap.size
            parts += [frame, gap]
        # one frame cut short by the end of the stream, one bogus hit in noise
        frame = builders[c % 5]()
        pos.append(at + 23)
        parts.append(frame[:frame.size // 2])
        tx = np.concatenate(parts).astype(np.uint8)
        pos.insert(0, 25)  # a "sync" inside the leading noise: NID fails or decodes to garbage, both must match the oracle
        d, rel, llr = _soft_from_dibits(rng, tx, flip=flip, coarse=coarse)
        streams.append((d, llr))
        positions.append(pos)
    frames, voices = _run(gpu, streams, positions)
    assert frames.size == sum(len(p) for p in positions)
    k, kinds = 0, {}
    for c in range(n_ch):
        d, llr = streams[c]
        for h, p in enumerate(positions[c]):
            f = frames[k]
            assert f["channel"] == c and f["position"] == p
            n, of, ov = H.oracle_p25_decode(d, llr, p, 0)
            if n == -1 and of["nid_status"] > 0:  # the oracle ran out of stream inside the payload: the device flags the same frame
                assert f["reserved"][0] == 1 and f["nid_status"] == of["nid_status"] and f["duid"] == of["duid"], (c, h)
                if of["duid"] == 7:  # the blocks that were complete are still delivered
                    nb = int(f["n_tsbk"])
                    assert np.array_equal(f["tsbk"][:nb], of["tsbk"][:nb])
            elif n == -1:
                assert f["nid_status"] == 0 or f["reserved"][0] == 1
            else:
                gv = _voice_of(frames, voices, k)
                for name in H.P25_FRAME_DTYPE.names:
                    if name in ("position", "channel", "voice_index", "reserved"):
                        continue
                    assert np.array_equal(f[name], of[name]), (c, h, p, name, f[name], of[name], int(of["duid"]))
                if of["duid"] in (5, 10):
                    assert f["voice_index"] >= 0 and np.array_equal(gv["bits"], ov["bits"]) and np.array_equal(gv["reliab"], ov["reliab"])
                else:
                    assert f["voice_index"] == -1
                kinds[int(of["duid"])] = kinds.get(int(of["duid"]), 0) + 1
            k += 1
    if flip == 0.0:
        assert all(kinds.get(d, 0) >= 2 * n_ch for d in (0, 5, 7, 10, 15)), kinds


@pytest.mark.parametrize("name", ["c1_p25p1_c4fm_cc", "c1_p25p1_c4fm_vc"])
def test_fixture_frames_equal_the_unmodified_reference_handlers(gpu, name):
    g = np.load(os.path.join(H.GOLDEN_DIR, name + ".npz"))
    d, llr = g["dibits"], g["llr"]
    pos, recs = g["frame_pos"], g["frame_ref"].view(H.REF_P25_DTYPE).reshape(-1)
    # the same stream on 3 channels with different hit subsets (slot / scan bookkeeping)
    positions = [list(map(int, pos)), list(map(int, pos[::2])), list(map(int, pos[1:]))]
    frames, voices = _run(gpu, [(d, llr)] * 3, positions)
    assert frames.size == sum(len(p) for p in positions)
    by_pos = {int(p): r for p, r in zip(pos, recs)}
    decoded = 0
    for i, f in enumerate(frames):
        bad = H.p25_frames_agree(by_pos[int(f["position"])], f, _voice_of(frames, voices, i))
        assert not bad, (name, int(f["channel"]), int(f["position"]), bad)
        decoded += int(f["nid_status"] == 1 and f["nac"] == int(g["expected_nac"]))
    assert decoded >= 8
