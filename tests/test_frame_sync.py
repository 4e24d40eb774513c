"""Frame-sync hunt: oracle semantics on CPU, CUDA kernel vs oracle on the GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _harness as H

REF_PATTERNS_H = "/root/reference/include/dsd-neo/core/sync_patterns.h"
REF_IDS_H = "/root/reference/include/dsd-neo/core/synctype_ids.h"


def _default_patterns():
    import __graft_entry__ as g

    return g.load_package().DEFAULT_SYNC_PATTERNS


def oracle_search(x, patterns, hist=None, count=None, max_hits=256):
    O = H.oracle_sym()
    O.oracle_frame_sync_search.restype = C.c_int
    pats = (C.c_char_p * len(patterns))(*[p.encode() for p, _ in patterns])
    ids = (C.c_int * len(patterns))(*[t for _, t in patterns])
    hist = C.create_string_buffer(b"\0" * 32, 32) if hist is None else hist
    count = C.c_int(0) if count is None else count
    pos, typ = np.zeros(max_hits, np.int32), np.zeros(max_hits, np.int32)
    x = np.ascontiguousarray(x, np.float32)
    n = O.oracle_frame_sync_search(H._ptr(x), x.size, pats, ids, len(patterns), hist, C.byref(count),
                                   pos.ctypes.data_as(H.i32p), typ.ctypes.data_as(H.i32p), max_hits)
    k = min(n, max_hits)
    return n, pos[:k].copy(), typ[:k].copy(), hist, count


@pytest.mark.skipif(not os.path.exists(REF_PATTERNS_H), reason="reference tree not present")
def test_default_patterns_are_the_reference_strings():
    """Every default pattern is a #define of include/dsd-neo/core/sync_patterns.h and its id a DSD_SYNC_* of synctype_ids.h."""
    defs = dict(re.findall(r'#define\s+(\w+)\s+"([13]+)"', open(REF_PATTERNS_H).read()))
    ids = {k: int(v) for k, v in re.findall(r"#define\s+(DSD_SYNC_\w+)\s+(\d+)\b", open(REF_IDS_H).read())}
    want = {"P25P1_SYNC": "DSD_SYNC_P25P1_POS", "INV_P25P1_SYNC": "DSD_SYNC_P25P1_NEG", "DMR_BS_DATA_SYNC": "DSD_SYNC_DMR_BS_DATA_POS",
            "DMR_BS_VOICE_SYNC": "DSD_SYNC_DMR_BS_VOICE_POS", "DMR_MS_DATA_SYNC": "DSD_SYNC_DMR_MS_DATA",
            "DMR_MS_VOICE_SYNC": "DSD_SYNC_DMR_MS_VOICE", "X2TDMA_BS_DATA_SYNC": "DSD_SYNC_X2TDMA_DATA_POS",
            "X2TDMA_BS_VOICE_SYNC": "DSD_SYNC_X2TDMA_VOICE_POS", "FUSION_SYNC": "DSD_SYNC_YSF_POS", "INV_FUSION_SYNC": "DSD_SYNC_YSF_NEG"}
    got = {p: t for p, t in _default_patterns()}
    assert len(got) == len(want)
    for name, sync in want.items():
        assert got[defs[name]] == ids[sync], name


def test_oracle_frame_sync_semantics():
    """symbol > 0 -> '1' else '3'; a pattern fires exactly when its last character arrives, never before enough symbols
    have been seen, and history carries across calls (a pattern split over two calls is still found)."""
    pats = _default_patterns()
    rng = np.random.default_rng(3)
    p25 = np.array([1.0 if c == "1" else -1.0 for c in pats[0][0]], np.float32) * 9000
    dmr = np.array([1.0 if c == "1" else -1.0 for c in pats[2][0]], np.float32) * 7000
    x = np.concatenate([p25, rng.choice([-3000.0, 4000.0], 100), dmr, [0.0] * 5, p25[:-1], [1.0]]).astype(np.float32)
    n, pos, typ, _, _ = oracle_search(x, pats)
    hits = list(zip(pos.tolist(), typ.tolist()))
    assert (23, 0) in hits and (24 + 100 + 23, 10) in hits
    assert not any(p > 24 + 100 + 24 and t == 0 for p, t in hits)  # last symbol flipped: no P25 sync at the end
    # split across two calls
    n1, pos1, typ1, hist, cnt = oracle_search(x[:10], pats)
    n2, pos2, typ2, _, _ = oracle_search(x[10:], pats, hist, cnt)
    assert n1 + n2 == n and [(p + 10, t) for p, t in zip(pos2.tolist(), typ2.tolist())] == hits[n1:]
    # fewer symbols than the pattern length never match, even if the zero-initialised history would compare equal
    n, _, _, _, _ = oracle_search(p25[:23], pats)
    assert n == 0


@pytest.mark.gpu
def test_frame_sync_kernel_bit_exact(gpu):
    """Batched kernel == oracle: random sign streams with embedded sync words, ragged lengths per channel, history
    carried over three launches, hit-list overflow reported through n_hits."""
    import torch

    pats = gpu.DEFAULT_SYNC_PATTERNS
    rng = np.random.default_rng(17)
    n_ch, n = 37, 1500
    x = rng.normal(0, 5000, (n_ch, 3 * n)).astype(np.float32)
    for c in range(n_ch):
        for _ in range(int(rng.integers(0, 12))):
            s, _ = pats[int(rng.integers(0, len(pats)))]
            at = int(rng.integers(0, 3 * n - 40))
            x[c, at:at + len(s)] = np.array([6000.0 if ch == "1" else -6000.0 for ch in s], np.float32)
    x[3] = np.where(np.arange(3 * n) % 2 == 0, 1.0, -1.0)  # alternating: exercises no-hit path
    x[4] = np.tile(np.array([6000.0 if ch == "1" else -6000.0 for ch in pats[0][0]], np.float32), 3 * n // 24 + 1)[:3 * n]
    fs = gpu.FrameSync(n_ch, pats)
    lens = [rng.integers(n // 2, n + 1, n_ch).astype(np.int32) for _ in range(3)]
    lens[1][5] = 0
    lens[2][6] = 7
    state = [(None, None)] * n_ch
    offs = np.zeros(n_ch, np.int64)
    for launch in range(3):
        tile = np.zeros((n_ch, n), np.float32)
        for c in range(n_ch):
            tile[c, :lens[launch][c]] = x[c, offs[c]:offs[c] + lens[launch][c]]
        hits, n_hits = fs.search(torch.from_numpy(tile).cuda(), torch.from_numpy(lens[launch]).cuda(), max_hits=16)
        torch.cuda.synchronize()
        hits, n_hits = hits.cpu().numpy(), n_hits.cpu().numpy()
        for c in range(n_ch):
            want_n, pos, typ, hist, cnt = oracle_search(tile[c, :lens[launch][c]], pats, state[c][0], state[c][1], max_hits=16)
            state[c] = (hist, cnt)
            assert n_hits[c] == want_n, (launch, c)
            k = min(want_n, 16)
            assert np.array_equal(hits[c, :k, 0], pos[:k]) and np.array_equal(hits[c, :k, 1], typ[:k]), (launch, c)
            offs[c] += lens[launch][c]
    assert state[4][1].value == 32
    with pytest.raises(gpu.B200Error):
        gpu.FrameSync(2, [("1312", 0)])
