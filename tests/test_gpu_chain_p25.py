"""Config C3 in miniature, end to end on the GPU: synthetic P25 Phase 1 C4FM channels (discriminator level) ->
matched filter + getSymbol + slicer (symbolizer) -> frame-sync hunt -> half-rate trellis (soft LLRs) and RS(36,20,17),
every stage bit-exact against the CPU oracle chain on the same samples, and the transmitted payloads recovered."""
import ctypes as C

import numpy as np
import pytest

import _harness as H
from test_frame_sync import oracle_search
from test_gpu_symbolizer import _oracle_dibits, _taps
from test_oracle_fec import _rs_words

pytestmark = pytest.mark.gpu

P25_SYNC = "111113113311333313133333"
WARMUP, GAP, N_FRAMES = 240, 16, 6
FRAME = 24 + 98 + 108 + GAP


def build_channel(rng):
    """Returns (dibits, payloads): payloads[i] = (49 trellis input dibits, 20 RS data symbols) of frame i."""
    O = H.oracle_fec()
    parts, payloads = [rng.integers(0, 4, WARMUP)], []
    for _ in range(N_FRAMES):
        d49, tx98 = H.p25_trellis_encode(rng)
        data = np.zeros(47, np.int32)
        data[:20] = rng.integers(0, 64, 20)
        cw = np.zeros(63, np.int32)
        O.oracle_rs63_encode(8, data.ctypes.data_as(H.i32p), cw.ctypes.data_as(H.i32p))
        hexw = cw[:36]  # 16 parity + 20 data hex words, 3 dibits each, MSB first
        rs_dibits = np.stack([(hexw >> 4) & 3, (hexw >> 2) & 3, hexw & 3], axis=1).reshape(-1)
        parts += [np.array([int(c) for c in P25_SYNC]), tx98, rs_dibits, rng.integers(0, 4, GAP)]
        payloads.append((d49, data[:20].copy()))
    return np.concatenate(parts), payloads


def cut_frames(hit_pos, dibits, llr, count):
    """Host-side frame cutter: for each sync that leaves room for a whole frame, the trellis LLRs and the RS hex-word bits."""
    blocks, rs_data, rs_par, keep = [], [], [], []
    for p in hit_pos:
        if p + 1 + 98 + 108 > count:
            continue
        blocks.append(llr[p + 1:p + 99].reshape(-1))
        d = dibits[p + 99:p + 207].astype(np.int64)
        bits = np.stack([(d >> 1) & 1, d & 1], axis=1).reshape(-1).astype(np.uint8)  # 216 bits = 36 hex words
        rs_par.append(bits[:96])
        rs_data.append(bits[96:])
        keep.append(p)
    return np.array(blocks, np.int16), np.array(rs_data, np.uint8), np.array(rs_par, np.uint8), keep


def test_p25_c4fm_chain_bit_exact_and_payloads_recovered(gpu):
    import torch

    rng = np.random.default_rng(2026)
    n_ch = 48
    chans = [build_channel(rng) for _ in range(n_ch)]
    noise = [0.0 if c % 3 == 0 else 600.0 + 25.0 * c for c in range(n_ch)]
    xs = np.stack([H.synth_c4fm_disc(rng, chans[c][0], 9000.0, noise[c]) for c in range(n_ch)])
    n = xs.shape[1]
    taps = _taps()
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_P25P1_POS, H.SYNC_P25P1_POS)] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), n)
    fs = gpu.FrameSync(n_ch, [(P25_SYNC, 0)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=32)
    torch.cuda.synchronize()
    cnt = res["count"].cpu().numpy()
    dib, llr = res["dibits"].cpu().numpy(), res["llr"].cpu().numpy()
    sym = res["symbols"].cpu().numpy()
    hits, n_hits = hits.cpu().numpy(), n_hits.cpu().numpy()
    O = H.oracle_fec()
    all_blocks, all_rsd, all_rsp, owner = [], [], [], []
    want_bytes, want_rs, want_rs_rc = [], [], []
    for c in range(n_ch):
        # ---- CPU oracle chain on the same samples
        wd, wr, wl, ws = _oracle_dibits(xs[c], H.SYNC_P25P1_POS, taps)
        assert cnt[c] == wd.size and np.array_equal(dib[c, :cnt[c]], wd) and np.array_equal(llr[c, :cnt[c]], wl)
        assert H.bits_equal(sym[c, :cnt[c]], ws)
        on, opos, otyp, _, _ = oracle_search(ws, [(P25_SYNC, 0)], max_hits=32)
        assert n_hits[c] == on and np.array_equal(hits[c, :on, 0], opos)
        blocks, rsd, rsp, keep = cut_frames(opos.tolist(), wd, wl, wd.size)
        assert len(keep) == N_FRAMES, (c, len(keep))  # every transmitted sync was found where a whole frame follows
        for b in range(len(keep)):
            out12 = np.zeros(12, np.uint8)
            O.oracle_p25_12_soft_llr(blocks[b].ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(out12, H.u8p))
            want_bytes.append(out12)
            d = rsd[b].copy()
            want_rs_rc.append(O.oracle_p25_rs_decode(36, 20, H._ptr(d, H.u8p), H._ptr(rsp[b].copy(), H.u8p)))
            want_rs.append(d)
        # ---- the GPU chain cuts its own frames from its own outputs
        gb, gd, gp, gkeep = cut_frames(hits[c, :n_hits[c], 0].tolist(), dib[c], llr[c], cnt[c])
        assert gkeep == keep
        all_blocks.append(gb); all_rsd.append(gd); all_rsp.append(gp); owner += [(c, i) for i in range(len(gkeep))]
    blocks, rsd, rsp = np.concatenate(all_blocks), np.concatenate(all_rsd), np.concatenate(all_rsp)
    got12, metric = gpu.p25_12_soft_llr(blocks)
    got_rs = np.ascontiguousarray(rsd)
    rs_rc = gpu.p25_rs_decode(0, got_rs, rsp)  # corrects in place
    assert np.array_equal(got12, np.array(want_bytes)) and np.array_equal(got_rs, np.array(want_rs))
    assert np.array_equal(rs_rc, np.array(want_rs_rc, np.uint8))
    # ---- payloads: 48 data dibits -> 12 bytes MSB first; RS data symbols
    ok_t = ok_r = 0
    for k, (c, i) in enumerate(owner):
        d49, rs20 = chans[c][1][i]
        by = np.zeros(12, np.uint8)
        for j in range(48):
            by[j // 4] |= int(d49[j]) << (6 - 2 * (j % 4))
        ok_t += np.array_equal(got12[k], by)
        ok_r += rs_rc[k] == 0 and np.array_equal(got_rs[k], _rs_words(rs20))
    assert ok_t == len(owner) and ok_r == len(owner), (ok_t, ok_r, len(owner))
