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


def test_p25_tsdu_chain_on_device_with_status_symbols(gpu):
    """Real air-interface framing, no host step between slicer and FEC: discriminator samples -> symbolizer -> frame-sync
    hunt -> device frame cutter (NID fields, status symbols stripped) -> p25p1_nid_decode (BCH + Chase) and the three
    half-rate trellis blocks of each TSDU.  The cutter output equals the sequential oracle cutter on the same stream, the
    decoders equal their oracles, and NAC / DUID / TSBK dibits are the transmitted ones."""
    import torch
    from test_oracle_fec import bch_63_16_encode, _oracle_cut

    rng = np.random.default_rng(77)
    n_ch, n_frames, max_hits = 24, 4, 8
    chans = []
    for c in range(n_ch):
        parts, truth = [rng.integers(0, 4, WARMUP)], []
        for f in range(n_frames):
            nac = int(rng.integers(1, 0xFFF))
            frame, payloads = H.p25p1_build_tsdu(rng, nac, 3, bch_63_16_encode)
            parts += [frame, rng.integers(0, 4, GAP)]
            truth.append((nac, payloads))
        chans.append((np.concatenate(parts), truth))
    n_dib = max(ch[0].size for ch in chans)
    noise = [0.0 if c % 3 == 0 else 500.0 + 30.0 * c for c in range(n_ch)]
    xs = np.stack([H.synth_c4fm_disc(rng, np.concatenate([ch[0], rng.integers(0, 4, n_dib - ch[0].size)]), 9000.0, noise[c])
                   for c, ch in enumerate(chans)])
    n = xs.shape[1]
    taps = _taps()
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_P25P1_POS, H.SYNC_P25P1_POS)] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), n)
    fs = gpu.FrameSync(n_ch, [(P25_SYNC, 0)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=max_hits)
    cut = gpu.p25p1_frame_cut(res["dibits"], res["llr"], res["count"], hits, n_hits, 3 * 98)
    torch.cuda.synchronize()
    cut_h = {k: v.cpu().numpy() for k, v in cut.items()}
    dib, llr, cnt = res["dibits"].cpu().numpy(), res["llr"].cpu().numpy(), res["count"].cpu().numpy()
    hits_h, n_hits_h = hits.cpu().numpy(), n_hits.cpu().numpy()
    # ---- cutter == sequential oracle cutter, slot by slot
    slots = []
    for c in range(n_ch):
        assert n_hits_h[c] >= n_frames
        for h in range(max_hits):
            s = c * max_hits + h
            if h >= min(n_hits_h[c], max_hits):
                assert cut_h["nid_valid"][s] == 0 and cut_h["payload_valid"][s] == 0
                continue
            flags, code, rel, par, prel, pd, pl = _oracle_cut(dib[c, :cnt[c]], llr[c, :cnt[c]], int(hits_h[c, h, 0]), 3 * 98)
            assert cut_h["nid_valid"][s] == (flags & 1) and cut_h["payload_valid"][s] == ((flags >> 1) & 1)
            assert np.array_equal(cut_h["nid_code63"][s], code) and np.array_equal(cut_h["nid_reliab63"][s], rel)
            assert cut_h["nid_parity"][s] == par and cut_h["nid_parity_reliab"][s] == prel
            assert np.array_equal(cut_h["payload_dibits"][s], pd) and np.array_equal(cut_h["payload_llr"][s], pl)
            if flags == 3:
                slots.append((c, h, s))
    assert len(slots) >= n_ch * n_frames
    # ---- decoders on the cut frames (device buffers straight from the cutter)
    sel = torch.tensor([s for _, _, s in slots], device="cuda")
    code_d = cut["nid_code63"][sel].contiguous()
    rel_d = cut["nid_reliab63"][sel].contiguous()
    par_d = cut["nid_parity"][sel].contiguous()
    prel_d = cut["nid_parity_reliab"][sel].contiguous()
    k = len(slots)
    st = torch.zeros(k, dtype=torch.int8, device="cuda")
    nac = torch.zeros(k, dtype=torch.int32, device="cuda")
    duid = torch.zeros(k, dtype=torch.uint8, device="cuda")
    errs = torch.zeros(k, dtype=torch.int32, device="cuda")
    gpu.check(gpu.lib().dsdneo_b200_p25p1_nid_decode_batch(code_d.data_ptr(), rel_d.data_ptr(), None, par_d.data_ptr(),
                                                           prel_d.data_ptr(), 64, st.data_ptr(), nac.data_ptr(), duid.data_ptr(),
                                                           errs.data_ptr(), k, None))
    blocks = cut["payload_llr"][sel].reshape(k * 3, 196).contiguous()
    out12 = torch.zeros((k * 3, 12), dtype=torch.uint8, device="cuda")
    met = torch.zeros(k * 3, dtype=torch.int32, device="cuda")
    gpu.check(gpu.lib().dsdneo_b200_p25_12_soft_llr_batch(blocks.data_ptr(), out12.data_ptr(), met.data_ptr(), k * 3, None))
    torch.cuda.synchronize()
    st, nac, duid, out12 = st.cpu().numpy(), nac.cpu().numpy(), duid.cpu().numpy(), out12.cpu().numpy()
    # ---- every transmitted frame is among the decoded ones, with its NAC and TSBK dibits
    recovered = 0
    for c in range(n_ch):
        mine = [(i, s) for i, (cc, h, s) in enumerate(slots) if cc == c]
        for nac_tx, payloads in chans[c][1]:
            for i, s in mine:
                if st[i] == 1 and nac[i] == nac_tx and duid[i] == 7:
                    ok = True
                    for b in range(3):
                        want = np.zeros(12, np.uint8)
                        for j in range(48):
                            want[j // 4] |= int(payloads[b][j]) << (6 - 2 * (j % 4))
                        ok = ok and np.array_equal(out12[3 * i + b], want)
                    recovered += ok
                    break
    assert recovered == n_ch * n_frames, recovered
