"""Configuration C4 in miniature: synthetic DMR-class 4FSK channels (discriminator level) -> RRC matched filter + getSymbol +
fixed-threshold slicer -> frame-sync hunt (DMR BS data sync) -> BPTC(196,96), every stage bit-exact against the CPU
oracle chain and the payloads recovered; then the MBE synthesis stage on the side (parity unpinned, +-1 LSB vs its oracle,
tests/test_mbe.py)."""
import ctypes as C

import numpy as np
import pytest

import _harness as H
from test_frame_sync import oracle_search
from test_gpu_symbolizer import _oracle_dibits, _taps

pytestmark = pytest.mark.gpu

DMR_SYNC = "313333111331131131331131"
WARMUP, GAP, N_FRAMES = 120, 14, 5


def test_bptc_test_encoder_makes_valid_codewords():
    O = H.oracle_fec()
    rng = np.random.default_rng(1)
    for _ in range(5):
        payload = rng.integers(0, 2, 96).astype(np.uint8)
        dei = H.bptc_196x96_encode(payload, interleave=False)
        out, r3, und = np.zeros(96, np.uint8), np.zeros(3, np.uint8), C.c_int(0)
        assert O.oracle_bptc_196x96_extract(H._ptr(dei, H.u8p), H._ptr(out, H.u8p), H._ptr(r3, H.u8p), C.byref(und)) == 0
        assert np.array_equal(out, payload)


def test_dmr_chain_bit_exact_and_payloads_recovered(gpu):
    import torch

    rng = np.random.default_rng(77)
    n_ch = 24
    taps = _taps()
    chans, xs = [], []
    for c in range(n_ch):
        parts, payloads = [rng.integers(0, 4, WARMUP)], []
        for _ in range(N_FRAMES):
            payload = rng.integers(0, 2, 96).astype(np.uint8)
            bits = H.bptc_196x96_encode(payload)
            parts += [np.array([int(ch) for ch in DMR_SYNC]), (bits[0::2] << 1) | bits[1::2], rng.integers(0, 4, GAP)]
            payloads.append(payload)
        dib = np.concatenate(parts)
        chans.append(payloads)
        xs.append(H.synth_dmr_disc(rng, dib, taps[1], 10000.0, 0.0 if c % 2 == 0 else 500.0 + 40.0 * c))
    xs = np.stack(xs)
    n = xs.shape[1]
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS)] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), n)
    fs = gpu.FrameSync(n_ch, [(DMR_SYNC, 10)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=32)
    torch.cuda.synchronize()
    cnt, dib_g = res["count"].cpu().numpy(), res["dibits"].cpu().numpy()
    sym_g, hits, n_hits = res["symbols"].cpu().numpy(), hits.cpu().numpy(), n_hits.cpu().numpy()
    O = H.oracle_fec()
    bursts, want_out, want_err, owner = [], [], [], []
    for c in range(n_ch):
        wd, wr, wl, ws = _oracle_dibits(xs[c], H.SYNC_DMR_BS_DATA_POS, taps)
        assert cnt[c] == wd.size and np.array_equal(dib_g[c, :cnt[c]], wd) and H.bits_equal(sym_g[c, :cnt[c]], ws)
        on, opos, otyp, _, _ = oracle_search(ws, [(DMR_SYNC, 10)], max_hits=32)
        assert n_hits[c] == on and np.array_equal(hits[c, :on, 0], opos) and (hits[c, :on, 1] == 10).all()
        frames = [p for p in opos.tolist() if p + 1 + 98 <= wd.size]
        assert len(frames) >= N_FRAMES, (c, len(frames))
        for i, p in enumerate(frames):
            d = dib_g[c, p + 1:p + 99].astype(np.uint8)
            bits = np.stack([(d >> 1) & 1, d & 1], axis=1).reshape(-1).astype(np.uint8)
            bursts.append(bits)
            dei, out, r3, und = np.zeros(196, np.uint8), np.zeros(96, np.uint8), np.zeros(3, np.uint8), C.c_int(0)
            O.oracle_bptc_deinterleave(H._ptr(bits, H.u8p), H._ptr(dei, H.u8p))
            want_err.append(O.oracle_bptc_196x96_extract(H._ptr(dei, H.u8p), H._ptr(out, H.u8p), H._ptr(r3, H.u8p), C.byref(und)))
            want_out.append(out)
            owner.append((c, i))
    out, r3, errs = gpu.bptc_196x96(np.array(bursts), interleaved=True)
    assert np.array_equal(out, np.array(want_out)) and np.array_equal(errs, np.array(want_err, np.uint32))
    good = sum(int(errs[k] == 0 and i < N_FRAMES and np.array_equal(out[k], chans[c][i])) for k, (c, i) in enumerate(owner))
    assert good >= n_ch * N_FRAMES - 2, (good, len(owner))  # every transmitted payload (a late false sync may add extras)


def test_dmr_bs_data_bursts_on_device(gpu):
    """Real DMR BS data burst framing (CACH + info + slot type around the sync), no host step between slicer and FEC:
    symbolizer -> BS DATA sync hunt -> device burst cutter -> BPTC(196,96), Golay(20,8) slot type, Hamming(7,4) TACT.
    The cutter equals the sequential oracle cutter on the same stream; payload, colour code and data type are recovered."""
    import torch

    rng = np.random.default_rng(99)
    n_ch, n_bursts, max_hits = 16, 4, 8
    taps = _taps()
    chans, xs = [], []
    for c in range(n_ch):
        parts, sent = [rng.integers(0, 4, WARMUP)], []
        for _ in range(n_bursts):
            payload = rng.integers(0, 2, 96).astype(np.uint8)
            cc, dt = int(rng.integers(0, 16)), int(rng.integers(0, 11))
            burst, _ = H.dmr_build_bs_data_burst(rng, payload, cc, dt)
            parts += [burst, rng.integers(0, 4, GAP)]
            sent.append((payload, cc, dt))
        chans.append(sent)
        xs.append(H.synth_dmr_disc(rng, np.concatenate(parts), taps[1], 10000.0, 0.0 if c % 2 == 0 else 400.0 + 30.0 * c))
    xs = np.stack(xs)
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS)] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), xs.shape[1])
    fs = gpu.FrameSync(n_ch, [(DMR_SYNC, 10)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=max_hits)
    cut = gpu.dmr_burst_cut(res["dibits"], res["reliability"], res["count"], hits, n_hits)
    torch.cuda.synchronize()
    cut_h = {k: v.cpu().numpy() for k, v in cut.items()}
    dib, rel, cnt = res["dibits"].cpu().numpy(), res["reliability"].cpu().numpy(), res["count"].cpu().numpy()
    hits_h, n_hits_h = hits.cpu().numpy(), n_hits.cpu().numpy()
    O = H.oracle_fec()
    O.oracle_dmr_burst_cut.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, H.u8p, H.u8p, H.u8p, H.u8p]
    good = []
    for c in range(n_ch):
        for h in range(max_hits):
            s = c * max_hits + h
            if h >= min(n_hits_h[c], max_hits):
                assert cut_h["valid"][s] == 0
                continue
            cach, info, r98, slot = np.zeros(24, np.uint8), np.zeros(196, np.uint8), np.zeros(98, np.uint8), np.zeros(20, np.uint8)
            ok = O.oracle_dmr_burst_cut(H._ptr(np.ascontiguousarray(dib[c, :cnt[c]]), H.u8p), H._ptr(np.ascontiguousarray(rel[c, :cnt[c]]), H.u8p),
                                        int(cnt[c]), int(hits_h[c, h, 0]), 0, *[H._ptr(a, H.u8p) for a in (cach, info, r98, slot)])
            assert cut_h["valid"][s] == ok
            assert np.array_equal(cut_h["cach24"][s], cach) and np.array_equal(cut_h["info196"][s], info)
            assert np.array_equal(cut_h["rel98"][s], r98) and np.array_equal(cut_h["slot_type20"][s], slot)
            if ok:
                good.append((c, s))
    assert len(good) >= n_ch * n_bursts
    sel = torch.tensor([s for _, s in good], device="cuda")
    k = len(good)
    info_d = cut["info196"][sel].contiguous()
    out96 = torch.zeros((k, 96), dtype=torch.uint8, device="cuda")
    r3 = torch.zeros((k, 3), dtype=torch.uint8, device="cuda")
    errs = torch.zeros(k, dtype=torch.int32, device="cuda")
    L = gpu.lib()
    gpu.check(L.dsdneo_b200_bptc_196x96_batch(info_d.data_ptr(), 1, out96.data_ptr(), r3.data_ptr(), errs.data_ptr(), k, None))
    slot_d = cut["slot_type20"][sel].contiguous()
    ok_slot = torch.zeros(k, dtype=torch.uint8, device="cuda")
    gpu.check(L.dsdneo_b200_fec_block_decode_batch(gpu.FEC_GOLAY_20_8, slot_d.data_ptr(), None, ok_slot.data_ptr(), k, None))
    tact_d = cut["cach24"][sel][:, :7].contiguous()
    tact_dec = torch.zeros((k, 4), dtype=torch.uint8, device="cuda")
    ok_tact = torch.zeros(k, dtype=torch.uint8, device="cuda")
    gpu.check(L.dsdneo_b200_fec_block_decode_batch(gpu.FEC_HAMMING_7_4, tact_d.data_ptr(), tact_dec.data_ptr(), ok_tact.data_ptr(), k, None))
    torch.cuda.synchronize()
    out96, errs, slot_h, ok_slot, ok_tact = out96.cpu().numpy(), errs.cpu().numpy(), slot_d.cpu().numpy(), ok_slot.cpu().numpy(), ok_tact.cpu().numpy()
    tact_h = tact_d.cpu().numpy()
    recovered = 0
    for c in range(n_ch):
        mine = [i for i, (cc_, s) in enumerate(good) if cc_ == c]
        for payload, cc, dt in chans[c]:
            for i in mine:
                if errs[i] == 0 and np.array_equal(out96[i], payload) and ok_slot[i] and ok_tact[i] \
                        and int("".join(map(str, slot_h[i, :4])), 2) == cc and int("".join(map(str, slot_h[i, 4:8])), 2) == dt \
                        and list(tact_h[i, :4]) == [1, 0, 1, 0]:
                    recovered += 1
                    break
    assert recovered == n_ch * n_bursts, recovered


def _voice_burst(rng, gpu, L, frames49, first, slot_bit):
    """One BS voice burst in dibits: CACH (TACT with the slot bit), three AMBE+2 frames (Golay-protected, PN-modulated) through the
    reference's interleave schedule, the BS VOICE sync (burst A) or an arbitrary EMB field."""
    from test_mbe_ecc import ambe_encode

    amap = gpu.ambe_2450_dibit_map()
    cach = np.zeros(24, np.uint8)
    cach[:7] = H.hamming_7_4_encode_bruteforce((1, slot_bit, 0, 0))
    cach[7:] = rng.integers(0, 2, 17)
    tx = np.array([cach[H.DMR_CACH_INTERLEAVE[i]] for i in range(24)], np.uint8)
    dib = np.zeros(144, np.int64)
    dib[:12] = (tx[0::2] << 1) | tx[1::2]
    frs = [ambe_encode(L, d) for d in frames49]
    for f, off, cnt, m0 in [(0, 12, 36, 0), (1, 48, 18, 0), (1, 90, 18, 18), (2, 108, 36, 0)]:
        for i in range(cnt):
            hr, hc, lr, lc = amap[m0 + i]
            dib[off + i] = (int(frs[f][hr, hc]) << 1) | int(frs[f][lr, lc])
    dib[66:90] = [int(c) for c in "131111333113313313113313"] if first else rng.integers(0, 4, 24)
    return dib


def test_c4_chain_voice_and_data_on_device(gpu):
    """BASELINE config C4 in miniature, no host step between slicer and vocoder bits: DMR base-station traffic (slot 1 voice
    superframes A..F, slot 2 data bursts, noise) at discriminator level -> dmr matched filter + slicer -> BS DATA / BS VOICE sync
    hunt -> data burst cutter -> BPTC(196,96) + Golay(20,8); voice burst cutter (6 bursts per superframe) -> AMBE+2 3600x2450
    frame ECC.  Every transmitted payload and every transmitted 49-bit parameter vector comes back."""
    import torch
    from test_mbe_ecc import _o

    L = _o()
    rng = np.random.default_rng(404)
    n_ch, n_bursts, max_hits, voice_hits, vb = 12, 26, 24, 4, 11
    taps = _taps()
    xs, sent_v, sent_d = [], [], []
    for c in range(n_ch):
        parts, sv, sd, k = [rng.integers(0, 4, 60)], [], [], 0
        for b in range(n_bursts):
            if b % 2 == 0:
                frames = rng.integers(0, 2, (3, 49)).astype(np.uint8)
                parts.append(_voice_burst(rng, gpu, L, frames, k % 6 == 0, 0))
                sv.append(frames)
                k += 1
            else:
                payload = rng.integers(0, 2, 96).astype(np.uint8)
                parts.append(H.dmr_build_bs_data_burst(rng, payload, 7, 3, tact4=(1, 1, 0, 0))[0])
                sd.append(payload)
        parts.append(rng.integers(0, 4, 40))
        xs.append(H.synth_dmr_disc(rng, np.concatenate(parts), taps[1], 10000.0, 0.0 if c % 3 == 0 else 300.0 + 25.0 * c))
        sent_v.append(sv)
        sent_d.append(sd)
    xs = np.stack(xs)
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS)] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), xs.shape[1])
    fs = gpu.FrameSync(n_ch, [(DMR_SYNC, 10), ("131111333113313313113313", 12)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=max_hits)
    ar = torch.arange(max_hits, device="cuda")[None, :]

    def split(typ, keep):
        m = (hits[..., 1] == typ) & (ar < n_hits[:, None])
        order = torch.argsort((~m).to(torch.int8), dim=1, stable=True)
        h = torch.gather(hits, 1, order[..., None].expand(-1, -1, 2))[:, :keep].contiguous()
        return h, m.sum(1).to(torch.int32).clamp(max=keep)

    dh, dn = split(10, max_hits)
    vh, vn = split(12, voice_hits)
    for typ, keep_n, (wh, wn) in ((10, max_hits, (dh, dn)), (12, voice_hits, (vh, vn))):  # the library's own selection kernel
        gh, gn = gpu.sync_hits_select(hits, n_hits, typ, keep_n)
        assert torch.equal(gn, wn)
        live = torch.arange(keep_n, device="cuda")[None, :] < gn[:, None]
        assert torch.equal(gh[live], wh[live])
    dh, dn = gpu.sync_hits_select(hits, n_hits, 10, max_hits)
    vh, vn = gpu.sync_hits_select(hits, n_hits, 12, voice_hits)
    cut = gpu.dmr_burst_cut(res["dibits"], res["reliability"], res["count"], dh, dn)
    k = cut["valid"].shape[0]
    out96 = torch.zeros((k, 96), dtype=torch.uint8, device="cuda")
    r3 = torch.zeros((k, 3), dtype=torch.uint8, device="cuda")
    errs = torch.zeros(k, dtype=torch.int32, device="cuda")
    Lb = gpu.lib()
    gpu.check(Lb.dsdneo_b200_bptc_196x96_batch(cut["info196"].data_ptr(), 1, out96.data_ptr(), r3.data_ptr(), errs.data_ptr(), k, None))
    cach, fr, sync, valid = gpu.dmr_voice_cut(res["dibits"], res["count"], vh, vn, voice_hits, vb)
    n_fr = fr.shape[0] * 3
    ambe_d = torch.zeros((n_fr, 49), dtype=torch.uint8, device="cuda")
    c0 = torch.zeros(n_fr, dtype=torch.int32, device="cuda")
    tot = torch.zeros(n_fr, dtype=torch.int32, device="cuda")
    gpu.check(Lb.dsdneo_b200_ambe3600x2450_decode_batch(fr.data_ptr(), ambe_d.data_ptr(), c0.data_ptr(), tot.data_ptr(), n_fr, None))
    torch.cuda.synchronize()
    ok_d = ((cut["valid"] == 1) & (errs == 0)).view(n_ch, max_hits).cpu().numpy()
    out96 = out96.view(n_ch, max_hits, 96).cpu().numpy()
    v = (valid.view(n_ch, voice_hits, vb).bool() & (torch.arange(vb, device="cuda") % 2 == 0)
         & (ar[:, :voice_hits] < vn[:, None])[..., None]).cpu().numpy()
    ambe_d = ambe_d.view(n_ch, voice_hits, vb, 3, 49).cpu().numpy()
    tot = tot.view(n_ch, voice_hits, vb, 3).cpu().numpy()
    # the TACT of a voice burst names slot 0 and decodes under Hamming(7,4)
    tact = cach.view(n_ch, voice_hits, vb, 24)[..., :7].cpu().numpy()
    for c in range(n_ch):
        got_p = {out96[c, h].tobytes() for h in range(max_hits) if ok_d[c, h]}
        assert {p.tobytes() for p in sent_d[c]} <= got_p, c
        got_f = {ambe_d[c, h, j, f].tobytes() for h in range(voice_hits) for j in range(vb) for f in range(3) if v[c, h, j] and tot[c, h, j, f] <= 3}
        want_f = {f.tobytes() for fr3 in sent_v[c] for f in fr3}
        assert len(want_f & got_f) >= len(want_f) - 3, (c, len(want_f & got_f), len(want_f))  # a superframe cut short by the stream's end
        for h in range(voice_hits):
            if v[c, h, 0]:
                assert np.array_equal(tact[c, h, 0], H.hamming_7_4_encode_bruteforce((1, 0, 0, 0)))


def test_dmr_bursts_straddling_launches_are_cut_once_and_whole(gpu):
    """A DMR BS data stream cut into launches in the middle of bursts.  On one launch's rows alone the cutter has to give up
    every burst that straddles a boundary; on the joined rows of a symbol stream (dsdneo_b200_symbol_stream_*: 256 symbols of
    history in front of every launch, sync hunt 64 symbols behind the slicer) every burst of the stream is cut exactly once, at
    the same stream position and with the same bits as in ONE launch over the whole stream, and its payload decodes."""
    import torch

    rng = np.random.default_rng(123)
    n_ch, n_bursts, max_hits, keep, delay = 8, 14, 24, 256, 64
    taps = _taps()
    chans, xs = [], []
    for c in range(n_ch):
        parts, sent = [rng.integers(0, 4, WARMUP + 7 * c)], []
        for _ in range(n_bursts):
            payload = rng.integers(0, 2, 96).astype(np.uint8)
            burst, _ = H.dmr_build_bs_data_burst(rng, payload, int(rng.integers(0, 16)), int(rng.integers(0, 11)))
            parts += [burst, rng.integers(0, 4, GAP)]
            sent.append(payload)
        parts.append(rng.integers(0, 4, 200))
        chans.append(sent)
        xs.append(H.synth_dmr_disc(rng, np.concatenate(parts), taps[1], 10000.0, 0.0 if c % 2 == 0 else 300.0 + 30.0 * c))
    n = min(x.size for x in xs)
    xs = np.stack([x[:n] for x in xs])
    cls = gpu.sym_class_from_synctype(H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS)
    L = gpu.lib()

    def bursts_of(cut_h, hits_h, n_hits_h, base):
        got = {}
        for c in range(n_ch):
            for h in range(min(int(n_hits_h[c]), max_hits)):
                s = c * max_hits + h
                if cut_h["valid"][s]:
                    got[(c, int(base[c]) + int(hits_h[c, h, 0]))] = (cut_h["info196"][s].copy(), cut_h["cach24"][s].copy(),
                                                                     cut_h["slot_type20"][s].copy(), cut_h["rel98"][s].copy())
        return got

    # one launch over the whole stream
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([cls] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), n)
    fs = gpu.FrameSync(n_ch, [(DMR_SYNC, 10)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=max_hits)
    cut = gpu.dmr_burst_cut(res["dibits"], res["reliability"], res["count"], hits, n_hits)
    torch.cuda.synchronize()
    whole = bursts_of({k: v.cpu().numpy() for k, v in cut.items()}, hits.cpu().numpy(), n_hits.cpu().numpy(), np.zeros(n_ch, np.int64))
    assert len(whole) >= n_ch * n_bursts

    # the same stream in launches that end in the middle of bursts
    edges = [0, 5010, 9990, 16470, 20010, n - 2000, n]
    assert all(b > a for a, b in zip(edges, edges[1:]))
    sy2 = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy2.set_class([cls] * n_ch)
    sy3 = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy3.set_class([cls] * n_ch)
    fs2, fs3 = gpu.FrameSync(n_ch, [(DMR_SYNC, 10)]), gpu.FrameSync(n_ch, [(DMR_SYNC, 10)])
    ss = gpu.SymbolStream(n_ch, keep, max(sy2.out_pitch(b - a) for a, b in zip(edges, edges[1:])))
    joined, alone, seen = {}, {}, np.zeros(n_ch, np.int64)
    for a, b in zip(edges, edges[1:]):
        d = torch.from_numpy(np.ascontiguousarray(xs[:, a:b])).cuda()
        # (a) joined rows
        view = ss.run(sy2, d, b - a)
        hits2 = torch.zeros((n_ch, max_hits, 2), dtype=torch.int32, device="cuda")
        n_hits2 = torch.zeros(n_ch, dtype=torch.int32, device="cuda")
        gpu.check(L.dsdneo_b200_frame_sync_search_batch(fs2._h, view.d_symbols + 4 * (keep - delay), view.pitch, view.d_new,
                                                        hits2.data_ptr(), max_hits, n_hits2.data_ptr(), None))
        gpu.check(L.dsdneo_b200_sync_hits_rebase(hits2.data_ptr(), n_hits2.data_ptr(), n_ch, max_hits, keep - delay, None))
        slots = n_ch * max_hits
        o = {k: torch.zeros((slots, w) if w else (slots,), dtype=torch.uint8, device="cuda")
             for k, w in (("cach24", 24), ("info196", 196), ("rel98", 98), ("slot_type20", 20), ("valid", 0))}
        gpu.check(L.dsdneo_b200_dmr_burst_cut_batch(view.d_dibits, view.pitch, view.d_reliability, view.pitch, view.d_valid,
                                                    hits2.data_ptr(), n_hits2.data_ptr(), n_ch, max_hits, 0, o["cach24"].data_ptr(),
                                                    o["info196"].data_ptr(), o["rel98"].data_ptr(), o["slot_type20"].data_ptr(),
                                                    o["valid"].data_ptr(), None))
        rows = ss.fetch(view)
        got = bursts_of({k: v.cpu().numpy() for k, v in o.items()}, hits2.cpu().numpy(), n_hits2.cpu().numpy(), rows["stream_base"])
        assert not (set(got) & set(joined)), "a burst was cut twice"
        joined.update(got)
        # (b) every launch on its own rows
        r3 = sy3.run(d, b - a)
        h3, nh3 = fs3.search(r3["symbols"], r3["count"], max_hits=max_hits)
        c3 = gpu.dmr_burst_cut(r3["dibits"], r3["reliability"], r3["count"], h3, nh3)
        torch.cuda.synchronize()
        alone.update(bursts_of({k: v.cpu().numpy() for k, v in c3.items()}, h3.cpu().numpy(), nh3.cpu().numpy(), seen))
        seen += r3["count"].cpu().numpy()
    # the hunt trails the slicer by `delay` symbols: bursts whose sync ends in the last `delay` symbols of the stream are still pending
    total = seen
    expect = {k: v for k, v in whole.items() if k[1] < total[k[0]] - delay}
    assert len(expect) >= n_ch * n_bursts
    assert set(joined) == set(expect), (sorted(set(expect) - set(joined))[:5], sorted(set(joined) - set(expect))[:5])
    for k in expect:
        for x, y in zip(joined[k], expect[k]):
            assert np.array_equal(x, y), k
    assert len(alone) < len(expect)  # launches on their own lose the bursts at their edges
    info = np.array([joined[k][0] for k in sorted(joined)])
    out, _, errs = gpu.bptc_196x96(info, interleaved=True)
    sent = {c: [p.tobytes() for p in chans[c]] for c in range(n_ch)}
    ok = sum(1 for (c, _), o96, e in zip(sorted(joined), out, errs) if e == 0 and o96.tobytes() in sent[c])
    assert ok == n_ch * n_bursts, ok


def test_dmr_voice_superframes_straddling_launches(gpu):
    """The voice burst cutter on the joined rows of a symbol stream (1024 symbols kept, sync hunt 1536 symbols behind the slicer:
    eleven 144-dibit bursts after the sync): a slot-1 voice stream cut into launches in the middle of superframes gives, for
    every BS VOICE sync, the same eleven burst records (CACH, ambe_fr, sync / EMB bits, validity) as ONE launch over the whole
    stream, exactly once."""
    import torch
    from test_mbe_ecc import _o

    Lo = _o()
    rng = np.random.default_rng(505)
    n_ch, n_bursts, max_hits, vb, keep, delay = 4, 40, 8, 11, 2048, 1536
    taps = _taps()
    xs = []
    for c in range(n_ch):
        parts, k = [rng.integers(0, 4, 60 + 11 * c)], 0
        for b in range(n_bursts):
            if b % 2 == 0:
                parts.append(_voice_burst(rng, gpu, Lo, rng.integers(0, 2, (3, 49)).astype(np.uint8), k % 6 == 0, 0))
                k += 1
            else:
                parts.append(H.dmr_build_bs_data_burst(rng, rng.integers(0, 2, 96).astype(np.uint8), 7, 3, tact4=(1, 1, 0, 0))[0])
        parts.append(rng.integers(0, 4, 1700))
        xs.append(H.synth_dmr_disc(rng, np.concatenate(parts), taps[1], 10000.0, 0.0 if c % 2 == 0 else 300.0))
    n = min(x.size for x in xs)
    xs = np.stack([x[:n] for x in xs])
    cls = gpu.sym_class_from_synctype(H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS)
    VOICE = ("131111333113313313113313", 12)
    L = gpu.lib()

    def records(cach, fr, sync, valid, hits_h, n_hits_h, base):
        cach, fr, sync, valid = (t.cpu().numpy() for t in (cach, fr, sync, valid))
        got = {}
        for c in range(n_ch):
            for h in range(min(int(n_hits_h[c]), max_hits)):
                r0 = (c * max_hits + h) * vb
                got[(c, int(base[c]) + int(hits_h[c, h, 0]))] = (valid[r0:r0 + vb].copy(), cach[r0:r0 + vb].copy(), fr[r0:r0 + vb].copy(),
                                                                  sync[r0:r0 + vb].copy())
        return got

    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([cls] * n_ch)
    res = sy.run(torch.from_numpy(xs).cuda(), n)
    fs = gpu.FrameSync(n_ch, [VOICE])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=max_hits)
    whole = records(*gpu.dmr_voice_cut(res["dibits"], res["count"], hits, n_hits, max_hits, vb), hits.cpu().numpy(), n_hits.cpu().numpy(),
                    np.zeros(n_ch, np.int64))
    total = res["count"].cpu().numpy()
    assert len(whole) >= n_ch * 3

    edges = [0, 7010, 15990, 23330, 30010, 41110, n - 3000, n]
    assert all(b > a for a, b in zip(edges, edges[1:]))
    sy2 = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy2.set_class([cls] * n_ch)
    fs2 = gpu.FrameSync(n_ch, [VOICE])
    ss = gpu.SymbolStream(n_ch, keep, max(sy2.out_pitch(b - a) for a, b in zip(edges, edges[1:])))
    joined = {}
    for a, b in zip(edges, edges[1:]):
        view = ss.run(sy2, torch.from_numpy(np.ascontiguousarray(xs[:, a:b])).cuda(), b - a)
        h2 = torch.zeros((n_ch, max_hits, 2), dtype=torch.int32, device="cuda")
        nh2 = torch.zeros(n_ch, dtype=torch.int32, device="cuda")
        gpu.check(L.dsdneo_b200_frame_sync_search_batch(fs2._h, view.d_symbols + 4 * (keep - delay), view.pitch, view.d_new, h2.data_ptr(),
                                                        max_hits, nh2.data_ptr(), None))
        gpu.check(L.dsdneo_b200_sync_hits_rebase(h2.data_ptr(), nh2.data_ptr(), n_ch, max_hits, keep - delay, None))
        R = n_ch * max_hits * vb
        cach = torch.zeros((R, 24), dtype=torch.uint8, device="cuda")
        fr = torch.zeros((R, 3, 4, 24), dtype=torch.uint8, device="cuda")
        sync = torch.zeros((R, 48), dtype=torch.uint8, device="cuda")
        valid = torch.zeros(R, dtype=torch.uint8, device="cuda")
        gpu.check(L.dsdneo_b200_dmr_voice_cut_batch(view.d_dibits, view.pitch, view.d_valid, h2.data_ptr(), nh2.data_ptr(), n_ch, max_hits, vb, 0,
                                                    cach.data_ptr(), fr.data_ptr(), sync.data_ptr(), valid.data_ptr(), None))
        rows = ss.fetch(view)
        got = records(cach, fr, sync, valid, h2.cpu().numpy(), nh2.cpu().numpy(), rows["stream_base"])
        assert not (set(got) & set(joined)), "a sync was reported twice"
        joined.update(got)
    expect = {k: v for k, v in whole.items() if k[1] < total[k[0]] - delay}
    assert len(expect) >= n_ch * 3 and set(joined) == set(expect), (sorted(expect)[:4], sorted(joined)[:4])
    full = 0
    for k in expect:
        for x, y in zip(joined[k], expect[k]):
            assert np.array_equal(x, y), k
        full += int(expect[k][0].all())
    assert full >= n_ch * 2  # superframes whose eleven bursts are all present
