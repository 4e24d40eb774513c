"""CPU tests: pin the sample-side oracle (oracle/oracle_symbol.c) against the UNMODIFIED reference getSymbol() /
getDibitSoft() driven through its own runtime hook seam (oracle/ref_shim_symbol.c)."""
import ctypes as C

import os

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_available("par"), reason="oracle/_ref not built (no /root/reference)")


def _oracle_chan(sync, lastsync, taps_by_filter, rate=48000, symrate=4800, use_cosine=1):
    cls = H.SYNC_CLASS[lastsync]
    neg = H.SYNC_CLASS[sync]["negative"]
    ch = H.OracleSymChan()
    taps = taps_by_filter.get(cls["filter"]) if (use_cosine and cls["filter"] is not None) else None
    tp = H._ptr(taps) if taps is not None else None
    H.oracle_sym().oracle_sym_init(C.byref(ch), rate, symrate, 1 if taps is not None else 0, cls["window_l"], cls["track"], neg, tp,
                                   0 if taps is not None else 0 if taps is None else taps.size, 128, 1024) if taps is None else \
        H.oracle_sym().oracle_sym_init(C.byref(ch), rate, symrate, 1, cls["window_l"], cls["track"], neg, tp, taps.size, 128, 1024)
    return ch


@needs_ref
@pytest.mark.parametrize("sync,lastsync", [(H.SYNC_P25P1_POS, H.SYNC_P25P1_POS), (H.SYNC_P25P1_NEG, H.SYNC_P25P1_NEG),
                                           (H.SYNC_DMR_BS_DATA_POS, H.SYNC_DMR_BS_DATA_POS), (H.SYNC_NONE, H.SYNC_NONE)])
@pytest.mark.parametrize("noise", [0.0, 2500.0])
def test_get_dibit_soft_matches_reference(sync, lastsync, noise):
    R, O = H.ref_sym(), H.oracle_sym()
    rng = np.random.default_rng(100 + sync * 7 + int(noise))
    nsym = 6000
    x, _ = H.synth_disc(rng, nsym, 10, 9000.0, noise, drift=1500.0)
    taps = {0: H.sps_fir_taps(0, 10), 1: H.sps_fir_taps(1, 10)}
    h = R.ref_sym_create(48000, 4800, sync, lastsync, 1, 128, 1024)
    R.ref_sym_feed(h, H._ptr(x), x.size)
    d = np.zeros(nsym, np.uint8); r = np.zeros(nsym, np.uint8); l = np.zeros(2 * nsym, np.int16); s = np.zeros(nsym, np.float32)
    n = R.ref_sym_get_dibits(h, nsym, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
    assert n > 5000
    consumed_ref = R.ref_sym_consumed(h)
    R.ref_sym_destroy(h)
    ch = _oracle_chan(sync, lastsync, taps)
    d2 = np.zeros(nsym, np.uint8); r2 = np.zeros(nsym, np.uint8); l2 = np.zeros(2 * nsym, np.int16); s2 = np.zeros(nsym, np.float32)
    cons = C.c_long(0)
    n2 = O.oracle_sym_run_dibits(C.byref(ch), H._ptr(x), x.size, 600, H._ptr(d2, H.u8p), H._ptr(r2, H.u8p),
                                 l2.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s2), n, C.byref(cons))
    assert n2 == n and cons.value == consumed_ref
    assert H.bits_equal(s[:n], s2[:n]), H.first_mismatch(s[:n], s2[:n])
    assert np.array_equal(d[:n], d2[:n])
    assert np.array_equal(r[:n], r2[:n]) and np.array_equal(l[:2 * n], l2[:2 * n])


@needs_ref
@pytest.mark.parametrize("lastsync", [H.SYNC_P25P1_POS, H.SYNC_NONE, H.SYNC_DMR_BS_DATA_POS])
def test_get_symbol_unsynced_timing_matches_reference(lastsync):
    """have_sync = 0: the +-1 sample jitter nudge is active, so symbols consume 9, 10 or 11 samples."""
    R, O = H.ref_sym(), H.oracle_sym()
    rng = np.random.default_rng(300 + lastsync)
    nsym = 5000
    # sampling-clock offset: 10.02 samples per symbol so the tracker has to keep nudging
    dib = rng.integers(0, 4, nsym)
    t = np.arange(int(nsym * 10.02))
    idx = np.minimum((t / 10.02).astype(np.int64), nsym - 1)
    x = (H.LEVELS[dib][idx] * 9000.0 + rng.standard_normal(t.size) * 800.0).astype(np.float32)
    k = np.hanning(12)[1:-1]; k /= k.sum()
    x = np.convolve(x, k, mode="same").astype(np.float32)
    taps = {0: H.sps_fir_taps(0, 10), 1: H.sps_fir_taps(1, 10)}
    h = R.ref_sym_create(48000, 4800, lastsync, lastsync, 1, 128, 1024)
    R.ref_sym_feed(h, H._ptr(x), x.size)
    s = np.zeros(nsym, np.float32)
    n = R.ref_sym_get_symbols(h, 0, nsym, 600, H._ptr(s))
    consumed_ref = R.ref_sym_consumed(h)
    R.ref_sym_destroy(h)
    assert n > 4000
    ch = _oracle_chan(lastsync, lastsync, taps)
    s2 = np.zeros(nsym, np.float32)
    cons = C.c_long(0)
    n2 = O.oracle_sym_run_symbols(C.byref(ch), 0, H._ptr(x), x.size, 600, H._ptr(s2), n, C.byref(cons))
    assert n2 == n and cons.value == consumed_ref, (n, n2, cons.value, consumed_ref)
    assert consumed_ref != 10 * n  # the nudge really fired
    assert H.bits_equal(s[:n], s2[:n]), H.first_mismatch(s[:n], s2[:n])


@needs_ref
def test_fractional_samples_per_symbol():
    """50 kS/s / 4800 sym/s = 10 + 2000/4800: the remainder accumulator alternates 10- and 11-sample symbols."""
    R, O = H.ref_sym(), H.oracle_sym()
    rng = np.random.default_rng(400)
    x, _ = H.synth_disc(rng, 3000, 10, 9000.0, 500.0)
    h = R.ref_sym_create(50000, 4800, H.SYNC_NONE, H.SYNC_NONE, 1, 128, 1024)
    R.ref_sym_feed(h, H._ptr(x), x.size)
    s = np.zeros(3000, np.float32)
    n = R.ref_sym_get_symbols(h, 1, 3000, 600, H._ptr(s))
    consumed_ref = R.ref_sym_consumed(h)
    R.ref_sym_destroy(h)
    ch = _oracle_chan(H.SYNC_NONE, H.SYNC_NONE, {}, rate=50000)
    s2 = np.zeros(3000, np.float32)
    cons = C.c_long(0)
    n2 = O.oracle_sym_run_symbols(C.byref(ch), 1, H._ptr(x), x.size, 600, H._ptr(s2), n, C.byref(cons))
    assert n2 == n and cons.value == consumed_ref
    assert H.bits_equal(s[:n], s2[:n])


def test_oracle_matches_committed_golden_vectors():
    """Reference outputs captured by tests/golden/make_golden.py: runs everywhere, also without oracle/_ref."""
    import os

    O = H.oracle_sym()
    g = np.load(os.path.join(H.GOLDEN_DIR, "symbols.npz"))
    t = np.load(os.path.join(H.GOLDEN_DIR, "sps_fir_taps.npz"))
    taps = {0: t["f0_sps10"], 1: t["f1_sps10"]}
    for name, sync in (("p25p1_pos", H.SYNC_P25P1_POS), ("dmr_bs_data", H.SYNC_DMR_BS_DATA_POS)):
        x = g[name + "_x"]
        n = g[name + "_dibits"].size
        ch = _oracle_chan(sync, sync, taps)
        d = np.zeros(n, np.uint8); r = np.zeros(n, np.uint8); l = np.zeros(2 * n, np.int16); s = np.zeros(n, np.float32)
        cons = C.c_long(0)
        k = O.oracle_sym_run_dibits(C.byref(ch), H._ptr(x), x.size, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p),
                                    l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s), n, C.byref(cons))
        assert k == n
        assert np.array_equal(d, g[name + "_dibits"]) and np.array_equal(r, g[name + "_rel"]) and np.array_equal(l, g[name + "_llr"])
        assert H.bits_equal(s, g[name + "_symbols"])


def cqpsk_symbol_stream(rng, n, noise=0.25, offset=0.0):
    """Symbol-rate CQPSK stream as the block side emits it: values near {-3,-1,+1,+3} (+ noise, + a slow centre drift)."""
    lv = H.LEVELS[rng.integers(0, 4, n)]
    drift = offset * np.sin(np.arange(n) / 700.0)
    return (lv + noise * rng.standard_normal(n) + drift).astype(np.float32)


@needs_ref
@pytest.mark.parametrize("sync,active,map_idx,snr", [(H.SYNC_P25P1_POS, 1, 0, -100.0), (H.SYNC_P25P1_NEG, 1, 0, 18.0),
                                                     (H.SYNC_P25P1_POS, 1, 3, 30.0), (H.SYNC_P25P1_POS, 1, 2, 7.5),
                                                     (H.SYNC_P25P1_NEG, 1, 4, -100.0), (H.SYNC_P25P1_POS, 0, 0, 12.0),
                                                     (H.SYNC_DMR_BS_DATA_POS, 1, 1, 3.0)])
def test_cqpsk_symbol_rate_slicer_matches_reference(sync, active, map_idx, snr):
    """Output kind 2 (symbol-rate CQPSK): the oracle's thresholds / tracker / CQPSK slicer / soft metric equal the unmodified
    reference getDibitSoft() driven through its hook seam with rf_mod = 1: dibits, reliabilities, LLRs, symbols, thresholds."""
    R = H.ref_sym()
    R.ref_sym_create_cqpsk.restype = C.c_void_p
    R.ref_sym_create_cqpsk.argtypes = [C.c_int] * 7 + [C.c_double]
    R.ref_sym_get_dibits_n.restype = C.c_long
    R.ref_sym_get_dibits_n.argtypes = [C.c_void_p, C.c_long, H.u8p, H.u8p, C.POINTER(C.c_int16), H.f32p]
    rng = np.random.default_rng(500 + sync * 11 + map_idx)
    n = 5000
    x = cqpsk_symbol_stream(rng, n + 600, noise=0.35, offset=0.3)
    h = R.ref_sym_create_cqpsk(4800, sync, sync, 128, 1024, map_idx, active, snr)
    R.ref_sym_feed(h, H._ptr(x), x.size)
    d = np.zeros(n, np.uint8); r = np.zeros(n, np.uint8); l = np.zeros(2 * n, np.int16); s = np.zeros(n, np.float32)
    assert R.ref_sym_get_dibits_n(h, n, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s)) == n
    f8, i5 = np.zeros(8, np.float32), np.zeros(5, np.int32)
    R.ref_sym_get_state(h, H._ptr(f8), i5.ctypes.data_as(H.i32p))
    R.ref_sym_destroy(h)
    assert H.bits_equal(s, x[:n])  # one stream float = one symbol
    is_p25 = sync in (H.SYNC_P25P1_POS, H.SYNC_P25P1_NEG)
    neg = H.SYNC_CLASS[sync]["negative"]
    d2, r2, l2, st = H.oracle_cqpsk_slicer_run(x[:n], negative=neg, p25_slice=1 if (active and is_p25) else 0, map_idx=map_idx,
                                               snr_db=snr)
    assert np.array_equal(d, d2), int(np.argmax(d != d2))
    assert np.array_equal(r, r2), int(np.argmax(r != r2))
    assert np.array_equal(l, l2.reshape(-1))
    b = st.base
    want = np.array([b.min, b.max, b.center, b.umid, b.lmid, b.minref, b.maxref, b.lastsample], np.float32)
    assert H.bits_equal(f8, want), (f8, want)
    assert len(set(d.tolist())) == 4


# ---- acquisition: the oracle's getFrameSync restatement against the UNMODIFIED getFrameSync() -------------------------------

def _ref_acquire(disc, frame_mask, rf_mod, n_after=600):
    """The real getFrameSync() once, then n_after getDibitSoft() calls.  Returns a dict of what the reference did."""
    R = H.ref_sym()
    R.ref_sym_configure_acquire.argtypes = [C.c_void_p, C.c_int, C.c_int]
    R.ref_sym_frame_sync.argtypes = [C.c_void_p]
    R.ref_sym_frame_sync.restype = C.c_int
    R.ref_sym_recent.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_sym_symbol_count.restype = C.c_long
    R.ref_sym_symbol_count.argtypes = [C.c_void_p]
    h = R.ref_sym_create(48000, 4800, -1, -1, 1, 128, 1024)
    R.ref_sym_configure_acquire(h, frame_mask, rf_mod)
    R.ref_sym_feed(h, H._ptr(disc), disc.size)
    st = -1
    for _ in range(64):  # a call gives up after 1800 symbols without sync; the decoder loop simply calls again
        st = R.ref_sym_frame_sync(h)
        if st >= 0 or disc.size - R.ref_sym_consumed(h) < 2000:
            break
    out = {"sync_type": st, "hunted": int(R.ref_sym_symbol_count(h)), "consumed": int(R.ref_sym_consumed(h))}
    f8, i5 = np.zeros(8, np.float32), np.zeros(5, np.int32)
    R.ref_sym_get_state(h, H._ptr(f8), i5.ctypes.data_as(C.POINTER(C.c_int)))
    out["f8"], out["i5"] = f8.copy(), i5.copy()
    n = min(out["hunted"], 200)
    sym, pd, pr = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    R.ref_sym_recent(h, n, H._ptr(sym), pd.ctypes.data, pr.ctypes.data)
    out["recent_sym"], out["recent_dib"], out["recent_rel"] = sym, pd, pr
    if st >= 0:
        d, r = np.zeros(n_after, np.uint8), np.zeros(n_after, np.uint8)
        l, s = np.zeros(2 * n_after, np.int16), np.zeros(n_after, np.float32)
        k = R.ref_sym_get_dibits(h, n_after, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
        out["after"] = (d[:k], r[:k], l[:2 * k].reshape(-1, 2), s[:k])
    R.ref_sym_destroy(h)
    return out


def _oracle_acquire(disc, frame_mask, rf_mod, taps, n_after=600):
    O = H.oracle_sym()
    ch = H.OracleSymChan()
    O.oracle_sym_init(C.byref(ch), 48000, 4800, 0, 2, 0, 0, None, 0, 128, 1024)
    ch.rf_mod = rf_mod
    pats, keep = H.acquire_patterns(frame_mask & 1, frame_mask & 2, taps, inverted_dmr=bool(frame_mask & 4))
    n = disc.size // 8 + 8
    sym, dib, rel = np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    res = H.OracleAcqResult()
    O.oracle_sym_acquire.restype = C.c_long
    k = O.oracle_sym_acquire(C.byref(ch), H._ptr(disc), C.c_long(disc.size), C.c_long(600), pats, len(pats), H._ptr(sym),
                             H._ptr(dib, H.u8p), H._ptr(rel, H.u8p), C.c_long(n), C.byref(res))
    out = {"sync_type": res.sync_type, "hunted": int(k), "consumed": int(res.consumed), "ch": ch, "sym": sym[:k], "dib": dib[:k],
           "rel": rel[:k], "warm": res.warm_start,
           "f8": np.array([ch.min, ch.max, ch.center, ch.umid, ch.lmid, ch.minref, ch.maxref, ch.lastsample], np.float32)}
    if res.sync_type >= 0:
        d, r = np.zeros(n_after, np.uint8), np.zeros(n_after, np.uint8)
        l, s = np.zeros(2 * n_after, np.int16), np.zeros(n_after, np.float32)
        cons = C.c_long(0)
        rest = np.ascontiguousarray(disc[res.consumed:])
        kk = O.oracle_sym_run_dibits(C.byref(ch), H._ptr(rest), rest.size, 600, H._ptr(d, H.u8p), H._ptr(r, H.u8p),
                                     l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s), n_after, C.byref(cons))
        out["after"] = (d[:kk], r[:kk], l[:2 * kk].reshape(-1, 2), s[:kk])
    return out


def _compare_acquire(ref, got):
    assert got["sync_type"] == ref["sync_type"], (got["sync_type"], ref["sync_type"])
    # the reference's symbol history saturates at DSD_SYMBOL_HISTORY_SIZE = 2048 entries
    assert min(got["hunted"], 2048) == ref["hunted"] and got["consumed"] == ref["consumed"], (got["hunted"], ref["hunted"], got["consumed"], ref["consumed"])
    n = ref["recent_sym"].size
    assert H.bits_equal(got["sym"][-n:], ref["recent_sym"])
    assert np.array_equal(got["dib"][-n:], ref["recent_dib"].astype(np.uint8)) and np.array_equal(got["rel"][-n:], ref["recent_rel"])
    assert H.bits_equal(got["f8"], ref["f8"]), (got["f8"], ref["f8"])
    if ref["sync_type"] >= 0:
        for a, b in zip(got["after"], ref["after"]):
            assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@needs_ref
@pytest.mark.parametrize("fixture,mask,rf_mod,profile,want_sync", [
    ("p25p1_c4fm_cc", 1, 0, 4, 0), ("p25p1_c4fm_vc", 1, 0, 4, 0), ("dmr_t3_cc", 2, 2, 2, 12), ("dmr_voice", 2, 2, 2, None),
    ("dmr_t3_cc", 2, 0, 2, 12), ("dmr_t3_cc", 3, 0, 2, 12)])
def test_acquisition_matches_the_unmodified_getframesync_on_the_reference_captures(fixture, mask, rf_mod, profile, want_sync):
    path = "/root/reference/tests/fixtures/iq/%s.iq" % fixture
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    x = H.widen_cu8(np.fromfile(path, dtype=np.uint8)).reshape(-1, 2)
    bp = 8000
    disc = H.RefDemod("par", rate=48000, symrate=4800, profile=profile).run(x, bp, x.shape[0] // bp)
    taps = {0: H.sps_fir_taps(0, 10), 1: H.sps_fir_taps(1, 10)}
    ref = _ref_acquire(disc, mask, rf_mod)
    got = _oracle_acquire(disc, mask, rf_mod, taps)
    if want_sync is not None:
        assert ref["sync_type"] == want_sync
    assert ref["sync_type"] >= 0
    _compare_acquire(ref, got)
