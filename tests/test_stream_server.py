"""The reference's runtime hook seam served by libdsdneo_b200: the UNMODIFIED reference
sample side (getDibitSoft -> getSymbol -> dsd_rtl_stream_io_hook_read) pulls its floats from dsdneo_b200_stream_hook_read
and its stream metadata from the dsdneo_b200_stream_hook_* metrics functions, and decodes the same dibits as through the
test shim's own hooks."""
import ctypes as C
import threading

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_available("par"), reason="oracle/_ref not built (no /root/reference)")


def _b200():
    import __graft_entry__ as g

    return g.load_package()


def _fn(lib, name):
    return C.cast(getattr(lib, name), C.c_void_p)


def _ref_dibits_via_shim(R, x, sync, n_sym):
    h = R.ref_sym_create(48000, 4800, sync, sync, 1, 128, 1024)
    R.ref_sym_feed(h, H._ptr(x), x.size)
    d = np.zeros(n_sym, np.uint8); r = np.zeros(n_sym, np.uint8); l = np.zeros(2 * n_sym, np.int16); s = np.zeros(n_sym, np.float32)
    n = R.ref_sym_get_dibits_n(h, n_sym, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
    R.ref_sym_destroy(h)
    assert n == n_sym
    return d, r, l.reshape(-1, 2), s


def _ref_dibits_via_server(R, b, x, sync, n_sym, ring=4096, threaded=True):
    L = b.lib()
    srv = L.dsdneo_b200_stream_server_create(ring, 48000, 4800, 4, 4)  # profile 4 = P25_C4FM (rtl_stream_metrics_hooks.h:49-56)
    assert srv
    L.dsdneo_b200_stream_server_make_current(srv)
    h = R.ref_sym_create(48000, 4800, sync, sync, 1, 128, 1024)
    R.ref_sym_use_external_hooks(C.c_void_p(h), _fn(L, "dsdneo_b200_stream_hook_read"), _fn(L, "dsdneo_b200_stream_hook_return_pwr"),
                                 C.c_void_p(srv), _fn(L, "dsdneo_b200_stream_hook_output_rate_hz"),
                                 _fn(L, "dsdneo_b200_stream_hook_output_kind"), _fn(L, "dsdneo_b200_stream_hook_symbol_profile"),
                                 _fn(L, "dsdneo_b200_stream_hook_stream_generation"))

    def produce():
        for lo in range(0, x.size, 1000):  # ingest loop: tile by tile, blocking on a full ring (back-pressure)
            chunk = np.ascontiguousarray(x[lo:lo + 1000])
            assert L.dsdneo_b200_stream_server_push(srv, chunk.ctypes.data, chunk.size, 1) == chunk.size
        L.dsdneo_b200_stream_server_close(srv)

    if threaded:
        t = threading.Thread(target=produce)
        t.start()
    else:
        produce()
    d = np.zeros(n_sym, np.uint8); r = np.zeros(n_sym, np.uint8); l = np.zeros(2 * n_sym, np.int16); s = np.zeros(n_sym, np.float32)
    n = R.ref_sym_get_dibits_n(C.c_void_p(h), n_sym, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
    if threaded:
        # drain whatever the decoder did not need so the producer can finish
        buf = np.zeros(4096, np.float32)
        got = C.c_int(0)
        while L.dsdneo_b200_stream_hook_read(srv, buf.ctypes.data, buf.size, C.byref(got)) == 0:
            pass
        t.join()
    R.ref_sym_destroy(C.c_void_p(h))
    L.dsdneo_b200_stream_server_destroy(srv)
    assert n == n_sym
    return d, r, l.reshape(-1, 2), s


def _bind(R):
    R.ref_sym_create.restype = C.c_void_p
    R.ref_sym_feed.argtypes = [C.c_void_p, H.f32p, C.c_long]
    R.ref_sym_destroy.argtypes = [C.c_void_p]
    R.ref_sym_get_dibits_n.restype = C.c_long
    R.ref_sym_get_dibits_n.argtypes = [C.c_void_p, C.c_long, H.u8p, H.u8p, C.POINTER(C.c_int16), H.f32p]
    R.ref_sym_use_external_hooks.argtypes = [C.c_void_p] * 8


@needs_ref
def test_reference_decoder_reads_through_the_stream_server():
    b = _b200()
    R = C.CDLL(H._ref_path("par"))
    _bind(R)
    rng = np.random.default_rng(21)
    dib = rng.integers(0, 4, 3000)
    x = H.synth_c4fm_disc(rng, dib, 9000.0, 900.0)
    n_sym = 2900
    want = _ref_dibits_via_shim(R, x, H.SYNC_P25P1_POS, n_sym)
    got = _ref_dibits_via_server(R, b, x, H.SYNC_P25P1_POS, n_sym, ring=2048, threaded=True)
    for a, w in zip(got, want):
        assert np.array_equal(a.view(np.uint8), w.view(np.uint8))
    assert (want[0][5:2800] == dib[:2795]).mean() > 0.99  # and it is the transmitted data


def test_stream_server_ring_semantics():
    """Blocking read returns what is there (<= count), wraps around the ring, reports end of stream after close, and a
    generation bump drops buffered samples."""
    b = _b200()
    L = b.lib()
    srv = L.dsdneo_b200_stream_server_create(1024, 48000, 4800, 4, 2)
    L.dsdneo_b200_stream_server_make_current(srv)
    assert L.dsdneo_b200_stream_hook_output_rate_hz() == 48000 and L.dsdneo_b200_stream_hook_output_kind() == 1
    r, lv, pr = C.c_int(0), C.c_int(0), C.c_int(0)
    assert L.dsdneo_b200_stream_hook_symbol_profile(C.byref(r), C.byref(lv), C.byref(pr)) == 0 and (r.value, lv.value, pr.value) == (4800, 4, 2)
    gen0 = L.dsdneo_b200_stream_hook_stream_generation()
    x = np.arange(3000, dtype=np.float32)
    out = np.zeros(512, np.float32)
    got = C.c_int(0)
    seen, pushed = [], 0
    while pushed < x.size:
        pushed += L.dsdneo_b200_stream_server_push(srv, x[pushed:].ctypes.data, x.size - pushed, 0)  # non-blocking: fills the ring
        assert L.dsdneo_b200_stream_hook_read(srv, out.ctypes.data, 512, C.byref(got)) == 0 and 0 < got.value <= 512
        seen.append(out[:got.value].copy())
    L.dsdneo_b200_stream_server_close(srv)
    while L.dsdneo_b200_stream_hook_read(srv, out.ctypes.data, 512, C.byref(got)) == 0:
        seen.append(out[:got.value].copy())
    assert np.array_equal(np.concatenate(seen), x)
    assert L.dsdneo_b200_stream_hook_read(srv, out.ctypes.data, 512, C.byref(got)) < 0 and got.value == 0
    L.dsdneo_b200_stream_server_destroy(srv)
    srv = L.dsdneo_b200_stream_server_create(1024, 24000, 2400, 4, 1)
    L.dsdneo_b200_stream_server_make_current(srv)
    L.dsdneo_b200_stream_server_push(srv, x.ctypes.data, 500, 0)
    L.dsdneo_b200_stream_server_bump_generation(srv)
    assert L.dsdneo_b200_stream_hook_stream_generation() == gen0 + 1
    L.dsdneo_b200_stream_server_push(srv, x[700:].ctypes.data, 10, 0)
    assert L.dsdneo_b200_stream_hook_read(srv, out.ctypes.data, 512, C.byref(got)) == 0 and got.value == 10 and out[0] == 700.0
    L.dsdneo_b200_stream_server_set_power(srv, 0.25)
    assert L.dsdneo_b200_stream_hook_return_pwr(srv) == 0.25
    L.dsdneo_b200_stream_server_destroy(srv)


@pytest.mark.gpu
@needs_ref
def test_gpu_front_end_feeds_the_unmodified_reference_decoder(gpu):
    """Drop-in at the seam: wideband IQ -> GPU channelizer + full_demod -> stream server -> the reference's own getDibitSoft;
    the dibits equal the GPU symbolizer's on the same channel (and therefore the oracle's)."""
    import torch
    from test_gpu_symbolizer import _taps

    R = C.CDLL(H._ref_path("par"))
    _bind(R)
    rng = np.random.default_rng(22)
    M, bp, nb = 256, 4096, 2
    x, _ = H.synth_wideband(rng, M, bp * nb, [40], snr_db=28.0)
    fe = gpu.Frontend(M, 8, False, 12_288_000, bp)
    disc = fe.process(torch.from_numpy(x).cuda())
    row = disc[40].cpu().numpy().copy()
    n_sym = row.size // 10 - 10
    got = _ref_dibits_via_server(R, gpu, row, H.SYNC_P25P1_POS, n_sym, ring=4096, threaded=True)
    sy = gpu.Symbolizer(1, 48000, 4800, filters=_taps())
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_P25P1_POS, H.SYNC_P25P1_POS)])
    res = sy.run(disc[40:41].contiguous(), row.size)
    torch.cuda.synchronize()
    assert int(res["count"][0]) >= n_sym
    assert np.array_equal(res["dibits"][0, :n_sym].cpu().numpy(), got[0])
    assert np.array_equal(res["llr"][0, :n_sym].cpu().numpy(), got[2])
    assert H.bits_equal(res["symbols"][0, :n_sym].cpu().numpy(), got[3])


def _ref_cqpsk_dibits_via_server(R, b, symbols, sync, map_idx, snr, n_sym):
    """The unmodified reference getDibitSoft() with rf_mod = 1 pulling symbol-rate CQPSK symbols (output kind 2) through
    the library's stream server and its cqpsk_status / snr_cqpsk_db hooks."""
    L = b.lib()
    srv = L.dsdneo_b200_stream_server_create(2048, 4800, 4800, 4, 5)  # profile 5 = P25_CQPSK
    assert srv
    L.dsdneo_b200_stream_server_set_output_kind(srv, 2, 1, snr)
    L.dsdneo_b200_stream_server_make_current(srv)
    assert L.dsdneo_b200_stream_hook_output_kind() == 2
    R.ref_sym_use_external_hooks_cqpsk.argtypes = [C.c_void_p] * 10 + [C.c_int]
    h = R.ref_sym_create(4800, 4800, sync, sync, 0, 128, 1024)
    R.ref_sym_use_external_hooks_cqpsk(C.c_void_p(h), _fn(L, "dsdneo_b200_stream_hook_read"), _fn(L, "dsdneo_b200_stream_hook_return_pwr"),
                                       C.c_void_p(srv), _fn(L, "dsdneo_b200_stream_hook_output_rate_hz"),
                                       _fn(L, "dsdneo_b200_stream_hook_output_kind"), _fn(L, "dsdneo_b200_stream_hook_symbol_profile"),
                                       _fn(L, "dsdneo_b200_stream_hook_stream_generation"), _fn(L, "dsdneo_b200_stream_hook_cqpsk_status"),
                                       _fn(L, "dsdneo_b200_stream_hook_snr_cqpsk_db"), map_idx)

    def produce():
        for lo in range(0, symbols.size, 700):
            chunk = np.ascontiguousarray(symbols[lo:lo + 700])
            assert L.dsdneo_b200_stream_server_push(srv, chunk.ctypes.data, chunk.size, 1) == chunk.size
        L.dsdneo_b200_stream_server_close(srv)

    t = threading.Thread(target=produce)
    t.start()
    d = np.zeros(n_sym, np.uint8); r = np.zeros(n_sym, np.uint8); l = np.zeros(2 * n_sym, np.int16); s = np.zeros(n_sym, np.float32)
    n = R.ref_sym_get_dibits_n(C.c_void_p(h), n_sym, H._ptr(d, H.u8p), H._ptr(r, H.u8p), l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s))
    buf = np.zeros(4096, np.float32)
    got = C.c_int(0)
    while L.dsdneo_b200_stream_hook_read(srv, buf.ctypes.data, buf.size, C.byref(got)) == 0:
        pass
    t.join()
    R.ref_sym_destroy(C.c_void_p(h))
    L.dsdneo_b200_stream_server_destroy(srv)
    assert n == n_sym
    return d, r, l.reshape(-1, 2), s


@needs_ref
def test_reference_decoder_reads_cqpsk_symbols_through_the_stream_server():
    """Output kind 2 at the seam: the reference decoder's symbol-rate path fed by the library's hooks produces exactly what the
    oracle slicer (pinned to the same reference through the test shim's own hooks) produces."""
    from test_oracle_symbol import cqpsk_symbol_stream

    b = _b200()
    R = C.CDLL(H._ref_path("par"))
    _bind(R)
    rng = np.random.default_rng(23)
    x = cqpsk_symbol_stream(rng, 3200, noise=0.3, offset=0.2)
    n_sym = 3000
    for sync, map_idx, snr in [(H.SYNC_P25P1_POS, 0, -100.0), (H.SYNC_P25P1_NEG, 3, 14.0)]:
        got = _ref_cqpsk_dibits_via_server(R, b, x, sync, map_idx, snr, n_sym)
        d, r, l, _ = H.oracle_cqpsk_slicer_run(x[:n_sym], negative=H.SYNC_CLASS[sync]["negative"], p25_slice=1, map_idx=map_idx, snr_db=snr)
        assert np.array_equal(got[0], d) and np.array_equal(got[1], r) and np.array_equal(got[2], l)
        assert H.bits_equal(got[3], x[:n_sym])


@pytest.mark.gpu
@needs_ref
def test_gpu_cqpsk_chain_feeds_the_unmodified_reference_decoder(gpu):
    """Drop-in at the seam for CQPSK channels: IQ -> GPU channel LPF + CQPSK chain -> stream server (output kind 2) -> the
    reference's own getDibitSoft; its dibits, reliabilities and LLRs equal the GPU symbol-rate slicer's on the same channel."""
    import torch

    R = C.CDLL(H._ref_path("par"))
    _bind(R)
    rng = np.random.default_rng(24)
    bp, nb = 2400, 6
    iq = H.synth_cqpsk_iq(rng, bp * nb // 5 + 2, sps=5, snr_db=20.0, cfo=0.015, timing=0.3)[0][:bp * nb][None]
    bank = gpu.CqpskBank(1, 24000)
    sym, counts = bank.full_demod(torch.from_numpy(np.ascontiguousarray(iq)).cuda(), bp, nb)
    total = counts.sum(dim=1, dtype=torch.int32).contiguous()
    n_all = int(total[0])
    row = sym[0, :n_all].cpu().numpy().copy()
    n_sym = n_all - 20
    got = _ref_cqpsk_dibits_via_server(R, gpu, row, H.SYNC_P25P1_POS, 0, -100.0, n_sym)
    sl = gpu.CqpskSlicer(1)
    res = sl.run(sym, total)
    torch.cuda.synchronize()
    assert np.array_equal(res["dibits"][0, :n_sym].cpu().numpy(), got[0])
    assert np.array_equal(res["reliability"][0, :n_sym].cpu().numpy(), got[1])
    assert np.array_equal(res["llr"][0, :n_sym].cpu().numpy(), got[2])
