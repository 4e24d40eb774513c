"""Configuration C1 (BASELINE.json configs[0]): the reference's own IQ-replay fixture p25p1_c4fm_cc.iq (P25 Phase 1 C4FM control
channel, cu8 at 48 kS/s; the reference's CLI test expects "NAC/CC: 140" from it).  tests/golden/c1_p25p1_c4fm_cc.npz holds the
capture and what the UNMODIFIED reference block side + sample side produce from it (tests/golden/make_c1_golden.py).
CPU test: the oracle chain reproduces the reference bit for bit and decodes NAC 0x140 from every TSDU.
GPU test: the CUDA chain (full_demod -> symbolizer hunt + dibits -> sync hunt -> frame cutter -> NID decode -> trellis) does."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import _harness as H

G = np.load(os.path.join(H.GOLDEN_DIR, "c1_p25p1_c4fm_cc.npz"))
SYNC = "111113113311333313133333"


def _crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def _widen(u8):
    return ((u8.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).reshape(-1, 2)


def test_c1_oracle_chain_reproduces_the_reference_and_decodes_nac_140():
    from test_frame_sync import oracle_search
    from test_oracle_fec import _oracle_cut

    x = _widen(G["iq_cu8"])
    bp = int(G["block_pairs"])
    disc = H.oracle_full_demod(x, bp, x.shape[0] // bp, fir_fma=0)
    assert _crc(disc) == G["disc_crc"] and H.bits_equal(disc[:64], G["disc_head"])
    taps = H.sps_fir_taps(0, 10)
    ch = H.OracleSymChan()
    O = H.oracle_sym()
    O.oracle_sym_init(C.byref(ch), 48000, 4800, 1, 2, 1, 0, H._ptr(taps), taps.size, 128, 1024)
    n_hunt = int(G["n_hunt_samples"])
    hunt = np.zeros(n_hunt // 8, np.float32)
    cons = C.c_long(0)
    k_a = O.oracle_sym_run_symbols(C.byref(ch), 0, H._ptr(disc), n_hunt, 12, H._ptr(hunt), hunt.size, C.byref(cons))
    assert k_a == int(G["hunt_count"]) and cons.value == int(G["hunt_consumed"]) and _crc(hunt[:k_a]) == G["hunt_crc"]
    rest = np.ascontiguousarray(disc[cons.value:])
    n = rest.size // 9
    d, r, l, s = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(2 * n, np.int16), np.zeros(n, np.float32)
    c2 = C.c_long(0)
    k_b = O.oracle_sym_run_dibits(C.byref(ch), H._ptr(rest), rest.size, 12, H._ptr(d, H.u8p), H._ptr(r, H.u8p),
                                  l.ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(s), n, C.byref(c2))
    m = G["dibits"].size
    assert k_b >= m  # the reference harness stops 600 samples early (its 512-float cache); the oracle runs to the end
    assert np.array_equal(d[:m], G["dibits"]) and np.array_equal(r[:m], G["reliab"]) and np.array_equal(l[:2 * m].reshape(-1, 2), G["llr"])
    assert _crc(s[:m]) == G["symbols_crc"]
    # frames: sync hunt on the symbols, sequential cutter, NID decode -> NAC 0x140, DUID 7 (TSDU) on every frame
    n_hits, pos, typ, _, _ = oracle_search(s[:k_b], [(SYNC, 0)], max_hits=64)
    Of = H.oracle_fec()
    Of.oracle_p25p1_nid_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    good = 0
    for p in pos[:n_hits].tolist():
        flags, code, rel, par, prel, pd, pl = _oracle_cut(d[:k_b], l[:2 * k_b].reshape(-1, 2), int(p), 3 * 98)
        if not (flags & 1):
            continue
        v = [C.c_int() for _ in range(3)]
        st = Of.oracle_p25p1_nid_decode(H._ptr(code, H.u8p), H._ptr(rel, H.u8p), 0, par, prel, 64, *[C.byref(a) for a in v])
        good += int(st == 1 and v[0].value == int(G["expected_nac"]) and v[1].value == 7)
    assert n_hits >= 25 and good >= n_hits - 1, (n_hits, good)


@pytest.mark.gpu
def test_c1_gpu_chain_reproduces_the_reference_and_decodes_nac_140(gpu):
    import torch

    x = _widen(G["iq_cu8"])
    bp = int(G["block_pairs"])
    nb = x.shape[0] // bp
    bank = gpu.DemodBank(1, 48000, True, profiles=[4], fir_arith=gpu.FIR_ARITH_NOFMA)  # the golden run used the SSE2 reference build
    disc = bank.full_demod(torch.from_numpy(np.ascontiguousarray(x[None])).cuda(), bp, nb)
    disc_h = disc[0].cpu().numpy()
    assert _crc(disc_h) == G["disc_crc"]
    taps = {0: H.sps_fir_taps(0, 10)}
    sy = gpu.Symbolizer(1, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_P25P1_POS, H.SYNC_P25P1_POS)])
    n_hunt = int(G["n_hunt_samples"])
    res_a = sy.run(disc[:, :n_hunt].contiguous(), n_hunt, mode=gpu.SYM_MODE_GET_SYMBOL, have_sync=0)
    k_a = int(res_a["count"][0])
    assert k_a == int(G["hunt_count"]) and _crc(res_a["symbols"][0, :k_a].cpu().numpy()) == G["hunt_crc"]
    rest = disc[:, n_hunt:].contiguous()
    res = sy.run(rest, rest.shape[1])  # getDibitSoft mode; the unconsumed tail of the hunt launch is carried over
    k_b = int(res["count"][0])
    m = G["dibits"].size
    assert k_b >= m
    assert np.array_equal(res["dibits"][0, :m].cpu().numpy(), G["dibits"])
    assert np.array_equal(res["reliability"][0, :m].cpu().numpy(), G["reliab"])
    assert np.array_equal(res["llr"][0, :m].cpu().numpy(), G["llr"])
    assert _crc(res["symbols"][0, :m].cpu().numpy()) == G["symbols_crc"]
    # frames on the device: sync hunt -> cutter -> NID decode + trellis blocks
    fs = gpu.FrameSync(1, [(SYNC, 0)])
    hits, n_hits = fs.search(res["symbols"], res["count"], max_hits=64)
    cut = gpu.p25p1_frame_cut(res["dibits"], res["llr"], res["count"], hits, n_hits, 3 * 98)
    k = 64
    st = torch.zeros(k, dtype=torch.int8, device="cuda")
    nac = torch.zeros(k, dtype=torch.int32, device="cuda")
    duid = torch.zeros(k, dtype=torch.uint8, device="cuda")
    errs = torch.zeros(k, dtype=torch.int32, device="cuda")
    L = gpu.lib()
    gpu.check(L.dsdneo_b200_p25p1_nid_decode_batch(cut["nid_code63"].data_ptr(), cut["nid_reliab63"].data_ptr(), None,
                                                   cut["nid_parity"].data_ptr(), cut["nid_parity_reliab"].data_ptr(), 64, st.data_ptr(),
                                                   nac.data_ptr(), duid.data_ptr(), errs.data_ptr(), k, None))
    blocks = cut["payload_llr"].reshape(k * 3, 196).contiguous()
    out12 = torch.zeros((k * 3, 12), dtype=torch.uint8, device="cuda")
    met = torch.zeros(k * 3, dtype=torch.int32, device="cuda")
    gpu.check(L.dsdneo_b200_p25_12_soft_llr_batch(blocks.data_ptr(), out12.data_ptr(), met.data_ptr(), k * 3, None))
    torch.cuda.synchronize()
    nh = int(n_hits[0])
    valid = cut["nid_valid"].cpu().numpy().astype(bool)
    st, nac, duid = st.cpu().numpy(), nac.cpu().numpy(), duid.cpu().numpy()
    good = int(((st[:nh] == 1) & (nac[:nh] == int(G["expected_nac"])) & (duid[:nh] == 7) & valid[:nh]).sum())
    assert nh >= 25 and good >= nh - 1, (nh, good)
    # the trellis blocks of complete frames decode to consistent words: re-running the oracle on the cut LLRs gives the same bytes
    O = H.oracle_fec()
    pv = cut["payload_valid"].cpu().numpy().astype(bool)
    blocks_h, out_h = blocks.cpu().numpy(), out12.cpu().numpy()
    for f in np.nonzero(pv[:nh])[0][:8]:
        for b in range(3):
            w = np.zeros(12, np.uint8)
            O.oracle_p25_12_soft_llr(blocks_h[3 * f + b].ctypes.data_as(C.POINTER(C.c_int16)), H._ptr(w, H.u8p))
            assert np.array_equal(out_h[3 * f + b], w)


# ---- the reference's CQPSK / LSM control-channel fixture through the CQPSK chain ----------------------------------------------

GQ = np.load(os.path.join(H.GOLDEN_DIR, "c1_p25p1_cqpsk_cc.npz"))


def _nids_ok(dib, llr, n_dib, hit_pos, expected_nac):
    from test_oracle_fec import _oracle_cut

    Of = H.oracle_fec()
    Of.oracle_p25p1_nid_decode.argtypes = [H.u8p, H.u8p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    good = 0
    for p in hit_pos:
        flags, code, rel, par, prel, _, _ = _oracle_cut(dib[:n_dib], llr[:n_dib], int(p), 0)
        if not (flags & 1):
            continue
        v = [C.c_int() for _ in range(3)]
        st = Of.oracle_p25p1_nid_decode(H._ptr(code, H.u8p), H._ptr(rel, H.u8p), 0, par, prel, 64, *[C.byref(a) for a in v])
        good += int(st == 1 and v[0].value == expected_nac and v[1].value == 7)
    return good


def test_c1_cqpsk_oracle_chain_reproduces_the_reference():
    x = _widen(GQ["iq_cu8"])
    bp = int(GQ["block_pairs"])
    sym, counts = H.OracleCqpsk(rate=48000, sps=10, fir_fma=0).run(x, bp, x.shape[0] // bp)
    assert np.array_equal(counts, GQ["counts"]) and _crc(sym) == GQ["symbols_crc"] and H.bits_equal(sym[:64], GQ["symbols_head"])
    d, r, l, _ = H.oracle_cqpsk_slicer_run(sym)
    m = GQ["dibits"].size
    assert np.array_equal(d[:m], GQ["dibits"]) and np.array_equal(r[:m], GQ["reliab"]) and np.array_equal(l[:m], GQ["llr"])
    sync = np.array(H.P25P1_SYNC_DIBITS)
    hits = [i + 23 for i in range(d.size - 24) if np.array_equal(d[i:i + 24], sync)]
    assert len(hits) >= 50 and _nids_ok(d, l, d.size, hits, int(GQ["expected_nac"])) >= len(hits) - 1


@pytest.mark.gpu
def test_c1_cqpsk_gpu_chain_reproduces_the_reference(gpu):
    import torch

    x = _widen(GQ["iq_cu8"])
    bp = int(GQ["block_pairs"])
    nb = x.shape[0] // bp
    bank = gpu.CqpskBank(1, 48000, ted_sps=[10], fir_arith=gpu.FIR_ARITH_NOFMA)
    sym, counts = bank.full_demod(torch.from_numpy(np.ascontiguousarray(x[None])).cuda(), bp, nb)
    total = counts.sum(dim=1, dtype=torch.int32).contiguous()
    n = int(total[0])
    sym_h = sym[0, :n].cpu().numpy()
    assert np.array_equal(counts[0].cpu().numpy(), GQ["counts"]) and _crc(sym_h) == GQ["symbols_crc"]
    res = gpu.CqpskSlicer(1).run(sym, total)
    m = GQ["dibits"].size
    d, l = res["dibits"][0, :n].cpu().numpy(), res["llr"][0, :n].cpu().numpy()
    assert np.array_equal(d[:m], GQ["dibits"]) and np.array_equal(res["reliability"][0, :m].cpu().numpy(), GQ["reliab"])
    assert np.array_equal(l[:m], GQ["llr"])
    # frames on the device: the sync hunt works on symbol signs, which CQPSK symbols near {-3,-1,+1,+3} satisfy as well
    fs = gpu.FrameSync(1, [(SYNC, 0)])
    hits, n_hits = fs.search(sym, total, max_hits=64)
    cut = gpu.p25p1_frame_cut(res["dibits"], res["llr"], total, hits, n_hits, 3 * 98)
    st, nac, duid, errs = gpu.p25p1_nid_decode(cut["nid_code63"].cpu().numpy(), cut["nid_reliab63"].cpu().numpy(), None,
                                               cut["nid_parity"].cpu().numpy(), cut["nid_parity_reliab"].cpu().numpy())
    nh = int(n_hits[0])
    valid = cut["nid_valid"].cpu().numpy().astype(bool)
    good = int(((st[:nh] == 1) & (nac[:nh] == int(GQ["expected_nac"])) & (duid[:nh] == 7) & valid[:nh]).sum())
    assert nh >= 50 and good >= nh - 1, (nh, good)
