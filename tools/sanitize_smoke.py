"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): tiny shapes, no timing."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import _harness as H

b200 = g.load_package()
b200.init(0)
rng = np.random.default_rng(1)
M, bp, nb = 256, 512, 2
x, _ = H.synth_wideband(rng, M, bp * nb, [5, 77], snr_db=25.0)
fe = b200.Frontend(M, 8, False, 12_288_000, bp)
out = torch.empty((M, bp * nb), device="cuda")
for _ in range(2):
    fe.process_async(torch.from_numpy(x).cuda(), out)
fe.join()
disc = fe.process(torch.from_numpy(x).cuda())
bank = b200.DemodBank(40, 48000, True, squelch_levels=[0.0] * 20 + [1e-3] * 20)
bank.full_demod(torch.randn((40, 700, 2), device="cuda") * 0.2, 350, 2)
hb = b200.HalfbandCascade(3, 3)
hb.decimate(torch.randn((3, 1024, 2), device="cuda"), 512, 2)
taps = {0: H.sps_fir_taps(0, 10), 1: H.sps_fir_taps(1, 10)}
sy = b200.Symbolizer(M, 48000, 4800, filters=taps)
sy.set_class([b200.sym_class_from_synctype(0, 0) if c % 2 else b200.sym_class_from_synctype(10, 10) for c in range(M)])
res = sy.run(disc, bp * nb)
res = sy.run(disc, bp * nb, mode=b200.SYM_MODE_GET_SYMBOL, have_sync=0)
fs = b200.FrameSync(M)
fs.search(res["symbols"], res["count"])
b200.bptc_196x96(rng.integers(0, 2, (64, 196)).astype(np.uint8), interleaved=True)
b200.bptc_128x77(rng.integers(0, 2, (64, 128)).astype(np.uint8))
b200.bptc_16x2(rng.integers(0, 2, (64, 32)).astype(np.uint8), True)
b200.p25_12_soft_llr(rng.integers(-300, 300, (32, 196)).astype(np.int16))
b200.p25_12_soft_llr_list(rng.integers(-300, 300, (32, 196)).astype(np.int16))
d, p = rng.integers(0, 2, (32, 120)).astype(np.uint8), rng.integers(0, 2, (32, 96)).astype(np.uint8)
b200.p25_rs_decode(0, d.copy(), p)
b200.p25_rs_soft_reliability(0, d, p, rng.integers(0, 255, (32, 20)).astype(np.uint8), rng.integers(0, 255, (32, 16)).astype(np.uint8))
b200.p25_word_decode(0, rng.integers(0, 2, (32, 6)).astype(np.uint8), rng.integers(0, 2, (32, 12)).astype(np.uint8))
b200.bch_63_16_decode(rng.integers(0, 2, (32, 63)).astype(np.uint8))
cur = (b200.MbeParms * 8)()
prev = (b200.MbeParms * 8)()
for i in range(8):
    for q in (cur[i], prev[i]):
        q.w0, q.L = 0.05 + 0.01 * i, 20 + i
        for l in range(1, q.L + 1):
            q.Vl[l], q.Ml[l] = l % 2, 10.0 + l
b200.mbe_synth(cur, prev)
# CQPSK block side + symbol-rate slicer (ragged warp, mixed sps, a squelched stretch, two launches)
n_cq = 37
iq = np.stack([H.synth_cqpsk_iq(rng, 700, sps=5 if c % 3 else 4, snr_db=15.0, cfo=0.01)[0][:2400] for c in range(n_cq)]).copy()
iq[3, 600:1800] *= np.float32(1e-4)
iq[4, 1000:1400] = 0.0
cq = b200.CqpskBank(n_cq, 24000, ted_sps=[5 if c % 3 else 4 for c in range(n_cq)], squelch_levels=[1e-3 if c == 3 else 0.0 for c in range(n_cq)])
sl = b200.CqpskSlicer(n_cq)
sl.set_class(np.arange(n_cq) % 2, np.ones(n_cq), np.arange(n_cq) % 5, 12.0)
for lo in (0, 1200):
    sym, counts = cq.full_demod(torch.from_numpy(np.ascontiguousarray(iq[:, lo:lo + 1200])).cuda(), 600, 2)
    sl.run(sym, counts.sum(dim=1, dtype=torch.int32).contiguous())
# P25p1 frame cutter + NID decode + soft word decoders, DMR burst cutter on the symbolizer outputs above
res2 = sy.run(disc, bp * nb)
hits, n_hits = fs.search(res2["symbols"], res2["count"], max_hits=8)
cut = b200.p25p1_frame_cut(res2["dibits"], res2["llr"], res2["count"], hits, n_hits, 3 * 98)
b200.dmr_burst_cut(res2["dibits"], res2["reliability"], res2["count"], hits, n_hits)
k = cut["nid_valid"].shape[0]
b200.p25p1_nid_decode(rng.integers(0, 2, (64, 63)).astype(np.uint8), rng.integers(0, 256, (64, 63)).astype(np.uint8),
                      rng.integers(0, 4096, 64).astype(np.int32), rng.integers(0, 2, 64).astype(np.uint8), rng.integers(0, 256, 64).astype(np.uint8))
b200.p25p1_nid_decode(rng.integers(0, 2, (5, 63)).astype(np.uint8), None, None, rng.integers(0, 2, 5).astype(np.uint8), None)
L = b200.lib()
bits10, rel10 = rng.integers(0, 2, (40, 10)).astype(np.uint8), rng.integers(0, 256, (40, 10)).astype(np.int32)
out10, st10 = np.zeros_like(bits10), np.zeros(40, np.uint8)
b200.check(L.dsdneo_b200_hamming_10_6_3_soft_batch_host(bits10.ctypes.data, rel10.ctypes.data, 1, 64, out10.ctypes.data, st10.ctypes.data, 40))
for code, ln in ((b200.P25_WORD_GOLAY_24_6, 6), (b200.P25_WORD_GOLAY_24_12, 12)):
    d, pbits = rng.integers(0, 2, (40, ln)).astype(np.uint8), rng.integers(0, 2, (40, 12)).astype(np.uint8)
    r = rng.integers(0, 256, (40, ln + 12)).astype(np.int32)
    stg, fxg = np.zeros(40, np.uint8), np.zeros(40, np.int32)
    b200.check(L.dsdneo_b200_p25_golay_soft_batch_host(code, d.ctypes.data, pbits.ctypes.data, r.ctypes.data, 1, 64, stg.ctypes.data, fxg.ctypes.data, 40))
# round 2: general / bin-pruned channelizer (cu8 and cf32, ragged launch lengths, cu8 output), vocoder frame ECC, DMR voice cutter,
# acquisition and the P25 receive bank (frames incl. TDULC, acquisition at stream start)
for Mx, R, cu8 in ((512, 2, True), (1024, 1, False), (2048, 8, True)):
    cz = b200.Channelizer(Mx, 8, cu8)
    for n_o in (37, 5):
        xin = (torch.randint(0, 256, (n_o * Mx, 2), dtype=torch.uint8, device="cuda") if cu8 else torch.randn((n_o * Mx, 2), device="cuda"))
        cz.channelize_bins(xin, R, R - 1)
    cz.prime(xin)
    xin = (torch.randint(0, 256, (40 * Mx, 2), dtype=torch.uint8, device="cuda") if cu8 else torch.randn((40 * Mx, 2), device="cuda"))
    cz.channelize_bins_cu8(xin, R, 0, 2.0)
b200.ambe3600x2450_decode(rng.integers(0, 2, (70, 4, 24)).astype(np.uint8))
b200.imbe7200x4400_decode(rng.integers(0, 2, (70, 8, 23)).astype(np.uint8))
b200.p25p1_voice_imbe_decode(torch.randint(0, 256, (5, 1944), dtype=torch.uint8, device="cuda"), 5)
b200.dmr_voice_cut(res2["dibits"], res2["count"], hits, n_hits, 8, 3)
import test_gpu_p25p1_rx as T

chans = [T._channel(rng, 3400, snr_db=22.0) for _ in range(3)]
p25_taps = H.sps_fir_taps(0, 10)
for acq, watch in ((0, 0), (1, 0), (0, 1)):
    # watch = 1 with a channel of noise: the loss-of-sync watch drops it (sync_watch_kernel, sym_drop_kernel) and the hunt kernel
    # runs inside the pipeline
    rx = b200.P25p1Rx(3, p25_taps, block_pairs=T.BP, max_pairs_per_call=2 * T.BP, acquire_tiles=acq, auto_reacquire_tiles=watch)
    rx_out = rx.alloc_device_out("cuda")
    for kk in range(4):
        tile = np.stack([u8[kk * 2 * T.BP:(kk + 1) * 2 * T.BP] for u8, _ in chans])
        if watch:
            tile[2] = rng.integers(96, 160, size=tile[2].shape, dtype=np.uint8)
        tk = rx.submit(torch.from_numpy(tile).cuda(), 2 * T.BP, rx_out)
        rx.wait(tk)
    if watch:
        rx.reacquire([1, 0, 1], tiles=1)
        tk = rx.submit(torch.from_numpy(tile).cuda(), 2 * T.BP, rx_out)
        rx.wait(tk)
    torch.cuda.synchronize()
torch.cuda.synchronize()
print("sanitize_smoke done, launches =", b200.launch_count())
