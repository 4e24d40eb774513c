"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle would take minutes per config):
  C2  256 channels x 49152 samples (12,582,912 wideband samples): a few channels against the oracle at full length, the rest
      through invariances -- one call == block-by-block streaming == the two-stream pipeline == the host-buffer path,
      and per-channel independence (a channel's output does not depend on which other channels share the launch).
  C3  1024 channels of sample side: channels fed identical signals produce identical outputs whatever lane / CTA they land
      in, and a few channels equal the oracle at full length."""
import ctypes as C
import zlib

import numpy as np
import pytest

import _harness as H
from test_gpu_symbolizer import _oracle_dibits, _taps

pytestmark = pytest.mark.gpu
M = 256


def _crc(t):
    return zlib.crc32(t.contiguous().cpu().numpy().tobytes())


def test_c2_full_size_invariances_and_spot_parity(gpu):
    import torch

    bp, nb = 8192, 6
    n_out = bp * nb
    g = torch.Generator(device="cuda").manual_seed(12)
    # 256 FSK-like carriers are expensive to synthesise exactly; a dense random-phase multitone + noise exercises the same
    # arithmetic (every channel non-trivial, discriminator in both atan branches)
    x = torch.randn((n_out * M, 2), device="cuda", generator=g) * 0.05
    t = torch.arange(n_out * M, device="cuda", dtype=torch.float32)
    for k in (3, 50, 77, 128, 200, 255):
        ph = 2 * torch.pi * ((k / M) * t % 1.0) + 0.4 * torch.sin(t * (2 * torch.pi * 1200.0 / 12_288_000) * (1 + k % 3))
        x[:, 0] += 0.2 * torch.cos(ph)
        x[:, 1] += 0.2 * torch.sin(ph)
    x = x.contiguous()
    fa = gpu.Frontend(M, 8, False, 12_288_000, bp)
    whole = fa.process(x)
    torch.cuda.synchronize()
    # (1) streaming invariance: six one-block calls == one six-block call (state carried, block padding identical)
    fb = gpu.Frontend(M, 8, False, 12_288_000, bp)
    parts = [fb.process(x[b * bp * M:(b + 1) * bp * M].contiguous()) for b in range(nb)]
    assert _crc(torch.cat(parts, dim=1)) == _crc(whole)
    # (2) pipelined and host-buffer paths
    fc = gpu.Frontend(M, 8, False, 12_288_000, bp)
    out_c = torch.empty_like(whole)
    fc.process_async(x, out_c)
    fc.join()
    torch.cuda.synchronize()
    assert _crc(out_c) == _crc(whole)
    fd = gpu.Frontend(M, 8, False, 12_288_000, bp)
    out_d = torch.from_numpy(fd.process_host(x.cpu().numpy()))
    assert _crc(out_d) == _crc(whole)
    # (3) spot parity at full length: full_demod of the GPU channelizer's output vs the oracle, three channels
    cz = gpu.Channelizer(M, 8)
    chan = cz.channelize(x)
    torch.cuda.synchronize()
    for k in (50, 128, 255):
        want = H.oracle_full_demod(chan[k].cpu().numpy(), bp, nb, fir_fma=1)
        assert H.bits_equal(whole[k].cpu().numpy(), want), k
    # (4) channel independence: a 32-channel bank over rows 96..127 reproduces those rows of the 256-channel result
    bank = gpu.DemodBank(32, 48000, True)
    sub = bank.full_demod(chan[96:128].contiguous(), bp, nb)
    assert _crc(sub) == _crc(whole[96:128])


def test_c3_full_size_sample_side_replicas_and_spot_parity(gpu):
    import torch

    rng = np.random.default_rng(13)
    n_ch, n_samp = 1024, 49150
    base = []
    for c in range(4):
        dib = rng.integers(0, 4, n_samp // 10 + 2)
        base.append(H.synth_c4fm_disc(rng, dib, 9000.0, 500.0 + 300.0 * c)[:n_samp])
    order = rng.integers(0, 4, n_ch)
    x = torch.from_numpy(np.stack([base[i] for i in order])).cuda()
    taps = _taps()
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(H.SYNC_P25P1_POS, H.SYNC_P25P1_POS)] * n_ch)
    res = sy.run(x, n_samp)
    torch.cuda.synchronize()
    cnt = res["count"].cpu().numpy()
    assert (cnt == cnt[0]).all()
    k = int(cnt[0])
    sym, dib, llr = res["symbols"][:, :k].cpu().numpy(), res["dibits"][:, :k].cpu().numpy(), res["llr"][:, :k].cpu().numpy()
    first = {i: int(np.nonzero(order == i)[0][0]) for i in range(4)}
    for c in range(n_ch):  # every replica equals the first channel that carries the same signal, bit for bit
        f = first[int(order[c])]
        assert np.array_equal(sym[c].view(np.uint32), sym[f].view(np.uint32)) and np.array_equal(dib[c], dib[f])
        assert np.array_equal(llr[c], llr[f])
    for i in range(4):  # and those four equal the oracle at full length
        wd, wr, wl, ws = _oracle_dibits(base[i], H.SYNC_P25P1_POS, taps)
        f = first[i]
        assert wd.size == k and np.array_equal(dib[f], wd) and np.array_equal(llr[f], wl) and H.bits_equal(sym[f], ws)


def test_c5_channel_count_mixed_classes_replicas_and_spot_parity(gpu):
    """8192 channels in one launch, protocol classes mixed per channel (P25 C4FM with threshold tracking, DMR/YSF class with
    fixed thresholds, no matched filter): replicas of six base signals are identical wherever they land, and the six equal
    the oracle."""
    import torch

    rng = np.random.default_rng(14)
    n_ch, n_samp = 8192, 24570
    taps = _taps()
    syncs = [H.SYNC_P25P1_POS, H.SYNC_DMR_BS_DATA_POS, H.SYNC_NONE, H.SYNC_P25P1_NEG, H.SYNC_DMR_BS_DATA_POS, H.SYNC_P25P1_POS]
    base = []
    for i, s in enumerate(syncs):
        dib = rng.integers(0, 4, n_samp // 10 + 2)
        if s == H.SYNC_DMR_BS_DATA_POS:
            base.append(H.synth_dmr_disc(rng, dib, taps[1], 10000.0, 300.0 * i)[:n_samp])
        else:
            base.append(H.synth_c4fm_disc(rng, dib, 9000.0, 250.0 * i)[:n_samp])
    order = rng.integers(0, 6, n_ch)
    x = torch.empty((n_ch, n_samp), dtype=torch.float32, device="cuda")
    d_base = torch.from_numpy(np.stack(base)).cuda()
    x.copy_(d_base[torch.from_numpy(order).cuda()])
    sy = gpu.Symbolizer(n_ch, 48000, 4800, filters=taps)
    sy.set_class([gpu.sym_class_from_synctype(syncs[i], syncs[i]) for i in order])
    res = sy.run(x, n_samp)
    torch.cuda.synchronize()
    cnt = res["count"].cpu().numpy()
    k = int(cnt.min())
    assert cnt.max() - k <= 1
    sym, dib = res["symbols"][:, :k].cpu().numpy(), res["dibits"][:, :k].cpu().numpy()
    first = {i: int(np.nonzero(order == i)[0][0]) for i in range(6)}
    ref_rows = np.array([first[int(i)] for i in order])
    assert np.array_equal(sym.view(np.uint32), sym[ref_rows].view(np.uint32)) and np.array_equal(dib, dib[ref_rows])
    for i in range(6):
        wd, wr, wl, ws = _oracle_dibits(base[i], syncs[i], taps)
        f = first[i]
        assert np.array_equal(dib[f], wd[:k]) and H.bits_equal(sym[f], ws[:k]), i


def test_c5_channel_count_cqpsk_streaming_and_spot_parity(gpu):
    """8192 CQPSK channels (mixed sps 5 / 4 = P25 LSM and Phase 2 at 24 kS/s) through channel LPF + chain + symbol-rate slicer:
    one launch of four blocks equals two launches of two blocks (symbols, counts, dibits, loop state), replicated channels
    agree bit for bit, and a few channels equal the CPU oracle chain."""
    import torch

    rng = np.random.default_rng(8192)
    n_ch, bp, nb = 8192, 1200, 4
    sps = np.array([5 if c % 4 else 4 for c in range(n_ch)])
    base = {}
    for s in (4, 5):
        base[s] = [H.synth_cqpsk_iq(rng, bp * nb // s + 2, sps=s, snr_db=[None, 18.0, 9.0][k % 3], cfo=0.01 * (k - 3), timing=0.1 * k)[0][:bp * nb]
                   for k in range(8)]
    iq = np.stack([base[int(sps[c])][c % 8] for c in range(n_ch)])
    d_iq = torch.from_numpy(iq).cuda()

    def run(splits):
        bank = gpu.CqpskBank(n_ch, 24000, ted_sps=sps.tolist())
        sl = gpu.CqpskSlicer(n_ch)
        syms, dibs, cnts = [], [], []
        for lo, hi in splits:
            sym, counts = bank.full_demod(d_iq[:, lo * bp:hi * bp].contiguous(), bp, hi - lo)
            tot = counts.sum(dim=1, dtype=torch.int32).contiguous()
            res = sl.run(sym, tot)
            syms.append(sym.cpu().numpy()); dibs.append(res["dibits"].cpu().numpy()); cnts.append(counts.cpu().numpy())
        return bank, syms, dibs, cnts

    bank1, s1, d1, c1 = run([(0, 4)])
    bank2, s2, d2, c2 = run([(0, 2), (2, 4)])
    assert np.array_equal(c1[0], np.concatenate(c2, axis=1))
    for c in list(range(0, n_ch, 97)) + [n_ch - 1]:
        n_a = int(c2[0][c].sum()); n_b = int(c2[1][c].sum())
        assert H.bits_equal(s1[0][c, :n_a + n_b], np.concatenate([s2[0][c, :n_a], s2[1][c, :n_b]]))
        assert np.array_equal(d1[0][c, :n_a + n_b], np.concatenate([d2[0][c, :n_a], d2[1][c, :n_b]]))
        a, b = bank1.state(c), bank2.state(c)
        for f, _ in a._fields_:
            assert getattr(a, f) == getattr(b, f) or (isinstance(getattr(a, f), float) and np.float32(getattr(a, f)).tobytes() == np.float32(getattr(b, f)).tobytes()), (c, f)
    # replicas: channels c and c + 32 carry the same signal whenever sps and (c % 8) agree -> compare c with c + 32 * 4 ... use c, c + 64
    for c in (0, 5, 1000, 4097):
        r = c + 64
        n = int(c1[0][c].sum())
        assert sps[c] == sps[r] and np.array_equal(c1[0][c], c1[0][r]) and H.bits_equal(s1[0][c, :n], s1[0][r, :n])
    for c in (3, 4, 8191):
        orc = H.OracleCqpsk(rate=24000, sps=int(sps[c]), fir_fma=1)
        want_sym, want_counts = orc.run(iq[c], bp, nb)
        n = int(want_counts.sum())
        assert np.array_equal(c1[0][c], want_counts) and H.bits_equal(s1[0][c, :n], want_sym)
        d, _, _, _ = H.oracle_cqpsk_slicer_run(want_sym)
        assert np.array_equal(d1[0][c, :n], d)


def test_c5_size_channelizer_classes_agree_with_the_unpruned_transform(gpu):
    """C5 size: one 8192-channel cu8 tile (6144 rows = 50.3 M wideband samples).  Two of the eight per-GPU channel classes equal
    the unpruned 8192-point transform within the stage's tolerance, every tone sits in its own channel, and streaming the tile
    in three launches is bit-identical to one launch."""
    import torch

    Mx, R, T, n_out = 8192, 8, 8, 6144
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((n_out * Mx, 2), device="cuda", generator=g) * 0.08
    t = torch.arange(n_out * Mx, device="cuda", dtype=torch.float64)
    tones = [11, 1234, 4096 + 77, 8190, 2 * 8 + 5, 777 * 8 + 5]
    for k in tones:
        ph = (2 * torch.pi * ((k / Mx) * t % 1.0)).to(torch.float32)
        x[:, 0] += 0.12 * torch.cos(ph)
        x[:, 1] += 0.12 * torch.sin(ph)
    u8 = torch.clamp(torch.round(x * 127.5 + 127.5), 0, 255).to(torch.uint8).contiguous()
    del x, t
    full = gpu.Channelizer(Mx, T, True).channelize(u8)
    scale = float(full.abs().max())
    for r0 in (5, 0):
        cz = gpu.Channelizer(Mx, T, True)
        y = cz.channelize_bins(u8, R, r0)
        assert float((y - full[r0::R]).abs().max()) <= 2e-5 * scale + 1e-7
        cz2 = gpu.Channelizer(Mx, T, True)
        parts = [cz2.channelize_bins(u8[a * Mx:b * Mx], R, r0) for a, b in ((0, 1000), (1000, 1003), (1003, n_out))]
        assert _crc(torch.cat(parts, dim=1)) == _crc(y)
        p = (y[..., 0] ** 2 + y[..., 1] ** 2).mean(dim=1)
        med = float(p.median())
        for k in tones:
            if k % R == r0:
                assert float(p[k // R]) > 30 * med, (k, float(p[k // R]), med)
