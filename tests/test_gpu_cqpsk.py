"""GPU parity tests (-m gpu): the batched CQPSK symbol output kind of full_demod (AGC -> FLL band-edge -> Gardner/MMSE ->
diff phasor -> Costas -> 4/pi atan) through the C-ABI against the CPU oracle (oracle/oracle_cqpsk.c, itself pinned to
the unmodified reference by tests/test_oracle_cqpsk.py) and the committed reference vectors (tests/golden/cqpsk.npz).
Bar: bit-exact symbols, per-block counts and carried loop state."""
import os

import numpy as np
import pytest

import _harness as H

pytestmark = pytest.mark.gpu

STATE_MAP = [("agc_avg", "cqpsk_agc_avg"), ("fll_phase", "fll_phase"), ("fll_freq", "fll_freq"), ("fll_alpha", "fll_alpha"),
             ("fll_beta", "fll_beta"), ("mu", "ted_mu"), ("omega", "ted_omega"), ("last_r", "ted_last_r"),
             ("last_j", "ted_last_j"), ("lock_accum", "ted_lock_accum"), ("lock_count", "ted_lock_count"),
             ("ted_effective_gain", "ted_effective_gain"), ("diff_prev_r", "cqpsk_diff_prev_r"),
             ("diff_prev_j", "cqpsk_diff_prev_j"), ("costas_phase", "costas_phase"), ("costas_freq", "costas_freq"),
             ("costas_error", "costas_error"), ("costas_err_smooth", "costas_error_smooth"),
             ("costas_err_avg_q14", "costas_err_avg_q14"), ("costas_err_raw_avg_q14", "costas_err_raw_avg_q14"),
             ("costas_conf_avg_q14", "costas_conf_avg_q14"), ("costas_zero_conf_pct", "costas_zero_conf_pct")]


def _check_state(bank, c, orc):
    st, lst, want = bank.state(c), bank.lpf.state(c), orc.state()
    bad = []
    for ok, gk in STATE_MAP:
        a, b = want[ok], getattr(st, gk)
        same = (a == b) if isinstance(a, int) else (np.float32(a).tobytes() == np.float32(b).tobytes())
        if not same:
            bad.append((ok, a, b))
    if np.float32(want["channel_pwr"]).tobytes() != np.float32(lst.channel_pwr).tobytes():
        bad.append(("channel_pwr", want["channel_pwr"], lst.channel_pwr))
    if want["channel_squelched"] != lst.channel_squelched:
        bad.append(("channel_squelched", want["channel_squelched"], lst.channel_squelched))
    assert not bad, (c, bad)
    assert st.overflow == 0


def _compare(sym, counts, c, want_sym, want_counts):
    assert np.array_equal(counts[c], want_counts), (c, counts[c], want_counts)
    n = int(want_counts.sum())
    assert H.bits_equal(sym[c, :n], want_sym), (c, H.first_mismatch(sym[c, :n], want_sym))


def _signals(rng, n_ch, sps, n, snrs, cfos):
    return np.stack([H.synth_cqpsk_iq(rng, n // sps + 2, sps=sps, snr_db=snrs[c % len(snrs)], cfo=cfos[c % len(cfos)],
                                      timing=0.1 * c, phase0=0.2 * c)[0][:n] for c in range(n_ch)])


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("sps,rate,bp,nb", [(5, 24000, 2400, 4), (5, 24000, 997, 5), (4, 24000, 1600, 8), (10, 48000, 4800, 2),
                                            (8, 48000, 3000, 5), (5, 24000, 36, 9), (5, 24000, 4, 40)])
def test_cqpsk_bit_exact_vs_oracle(gpu, arith, sps, rate, bp, nb):
    import torch

    rng = np.random.default_rng(1000 + bp)
    n_ch = 37  # two warps, the second one ragged
    iq = _signals(rng, n_ch, sps, bp * nb, [None, 20.0, 12.0, 6.0, 2.0], [0.0, 0.02, -0.03, 0.08])
    bank = gpu.CqpskBank(n_ch, rate, ted_sps=[sps] * n_ch, fir_arith=arith)
    sym, counts = bank.full_demod(torch.from_numpy(iq).cuda(), bp, nb)
    sym, counts = sym.cpu().numpy(), counts.cpu().numpy()
    for c in range(n_ch):
        orc = H.OracleCqpsk(rate=rate, sps=sps, fir_fma=1 - arith)
        want_sym, want_counts = orc.run(iq[c], bp, nb)
        _compare(sym, counts, c, want_sym, want_counts)
        _check_state(bank, c, orc)


def test_cqpsk_mixed_sps_squelch_and_launch_split(gpu):
    """Channels of different sps in one warp, squelched blocks in some channels, and the same stream fed as three
    launches: equals the oracle run block by block."""
    import torch

    rng = np.random.default_rng(77)
    n_ch, bp, nb, rate = 40, 1200, 6, 24000
    sps = [5 if c % 3 else 4 for c in range(n_ch)]
    iq = np.stack([H.synth_cqpsk_iq(rng, bp * nb // sps[c] + 2, sps=sps[c], snr_db=18.0, cfo=0.01 * (c % 5 - 2),
                                    timing=0.07 * c)[0][:bp * nb] for c in range(n_ch)]).copy()
    levels = [1e-3 if c % 4 == 1 else 0.0 for c in range(n_ch)]
    for c in range(n_ch):
        if c % 4 == 1:
            iq[c, bp * (1 + c % 3):bp * (3 + c % 3)] *= np.float32(1e-4)
    bank = gpu.CqpskBank(n_ch, rate, ted_sps=sps, squelch_levels=levels)
    syms, cnts = [], []
    for lo, hi in [(0, 1), (1, 4), (4, 6)]:
        s, k = bank.full_demod(torch.from_numpy(np.ascontiguousarray(iq[:, lo * bp:hi * bp])).cuda(), bp, hi - lo)
        syms.append(s.cpu().numpy())
        cnts.append(k.cpu().numpy())
    seen_sq = False
    for c in range(n_ch):
        orc = H.OracleCqpsk(rate=rate, sps=sps[c], squelch=levels[c], fir_fma=1)
        for (lo, hi), s, k in zip([(0, 1), (1, 4), (4, 6)], syms, cnts):
            want_sym, want_counts = orc.run(iq[c, lo * bp:hi * bp], bp, hi - lo)
            _compare(s, k, c, want_sym, want_counts)
            seen_sq = seen_sq or orc.state()["channel_squelched"] == 1 or bool((want_sym == 0).all() and want_sym.size)
        _check_state(bank, c, orc)
    assert seen_sq


def test_cqpsk_golden_vectors(gpu):
    """The committed outputs of the unmodified reference (tests/golden/cqpsk.npz), reference scalar/SSE2 FIR arithmetic."""
    import torch

    g = np.load(os.path.join(H.GOLDEN_DIR, "cqpsk.npz"))
    for i in range(int(g["n_cases"])):
        sps, rate, bp, nb = [int(v) for v in g[f"cfg{i}"]]
        iq = g[f"iq{i}"].astype(np.float32)[None]
        bank = gpu.CqpskBank(1, rate, ted_sps=[sps], squelch_levels=[float(g[f"squelch{i}"])], fir_arith=1)
        sym, counts = bank.full_demod(torch.from_numpy(iq).cuda(), bp, nb)
        _compare(sym.cpu().numpy(), counts.cpu().numpy(), 0, g[f"sym{i}"], g[f"counts{i}"])


def test_cqpsk_host_entry_point_and_taps(gpu):
    rng = np.random.default_rng(3)
    n_ch, bp, nb = 3, 1000, 3
    iq = _signals(rng, n_ch, 5, bp * nb, [15.0], [0.01])
    bank = gpu.CqpskBank(n_ch, 24000)
    sym, counts = bank.full_demod_host(iq, bp, nb)
    L = H.oracle_cqpsk()
    for c in range(n_ch):
        orc = H.OracleCqpsk(fir_fma=1)
        want_sym, want_counts = orc.run(iq[c], bp, nb)
        _compare(sym, counts, c, want_sym, want_counts)
    want = [np.zeros(H.FLL_MAX_TAPS, np.float32) for _ in range(4)]
    n = L.oracle_fll_band_edge_design(5, *[H._ptr(w) for w in want], H.FLL_MAX_TAPS)
    for got, w in zip(bank.fll_taps(0), want):
        assert got.size == n and H.bits_equal(got, w[:n])


def test_cqpsk_many_channels_independent(gpu):
    """1024 channels (C3-sized bank): every channel equals the same channel run alone in a one-channel bank."""
    import torch

    rng = np.random.default_rng(9)
    n_ch, bp, nb = 1024, 2400, 2
    base = _signals(rng, 16, 5, bp * nb, [None, 14.0, 8.0], [0.0, 0.03, -0.02])
    iq = np.ascontiguousarray(base[np.arange(n_ch) % 16] * (0.5 + (np.arange(n_ch) % 7)[:, None, None] * 0.1)).astype(np.float32)
    bank = gpu.CqpskBank(n_ch, 24000)
    sym, counts = bank.full_demod(torch.from_numpy(iq).cuda(), bp, nb)
    sym, counts = sym.cpu().numpy(), counts.cpu().numpy()
    for c in [0, 31, 32, 500, 777, 1023]:
        orc = H.OracleCqpsk(fir_fma=1)
        want_sym, want_counts = orc.run(iq[c], bp, nb)
        _compare(sym, counts, c, want_sym, want_counts)


def test_cqpsk_slow_path_inputs(gpu):
    """Inputs that push the speculative fast paths out of their safe range -- silent (all-zero) stretches, amplitudes
    around 1e-20 and 1e15, a channel that is zero from the start -- still match the oracle bit for bit (the kernel
    re-runs those chunks / symbols with the plain operators)."""
    import torch

    rng = np.random.default_rng(21)
    n_ch, bp, nb = 8, 1200, 5
    iq = _signals(rng, n_ch, 5, bp * nb, [20.0, 10.0], [0.01, -0.02]).copy()
    iq[1] *= np.float32(1e-20)
    iq[2] *= np.float32(1e15)
    iq[3][:] = 0.0
    iq[4][1500:3200] = 0.0
    iq[5][::7] = 0.0
    iq[6] *= np.float32(3e-12)
    iq[7][4000:] *= np.float32(1e-30)
    bank = gpu.CqpskBank(n_ch, 24000, channel_lpf_enable=False)
    sym, counts = bank.full_demod(torch.from_numpy(iq).cuda(), bp, nb)
    sym, counts = sym.cpu().numpy(), counts.cpu().numpy()
    for c in range(n_ch):
        orc = H.OracleCqpsk(lpf_enable=0, fir_fma=1)
        want_sym, want_counts = orc.run(iq[c], bp, nb)
        _compare(sym, counts, c, want_sym, want_counts)
        _check_state(bank, c, orc)


def test_branch_free_div_sqrt_match_ieee(gpu):
    """The kernel's branch-free a / b and sqrt(a) equal the IEEE operators bit for bit wherever they declare their
    operands safe (2^24 random pairs across the whole exponent range plus the values the chain produces)."""
    import torch

    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(5)
    bits = torch.randint(0, 2 ** 31 - 1, (2, n), generator=g, device="cuda", dtype=torch.int32)
    a = bits[0].view(torch.float32).clone()
    b = bits[1].view(torch.float32).clone()
    a[: n // 2] = torch.rand(n // 2, generator=g, device="cuda") * 2.0          # the chain's own range
    b[: n // 2] = torch.rand(n // 2, generator=g, device="cuda") * 2.0 + 1e-3
    b[::2] = -b[::2]
    outs = [torch.empty(n, device="cuda") for _ in range(4)]
    flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    gpu.check(gpu.lib().dsdneo_b200_selftest_divsqrt(a.data_ptr(), b.data_ptr(), *[o.data_ptr() for o in outs],
                                                     flags.data_ptr(), n, None))
    torch.cuda.synchronize()
    qf, qi, sf, si = [o.view(torch.int32) for o in outs]
    ok_d, ok_s = (flags & 1) != 0, (flags & 2) != 0
    assert int(ok_d.sum()) > n // 3 and int(ok_s.sum()) > n // 3
    assert int(((qf != qi) & ok_d).sum()) == 0
    assert int(((sf != si) & ok_s).sum()) == 0


def test_cqpsk_vs_compiled_reference(gpu):
    """Directly against the unmodified reference full_demod() (oracle/_ref) when it travelled with the snapshot: symbols,
    counts and loop state, reference scalar/SSE2 FIR arithmetic."""
    import torch

    if not H.ref_available("par"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(31)
    for sps, rate, bp, nb in [(5, 24000, 1500, 4), (8, 48000, 2000, 4)]:
        n_ch = 4
        iq = _signals(rng, n_ch, sps, bp * nb, [None, 15.0, 6.0], [0.0, 0.03, -0.01])
        bank = gpu.CqpskBank(n_ch, rate, ted_sps=[sps] * n_ch, fir_arith=1)
        sym, counts = bank.full_demod(torch.from_numpy(iq).cuda(), bp, nb)
        sym, counts = sym.cpu().numpy(), counts.cpu().numpy()
        for c in range(n_ch):
            r = H.RefCqpsk("par", rate=rate, symrate=rate // sps, sps=sps)
            want_sym, want_counts = r.run(iq[c], bp, nb)
            _compare(sym, counts, c, want_sym, want_counts)
            want, st = r.state(), bank.state(c)
            for ok, gk in STATE_MAP:
                a, b = want[ok], getattr(st, gk)
                same = (a == b) if isinstance(a, int) else (np.float32(a).tobytes() == np.float32(b).tobytes())
                assert same, (c, ok, a, b)


def test_cqpsk_symbol_rate_slicer_bit_exact_vs_oracle(gpu):
    """The sample side behind the chain (output kind 2): tracker thresholds, CQPSK slicer with the OP25 orientation maps,
    soft metrics -- dibits, reliabilities, LLRs and final thresholds equal the oracle (pinned to the reference's
    getDibitSoft by tests/test_oracle_symbol.py) for every class, across two launches with ragged symbol counts."""
    import torch
    from test_oracle_symbol import cqpsk_symbol_stream

    rng = np.random.default_rng(900)
    n_ch, n = 70, 3000
    neg = np.array([c % 2 for c in range(n_ch)], np.uint8)
    p25 = np.array([0 if c % 7 == 3 else 1 for c in range(n_ch)], np.uint8)
    mp = np.array([c % 5 for c in range(n_ch)], np.uint8)
    x = np.stack([cqpsk_symbol_stream(rng, n, noise=0.2 + 0.01 * c, offset=0.02 * (c % 9)) for c in range(n_ch)])
    x[5, 700:1400] = 0.0  # a squelched stretch: the block side emits exact zeros
    for snr in (-100.0, 17.0):
        sl = gpu.CqpskSlicer(n_ch)
        sl.set_class(neg, p25, mp, snr)
        counts = [np.array([n // 2 - (c % 13) for c in range(n_ch)], np.int32), None]
        counts[1] = (n - counts[0]).astype(np.int32)
        got_d, got_r, got_l = [], [], []
        pos = np.zeros(n_ch, np.int64)
        for part in range(2):
            buf = np.zeros((n_ch, n), np.float32)
            for c in range(n_ch):
                buf[c, :counts[part][c]] = x[c, pos[c]:pos[c] + counts[part][c]]
            res = sl.run(torch.from_numpy(buf).cuda(), torch.from_numpy(counts[part]).cuda())
            got_d.append(res["dibits"].cpu().numpy()); got_r.append(res["reliability"].cpu().numpy()); got_l.append(res["llr"].cpu().numpy())
            pos += counts[part]
        for c in range(n_ch):
            d, r, l, st = H.oracle_cqpsk_slicer_run(x[c], negative=int(neg[c]), p25_slice=int(p25[c]), map_idx=int(mp[c]), snr_db=snr)
            a, b = counts[0][c], counts[1][c]
            gd = np.concatenate([got_d[0][c, :a], got_d[1][c, :b]])
            gr = np.concatenate([got_r[0][c, :a], got_r[1][c, :b]])
            gl = np.concatenate([got_l[0][c, :a], got_l[1][c, :b]])
            assert np.array_equal(gd, d), (c, int(np.argmax(gd != d)))
            assert np.array_equal(gr, r) and np.array_equal(gl, l), c
            bs = st.base
            want = np.array([bs.min, bs.max, bs.center, bs.umid, bs.lmid, bs.minref, bs.maxref, bs.lastsample], np.float32)
            assert H.bits_equal(sl.state(c), want), (c, sl.state(c), want)


def test_cqpsk_iq_to_dibits(gpu):
    """IQ -> CQPSK chain -> symbol-rate slicer on the device: the dibits equal the oracle chain's and, after acquisition, the
    transmitted ones."""
    import torch

    rng = np.random.default_rng(901)
    n_ch, bp, nb = 6, 2400, 5
    sigs = [H.synth_cqpsk_iq(rng, bp * nb // 5 + 2, sps=5, snr_db=22.0, cfo=0.01 * (c - 2), timing=0.13 * c) for c in range(n_ch)]
    iq = np.stack([s[0][:bp * nb] for s in sigs])
    bank = gpu.CqpskBank(n_ch, 24000)
    sym, counts = bank.full_demod(torch.from_numpy(iq).cuda(), bp, nb)
    total = counts.sum(dim=1, dtype=torch.int32).contiguous()
    sl = gpu.CqpskSlicer(n_ch)
    res = sl.run(sym, total)
    dib = res["dibits"].cpu().numpy()
    tot = total.cpu().numpy()
    for c in range(n_ch):
        orc = H.OracleCqpsk(fir_fma=1)
        want_sym, _ = orc.run(iq[c], bp, nb)
        d, r, l, _ = H.oracle_cqpsk_slicer_run(want_sym)
        assert tot[c] == want_sym.size and np.array_equal(dib[c, :tot[c]], d)
        tx = sigs[c][1]
        tail = dib[c, tot[c] - 600:tot[c]]
        best = max(int((tx[tx.size - 600 - lag: tx.size - lag] == tail).sum()) for lag in range(0, 14))
        assert best >= 594, (c, best)
