"""CPU tests: pin the CQPSK block-side oracle (oracle/oracle_cqpsk.c) against the unmodified reference full_demod()
compiled into oracle/_ref (AGC -> FLL band-edge -> Gardner/MMSE -> diff phasor -> Costas ->
4/pi atan), and against committed golden vectors generated from that reference (tests/golden/make_cqpsk_golden.py)."""
import os

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_available("par"), reason="oracle/_ref not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("sps,rate", [(5, 24000), (4, 24000), (10, 48000), (8, 48000)])
def test_fll_band_edge_taps_match_reference(sps, rate):
    """fll_band_edge_design_filter (costas.cpp:936-1024) == oracle design, bit for bit."""
    r = H.RefCqpsk(rate=rate, symrate=rate // sps, sps=sps)
    r.block(np.full((64, 2), 0.1, np.float32))  # first block designs the filters
    want = r.fll_taps()
    L = H.oracle_cqpsk()
    got = [np.zeros(H.FLL_MAX_TAPS, np.float32) for _ in range(4)]
    n = L.oracle_fll_band_edge_design(sps, *[H._ptr(g) for g in got], H.FLL_MAX_TAPS)
    assert n == 2 * sps + 1 == want[0].size
    for g, w in zip(got, want):
        assert H.bits_equal(g[:n], w)


CASES = [
    # sps, rate, symrate, n_symbols, block sizes, snr, cfo (rad/sample), squelch, ted_gain, is_set
    dict(sps=5, rate=24000, n_sym=3000, blocks=[2400] * 6, snr=None, cfo=0.0),
    dict(sps=5, rate=24000, n_sym=3000, blocks=[2400] * 6, snr=18.0, cfo=0.02),
    dict(sps=5, rate=24000, n_sym=2000, blocks=[997, 4, 1500, 13, 333, 2048, 7, 1200], snr=12.0, cfo=-0.03),
    dict(sps=4, rate=24000, n_sym=4000, blocks=[1600] * 10, snr=20.0, cfo=0.01),           # P25p2-like: 6000 sym/s, gain switch
    dict(sps=10, rate=48000, n_sym=1500, blocks=[4800] * 3, snr=15.0, cfo=0.005),
    dict(sps=8, rate=48000, n_sym=2500, blocks=[3000] * 6, snr=25.0, cfo=-0.004),          # 6000 sym/s at 48 k
    dict(sps=5, rate=24000, n_sym=2000, blocks=[2000] * 5, snr=10.0, cfo=0.0, ted_gain=0.05, is_set=1),
    dict(sps=5, rate=24000, n_sym=2000, blocks=[2000] * 5, snr=3.0, cfo=0.1),              # low confidence branches
]


def _signal(case, seed):
    rng = np.random.default_rng(seed)
    x, _ = H.synth_cqpsk_iq(rng, case["n_sym"], sps=case["sps"], snr_db=case["snr"], cfo=case["cfo"], timing=0.37)
    return x


@needs_ref
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_cqpsk_oracle_matches_reference(idx):
    case = CASES[idx]
    x = _signal(case, 100 + idx)
    symrate = case["rate"] // case["sps"]
    r = H.RefCqpsk(rate=case["rate"], symrate=symrate, sps=case["sps"], ted_gain=case.get("ted_gain", 0.0),
                   ted_gain_is_set=case.get("is_set", 0))
    o = H.OracleCqpsk(rate=case["rate"], sps=case["sps"], ted_gain=case.get("ted_gain", 0.0),
                      ted_gain_is_set=case.get("is_set", 0))
    pos = 0
    total = 0
    for n in case["blocks"]:
        blk = x[pos:pos + n]
        pos += n
        want, got = r.block(blk), o.block(blk)
        assert want.size == got.size, (idx, n, want.size, got.size)
        assert H.bits_equal(got, want), (idx, n, H.first_mismatch(got, want))
        bad = H.cqpsk_state_equal(o.state(), r.state())
        assert not bad, (idx, n, bad)
        total += want.size
    assert total > 0.9 * pos / case["sps"]


@needs_ref
def test_cqpsk_demodulates_the_transmitted_dibits():
    """Sanity of the synthetic signal and of the chain: after acquisition the symbols sit near {-3,-1,+1,+3} and slice to
    the transmitted dibits (4-level map of dsd_dibit.c:963-976)."""
    rng = np.random.default_rng(7)
    x, dib = H.synth_cqpsk_iq(rng, 4000, sps=5, snr_db=25.0, cfo=0.01, timing=0.2)
    o = H.OracleCqpsk()
    sym, _ = o.run(x, 4000, 5)
    tail = sym[-1500:]
    lv = np.where(tail > 2, 3, np.where(tail > 0, 1, np.where(tail > -2, -1, -3)))
    want_lv = H.LEVELS[dib].astype(int)
    best = 0
    for lag in range(0, 12):
        seg = want_lv[len(want_lv) - 1500 - lag: len(want_lv) - lag]
        best = max(best, int((seg == lv).sum()))
    assert best > 1490, best


@needs_ref
def test_cqpsk_squelch_transitions_match_reference():
    """Squelched blocks emit ceil(pairs / sps) zeros and leave every loop untouched (demod_pipeline.cpp:1022-1040)."""
    rng = np.random.default_rng(11)
    x, _ = H.synth_cqpsk_iq(rng, 2400, sps=5, snr_db=20.0, cfo=0.01)
    x[2000:5000] *= 1e-4
    x[9001:10000] *= 1e-4
    r = H.RefCqpsk(squelch=1e-3)
    o = H.OracleCqpsk(squelch=1e-3)
    seen = set()
    for b in range(12):
        blk = x[b * 1000:(b + 1) * 1000 - (b % 3)]
        want, got = r.block(blk), o.block(blk)
        assert want.size == got.size and H.bits_equal(got, want)
        bad = H.cqpsk_state_equal(o.state(), r.state())
        assert not bad, (b, bad)
        seen.add(r.state()["channel_squelched"])
    assert seen == {0, 1}


def test_cqpsk_golden_vectors():
    """The oracle reproduces the committed reference outputs (tests/golden/cqpsk.npz, made by make_cqpsk_golden.py from
    the compiled reference); runs on the GPU box too, where /root/reference does not exist."""
    g = np.load(os.path.join(H.GOLDEN_DIR, "cqpsk.npz"))
    for i in range(int(g["n_cases"])):
        sps, rate, bp, nb = [int(v) for v in g[f"cfg{i}"]]
        o = H.OracleCqpsk(rate=rate, sps=sps, squelch=float(g[f"squelch{i}"]))
        sym, counts = o.run(g[f"iq{i}"], bp, nb)
        assert np.array_equal(counts, g[f"counts{i}"])
        assert H.bits_equal(sym, g[f"sym{i}"])
