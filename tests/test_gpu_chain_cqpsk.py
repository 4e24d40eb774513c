"""GPU chain test (-m gpu) for the CQPSK output kind: wideband IQ -> polyphase channelizer -> channel LPF (P25_CQPSK
profile) -> AGC / FLL / Gardner / Costas -> symbols.  Every occupied channel's symbol stream is bit-identical to the CPU
oracle run on the GPU channelizer's output, and slices back to the transmitted dibits."""
import numpy as np
import pytest

import _harness as H

pytestmark = pytest.mark.gpu

M = 256


def test_wideband_to_cqpsk_symbols(gpu):
    import torch

    rng = np.random.default_rng(44)
    n_out, bp = 9600, 2400  # 0.2 s at 48 kS/s per channel, four reference blocks
    nb = n_out // bp
    active = [5, 64, 131, 250]
    x, truth = H.synth_wideband_cqpsk(rng, M, n_out, active, sps=10, snr_db=28.0, cfo_hz_frac=0.004)
    cz = gpu.Channelizer(M, 8)
    chan = cz.channelize(torch.from_numpy(x).cuda())
    bank = gpu.CqpskBank(M, 48000, ted_sps=[10] * M)
    sym, counts = bank.full_demod(chan, bp, nb)
    sym, counts, chan_h = sym.cpu().numpy(), counts.cpu().numpy(), chan.cpu().numpy()
    for k in active:
        orc = H.OracleCqpsk(rate=48000, sps=10, fir_fma=1)
        want_sym, want_counts = orc.run(chan_h[k], bp, nb)
        assert np.array_equal(counts[k], want_counts)
        n = int(want_counts.sum())
        assert H.bits_equal(sym[k, :n], want_sym), H.first_mismatch(sym[k, :n], want_sym)
        # after acquisition the symbols sit near {-3,-1,+1,+3}: slice with the reference's fixed CQPSK thresholds (+-2, 0)
        tail = sym[k, n - 400:n]
        lv = np.where(tail > 2, 3, np.where(tail > 0, 1, np.where(tail > -2, -1, -3)))
        want_lv = H.LEVELS[truth[k]].astype(int)
        best = 0
        for lag in range(0, 40):
            seg = want_lv[want_lv.size - 400 - lag: want_lv.size - lag]
            if seg.size == 400:
                best = max(best, int((seg == lv).sum()))
        assert best >= 392, (k, best)
    # an empty channel stays un-squelched noise: finite output, counts near n / sps
    assert np.isfinite(sym[3]).all() and abs(int(counts[3].sum()) - n_out // 10) <= 2
