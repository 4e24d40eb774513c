"""Host-side pieces of the one-stream multi-GPU form (dsd-neo_b200/shard.py) that need no GPU: channel classes partition the
band, and the polyphase SYNTHESIS bank used as test / bench signal source is the transpose of the channelizer -- a signal
pushed through it and through the float64 direct-form channelizer oracle comes back delayed by taps_per_branch - 1 channel
samples."""
import importlib.util
import os

import numpy as np
import pytest

import _harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard():
    spec = importlib.util.spec_from_file_location("b200_shard_host", os.path.join(ROOT, "dsd-neo_b200", "shard.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_channel_classes_partition_the_band():
    sh = _shard()
    for M, world in ((1024, 1), (2048, 2), (8192, 8), (4096, 4)):
        seen = []
        for r in range(world):
            cls = list(sh.channel_class(r, world, M))
            assert len(cls) == M // world and all(k % world == r for k in cls)
            seen += cls
        assert sorted(seen) == list(range(M))
    with pytest.raises(ValueError):
        sh.channel_class(0, 3, 1024)
    with pytest.raises(ValueError):
        sh.channel_class(4, 4, 1024)


def test_synthesis_bank_is_the_transpose_of_the_channelizer():
    import torch

    sh = _shard()
    M, T, n = 32, 8, 192
    rng = np.random.default_rng(3)
    # band-limited channel signals (a quarter of the channel bandwidth), two base signals spread over the bins, some bins empty
    base = []
    for _ in range(2):
        z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        Z = np.fft.fft(z)
        Z[n // 8:-n // 8] = 0
        base.append(np.fft.ifft(Z))
    base = torch.from_numpy(np.stack(base).astype(np.complex64))
    idx = torch.tensor([(k % 2) if k % 5 else -1 for k in range(M)])
    L = M * T
    m = np.arange(L)
    x = 2 * (0.5 / M) * (m - (L - 1) / 2)
    h = (2 * (0.5 / M) * np.sinc(x) * (0.42 - 0.5 * np.cos(2 * np.pi * m / (L - 1)) + 0.08 * np.cos(4 * np.pi * m / (L - 1))))
    h = (h / h.sum()).astype(np.float32)
    wide = sh.synthesize_wideband(torch, base, idx, M, h, T, rows_per_block=50, rms=0.25).numpy()
    assert wide.shape == (n * M, 2) and wide.dtype == np.uint8 and wide.min() > 0 and wide.max() < 255
    w = ((wide.astype(np.float32) - np.float32(127.5)) * np.float32(1 / 127.5))
    hist = w[-(T - 1) * M:]  # the axis is circular
    sel = [1, 2, 6, 17, 31, 5, 10]
    y = H.oracle_pfb(np.concatenate([hist, w]), (T - 1) * M, h, M, sel, n)
    for row, k in enumerate(sel):
        if idx[k] < 0:
            assert np.abs(y[row]).max() < 0.05 * np.abs(y).max()  # an empty bin stays empty (quantisation noise only)
            continue
        ref = np.roll(base[idx[k]].numpy(), T - 1)
        g = np.vdot(ref, y[row]) / np.vdot(ref, ref)
        err = y[row] - g * ref
        snr = 10 * np.log10(np.vdot(g * ref, g * ref).real / np.vdot(err, err).real)
        assert snr > 25.0, (k, snr)
