"""GPU parity tests (run on the B200 box with -m gpu): the CUDA block side, called through the C-ABI,
against the oracle (and, where oracle/_ref travelled with the snapshot, the compiled reference)."""
import ctypes as C

import numpy as np
import pytest

import _harness as H

pytestmark = pytest.mark.gpu


def _run_gpu(b200, iq, bp, nb, **kw):
    import torch

    bank = b200.DemodBank(iq.shape[0], kw.pop("rate", 48000), kw.pop("lpf", True), **kw)
    d_iq = torch.from_numpy(np.ascontiguousarray(iq)).cuda()
    out = bank.full_demod(d_iq, bp, nb).cpu().numpy()
    return out, bank


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("bp,nb", [(8192, 2), (2048, 3), (1000, 4), (4100, 2), (136, 3), (52, 5)])
def test_full_demod_bit_exact_vs_oracle(gpu, arith, bp, nb):
    rng = np.random.default_rng(100 + bp)
    n_ch = 5
    iq = np.stack([H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=[None, 20.0, 12.0, 6.0, 0.0][c])[: bp * nb]
                   for c in range(n_ch)])
    got, bank = _run_gpu(gpu, iq, bp, nb, fir_arith=arith)
    for c in range(n_ch):
        want, chan = H.oracle_full_demod(iq[c], bp, nb, fir_fma=1 - arith, return_chan=True)
        assert H.bits_equal(got[c], want), (c, H.first_mismatch(got[c], want))
        st = bank.state(c)
        for a, b in [(st.dc_est, chan.dc_est), (st.discriminator_peak_est, chan.peak_est), (st.prev_i, chan.prev_i),
                     (st.prev_q, chan.prev_q), (st.channel_pwr, chan.channel_pwr)]:
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)
        assert st.have_prev == chan.have_prev


def test_state_carries_across_launches(gpu):
    import torch

    rng = np.random.default_rng(5)
    bp, nb, n_ch = 1024, 6, 3
    iq = np.stack([H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=15.0)[: bp * nb] for _ in range(n_ch)])
    bank = gpu.DemodBank(n_ch)
    parts = []
    for lo, hi in [(0, 1), (1, 3), (3, 6)]:  # three launches with 1, 2, 3 blocks
        d = torch.from_numpy(np.ascontiguousarray(iq[:, lo * bp:hi * bp])).cuda()
        parts.append(bank.full_demod(d, bp, hi - lo).cpu().numpy())
    got = np.concatenate(parts, axis=1)
    for c in range(n_ch):
        want = H.oracle_full_demod(iq[c], bp, nb, fir_fma=1)
        assert H.bits_equal(got[c], want), (c, H.first_mismatch(got[c], want))


def test_squelch_and_profiles(gpu):
    rng = np.random.default_rng(6)
    bp, nb, n_ch = 1024, 6, 6
    iq = np.stack([H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=20.0)[: bp * nb] for _ in range(n_ch)]).copy()
    iq[:, 1 * bp:2 * bp] *= 1e-3
    iq[:, 4 * bp:5 * bp] *= 1e-3
    profiles = [0, 1, 2, 3, 4, 5]
    levels = [0.01, 0.0, 0.01, 0.01, 0.0, 0.01]
    got, bank = _run_gpu(gpu, iq, bp, nb, profiles=profiles, squelch_levels=levels, fir_arith=1)
    for c in range(n_ch):
        want, chan = H.oracle_full_demod(iq[c], bp, nb, fir_fma=0, profile=profiles[c], squelch=levels[c], return_chan=True)
        assert H.bits_equal(got[c], want), (c, H.first_mismatch(got[c], want))
        assert bank.state(c).channel_squelched == chan.channel_squelched
    assert np.all(got[0, bp:2 * bp] == 0)


def test_lpf_disabled_and_24k(gpu):
    rng = np.random.default_rng(8)
    bp, nb = 1024, 2
    iq = np.stack([H.synth_fsk_iq(rng, bp * nb // 5 + 1, 5, dev_per_level=0.157, snr_db=15.0)[: bp * nb] for _ in range(2)])
    got, _ = _run_gpu(gpu, iq, bp, nb, lpf=False)
    for c in range(2):
        assert H.bits_equal(got[c], H.oracle_full_demod(iq[c], bp, nb, lpf_enable=0))
    got, _ = _run_gpu(gpu, iq, bp, nb, rate=24000)  # 67-tap plan, unrolled C=33 kernel
    for c in range(2):
        assert H.bits_equal(got[c], H.oracle_full_demod(iq[c], bp, nb, rate=24000, fir_fma=1))
    got, _ = _run_gpu(gpu, iq, bp, nb, rate=50000)  # 141-tap plan, generic kernel
    for c in range(2):
        assert H.bits_equal(got[c], H.oracle_full_demod(iq[c], bp, nb, rate=50000, fir_fma=1))


def test_full_demod_vs_compiled_reference(gpu):
    """Directly against the unmodified reference (oracle/_ref) when it travelled with the snapshot."""
    if not H.ref_available("avx2") or H.ref("avx2").simd_fir_get_impl_name() != b"avx2":
        pytest.skip("oracle/_ref avx2 build or AVX2 host not available")
    rng = np.random.default_rng(9)
    bp, nb = 8192, 3
    iq = np.stack([H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=s)[: bp * nb] for s in (None, 10.0)])
    got, _ = _run_gpu(gpu, iq, bp, nb, fir_arith=0)
    for c in range(2):
        want = H.RefDemod("avx2").run(iq[c], bp, nb)
        assert H.bits_equal(got[c], want), H.first_mismatch(got[c], want)
    got, _ = _run_gpu(gpu, iq, bp, nb, fir_arith=1)
    for c in range(2):
        want = H.RefDemod("par").run(iq[c], bp, nb)
        assert H.bits_equal(got[c], want), H.first_mismatch(got[c], want)


def test_cu8_entry_equals_widen_then_full_demod(gpu):
    """dsdneo_b200_full_demod_batch_cu8 == widen_u8_to_f32_bias127 + dsdneo_b200_full_demod_batch, bit for bit, state included."""
    import torch

    rng = np.random.default_rng(61)
    n_ch, bp, nb = 7, 1024, 3
    u8 = np.stack([H.synth_c4fm_iq(rng, rng.integers(0, 4, bp * nb // 10 + 4), snr_db=18.0)[:bp * nb] for _ in range(n_ch)])
    a, b = gpu.DemodBank(n_ch, 48000, True), gpu.DemodBank(n_ch, 48000, True)
    for k in range(2):
        want = a.full_demod(torch.from_numpy(np.stack([H.widen_cu8(x) for x in u8])).cuda(), bp, nb).cpu().numpy()
        got = b.full_demod_cu8(torch.from_numpy(u8).cuda(), bp, nb).cpu().numpy()
        assert H.bits_equal(got, want), k


def test_host_buffer_entry_point(gpu):
    rng = np.random.default_rng(10)
    bp, nb = 2048, 2
    iq = np.stack([H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=18.0)[: bp * nb] for _ in range(3)])
    bank = gpu.DemodBank(3)
    got = bank.full_demod_host(iq, bp, nb)
    for c in range(3):
        assert H.bits_equal(got[c], H.oracle_full_demod(iq[c], bp, nb, fir_fma=1))


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("passes,bp,nb", [(1, 512, 3), (3, 512, 3), (8, 8192, 2), (2, 16, 5), (1, 2, 9)])
def test_halfband_cascade_bit_exact_vs_oracle(gpu, arith, passes, bp, nb):
    """Batched half-band cascade == oracle (== reference simd_hb_decim2_complex chain) per channel, incl. blocks shorter
    than the filter (scalar-kernel rule), history carried across two launches, host-buffer entry point."""
    import torch

    rng = np.random.default_rng(50 + passes)
    n_ch = 5
    x = rng.standard_normal((n_ch, 2 * bp * nb, 2)).astype(np.float32)
    hb = gpu.HalfbandCascade(n_ch, passes, fir_arith=arith)
    d = torch.from_numpy(x).cuda()
    got = torch.cat([hb.decimate(d[:, : bp * nb].contiguous(), bp, nb), hb.decimate(d[:, bp * nb:].contiguous(), bp, nb)], dim=1)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    for c in range(n_ch):
        want = H.oracle_hb_cascade(x[c], bp, 2 * nb, passes, fma=1 - arith)
        assert H.bits_equal(got[c], want), (c, passes, bp)
    hb2 = gpu.HalfbandCascade(n_ch, passes, fir_arith=arith)
    got_h = hb2.decimate_host(x[:, : bp * nb], bp, nb)
    assert H.bits_equal(got_h, got[:, : got_h.shape[1]])
    with pytest.raises(gpu.B200Error):
        hb2.decimate(d[:, :3].contiguous(), 3, 1)  # not a multiple of 2^passes (or odd)


def test_device_atan2f(gpu):
    """Device fd_atan2f == host libm atan2f, bit for bit, over 4M random pairs incl. raw bit patterns."""
    import torch

    rng = np.random.default_rng(77)
    n = 1 << 22
    y = rng.uniform(-1, 1, n).astype(np.float32)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y[: n // 4] = rng.integers(0, 2 ** 32, n // 4, dtype=np.uint64).astype(np.uint32).view(np.float32)
    x[: n // 4] = rng.integers(0, 2 ** 32, n // 4, dtype=np.uint64).astype(np.uint32).view(np.float32)
    y[n // 4: n // 2] *= 1e-3
    dy, dx = torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda()
    do = torch.empty_like(dy)
    gpu.check(gpu.lib().dsdneo_b200_selftest_atan2f(dy.data_ptr(), dx.data_ptr(), do.data_ptr(), n, None))
    got = do.cpu().numpy()
    want = np.empty_like(y)  # libm atan2f (numpy's own arctan2 may use SVML, which differs in the last bit)
    H.oracle().oracle_libm_atan2f_array(H._ptr(y), H._ptr(x), H._ptr(want), n)
    ok = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert ok.all(), (y[~ok][:3], x[~ok][:3], got[~ok][:3], want[~ok][:3])


def test_device_scale_division(gpu):
    """The branch-free 30000/peak of the recurrence kernel == IEEE division (device and host), bit for bit, on every
    kind of peak the tracker can produce: log-uniform over (1e-7, 2^20), dense around typical values, exact powers of two."""
    import torch

    rng = np.random.default_rng(78)
    n = 1 << 23
    pk = np.exp(rng.uniform(np.log(1.0e-7), np.log(1048576.0), n)).astype(np.float32)
    pk[: n // 4] = rng.uniform(0.01, 4.0, n // 4).astype(np.float32)
    pk[n // 4: n // 4 + 64] = (2.0 ** np.arange(-23, 41, dtype=np.float64)).astype(np.float32)[:64]
    lo = np.float32(0.3).view(np.uint32)
    pk[n // 2: n // 2 + (1 << 20)] = (lo + np.arange(1 << 20, dtype=np.uint32)).view(np.float32)  # 2^20 consecutive floats
    pk = np.clip(pk, np.nextafter(np.float32(1.0e-7), np.float32(1)), np.nextafter(np.float32(1048576.0), np.float32(0)))
    dp = torch.from_numpy(pk).cuda()
    df, di = torch.empty_like(dp), torch.empty_like(dp)
    gpu.check(gpu.lib().dsdneo_b200_selftest_scale(dp.data_ptr(), df.data_ptr(), di.data_ptr(), n, None))
    fast, ieee = df.cpu().numpy(), di.cpu().numpy()
    host = (np.float32(30000.0) / pk).astype(np.float32)
    assert np.array_equal(ieee.view(np.uint32), host.view(np.uint32))
    bad = fast.view(np.uint32) != host.view(np.uint32)
    assert not bad.any(), (pk[bad][:4], fast[bad][:4], host[bad][:4])


def test_no_cpu_fallback_symbols(b200):
    """The product library exports no oracle symbols and links no oracle code."""
    import subprocess

    out = subprocess.run(["nm", "-D", b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "ref_demod" not in out
