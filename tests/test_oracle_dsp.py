"""CPU tests: pin the block-side oracle (oracle/oracle_dsp.c) against the compiled reference
(oracle/_ref, built from /root/reference by oracle/Makefile) and the reference's own test contracts."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import _harness as H

needs_ref = pytest.mark.skipif(not H.ref_available("par"), reason="oracle/_ref not built (no /root/reference)")
needs_ref_avx2 = pytest.mark.skipif(not H.ref_available("avx2"), reason="oracle/_ref avx2 variant not built")


@needs_ref
@pytest.mark.parametrize("rate", [48000, 24000, 50000])
@pytest.mark.parametrize("profile", [0, 1, 2, 3, 4, 5])
def test_lpf_taps_match_reference(rate, profile):
    """channel_lpf_ensure_plan taps (demod_pipeline.cpp:498-523) == oracle taps, bit for bit."""
    r = H.RefDemod("par", rate=rate, profile=profile)
    r.block(np.zeros((64, 2), np.float32))  # first block designs the plan
    want = r.taps()
    got = H.oracle_lpf_taps(rate, profile)
    assert got.size == want.size
    assert H.bits_equal(got, want)
    if rate == 48000:
        assert got.size == 135  # demod_pipeline.cpp:130-131


def test_lpf_taps_unit_dc_gain_and_symmetry():
    t = H.oracle_lpf_taps(48000, 4)
    assert abs(float(t.astype(np.float64).sum()) - 1.0) < 1e-6
    # float cosf() makes the window only approximately symmetric; the FIR uses taps[0..centre] only
    assert np.allclose(t, t[::-1], rtol=0, atol=1e-7)


@needs_ref
def test_fir_scalar_matches_reference_scalar_and_sse2():
    """oracle FIR (fma=0) == simd_fir_complex_apply_scalar == SSE2 dispatch, incl. history continuity
    and short blocks (reference test: tests/dsp/test_dsp_simd_fir.cpp)."""
    L, R = H.oracle(), H.ref("par")
    assert R.simd_fir_get_impl_name() == b"sse2"
    rng = np.random.default_rng(1)
    taps = H.oracle_lpf_taps(48000, 4)
    for sizes in ([4096, 4096], [300, 20, 7, 134, 135, 1000], [8192]):
        hi = [np.zeros(144, np.float32) for _ in range(3)]
        hq = [np.zeros(144, np.float32) for _ in range(3)]
        for n in sizes:
            x = rng.standard_normal(2 * n).astype(np.float32)
            outs = [np.zeros(2 * n, np.float32) for _ in range(3)]
            L.oracle_fir_complex(H._ptr(x), 2 * n, H._ptr(outs[0]), H._ptr(hi[0]), H._ptr(hq[0]), H._ptr(taps), taps.size, 0)
            R.ref_fir_complex_scalar(H._ptr(x), 2 * n, H._ptr(outs[1]), H._ptr(hi[1]), H._ptr(hq[1]), H._ptr(taps), taps.size)
            R.simd_fir_complex_apply(H._ptr(x), 2 * n, H._ptr(outs[2]), H._ptr(hi[2]), H._ptr(hq[2]), H._ptr(taps), taps.size)
            assert H.bits_equal(outs[0], outs[1]), H.first_mismatch(outs[0], outs[1])
            assert H.bits_equal(outs[0], outs[2]), H.first_mismatch(outs[0], outs[2])
            assert H.bits_equal(hi[0], hi[1]) and H.bits_equal(hq[0], hq[1])


@needs_ref_avx2
def test_fir_fma_matches_reference_avx2():
    """oracle FIR (fma=1) == what simd_fir_complex_apply runs on an AVX2 host: the AVX2 kernel (vector body and its
    contracted scalar epilogue) for blocks of at least taps_len pairs, the scalar kernel for shorter ones
    (simd_fir_prefer_scalar_for_block, src/dsp/simd_fir.cpp:302-305)."""
    L, R = H.oracle(), H.ref("avx2")
    if R.simd_fir_get_impl_name() != b"avx2":
        pytest.skip("host CPU has no AVX2")
    rng = np.random.default_rng(2)
    taps = H.oracle_lpf_taps(48000, 2)
    hi = [np.zeros(144, np.float32) for _ in range(2)]
    hq = [np.zeros(144, np.float32) for _ in range(2)]
    for n in [4096, 512, 8192, 272, 100, 134, 135, 137, 141, 1001, 64]:
        x = rng.standard_normal(2 * n).astype(np.float32)
        a, b = np.zeros(2 * n, np.float32), np.zeros(2 * n, np.float32)
        L.oracle_fir_complex(H._ptr(x), 2 * n, H._ptr(a), H._ptr(hi[0]), H._ptr(hq[0]), H._ptr(taps), taps.size, 1)
        R.simd_fir_complex_apply(H._ptr(x), 2 * n, H._ptr(b), H._ptr(hi[1]), H._ptr(hq[1]), H._ptr(taps), taps.size)
        assert H.bits_equal(a, b), H.first_mismatch(a, b)


@needs_ref
def test_mean_power_matches_reference():
    L, R = H.oracle(), H.ref("par")
    rng = np.random.default_rng(3)
    for n in [2, 17, 512, 511]:
        x = (rng.standard_normal(n) * 3 + 0.7).astype(np.float32)
        a = L.oracle_mean_power(H._ptr(x), n, 1)
        b = R.mean_power(H._ptr(x), n, 1)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)


@needs_ref
@pytest.mark.parametrize("variant,fma", [("par", 0), ("avx2", 1)])
@pytest.mark.parametrize("snr", [None, 12.0])
def test_full_demod_matches_reference(variant, fma, snr):
    """oracle_full_demod_block == reference full_demod() (FSK discriminator kind), multi-block with carried state."""
    if not H.ref_available(variant):
        pytest.skip("variant not built")
    if variant == "avx2" and H.ref("avx2").simd_fir_get_impl_name() != b"avx2":
        pytest.skip("host CPU has no AVX2")
    rng = np.random.default_rng(11)
    bp, nb = 2048, 5
    iq = H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=snr)[: bp * nb]
    r = H.RefDemod(variant)
    want = r.run(iq, bp, nb)
    got, chan = H.oracle_full_demod(iq, bp, nb, fir_fma=fma, return_chan=True)
    assert H.bits_equal(got, want), H.first_mismatch(got, want)
    st = r.state()
    assert np.float32(chan.dc_est).view(np.uint32) == np.float32(st["dc_est"]).view(np.uint32)
    assert np.float32(chan.peak_est).view(np.uint32) == np.float32(st["peak"]).view(np.uint32)
    assert np.float32(chan.channel_pwr).view(np.uint32) == np.float32(st["pwr"]).view(np.uint32)
    # reference contract (tests/dsp/test_fsk_modem.c): one output per pair, first sample 0, within int16 range
    assert got[0] == 0.0 and np.all(got <= 32767.0) and np.all(got >= -32768.0)


@needs_ref
def test_full_demod_squelch_matches_reference():
    """Per-block squelch: blocks below channel_squelch_level are zeroed and reset the modem
    (demod_pipeline.cpp:1009-1017,1179-1184)."""
    rng = np.random.default_rng(12)
    bp, nb = 1024, 6
    iq = H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=20.0)[: bp * nb].copy()
    iq[1 * bp:2 * bp] *= 1e-3
    iq[4 * bp:5 * bp] *= 1e-3
    level = 0.01
    r = H.RefDemod("par", squelch=level)
    want = r.run(iq, bp, nb)
    got = H.oracle_full_demod(iq, bp, nb, fir_fma=0, squelch=level)
    assert H.bits_equal(got, want), H.first_mismatch(got, want)
    assert np.all(want[1 * bp:2 * bp] == 0) and np.any(want[2 * bp:3 * bp] != 0)


@needs_ref
def test_full_demod_lpf_off_matches_reference():
    rng = np.random.default_rng(13)
    bp, nb = 512, 3
    iq = H.synth_fsk_iq(rng, bp * nb // 10 + 1, 10, snr_db=15.0)[: bp * nb]
    r = H.RefDemod("par", lpf_enable=0)
    want = r.run(iq, bp, nb)
    got = H.oracle_full_demod(iq, bp, nb, fir_fma=0, lpf_enable=0)
    assert H.bits_equal(got, want)


def test_atan2f_matches_libm(tmp_path):
    """The device atan2f (dsd-neo_b200/csrc/fdlibm_atan2f.cuh), compiled for the host, is bit-identical to
    this image's libm atan2f on 2e7 random inputs (uniform, small-ratio and raw-bit-pattern draws)."""
    src = tmp_path / "a2.cpp"
    src.write_text(
        '#include <math.h>\n#include <stdint.h>\n#include "%s/dsd-neo_b200/csrc/fdlibm_atan2f.cuh"\n'
        'extern "C" long a2_check(long n, uint64_t seed, float* bad) {\n'
        '  uint64_t s = seed; long nbad = 0;\n'
        '  for (long i = 0; i < n; i++) {\n'
        '    s ^= s << 13; s ^= s >> 7; s ^= s << 17; float y, x;\n'
        '    if (i & 1) { y = dsdneo::bits_f32((uint32_t)s); x = dsdneo::bits_f32((uint32_t)(s >> 32)); }\n'
        '    else { y = ((int32_t)(uint32_t)s) * (1.0f / 2147483648.0f); x = ((int32_t)(uint32_t)(s >> 32)) * (1.0f / 2147483648.0f);\n'
        '           if (i & 2) y *= 1e-3f; }\n'
        '    float a = atan2f(y, x), b = dsdneo::fd_atan2f(y, x);\n'
        '    if (dsdneo::f32_bits(a) != dsdneo::f32_bits(b) && !(a != a && b != b)) { if (!nbad) { bad[0] = y; bad[1] = x; } nbad++; }\n'
        '  }\n  return nbad;\n}\n' % H.ROOT
    )
    so = tmp_path / "a2.so"
    subprocess.run(["g++", "-O2", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(so), "-lm"],
                   check=True)
    L = C.CDLL(str(so))
    L.a2_check.restype = C.c_long
    L.a2_check.argtypes = [C.c_long, C.c_uint64, H.f32p]
    bad = np.zeros(2, np.float32)
    nbad = L.a2_check(20_000_000, 88172645463325252, H._ptr(bad))
    assert nbad == 0, (nbad, bad)


def test_oracle_full_demod_matches_committed_golden_vectors():
    """Reference full_demod() outputs captured by tests/golden/make_golden.py (both FIR arithmetic variants)."""
    import os

    g = np.load(os.path.join(H.GOLDEN_DIR, "full_demod.npz"))
    bp, nb = int(g["block_pairs"]), int(g["n_blocks"])
    assert H.bits_equal(H.oracle_full_demod(g["iq"], bp, nb, fir_fma=0), g["ref_par"])
    if "ref_avx2" in g:
        assert H.bits_equal(H.oracle_full_demod(g["iq"], bp, nb, fir_fma=1), g["ref_avx2"])


@needs_ref
@pytest.mark.parametrize("variant,fma", [("par", 0), ("avx2", 1)])
def test_halfband_matches_reference(variant, fma):
    """oracle_hb_decim2_complex == the reference's simd_hb_decim2_complex (src/dsp/simd_fir.cpp:363-373), bit for bit:
    both tap sets, long/odd/tiny blocks (tiny ones take the scalar kernel even on AVX2 hosts), history carried over 4 blocks."""
    if not H.ref_available(variant):
        pytest.skip("variant not built")
    R = C.CDLL(H._ref_path(variant))
    R.simd_hb_decim2_complex.restype = C.c_int
    O = H.oracle()
    O.oracle_hb_decim2_complex.restype = C.c_int
    rng = np.random.default_rng(3)
    for name, tl in (("hb31_q15_taps", 31), ("hb_q15_taps", 15)):
        taps = (C.c_float * tl).in_dll(R, name)
        otaps = (C.c_float * tl).in_dll(O, "oracle_hb31_taps" if tl == 31 else "oracle_hb15_taps")
        assert list(taps) == list(otaps)
        for npairs in (4096, 1000, 37, 8, 62, 2, 30, 31, 32, 15, 14, 16):
            hr = np.zeros((2, 30), np.float32)
            ho = np.zeros((2, 30), np.float32)
            for _ in range(4):
                x = rng.standard_normal(2 * npairs).astype(np.float32)
                out_r, out_o = np.zeros(npairs + 2, np.float32), np.zeros(npairs + 2, np.float32)
                nr = R.simd_hb_decim2_complex(H._ptr(x), 2 * npairs, H._ptr(out_r), H._ptr(hr[0]), H._ptr(hr[1]), taps, tl)
                no = O.oracle_hb_decim2_complex(H._ptr(x), 2 * npairs, H._ptr(out_o), H._ptr(ho[0]), H._ptr(ho[1]), otaps, tl, fma)
                assert nr == no == 2 * (npairs // 2)
                assert H.bits_equal(out_r, out_o) and H.bits_equal(hr, ho)


def test_oracle_halfband_matches_committed_golden_vectors():
    """Reference half-band cascade outputs captured by tests/golden/make_golden.py (3 passes: 31, 15, 15 taps)."""
    import os

    g = np.load(os.path.join(H.GOLDEN_DIR, "halfband.npz"))
    bp, nb, passes = int(g["block_pairs"]), int(g["n_blocks"]), int(g["passes"])
    assert H.bits_equal(H.oracle_hb_cascade(g["x"], bp, nb, passes, fma=0), g["ref_par"])
    if "ref_avx2" in g:
        assert H.bits_equal(H.oracle_hb_cascade(g["x"], bp, nb, passes, fma=1), g["ref_avx2"])
