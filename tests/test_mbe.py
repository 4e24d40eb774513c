"""MBE synthesis stage -- PARITY UNPINNED: the reference's vocoder (mbelib-neo) is not in its
tree, so these tests pin the CPU restatement (oracle/oracle_mbe.c, published mbelib 1.3.0 algorithm) to properties of that
algorithm and the CUDA kernel to the restatement (+-1 LSB of int16, the north star's tolerance)."""
import ctypes as C

import numpy as np
import pytest

import _harness as H


class OracleMbeParms(C.Structure):
    _fields_ = [("w0", C.c_float), ("L", C.c_int), ("K", C.c_int), ("Vl", C.c_int * 57), ("Ml", C.c_float * 57),
                ("log2Ml", C.c_float * 57), ("PHIl", C.c_float * 57), ("PSIl", C.c_float * 57), ("gamma", C.c_float),
                ("un", C.c_int), ("repeat", C.c_int)]


def O():
    L = H.oracle()
    L.oracle_mbe_uniform.restype = C.c_float
    L.oracle_mbe_uniform.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_mbe_synthesize_speechf.argtypes = [H.f32p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64]
    L.oracle_mbe_synth_frame.argtypes = [H.f32p, C.POINTER(C.c_int16), C.c_void_p, C.c_void_p, C.c_int, C.c_uint64]
    L.oracle_mbe_spectral_amp_enhance.argtypes = [C.c_void_p]
    return L


def random_parms(rng, cls=OracleMbeParms, voiced_prob=0.6):
    p = cls()
    pitch = rng.uniform(20.0, 120.0)
    p.w0 = np.float32(2 * np.pi / pitch)
    p.L = int(min(56, max(9, int(0.9254 * int(np.pi / p.w0 + 0.25)))))
    p.K = (p.L + 2) // 3 if p.L < 37 else 12
    for l in range(1, p.L + 1):
        p.Vl[l] = int(rng.random() < voiced_prob)
        p.Ml[l] = np.float32(rng.uniform(0.0, 900.0) * np.exp(-l / 25.0))
        p.PSIl[l] = np.float32(rng.uniform(-3.0, 3.0))
        p.PHIl[l] = p.PSIl[l]
    return p


def test_layout_matches_mbelib_1_3_0():
    assert C.sizeof(OracleMbeParms) == 4 * (3 + 5 * 57 + 3)


def test_window_overlap_add_reconstructs_a_steady_voiced_harmonic():
    """Same parameters in both frames, everything voiced, one band: ws(n) + ws(n-160) = 1, phases advance by w0*l*160, so the
    output is exactly M cos(w0 l n + PHI_prev) (eq. 133 with eq. 139)."""
    L = O()
    cur, prev = OracleMbeParms(), OracleMbeParms()
    for p in (cur, prev):
        p.w0, p.L = np.float32(0.1), 12
        for l in range(1, 13):
            p.Vl[l] = 1
        p.Ml[3] = 100.0
    prev.PSIl[3] = prev.PHIl[3] = 0.7
    out = np.zeros(160, np.float32)
    L.oracle_mbe_synthesize_speechf(H._ptr(out), C.byref(cur), C.byref(prev), 3, 5)
    want = 100.0 * np.cos(np.float32(0.1) * 3 * np.arange(160) + 0.7)
    assert np.max(np.abs(out - want)) < 2e-2
    assert abs(cur.PSIl[3] - (0.7 + 0.1 * 3 * 160)) < 1e-3 and cur.PHIl[3] == cur.PSIl[3]  # l <= L/4: no random phase


def test_enhancement_preserves_energy_and_bounds_the_weights():
    L = O()
    rng = np.random.default_rng(4)
    for _ in range(50):
        p = random_parms(rng)
        before = np.array(p.Ml[1:p.L + 1], np.float64)
        L.oracle_mbe_spectral_amp_enhance(C.byref(p))
        after = np.array(p.Ml[1:p.L + 1], np.float64)
        assert abs((after ** 2).sum() / (before ** 2).sum() - 1.0) < 1e-4          # gamma rescales to the input energy
        low = np.arange(1, p.L + 1) * 8 <= p.L
        ratio = after[low] / before[low]
        assert np.allclose(ratio, ratio[0], rtol=1e-5)                              # low bands only see gamma


def test_unvoiced_excitation_is_deterministic_and_keyed():
    L = O()
    u = np.array([L.oracle_mbe_uniform(9, l, n, i, 3) for l in range(1, 20) for n in range(0, 160, 7) for i in range(3)])
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03
    rng = np.random.default_rng(6)
    cur, prev = random_parms(rng, voiced_prob=0.0), random_parms(rng, voiced_prob=0.0)
    outs = []
    for key in (1, 1, 2):
        c, p = OracleMbeParms.from_buffer_copy(cur), OracleMbeParms.from_buffer_copy(prev)
        out = np.zeros(160, np.float32)
        L.oracle_mbe_synthesize_speechf(H._ptr(out), C.byref(c), C.byref(p), 3, key)
        outs.append(out)
    assert np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])


def test_float_to_short_gain_and_clip():
    L = O()
    x = np.zeros(160, np.float32)
    x[:4] = [1.0, -1.5, 5000.0, -5000.0]
    s = np.zeros(160, np.int16)
    L.oracle_mbe_floattoshort(H._ptr(x), s.ctypes.data_as(C.POINTER(C.c_int16)))
    assert list(s[:4]) == [7, -10, 32760, -32760]


@pytest.mark.gpu
def test_mbe_synth_kernel_matches_oracle_within_one_lsb(gpu):
    """Three consecutive frames per voice (state carried through prev_enhanced), mixed voiced / unvoiced bands, both arms
    fed the same parameters and keys: int16 PCM within +-1 LSB, float PCM within 1e-3 of full scale, phases equal."""
    L = O()
    rng = np.random.default_rng(8)
    n = 256
    prev_o = [random_parms(rng) for _ in range(n)]
    prev_g = (gpu.MbeParms * n)(*[gpu.MbeParms.from_buffer_copy(p) for p in prev_o])
    worst = 0
    for frame in range(3):
        cur_o = [random_parms(rng, voiced_prob=[0.7, 0.2, 1.0][frame]) for _ in range(n)]
        cur_g = (gpu.MbeParms * n)(*[gpu.MbeParms.from_buffer_copy(p) for p in cur_o])
        keys = np.arange(n, dtype=np.uint64) + 1000 * frame
        pf, ps = gpu.mbe_synth(cur_g, prev_g, keys, uvquality=3)
        for i in range(n):
            of, os_ = np.zeros(160, np.float32), np.zeros(160, np.int16)
            L.oracle_mbe_synth_frame(H._ptr(of), os_.ctypes.data_as(C.POINTER(C.c_int16)), C.byref(cur_o[i]), C.byref(prev_o[i]), 3,
                                     int(keys[i]))
            scale = max(1.0, float(np.abs(of).max()))
            assert np.max(np.abs(pf[i] - of)) <= 1e-3 * scale, (frame, i)
            d = int(np.max(np.abs(ps[i].astype(np.int32) - os_.astype(np.int32))))
            worst = max(worst, d)
            assert d <= 1, (frame, i, d)
            assert np.allclose(np.array(prev_g[i].PSIl[1:57]), np.array(prev_o[i].PSIl[1:57]), rtol=0, atol=1e-2)
            assert list(prev_g[i].Vl[1:57]) == list(prev_o[i].Vl[1:57])
    assert worst <= 1
