"""Test infrastructure: loaders for the oracle (oracle/liboracle.so), the compiled reference
(oracle/_ref/*.so, when present) and seeded synthetic-signal generators.

Nothing in here is product code; nothing in dsd-neo_b200/ imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
LPF_MAX_TAPS = 144

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
u8p = C.POINTER(C.c_uint8)


def _ptr(a, typ=f32p):
    return a.ctypes.data_as(typ)


# --------------------------------------------------------------------------- oracle (restatement)

class OracleDemodChan(C.Structure):
    _fields_ = [
        ("rate_out_hz", C.c_int),
        ("lpf_enable", C.c_int),
        ("lpf_profile", C.c_int),
        ("fir_fma", C.c_int),
        ("squelch_level", C.c_float),
        ("taps_len", C.c_int),
        ("taps", C.c_float * LPF_MAX_TAPS),
        ("hist_i", C.c_float * LPF_MAX_TAPS),
        ("hist_q", C.c_float * LPF_MAX_TAPS),
        ("prev_i", C.c_float),
        ("prev_q", C.c_float),
        ("have_prev", C.c_int),
        ("dc_est", C.c_float),
        ("peak_est", C.c_float),
        ("channel_pwr", C.c_float),
        ("channel_squelched", C.c_int),
    ]


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.run(["make", "oracle"], cwd=ORACLE_DIR, check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        L.oracle_channel_lpf_design.argtypes = [C.c_int, C.c_int, f32p, C.c_int]
        L.oracle_fir_complex.argtypes = [f32p, C.c_int, f32p, f32p, f32p, f32p, C.c_int, C.c_int]
        L.oracle_mean_power.restype = C.c_float
        L.oracle_mean_power.argtypes = [f32p, C.c_int, C.c_int]
        L.oracle_demod_chan_init.argtypes = [C.POINTER(OracleDemodChan), C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.oracle_full_demod_block.argtypes = [C.POINTER(OracleDemodChan), f32p, C.c_int, f32p, f32p]
        L.oracle_pfb_direct.argtypes = [f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int, i32p, C.c_int,
                                        C.POINTER(C.c_double), C.c_int]
        L.oracle_libm_atan2f_array.argtypes = [f32p, f32p, f32p, C.c_long]
        _oracle = L
    return _oracle


def oracle_lpf_taps(rate, profile):
    buf = np.zeros(LPF_MAX_TAPS, dtype=np.float32)
    n = oracle().oracle_channel_lpf_design(rate, profile, _ptr(buf), LPF_MAX_TAPS)
    assert n > 0
    return buf[:n].copy()


def oracle_full_demod(iq_ch, block_pairs, n_blocks, fir_fma=1, rate=48000, profile=4, lpf_enable=1, squelch=0.0,
                      chan=None, return_chan=False):
    """iq_ch: [n, 2] float32 for one channel. Runs n_blocks full_demod() blocks through the oracle."""
    L = oracle()
    if chan is None:
        chan = OracleDemodChan()
        assert L.oracle_demod_chan_init(C.byref(chan), rate, profile, lpf_enable, squelch, fir_fma) == 0
    iq_ch = np.ascontiguousarray(iq_ch, dtype=np.float32)
    out = np.empty(block_pairs * n_blocks, dtype=np.float32)
    scratch = np.empty(2 * block_pairs, dtype=np.float32)
    for b in range(n_blocks):
        blk = np.ascontiguousarray(iq_ch[b * block_pairs:(b + 1) * block_pairs]).reshape(-1)
        o = out[b * block_pairs:(b + 1) * block_pairs]
        n = L.oracle_full_demod_block(C.byref(chan), _ptr(blk), 2 * block_pairs, _ptr(scratch), _ptr(o))
        assert n == block_pairs
    return (out, chan) if return_chan else out


def oracle_hb_cascade(x, block_pairs, n_blocks, passes, fma=1, hist=None):
    """x: [n, 2] float32 of one channel; runs the oracle's half-band cascade block by block.  Returns [n >> passes, 2]."""
    L = oracle()
    L.oracle_hb_cascade.restype = C.c_int
    if hist is None:
        hist = np.zeros((passes, 2, 30), np.float32)
    outs = []
    work = np.zeros(2 * block_pairs + 8, np.float32)
    for b in range(n_blocks):
        blk = np.ascontiguousarray(x[b * block_pairs:(b + 1) * block_pairs], dtype=np.float32).reshape(-1)
        out = np.zeros(max(2, (2 * block_pairs) >> passes), np.float32)
        n = L.oracle_hb_cascade(_ptr(blk), C.c_int(blk.size), C.c_int(passes), _ptr(hist), _ptr(work), _ptr(out), C.c_int(fma))
        outs.append(out[:n].reshape(-1, 2).copy())
    return np.concatenate(outs)


# --------------------------------------------------------------------------- compiled reference

_refs = {}


def ref_available(variant="par"):
    return os.path.exists(_ref_path(variant))


def _ref_path(variant):
    name = {"par": "libdsdneo_ref.so", "avx2": "libdsdneo_ref_avx2.so", "fast": "libdsdneo_ref_fast.so"}[variant]
    return os.path.join(REF_DIR, name)


def ref(variant="par"):
    """The unmodified reference TUs compiled by oracle/Makefile (None when not built)."""
    if variant not in _refs:
        path = _ref_path(variant)
        if not os.path.exists(path):
            _refs[variant] = None
        else:
            L = C.CDLL(path)
            L.ref_demod_create.restype = C.c_void_p
            L.ref_demod_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
            L.ref_demod_create_wideband.restype = C.c_void_p
            L.ref_demod_create_wideband.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
            L.ref_demod_destroy.argtypes = [C.c_void_p]
            L.ref_demod_block.argtypes = [C.c_void_p, f32p, C.c_int, f32p, C.c_int]
            L.ref_demod_get_state.argtypes = [C.c_void_p, f32p]
            L.ref_demod_get_lpf_taps.argtypes = [C.c_void_p, f32p, C.c_int]
            L.ref_fir_complex_scalar.argtypes = [f32p, C.c_int, f32p, f32p, f32p, f32p, C.c_int]
            L.simd_fir_complex_apply.argtypes = [f32p, C.c_int, f32p, f32p, f32p, f32p, C.c_int]
            L.simd_fir_get_impl_name.restype = C.c_char_p
            L.mean_power.restype = C.c_float
            L.mean_power.argtypes = [f32p, C.c_int, C.c_int]
            _refs[variant] = L
    return _refs[variant]


class RefDemod:
    """One reference `struct demod_state` driven through the reference's own full_demod()."""

    def __init__(self, variant="par", rate=48000, symrate=4800, profile=4, lpf_enable=1, squelch=0.0):
        self.L = ref(variant)
        assert self.L is not None
        self.h = self.L.ref_demod_create(rate, symrate, profile, lpf_enable, squelch)
        assert self.h

    def block(self, iq_block):
        blk = np.ascontiguousarray(iq_block, dtype=np.float32).reshape(-1)
        out = np.empty(blk.size // 2, dtype=np.float32)
        n = self.L.ref_demod_block(self.h, _ptr(blk), blk.size, _ptr(out), out.size)
        assert n == out.size, n
        return out

    def run(self, iq_ch, block_pairs, n_blocks):
        return np.concatenate([self.block(iq_ch[b * block_pairs:(b + 1) * block_pairs]) for b in range(n_blocks)])

    def state(self):
        s = np.zeros(7, dtype=np.float32)
        self.L.ref_demod_get_state(self.h, _ptr(s))
        return dict(prev_i=s[0], prev_q=s[1], have_prev=int(s[2]), dc_est=s[3], peak=s[4], pwr=s[5], squelched=int(s[6]))

    def taps(self):
        buf = np.zeros(LPF_MAX_TAPS, dtype=np.float32)
        n = self.L.ref_demod_get_lpf_taps(self.h, _ptr(buf), LPF_MAX_TAPS)
        return buf[:n].copy()

    def close(self):
        if self.h:
            self.L.ref_demod_destroy(self.h)
            self.h = None

    __del__ = close


# --------------------------------------------------------------------------- synthetic signals

LEVELS = np.array([1.0, 3.0, -1.0, -3.0], dtype=np.float64)  # dibit 0,1,2,3 -> +1,+3,-1,-3 (dsd_dibit.c:963-976)


def synth_fsk_iq(rng, n_symbols, sps=10, dev_per_level=0.0785, amp=0.85, snr_db=None, dibits=None, phase0=0.19,
                 shape=True):
    """Phase-integrated 4-level FSK at `sps` samples/symbol -> [n_symbols*sps, 2] float32 (cf32).

    Same construction as the reference's synthesize_fsk_iq (tests/dsp/test_rtl_symbol_pipeline.cpp:39-54):
    per-sample phase step = dev_per_level * level; default 0.0785 rad = 2*pi*600 Hz/48 kHz per level
    (+-1.8 kHz at the outer levels, TIA-102 C4FM).  With shape=True the level sequence is smoothed with a
    raised-cosine-ish 1-symbol Hann kernel so the spectrum stays inside a 12.5 kHz channel.
    """
    if dibits is None:
        dibits = rng.integers(0, 4, size=n_symbols)
    lv = LEVELS[np.asarray(dibits)]
    f = np.repeat(lv, sps)
    if shape:
        k = np.hanning(sps + 2)[1:-1]
        k /= k.sum()
        f = np.convolve(f, k, mode="same")
    ph = phase0 + np.cumsum(f * dev_per_level)
    z = amp * np.exp(1j * ph)
    if snr_db is not None:
        sigma = amp * 10 ** (-snr_db / 20.0) / np.sqrt(2.0)
        z = z + sigma * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))
    out = np.empty((z.size, 2), dtype=np.float32)
    out[:, 0] = z.real
    out[:, 1] = z.imag
    return out


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def first_mismatch(a, b):
    d = np.nonzero(a.reshape(-1).view(np.uint32) != b.reshape(-1).view(np.uint32))[0]
    return None if d.size == 0 else (int(d[0]), float(a.reshape(-1)[d[0]]), float(b.reshape(-1)[d[0]]), int(d.size))


def oracle_pfb(x_with_hist, hist_len, h, M, channels, n_out, D=None):
    """Float64 direct-form channelizer oracle. x_with_hist: [hist_len + n, 2] float32. Returns complex128 [len(channels), n_out]."""
    D = M if D is None else D
    xh = np.ascontiguousarray(x_with_hist, dtype=np.float32).reshape(-1)
    h = np.ascontiguousarray(h, dtype=np.float32)
    ch = np.ascontiguousarray(channels, dtype=np.int32)
    out = np.zeros((ch.size, n_out, 2), dtype=np.float64)
    oracle().oracle_pfb_direct(_ptr(xh), hist_len, xh.size // 2 - hist_len, _ptr(h), h.size, M, D, _ptr(ch, i32p), ch.size,
                               out.ctypes.data_as(C.POINTER(C.c_double)), n_out)
    return out[..., 0] + 1j * out[..., 1]


def synth_wideband(rng, M, n_out, active, fs_ratio_dev=0.0785, snr_db=30.0, amp=0.5):
    """Sum of FSK carriers on the channelizer grid: channel k at k/M cycles/sample, each a 4-level FSK at
    10 samples/symbol of the CHANNEL rate.  Returns ([n_out*M, 2] float32, {k: dibits})."""
    n = n_out * M
    t = np.arange(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    truth = {}
    for k in active:
        nsym = n_out // 10 + 2
        dib = rng.integers(0, 4, size=nsym)
        lv = np.repeat(LEVELS[dib], 10 * M)[:n]
        ph = np.cumsum(lv * (fs_ratio_dev / M)) + rng.uniform(0, 2 * np.pi)
        x += np.exp(1j * (ph + 2 * np.pi * (k / M) * t))
        truth[k] = dib
    x *= amp / max(1, len(active)) ** 0.5
    sigma = amp * 10 ** (-snr_db / 20.0) / np.sqrt(2.0)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    out = np.empty((n, 2), dtype=np.float32)
    out[:, 0] = x.real
    out[:, 1] = x.imag
    return out, truth


# --------------------------------------------------------------------------- FEC bindings

class P25Candidate(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 12), ("metric", C.c_uint32)]


_fec_bound = {}


def oracle_fec():
    L = oracle()
    if "o" not in _fec_bound:
        i16p, u32p = C.POINTER(C.c_int16), C.POINTER(C.c_uint32)
        L.oracle_hamming_decode.argtypes = [C.c_int, u8p, u8p]
        L.oracle_golay_24_12_decode.argtypes = [u8p]
        L.oracle_golay_24_12_encode.argtypes = [u8p, u8p]
        L.oracle_golay_20_8_decode.argtypes = [u8p]
        L.oracle_qr_16_7_6_decode.argtypes = [u8p]
        L.oracle_bptc_deinterleave.argtypes = [u8p, u8p]
        L.oracle_bptc_196x96_extract.restype = C.c_uint
        L.oracle_bptc_196x96_extract.argtypes = [u8p, u8p, u8p, i32p]
        L.oracle_p25_12_soft_llr.argtypes = [i16p, u8p]
        L.oracle_p25_12_soft_llr_list.argtypes = [i16p, u8p, u32p, C.c_int]
        L.oracle_rs63_decode.argtypes = [C.c_int, i32p, i32p]
        L.oracle_rs63_encode.argtypes = [C.c_int, i32p, i32p]
        L.oracle_p25_rs_decode.argtypes = [C.c_int, C.c_int, u8p, u8p]
        L.oracle_fec_init()
        _fec_bound["o"] = True
    return L


def ref_fec(variant="par"):
    L = ref(variant)
    if L is None:
        return None
    key = "r" + variant
    if key not in _fec_bound:
        i16p = C.POINTER(C.c_int16)
        L.Hamming_7_4_decode.restype = C.c_bool
        L.Hamming_7_4_decode.argtypes = [u8p]
        for nm in ("Hamming_12_8_decode", "Hamming_13_9_decode", "Hamming_15_11_decode", "Hamming_16_11_4_decode"):
            getattr(L, nm).restype = C.c_bool
            getattr(L, nm).argtypes = [u8p, u8p, C.c_int]
        for nm in ("Golay_24_12_decode", "Golay_20_8_decode", "QR_16_7_6_decode"):
            getattr(L, nm).restype = C.c_bool
            getattr(L, nm).argtypes = [u8p]
        L.Golay_24_12_encode.argtypes = [u8p, u8p]
        L.BPTCDeInterleaveDMRData.argtypes = [u8p, u8p]
        L.BPTC_196x96_Extract_Data.restype = C.c_uint32
        L.BPTC_196x96_Extract_Data.argtypes = [u8p, u8p, u8p]
        L.p25_12_soft_llr.argtypes = [u8p, i16p, u8p]
        L.p25_12_soft_llr_list.argtypes = [u8p, i16p, C.POINTER(P25Candidate), C.c_int]
        for nm in ("check_and_fix_redsolomon_36_20_17", "check_and_fix_reedsolomon_24_12_13", "check_and_fix_reedsolomon_24_16_9"):
            getattr(L, nm).argtypes = [u8p, u8p]
        L.ref_rs63_decode.argtypes = [C.c_int, i32p, i32p]
        L.InitAllFecFunction()
        _fec_bound[key] = True
    return L


def u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def p25_trellis_encode(rng, dibits49=None):
    """Encode 49 dibits with the P25 half-rate trellis (p25_12.c:19 transition table) and interleave them
    (trellis34.c:8-13).  Returns (dibits49, tx_dibits98)."""
    dtm = [2, 12, 1, 15, 14, 0, 13, 3, 9, 7, 10, 4, 5, 11, 6, 8]
    if dibits49 is None:
        dibits49 = rng.integers(0, 4, 49)
        dibits49[48] = 0
    st = 0
    dei = np.zeros(98, dtype=np.int64)
    for i, d in enumerate(dibits49):
        nib = dtm[(st << 2) | int(d)]
        dei[2 * i] = (nib >> 2) & 3
        dei[2 * i + 1] = nib & 3
        st = int(d)
    tbl = []
    for g in range(4):
        for j in range(2 * g, 98, 8):
            tbl += [j, j + 1]
    tx = np.zeros(98, dtype=np.int64)
    for i in range(98):
        tx[i] = dei[tbl[i]]
    return np.asarray(dibits49), tx


def dibits_to_llr(tx98, mag=200, rng=None, noise=0.0):
    """int16 LLR pairs (positive = bit 1) for 98 dibits, optionally with Gaussian noise."""
    bits = np.stack([(tx98 >> 1) & 1, tx98 & 1], axis=1).reshape(-1)
    llr = np.where(bits == 1, mag, -mag).astype(np.float64)
    if rng is not None and noise > 0:
        llr = llr + rng.standard_normal(llr.size) * noise
    return np.clip(np.rint(llr), -32768, 32767).astype(np.int16)


# --------------------------------------------------------------------------- sample side

class OracleSymChan(C.Structure):
    _fields_ = [
        ("output_rate_hz", C.c_int), ("symbol_rate_hz", C.c_int), ("use_filter", C.c_int), ("window_l", C.c_int),
        ("track_minmax", C.c_int), ("negative", C.c_int), ("rf_mod", C.c_int), ("ssize", C.c_int), ("msize", C.c_int), ("taps_len", C.c_int),
        ("taps", C.c_float * 256), ("fir_hist", C.c_float * 256), ("fir_head", C.c_int),
        ("sps_num", C.c_int), ("sps_den", C.c_int), ("sps_accum", C.c_int),
        ("sps", C.c_int), ("center_idx", C.c_int), ("jitter", C.c_int), ("lastsample", C.c_float),
        ("min", C.c_float), ("max", C.c_float), ("center", C.c_float), ("umid", C.c_float), ("lmid", C.c_float),
        ("minref", C.c_float), ("maxref", C.c_float), ("sbuf", C.c_float * 128), ("sidx", C.c_int),
        ("minbuf", C.c_float * 1024), ("maxbuf", C.c_float * 1024), ("midx", C.c_int), ("sum_window", C.c_int),
        ("minbuf_sum", C.c_double), ("maxbuf_sum", C.c_double), ("symbolcnt", C.c_long),
    ]


# reference sync-type ids used by the tests (include/dsd-neo/core/synctype_ids.h:30-31,68,123)
SYNC_NONE, SYNC_P25P1_POS, SYNC_P25P1_NEG, SYNC_DMR_BS_DATA_POS = -1, 0, 1, 10
# which of the reference's matched filters a sync class selects (dsd_symbol.c:301-337) and its window / tracking
SYNC_CLASS = {
    SYNC_NONE: dict(filter=None, window_l=2, track=0, negative=0),
    SYNC_P25P1_POS: dict(filter=0, window_l=2, track=1, negative=0),
    SYNC_P25P1_NEG: dict(filter=0, window_l=2, track=1, negative=1),
    SYNC_DMR_BS_DATA_POS: dict(filter=1, window_l=1, track=0, negative=0),
}


def oracle_sym():
    L = oracle()
    if "sym" not in _fec_bound:
        i16p = C.POINTER(C.c_int16)
        lp = C.POINTER(C.c_long)
        L.oracle_sym_init.argtypes = [C.POINTER(OracleSymChan), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int,
                                      C.c_int, C.c_int]
        L.oracle_sym_run_symbols.restype = C.c_long
        L.oracle_sym_run_symbols.argtypes = [C.POINTER(OracleSymChan), C.c_int, f32p, C.c_long, C.c_long, f32p, C.c_long, lp]
        L.oracle_sym_run_dibits.restype = C.c_long
        L.oracle_sym_run_dibits.argtypes = [C.POINTER(OracleSymChan), f32p, C.c_long, C.c_long, u8p, u8p, i16p, f32p, C.c_long, lp]
        _fec_bound["sym"] = True
    return L


def ref_sym(variant="par"):
    L = ref(variant)
    if L is None:
        return None
    key = "s" + variant
    if key not in _fec_bound:
        i16p = C.POINTER(C.c_int16)
        L.ref_sym_create.restype = C.c_void_p
        L.ref_sym_create.argtypes = [C.c_int] * 7
        L.ref_sym_destroy.argtypes = [C.c_void_p]
        L.ref_sym_feed.argtypes = [C.c_void_p, f32p, C.c_long]
        L.ref_sym_set_sync.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_sym_get_symbols.restype = C.c_long
        L.ref_sym_get_symbols.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long, f32p]
        L.ref_sym_get_dibits.restype = C.c_long
        L.ref_sym_get_dibits.argtypes = [C.c_void_p, C.c_long, C.c_long, u8p, u8p, i16p, f32p]
        L.ref_sym_get_state.argtypes = [C.c_void_p, f32p, i32p]
        L.ref_sym_consumed.restype = C.c_long
        L.ref_sym_consumed.argtypes = [C.c_void_p]
        L.ref_sps_fir_taps.argtypes = [C.c_int, C.c_int, f32p, C.c_int]
        _fec_bound[key] = True
    return L


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def sps_fir_taps(which, sps):
    """Normalised matched-filter taps the reference uses for filter `which` at `sps` (0 p25, 1 dmr, ...).
    From the compiled reference when present, else from the committed golden fixture (tests/golden/make_golden.py)."""
    R = ref_sym()
    if R is not None:
        buf = np.zeros(1024, np.float32)
        n = R.ref_sps_fir_taps(which, sps, _ptr(buf), 1024)
        assert n > 0
        return buf[:n].copy()
    z = np.load(os.path.join(GOLDEN_DIR, "sps_fir_taps.npz"))
    return z["f%d_sps%d" % (which, sps)]


def synth_disc(rng, n_symbols, sps=10, level=9000.0, noise=0.0, drift=0.0, dibits=None):
    """Discriminator-like sample stream (what full_demod emits): 4-level, Hann-smoothed transitions, +-3*level peaks."""
    if dibits is None:
        dibits = rng.integers(0, 4, n_symbols)
    x = np.repeat(LEVELS[np.asarray(dibits)] * level, sps)
    k = np.hanning(sps + 2)[1:-1]
    k /= k.sum()
    x = np.convolve(x, k, mode="same")
    if drift:
        x = x + drift * np.sin(np.arange(x.size) * 2e-4)
    if noise:
        x = x + rng.standard_normal(x.size) * noise
    return x.astype(np.float32), np.asarray(dibits)


def c4fm_shaping_taps(ntaps=241, fs=48000.0, rs=4800.0, alpha=0.2):
    """P25 C4FM transmit shaping (TIA-102.BAAA: raised cosine alpha 0.2 times x/sin(x) pre-emphasis), frequency-sampled;
    the reference's p25_filter (de-emphasis sinc) brings it back to a raised cosine.  Gain 10 = one impulse per 10 samples."""
    N = 4096
    f = np.fft.rfftfreq(N, 1 / fs)
    T = 1 / rs
    f1, f2 = (1 - alpha) / (2 * T), (1 + alpha) / (2 * T)
    rc = np.zeros_like(f)
    rc[f <= f1] = 1.0
    m = (f > f1) & (f <= f2)
    rc[m] = 0.5 * (1 + np.cos(np.pi * T / alpha * (f[m] - f1)))
    x = np.pi * f * T
    shp = np.ones_like(f)
    shp[x > 0] = x[x > 0] / np.sin(x[x > 0])
    h = np.fft.irfft(rc * np.where(f <= f2, shp, 0.0), N)
    return np.roll(h, ntaps // 2)[:ntaps] * 10.0


def synth_c4fm_disc(rng, dibits, level=9000.0, noise=0.0, pre=9):
    """Discriminator-level C4FM stream at 10 samples/symbol whose symbol centres line up with getSymbol's window after the
    91-tap p25_filter (45-sample delay + `pre` = 54 = 5 symbols + centre index 4)."""
    imp = np.zeros(len(dibits) * 10)
    imp[::10] = LEVELS[np.asarray(dibits)] * level
    x = np.convolve(imp, c4fm_shaping_taps(), mode="same")
    if noise:
        x = x + rng.standard_normal(x.size) * noise
    return np.concatenate([np.zeros(pre), x]).astype(np.float32)


def nyquist_tx_taps(rx_taps, box, ntaps=241, fs=48000.0, rs=4800.0, alpha=0.2):
    """Transmit shaping that makes (tx * rx * box-average of `box` samples) a raised-cosine Nyquist pulse, so getSymbol's
    window mean after the reference's matched filter `rx_taps` lands on the transmitted level.  Gain 10 = one impulse per
    10 samples."""
    N = 4096
    f = np.fft.rfftfreq(N, 1 / fs)
    T = 1 / rs
    f1, f2 = (1 - alpha) / (2 * T), (1 + alpha) / (2 * T)
    rc = np.zeros_like(f)
    rc[f <= f1] = 1.0
    m = (f > f1) & (f <= f2)
    rc[m] = 0.5 * (1 + np.cos(np.pi * T / alpha * (f[m] - f1)))
    RX = np.abs(np.fft.rfft(np.asarray(rx_taps, np.float64), N)) * np.abs(np.fft.rfft(np.ones(box) / box, N))
    h = np.fft.irfft(np.where(f <= f2, rc / np.maximum(RX, 1e-3), 0.0), N)
    return np.roll(h, ntaps // 2)[:ntaps] * 10.0


def synth_dmr_disc(rng, dibits, rx_taps, level=10000.0, noise=0.0, pre=5):
    """Discriminator-level 4FSK stream for the DMR/YSF class (61-tap RRC matched filter, window centre-1..centre+2, fixed
    +-20000 thresholds): symbol centres aligned for `pre` = 5 leading samples."""
    imp = np.zeros(len(dibits) * 10)
    imp[::10] = LEVELS[np.asarray(dibits)] * level
    x = np.convolve(imp, nyquist_tx_taps(rx_taps, 4), mode="same")
    if noise:
        x = x + rng.standard_normal(x.size) * noise
    return np.concatenate([np.zeros(pre), x]).astype(np.float32)


def hamming_parity_bruteforce(code, data_bits, n):
    """Parity bits that make [data | parity] a codeword of the reference's Hamming code `code` (oracle_hamming_decode
    accepts it unchanged); test-side encoder for BPTC construction."""
    O = oracle_fec()
    k = len(data_bits)
    for v in range(1 << (n - k)):
        word = np.array(list(data_bits) + [(v >> (n - k - 1 - i)) & 1 for i in range(n - k)], np.uint8)
        probe, dec = word.copy(), np.zeros(k, np.uint8)
        if oracle_fec().oracle_hamming_decode(code, _ptr(probe, u8p), _ptr(dec, u8p)) and np.array_equal(probe, word):
            return word[k:]
    raise AssertionError("no parity found")


def bptc_196x96_encode(payload96, interleave=True):
    """DMR BPTC(196,96): 13 x 15 product code (rows 0-8 Hamming(15,11), all columns Hamming(13,9)), reserved bits zero,
    interleaved as the transmitter does (received[t] = deinterleaved[13 t mod 196], the inverse of bptc.c:51-59)."""
    m = np.zeros((13, 15), np.uint8)
    p = list(payload96)
    m[0, 3:11] = p[:8]
    for r in range(1, 9):
        m[r, :11] = p[8 + 11 * (r - 1):8 + 11 * r]
    for r in range(9):
        m[r, 11:] = hamming_parity_bruteforce(3, m[r, :11], 15)   # ORACLE_HAMMING_15_11
    for j in range(15):
        m[9:, j] = hamming_parity_bruteforce(2, m[:9, j], 13)     # ORACLE_HAMMING_13_9
    dei = np.zeros(196, np.uint8)
    dei[1:] = m.reshape(-1)
    if not interleave:
        return dei
    return np.array([dei[(13 * t) % 196] for t in range(196)], np.uint8)


def conv_k5_encode(bits):
    """Rate-1/2 K=5 encoder used by M17 / NXDN / YSF (G1 = 1+D^3+D^4, G2 = 1+D+D^2+D^4), returns 2*len(bits) bits."""
    sr = 0
    out = []
    for b in bits:
        sr = ((sr << 1) | int(b)) & 0x1F
        g1 = ((sr >> 0) ^ (sr >> 3) ^ (sr >> 4)) & 1
        g2 = ((sr >> 0) ^ (sr >> 1) ^ (sr >> 2) ^ (sr >> 4)) & 1
        out += [g1, g2]
    return np.array(out, dtype=np.uint8)


# --------------------------------------------------------------------------- CQPSK block side (section 8f rank 3)

FLL_MAX_TAPS = 48
CQPSK_RING = 64


class OracleCqpskChan(C.Structure):
    _fields_ = [
        ("rate_out_hz", C.c_int), ("sps", C.c_int), ("ted_gain", C.c_float), ("ted_gain_is_set", C.c_int),
        ("agc_avg", C.c_float),
        ("fll_ntaps", C.c_int),
        ("fll_alpha", C.c_float), ("fll_beta", C.c_float), ("fll_phase", C.c_float), ("fll_freq", C.c_float),
        ("fll_lower_r", C.c_float * FLL_MAX_TAPS), ("fll_lower_i", C.c_float * FLL_MAX_TAPS),
        ("fll_upper_r", C.c_float * FLL_MAX_TAPS), ("fll_upper_i", C.c_float * FLL_MAX_TAPS),
        ("ring_r", C.c_float * CQPSK_RING), ("ring_j", C.c_float * CQPSK_RING),
        ("pushed", C.c_long), ("consumed", C.c_long),
        ("mu", C.c_float), ("omega", C.c_float), ("omega_mid", C.c_float), ("omega_rel", C.c_float),
        ("last_r", C.c_float), ("last_j", C.c_float), ("lock_accum", C.c_float), ("ted_effective_gain", C.c_float),
        ("lock_count", C.c_int), ("ted_span", C.c_int),
        ("diff_prev_r", C.c_float), ("diff_prev_j", C.c_float),
        ("costas_alpha", C.c_float), ("costas_beta", C.c_float), ("costas_phase", C.c_float), ("costas_freq", C.c_float),
        ("costas_err_smooth", C.c_float), ("costas_error", C.c_float),
        ("m_err_abs", C.c_float), ("m_err_raw_abs", C.c_float), ("m_conf_acc", C.c_float), ("m_zero_conf", C.c_int),
        ("costas_err_avg_q14", C.c_int), ("costas_err_raw_avg_q14", C.c_int), ("costas_conf_avg_q14", C.c_int),
        ("costas_zero_conf_pct", C.c_int),
    ]


CQPSK_STATE_KEYS = ["agc_avg", "fll_phase", "fll_freq", "fll_alpha", "fll_beta", "mu", "omega", "last_r", "last_j",
                    "lock_accum", "lock_count", "ted_effective_gain", "diff_prev_r", "diff_prev_j", "costas_phase",
                    "costas_freq", "costas_error", "costas_err_smooth", "costas_err_avg_q14", "costas_err_raw_avg_q14",
                    "costas_conf_avg_q14", "costas_zero_conf_pct", "channel_pwr", "channel_squelched"]


def oracle_cqpsk():
    L = oracle()
    L.oracle_cqpsk_chan_init.argtypes = [C.POINTER(OracleCqpskChan), C.c_int, C.c_int, C.c_float, C.c_int]
    L.oracle_cqpsk_block.argtypes = [C.POINTER(OracleCqpskChan), f32p, C.c_int, f32p]
    L.oracle_full_demod_cqpsk_block.argtypes = [C.POINTER(OracleDemodChan), C.POINTER(OracleCqpskChan), f32p, C.c_int,
                                                f32p, f32p]
    L.oracle_fll_band_edge_design.argtypes = [C.c_int, f32p, f32p, f32p, f32p, C.c_int]
    L.oracle_cqpsk_mmse_table.restype = C.POINTER(C.c_float)
    return L


class OracleCqpsk:
    """One channel of the CQPSK oracle: full_demod() with output_kind == SYMBOL_CQPSK, block by block."""

    def __init__(self, rate=24000, sps=5, lpf_enable=1, squelch=0.0, fir_fma=0, ted_gain=0.0, ted_gain_is_set=0):
        self.L = oracle_cqpsk()
        self.c, self.q = OracleDemodChan(), OracleCqpskChan()
        assert self.L.oracle_demod_chan_init(C.byref(self.c), rate, 5, lpf_enable, squelch, fir_fma) == 0
        assert self.L.oracle_cqpsk_chan_init(C.byref(self.q), rate, sps, ted_gain, ted_gain_is_set) == 0

    def block(self, iq_block):
        blk = np.ascontiguousarray(iq_block, dtype=np.float32).reshape(-1)
        out = np.empty(blk.size // 2 + 2, dtype=np.float32)
        scratch = np.empty(blk.size, dtype=np.float32)
        n = self.L.oracle_full_demod_cqpsk_block(C.byref(self.c), C.byref(self.q), _ptr(blk), blk.size, _ptr(scratch),
                                                 _ptr(out))
        assert n >= 0
        return out[:n].copy()

    def run(self, iq_ch, block_pairs, n_blocks):
        outs = [self.block(iq_ch[b * block_pairs:(b + 1) * block_pairs]) for b in range(n_blocks)]
        return np.concatenate(outs), np.array([o.size for o in outs], np.int32)

    def state(self):
        q, c = self.q, self.c
        d = {k: getattr(q, k) for k in CQPSK_STATE_KEYS if hasattr(q, k)}
        d["channel_pwr"], d["channel_squelched"] = c.channel_pwr, c.channel_squelched
        return d


class RefCqpsk:
    """One reference `struct demod_state` in CQPSK symbol mode, driven through the reference's own full_demod()."""

    def __init__(self, variant="par", rate=24000, symrate=4800, sps=5, lpf_enable=1, squelch=0.0, ted_gain=0.0,
                 ted_gain_is_set=0):
        self.L = ref(variant)
        assert self.L is not None
        self.L.ref_demod_create_cqpsk.restype = C.c_void_p
        self.L.ref_demod_create_cqpsk.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        self.L.ref_demod_get_cqpsk_state.argtypes = [C.c_void_p, f32p]
        self.L.ref_demod_get_fll_taps.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, C.c_int]
        self.h = self.L.ref_demod_create_cqpsk(rate, symrate, sps, lpf_enable, squelch, ted_gain, ted_gain_is_set)
        assert self.h

    def block(self, iq_block):
        blk = np.ascontiguousarray(iq_block, dtype=np.float32).reshape(-1)
        out = np.empty(blk.size // 2 + 2, dtype=np.float32)
        n = self.L.ref_demod_block(self.h, _ptr(blk), blk.size, _ptr(out), out.size)
        assert 0 <= n <= out.size, n
        return out[:n].copy()

    def run(self, iq_ch, block_pairs, n_blocks):
        outs = [self.block(iq_ch[b * block_pairs:(b + 1) * block_pairs]) for b in range(n_blocks)]
        return np.concatenate(outs), np.array([o.size for o in outs], np.int32)

    def state(self):
        s = np.zeros(24, dtype=np.float32)
        self.L.ref_demod_get_cqpsk_state(self.h, _ptr(s))
        d = dict(zip(CQPSK_STATE_KEYS, s.tolist()))
        for k in ("lock_count", "costas_err_avg_q14", "costas_err_raw_avg_q14", "costas_conf_avg_q14",
                  "costas_zero_conf_pct", "channel_squelched"):
            d[k] = int(d[k])
        return d

    def fll_taps(self):
        bufs = [np.zeros(FLL_MAX_TAPS, np.float32) for _ in range(4)]
        n = self.L.ref_demod_get_fll_taps(self.h, *[_ptr(b) for b in bufs], FLL_MAX_TAPS)
        return [b[:n].copy() for b in bufs]

    def close(self):
        if self.h:
            self.L.ref_demod_destroy(self.h)
            self.h = None

    __del__ = close


def cqpsk_state_equal(a, b):
    """Bit-level comparison of two CQPSK state dicts (floats compared as float32 bit patterns)."""
    bad = []
    for k in CQPSK_STATE_KEYS:
        if k not in a or k not in b:
            continue
        x, y = a[k], b[k]
        if isinstance(x, int) and isinstance(y, int):
            same = x == y
        else:
            same = np.float32(x).tobytes() == np.float32(y).tobytes()
        if not same:
            bad.append((k, x, y))
    return bad


def rrc_taps(sps, span=8, alpha=0.2):
    """Root-raised-cosine pulse, unit energy (standard closed form; signal synthesis only)."""
    n = np.arange(-span * sps, span * sps + 1, dtype=np.float64) / sps
    h = np.zeros_like(n)
    for i, t in enumerate(n):
        if abs(t) < 1e-12:
            h[i] = 1.0 - alpha + 4 * alpha / np.pi
        elif abs(abs(t) - 1 / (4 * alpha)) < 1e-9:
            h[i] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
        else:
            h[i] = (np.sin(np.pi * t * (1 - alpha)) + 4 * alpha * t * np.cos(np.pi * t * (1 + alpha))) / (np.pi * t * (1 - (4 * alpha * t) ** 2))
    return h / np.sqrt((h ** 2).sum())


def synth_cqpsk_iq(rng, n_symbols, sps=5, amp=0.6, snr_db=None, cfo=0.0, dibits=None, timing=0.0, phase0=0.3):
    """pi/4-DQPSK (P25 LSM-like): each dibit advances the carrier phase by {+1,+3,-1,-3} x pi/4 (same dibit map as
    the 4-level slicer), impulses shaped by a raised-cosine pulse (RRC twice), `cfo` rad/sample carrier offset, a
    fractional `timing` offset in samples, complex AWGN at `snr_db` Es/N0.  Returns ([n, 2] float32, dibits)."""
    if dibits is None:
        dibits = rng.integers(0, 4, n_symbols)
    steps = LEVELS[dibits] * (np.pi / 4)
    ph = phase0 + np.cumsum(steps)
    sym = np.exp(1j * ph)
    up = np.zeros(n_symbols * sps, np.complex128)
    up[::sps] = sym
    h = rrc_taps(sps)
    rc = np.convolve(h, h)
    if timing:
        # fractional delay by linear-phase interpolation of the pulse
        t = np.arange(rc.size, dtype=np.float64)
        rc = np.interp(t - timing, t, rc, left=0.0, right=0.0)
    x = np.convolve(up, rc)[rc.size // 2: rc.size // 2 + n_symbols * sps]
    x = x / np.sqrt(np.mean(np.abs(x) ** 2)) * amp
    n = np.arange(x.size)
    x = x * np.exp(1j * cfo * n)
    if snr_db is not None:
        es = amp * amp * sps
        n0 = es / (10 ** (snr_db / 10))
        x = x + (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size)) * np.sqrt(n0 / 2)
    out = np.stack([x.real, x.imag], axis=1).astype(np.float32)
    return out, dibits


def synth_wideband_cqpsk(rng, M, n_out, active, sps=10, snr_db=30.0, amp=0.5, cfo_hz_frac=0.0):
    """Sum of pi/4-DQPSK carriers on the channelizer grid: channel k at k/M cycles/sample, each shaped at `sps` samples
    per symbol of the CHANNEL rate and interpolated to the wideband rate.  `cfo_hz_frac` offsets every carrier by that
    fraction of the channel spacing.  Returns ([n_out*M, 2] float32, {k: dibits})."""
    n = n_out * M
    t = np.arange(n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    truth = {}
    for k in active:
        nsym = n_out // sps + 2
        base, dib = synth_cqpsk_iq(rng, nsym, sps=sps, amp=1.0, snr_db=None, cfo=0.0, timing=0.0, phase0=rng.uniform(0, 6.28))
        b = (base[:, 0] + 1j * base[:, 1])[:n_out + 1]
        # linear interpolation from the channel rate to the wideband rate (the images fall outside the channel filter)
        pos = np.arange(n, dtype=np.float64) / M
        i0 = np.minimum(pos.astype(np.int64), b.size - 2)
        fr = pos - i0
        up = b[i0] * (1 - fr) + b[i0 + 1] * fr
        x += up * np.exp(2j * np.pi * ((k + cfo_hz_frac) / M) * t)
        truth[k] = dib
    x *= amp / max(1, len(active)) ** 0.5
    sigma = amp * 10 ** (-snr_db / 20.0) / np.sqrt(2.0)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    out = np.empty((n, 2), dtype=np.float32)
    out[:, 0] = x.real
    out[:, 1] = x.imag
    return out, truth


# --------------------------------------------------------------------------- P25 Phase 1 air-interface framing (section 8f rank 4)

P25P1_SYNC_DIBITS = [int(c) for c in "111113113311333313133333"]


def p25p1_insert_status(frame_dibits_no_status, status_dibit=2):
    """Insert a status symbol after every 35 dibits of a frame (frame offsets 35, 71, ...), as on the air interface."""
    out = []
    for i, d in enumerate(frame_dibits_no_status):
        out.append(int(d))
        if len(out) % 36 == 35:
            out.append(status_dibit)
    return np.array(out, dtype=np.int64)


def p25_crc16_ccitt_inverted(bits):
    """ComputeCrcCCITT16b (src/protocol/p25/p25_crc.c:11-29): polynomial 0x1021, zero preset, inverted."""
    crc = 0
    for b in bits:
        crc = ((crc << 1) ^ 0x1021) & 0xFFFF if ((crc >> 15) & 1) ^ int(b) else (crc << 1) & 0xFFFF
    return crc ^ 0xFFFF


def p25_tsbk_dibits49(rng, last_block):
    """One TSBK as the 49 trellis input dibits: 80 random bits with the last-block flag as bit 0, CRC-16 over them, flush."""
    bits = rng.integers(0, 2, 80)
    bits[0] = 1 if last_block else 0
    crc = p25_crc16_ccitt_inverted(bits)
    bits96 = np.concatenate([bits, [(crc >> (15 - i)) & 1 for i in range(16)]])
    return np.concatenate([bits96[0::2] * 2 + bits96[1::2], [0]])


def p25p1_build_tsdu(rng, nac, n_blocks=3, bch_encode=None, valid_crc=False):
    """One TSDU: sync + NID(NAC, DUID 7, BCH(63,16) + parity 0) + n_blocks half-rate trellis blocks, status symbols
    inserted.  Returns (dibits incl. status, [49-dibit payloads]).  valid_crc: blocks carry a CRC-16 and the last-block flag
    on the final block only (what a control channel sends); otherwise random dibits."""
    duid = 0x7
    info = np.array([(nac >> (11 - i)) & 1 for i in range(12)] + [(duid >> (3 - i)) & 1 for i in range(4)], np.uint8)
    cw = bch_encode(info).astype(np.int64)
    bits = np.concatenate([cw, [0]])  # parity bit 0 for TSDU
    nid = bits[0::2] * 2 + bits[1::2]
    body, payloads = [np.array(P25P1_SYNC_DIBITS), nid], []
    for b in range(n_blocks):
        d49, tx98 = p25_trellis_encode(rng, p25_tsbk_dibits49(rng, b == n_blocks - 1) if valid_crc else None)
        body.append(tx98)
        payloads.append(d49)
    return p25p1_insert_status(np.concatenate(body)), payloads


class OracleCqpskSlicer(C.Structure):
    _fields_ = [("base", OracleSymChan), ("p25_slice", C.c_int), ("map_idx", C.c_int), ("snr_db", C.c_double)]


def oracle_cqpsk_slicer_run(symbols, negative=0, p25_slice=1, map_idx=0, snr_db=-100.0, ssize=128, msize=1024, state=None):
    """Symbol-rate CQPSK sample side (oracle/oracle_symbol.c): returns (dibits, reliability, llr [n, 2], state)."""
    L = oracle()
    L.oracle_cqpsk_slicer_init.argtypes = [C.POINTER(OracleCqpskSlicer), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    L.oracle_cqpsk_slicer_run.restype = C.c_long
    L.oracle_cqpsk_slicer_run.argtypes = [C.POINTER(OracleCqpskSlicer), f32p, C.c_long, u8p, u8p, C.POINTER(C.c_int16)]
    if state is None:
        state = OracleCqpskSlicer()
        L.oracle_cqpsk_slicer_init(C.byref(state), negative, p25_slice, map_idx, snr_db, ssize, msize)
    symbols = np.ascontiguousarray(symbols, np.float32)
    n = symbols.size
    d, r, l = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros((n, 2), np.int16)
    L.oracle_cqpsk_slicer_run(C.byref(state), _ptr(symbols), n, _ptr(d, u8p), _ptr(r, u8p), l.ctypes.data_as(C.POINTER(C.c_int16)))
    return d, r, l, state


# --------------------------------------------------------------------------- DMR BS data burst framing

DMR_BS_DATA_SYNC_DIBITS = [int(c) for c in "313333111331131131331131"]
DMR_CACH_INTERLEAVE = [0, 7, 8, 9, 1, 10, 11, 12, 2, 13, 14, 15, 3, 16, 4, 17, 18, 19, 5, 20, 21, 22, 6, 23]


def golay_20_8_encode_bruteforce(data8):
    """Golay(20,8) codeword for 8 data bits: the 12 parity bits the (pinned) oracle decoder accepts without a change."""
    O = oracle_fec()
    d = [int(b) for b in data8]
    for par in range(4096):
        w = np.array(d + [(par >> (11 - i)) & 1 for i in range(12)], np.uint8)
        ref = w.copy()
        if O.oracle_golay_20_8_decode(_ptr(w, u8p)) and np.array_equal(w, ref):
            # a clean codeword is one whose single-bit flips all decode back to it
            t = ref.copy()
            t[3] ^= 1
            if O.oracle_golay_20_8_decode(_ptr(t, u8p)) and np.array_equal(t, ref):
                return ref
    raise AssertionError("no Golay(20,8) codeword found")


def hamming_7_4_encode_bruteforce(data4):
    O = oracle_fec()
    for par in range(8):
        w = np.array([int(b) for b in data4] + [(par >> (2 - i)) & 1 for i in range(3)], np.uint8)
        ref = w.copy()
        dec = np.zeros(4, np.uint8)
        t = ref.copy()
        t[1] ^= 1
        if O.oracle_hamming_decode(0, _ptr(t, u8p), _ptr(dec, u8p)) and np.array_equal(t, ref):
            return ref
    raise AssertionError("no Hamming(7,4) codeword found")


def dmr_build_bs_data_burst(rng, payload96, color_code, data_type, tact4=(1, 0, 1, 0)):
    """One DMR BS data burst in dibits: CACH (TACT Hamming(7,4) + 17 fragment bits, interleaved), 98 info bits, 10 slot-type
    bits, BS DATA sync, 10 slot-type bits, 98 info bits.  Returns (144 dibits, dict of what was sent)."""
    info = bptc_196x96_encode(payload96)
    st8 = [(color_code >> (3 - i)) & 1 for i in range(4)] + [(data_type >> (3 - i)) & 1 for i in range(4)]
    slot = golay_20_8_encode_bruteforce(st8)
    cach = np.zeros(24, np.uint8)
    cach[:7] = hamming_7_4_encode_bruteforce(tact4)
    cach[7:] = rng.integers(0, 2, 17)
    tx_cach = np.zeros(24, np.uint8)
    for i in range(24):
        tx_cach[i] = cach[DMR_CACH_INTERLEAVE[i]]
    bits = np.concatenate([tx_cach, info[:98], slot[:10]])
    first = (bits[0::2] << 1) | bits[1::2]
    second_bits = np.concatenate([slot[10:], info[98:]])
    second = (second_bits[0::2] << 1) | second_bits[1::2]
    dib = np.concatenate([first, np.array(DMR_BS_DATA_SYNC_DIBITS), second]).astype(np.int64)
    return dib, {"cach": cach, "info": info, "slot": slot}


# ---- P25 Phase 1 frames: record layouts (include/dsdneo_b200.h), reference frame harness, synthetic frame builders ----------

P25_FRAME_DTYPE = np.dtype([
    ("position", "<i8"), ("channel", "<i4"), ("voice_index", "<i4"), ("nac", "<i2"), ("nid_errs", "<i2"), ("nid_status", "i1"),
    ("duid", "u1"), ("n_tsbk", "u1"), ("tsbk_crc_ok", "u1"), ("rs_kind", "u1"), ("rs_status", "u1"), ("lsd_ok", "u1"),
    ("n_word_soft", "u1"), ("lsd", "u1", (2,)), ("reserved", "u1", (6,)), ("tsbk", "u1", (3, 12)), ("rs_data", "u1", (20,)),
    ("rs_in_data", "u1", (20,)), ("rs_in_parity", "u1", (16,))])
P25_VOICE_DTYPE = np.dtype([("bits", "<u4", (9, 8)), ("reliab", "u1", (9, 8, 23))])
assert P25_FRAME_DTYPE.itemsize == 128 and P25_VOICE_DTYPE.itemsize == 1944

REF_P25_DTYPE = np.dtype([
    ("consumed", "<i4"), ("overrun", "<i4"), ("nid_status", "<i4"), ("nac", "<i4"), ("duid", "<i4"), ("nid_errs", "<i4"),
    ("n_tsbk", "<i4"), ("tsbk_bytes", "u1", (3, 12)), ("tsbk_crc_err", "<i4", (3,)), ("tsbk_dibits", "u1", (3, 98)),
    ("tsbk_llr", "<i2", (3, 196)), ("n_imbe", "<i4"), ("imbe_bit", "u1", (9, 8, 23)), ("imbe_rel", "u1", (9, 8, 23)),
    ("n_words", "<i4"), ("word_code", "<i4", (48,)), ("word_rc", "<i4", (48,)), ("word_fixed", "<i4", (48,)),
    ("word_in", "u1", (48, 24)), ("word_out", "u1", (48, 12)), ("word_soft_called", "<i4", (48,)), ("word_soft_rc", "<i4", (48,)),
    ("rs_kind", "<i4"), ("rs_hard_rc", "<i4"), ("rs_soft_called", "<i4"), ("rs_soft_rc", "<i4"), ("rs_in_data", "u1", (120,)),
    ("rs_in_parity", "u1", (96,)), ("rs_out_data", "u1", (120,)), ("rs_data_reliab", "u1", (20,)), ("rs_parity_reliab", "u1", (16,)),
    ("n_lsd", "<i4"), ("lsd_in", "u1", (2, 16)), ("lsd_out", "u1", (2, 16)), ("lsd_llr", "<i2", (2, 16)), ("lsd_ok", "<i4", (2,))],
    align=True)

_ref_p25 = None


def ref_p25_available():
    return os.path.exists(os.path.join(REF_DIR, "libdsdneo_ref_p25.so"))


def ref_p25():
    """The UNMODIFIED reference P25p1 frame handlers replaying a dibit stream (oracle/ref_shim_p25.c)."""
    global _ref_p25
    if _ref_p25 is None:
        L = C.CDLL(os.path.join(REF_DIR, "libdsdneo_ref_p25.so"))
        assert L.ref_p25_frame_size() == REF_P25_DTYPE.itemsize, (L.ref_p25_frame_size(), REF_P25_DTYPE.itemsize)
        L.ref_p25_decode_frame.restype = C.c_long
        L.ref_p25_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_void_p]
        _ref_p25 = L
    return _ref_p25


def ref_p25_decode(dibits, reliab, llr, pos_last_sync, observed_nac=0):
    """One frame through dsd_dispatch_handle_p25p1; returns a REF_P25_DTYPE scalar."""
    rec = np.zeros(1, REF_P25_DTYPE)
    d, r, l = np.ascontiguousarray(dibits, np.uint8), np.ascontiguousarray(reliab, np.uint8), np.ascontiguousarray(llr, np.int16)
    ref_p25().ref_p25_decode_frame(d.ctypes.data, r.ctypes.data, l.ctypes.data, d.size, pos_last_sync + 1, observed_nac, rec.ctypes.data)
    return rec[0]


def oracle_p25_decode(dibits, llr, pos_last_sync, observed_nac=0, threshold=64):
    """oracle_p25p1_decode_frame; returns (consumed, frame record, voice record)."""
    O = oracle_fec()
    f, v = np.zeros(1, P25_FRAME_DTYPE), np.zeros(1, P25_VOICE_DTYPE)
    d, l = np.ascontiguousarray(dibits, np.uint8), np.ascontiguousarray(llr, np.int16)
    O.oracle_p25p1_decode_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = O.oracle_p25p1_decode_frame(d.ctypes.data, l.ctypes.data, d.size, pos_last_sync, observed_nac, threshold, f.ctypes.data,
                                    v.ctypes.data)
    return n, f[0], v[0]


def p25_words_from_bits(bits, n):
    b = np.asarray(bits[:6 * n], np.int64).reshape(n, 6)
    return (b * (1 << np.arange(5, -1, -1))).sum(axis=1).astype(np.uint8)


def p25_frames_agree(ref, f, v):
    """Field-by-field comparison of a reference harness record with an (oracle or device) frame + voice record.
    Returns a list of mismatching field names (empty = equal)."""
    bad = []
    if int(ref["nid_status"]) != int(f["nid_status"]):
        bad.append("nid_status")
    if ref["nid_status"] > 0:
        if int(ref["nac"]) != int(f["nac"]) or int(ref["duid"]) != int(f["duid"]) or int(ref["nid_errs"]) != int(f["nid_errs"]):
            bad.append("nid")
    else:
        return bad + (["duid"] if int(f["duid"]) != 0xFF else [])
    duid = int(ref["duid"])
    if duid == 7:
        n = int(ref["n_tsbk"])
        if n != int(f["n_tsbk"]):
            bad.append("n_tsbk")
        for b in range(min(n, int(f["n_tsbk"]))):
            if not np.array_equal(ref["tsbk_bytes"][b], f["tsbk"][b]):
                bad.append("tsbk%d" % b)
            if (ref["tsbk_crc_err"][b] == 0) != bool((f["tsbk_crc_ok"] >> b) & 1):
                bad.append("tsbk_crc%d" % b)
    if duid in (0x0, 0x5, 0xA, 0xF):
        kind = {0x0: 1, 0x5: 2, 0xA: 3, 0xF: 2}[duid]  # TDULC: RS(24,12,13) like LDU1, over Golay(24,12) dodeca words
        n_data, n_par = {1: (20, 16), 2: (12, 12), 3: (16, 8)}[kind]
        if int(ref["rs_kind"]) != kind or int(f["rs_kind"]) != kind:
            bad.append("rs_kind")
        if not np.array_equal(p25_words_from_bits(ref["rs_in_data"], n_data), f["rs_in_data"][:n_data]):
            bad.append("rs_in_data")
        if not np.array_equal(p25_words_from_bits(ref["rs_in_parity"], n_par), f["rs_in_parity"][:n_par]):
            bad.append("rs_in_parity")
        want = 0 if ref["rs_hard_rc"] == 0 else (1 if (ref["rs_soft_called"] and ref["rs_soft_rc"] == 0) else 2)
        if want != int(f["rs_status"]):
            bad.append("rs_status")
        if not np.array_equal(p25_words_from_bits(ref["rs_out_data"], n_data), f["rs_data"][:n_data]):
            bad.append("rs_data")
    if duid in (0x5, 0xA):
        if int(ref["n_imbe"]) != 9:
            bad.append("n_imbe")
        bits = ((v["bits"][:, :, None] >> np.arange(23, dtype=np.uint32)) & 1).astype(np.uint8)
        if not np.array_equal(bits, ref["imbe_bit"]):
            bad.append("imbe_bit")
        if not np.array_equal(v["reliab"], ref["imbe_rel"]):
            bad.append("imbe_rel")
        for k in range(2):
            val = int((ref["lsd_out"][k][:8].astype(np.int64) * (1 << np.arange(7, -1, -1))).sum())
            if val != int(f["lsd"][k]) or bool(ref["lsd_ok"][k]) != bool((f["lsd_ok"] >> k) & 1):
                bad.append("lsd%d" % k)
    return bad


IMBE_HI = [22, 66, 102, 43, 87, 115, 20, 64, 100, 41, 85, 151, 18, 62, 98, 39, 83, 149, 16, 60, 96, 37, 81, 147, 14, 58, 94, 35, 79,
           145, 12, 56, 92, 33, 77, 143, 10, 54, 128, 31, 75, 141, 8, 52, 126, 29, 73, 139, 6, 50, 124, 27, 71, 167, 4, 48, 122, 25,
           69, 165, 2, 46, 120, 23, 105, 163, 0, 90, 118, 67, 103, 161]
IMBE_LO = [44, 88, 116, 21, 65, 101, 42, 86, 152, 19, 63, 99, 40, 84, 150, 17, 61, 97, 38, 82, 148, 15, 59, 95, 36, 80, 146, 13, 57,
           93, 34, 78, 144, 11, 55, 129, 32, 76, 142, 9, 53, 127, 30, 74, 140, 7, 51, 125, 28, 72, 138, 5, 49, 123, 26, 70, 166, 3,
           47, 121, 24, 106, 164, 1, 91, 119, 68, 104, 162, 45, 89, 117]


def _bch_nid_encoder():
    from test_oracle_fec import bch_63_16_encode

    return bch_63_16_encode


def p25p1_nid_dibits(nac, duid):
    info = np.array([(nac >> (11 - i)) & 1 for i in range(12)] + [(duid >> (3 - i)) & 1 for i in range(4)], np.uint8)
    cw = _bch_nid_encoder()(info).astype(np.int64)
    parity = 1 if duid in (0x5, 0xA) else 0
    bits = np.concatenate([cw, [parity]])
    return bits[0::2] * 2 + bits[1::2]


def hamming_10_6_3_encode(word6):
    d = [(word6 >> (5 - i)) & 1 for i in range(6)]
    p = [d[0] ^ d[1] ^ d[2] ^ d[5], d[0] ^ d[1] ^ d[3] ^ d[5], d[0] ^ d[2] ^ d[3] ^ d[4], d[1] ^ d[2] ^ d[3] ^ d[4]]
    return d + p


_golay6_parity = None


def golay_24_6_encode(word6):
    """6 data bits + the 12 parity bits the reference decoder accepts with zero corrections (found once by search)."""
    global _golay6_parity
    if _golay6_parity is None:
        O = oracle_fec()
        # linear code: find the parity of each unit data word, the rest follows by XOR
        basis = []
        for k in range(6):
            data = np.zeros(6, np.uint8)
            data[k] = 1
            found = None
            for par in range(4096):
                p = np.array([(par >> (11 - i)) & 1 for i in range(12)], np.uint8)
                d = data.copy()
                fixed = C.c_int(0)
                if O.oracle_p25_golay24_decode(6, _ptr(d, u8p), _ptr(p, u8p), C.byref(fixed)) == 0 and fixed.value == 0 \
                        and np.array_equal(d, data):
                    found = par
                    break
            assert found is not None
            basis.append(found)
        _golay6_parity = basis
    par = 0
    for k in range(6):
        if (word6 >> (5 - k)) & 1:
            par ^= _golay6_parity[k]
    return [(word6 >> (5 - i)) & 1 for i in range(6)] + [(par >> (11 - i)) & 1 for i in range(12)]


def rs63_shortened_encode(data_syms, n_par):
    """Shortened RS(63, 63 - n_par) over GF(64): returns (parity symbols, data symbols) as the reference orders them."""
    O = oracle_fec()
    tt = n_par // 2
    data = np.zeros(63 - n_par, np.int32)
    data[:len(data_syms)] = data_syms
    cw = np.zeros(63, np.int32)
    O.oracle_rs63_encode(tt, data.ctypes.data_as(i32p), cw.ctypes.data_as(i32p))
    return cw[:n_par].copy(), cw[n_par:n_par + len(data_syms)].copy()


def _bits_to_dibits(bits):
    b = np.asarray(bits, np.int64)
    return b[0::2] * 2 + b[1::2]


def lsd_16_8_encode(byte):
    r = byte << 8
    for i in range(15, 7, -1):
        if (r >> i) & 1:
            r ^= 0x139 << (i - 8)
    par = r & 0xFF
    return [(byte >> (7 - i)) & 1 for i in range(8)] + [(par >> (7 - i)) & 1 for i in range(8)]


def p25p1_build_ldu(rng, nac, ldu2=False, voice=None):
    """One LDU1 / LDU2: sync + NID + 9 IMBE frames, 24 Hamming(10,6,3) hex words carrying an RS codeword, LSD.  Returns
    (dibits incl. status, truth dict)."""
    n_data = 16 if ldu2 else 12
    n_par = 24 - n_data
    data = rng.integers(0, 64, n_data)
    par, dat = rs63_shortened_encode(data, n_par)
    voice = rng.integers(0, 2, (9, 184)).astype(np.int64) if voice is None else np.asarray(voice, np.int64).reshape(9, 184)
    lsd = rng.integers(0, 256, 2)

    def imbe(k):
        return voice[k][IMBE_HI] * 2 + voice[k][IMBE_LO]

    def words(block):  # four air words starting at air index 4 * block
        out = []
        for w in range(4 * block, 4 * block + 4):
            sym = dat[n_data - 1 - w] if w < n_data else par[23 - w]
            out.append(_bits_to_dibits(hamming_10_6_3_encode(int(sym))))
        return np.concatenate(out)

    body = [np.array(P25P1_SYNC_DIBITS), p25p1_nid_dibits(nac, 0xA if ldu2 else 0x5), imbe(0), imbe(1), words(0), imbe(2), words(1),
            imbe(3), words(2), imbe(4), words(3), imbe(5), words(4), imbe(6), words(5), imbe(7),
            _bits_to_dibits(lsd_16_8_encode(int(lsd[0])) + lsd_16_8_encode(int(lsd[1]))), imbe(8)]
    mask = np.zeros(184, bool)
    mask[IMBE_HI + IMBE_LO] = True
    return p25p1_insert_status(np.concatenate(body)), {"rs_data": dat.astype(np.uint8), "voice": voice * mask, "lsd": lsd, "duid": 0xA if ldu2 else 0x5}


_golay12_parity = None


def golay_24_12_encode(word12):
    """12 data bits + the 12 parity bits check_and_fix_golay_24_12 accepts with zero corrections (basis found once by search)."""
    global _golay12_parity
    if _golay12_parity is None:
        O = oracle_fec()
        basis = []
        for k in range(12):
            data = np.zeros(12, np.uint8)
            data[k] = 1
            found = None
            for par in range(4096):
                p = np.array([(par >> (11 - i)) & 1 for i in range(12)], np.uint8)
                d = data.copy()
                fixed = C.c_int(0)
                # the decoder also accepts a word whose only error sits in a parity bit: a true codeword of the extended Golay
                # code has a weight divisible by four
                if (1 + bin(par).count("1")) % 4 == 0 and O.oracle_p25_golay24_decode(12, _ptr(d, u8p), _ptr(p, u8p), C.byref(fixed)) == 0 \
                        and fixed.value == 0 and np.array_equal(d, data):
                    found = par
                    break
            assert found is not None
            basis.append(found)
        _golay12_parity = basis
    par = 0
    for k in range(12):
        if (word12 >> (11 - k)) & 1:
            par ^= _golay12_parity[k]
    return [(word12 >> (11 - i)) & 1 for i in range(12)] + [(par >> (11 - i)) & 1 for i in range(12)]


def p25p1_build_tdulc(rng, nac):
    """Terminator with link control: sync + NID(0xF) + 12 Golay(24,12) dodeca words (six data, six parity, each two RS hex
    symbols with the halves swapped) + ten null dibits."""
    data = rng.integers(0, 64, 12)
    par, dat = rs63_shortened_encode(data, 12)
    out = []
    for syms in (dat, par):
        for i in range(5, -1, -1):  # air order: word 5 first; hex 2i = bits 6..11, hex 2i+1 = bits 0..5
            out.append(_bits_to_dibits(golay_24_12_encode((int(syms[2 * i + 1]) << 6) | int(syms[2 * i]))))
    body = [np.array(P25P1_SYNC_DIBITS), p25p1_nid_dibits(nac, 0xF)] + out + [np.zeros(10, np.int64)]
    return p25p1_insert_status(np.concatenate(body)), {"rs_data": dat.astype(np.uint8), "duid": 0xF}


def p25p1_build_hdu(rng, nac):
    data = rng.integers(0, 64, 20)
    par, dat = rs63_shortened_encode(data, 16)
    out = []
    for w in range(36):
        sym = dat[19 - w] if w < 20 else par[35 - w]
        out.append(_bits_to_dibits(golay_24_6_encode(int(sym))))
    body = [np.array(P25P1_SYNC_DIBITS), p25p1_nid_dibits(nac, 0x0)] + out + [np.zeros(5, np.int64)]
    return p25p1_insert_status(np.concatenate(body)), {"rs_data": dat.astype(np.uint8), "duid": 0}


def synth_c4fm_iq(rng, dibits, snr_db=None, amp=0.6, pre=9, cu8=True):
    """P25 C4FM on a complex carrier at 48 kS/s, 10 samples per symbol: TIA-102 shaping (raised cosine x inverse sinc), +-1.8 kHz
    at the outer levels, phase-integrated, AWGN, optionally quantised like an RTL-SDR (cu8, the reference's --iq-replay
    format).  `pre` = 9 lines the symbol centres up with getSymbol's window behind the 135-tap channel LPF (67 samples), the
    discriminator and the 91-tap p25_filter (45 samples); found by scanning (the locked slicer never moves its window)."""
    imp = np.zeros(len(dibits) * 10)
    imp[::10] = LEVELS[np.asarray(dibits)]
    f = np.concatenate([np.zeros(pre), np.convolve(imp, c4fm_shaping_taps(), mode="same")])
    ph = 0.3 + np.cumsum(f * (2 * np.pi * 600.0 / 48000.0))
    z = amp * np.exp(1j * ph)
    if snr_db is not None:
        sigma = amp * 10 ** (-snr_db / 20.0) / np.sqrt(2.0)
        z = z + sigma * (rng.standard_normal(z.size) + 1j * rng.standard_normal(z.size))
    iq = np.stack([z.real, z.imag], axis=1)
    if cu8:
        return np.clip(np.rint(iq * 127.5 + 127.5), 0, 255).astype(np.uint8)
    return iq.astype(np.float32)


def widen_cu8(u8):
    """widen_u8_to_f32_bias127 (src/dsp/simd_widen.cpp:139-147)"""
    return ((u8.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).astype(np.float32)


# ---- acquisition (getFrameSync from the never-synchronised state) -------------------------------------------------------------

class OracleAcqPattern(C.Structure):
    _fields_ = [("symbols", C.c_char_p), ("sync_type", C.c_int), ("kind", C.c_int), ("use_filter", C.c_int),
                ("taps", C.POINTER(C.c_float)), ("taps_len", C.c_int), ("window_l", C.c_int), ("track_minmax", C.c_int),
                ("negative", C.c_int)]


class OracleAcqResult(C.Structure):
    _fields_ = [("sync_type", C.c_int), ("warm_start", C.c_int), ("resample_ok", C.c_int), ("hunt_symbols", C.c_long),
                ("consumed", C.c_long), ("lmin", C.c_float), ("lmax", C.c_float), ("resampled", C.c_uint8 * 66)]


P25P1_SYNC_STR, P25P1_SYNC_INV_STR = "111113113311333313133333", "333331331133111131311111"
DMR_BS_DATA_STR, DMR_BS_VOICE_STR = "313333111331131131331131", "131111333113313313113313"
DMR_MS_DATA_STR, DMR_MS_VOICE_STR = "311131133313133331131113", "133313311131311113313331"


def acquire_patterns(frame_p25p1, frame_dmr, taps, use_cosine_filter=True, inverted_dmr=False):
    """The reference's matcher order for the enabled protocols (frame_sync_try_protocol_matches, dsd_frame_sync.c:1636-1700:
    P25 Phase 1 before DMR; inside DMR: MS data, MS voice, BS data, BS voice) with the class each sync type gives the decoder.
    Returns (ctypes array, keep-alive list)."""
    pats, keep = [], []

    def add(sym, st, kind, filt, window_l, track, negative):
        tp = taps.get(filt) if (use_cosine_filter and filt is not None) else None
        arr = np.ascontiguousarray(tp, np.float32) if tp is not None else None
        keep.append(arr)
        pats.append(OracleAcqPattern(sym.encode(), st, kind, 1 if arr is not None else 0, _ptr(arr) if arr is not None else None,
                                     arr.size if arr is not None else 0, window_l, track, negative))

    if frame_p25p1:
        add(P25P1_SYNC_STR, 0, 0, 0, 2, 1, 0)
        add(P25P1_SYNC_INV_STR, 1, 0, 0, 2, 1, 1)
    if frame_dmr and not inverted_dmr:  # src/dsp/dsd_frame_sync.c:1108-1340
        add(DMR_MS_DATA_STR, 33, 1, 1, 1, 0, 0)
        add(DMR_MS_VOICE_STR, 32, 1, 1, 1, 0, 0)
        add(DMR_BS_DATA_STR, 10, 1, 1, 1, 0, 0)
        add(DMR_BS_VOICE_STR, 12, 1, 1, 1, 0, 0)
    elif frame_dmr:  # opts->inverted_dmr == 1 (-xr): the same patterns name the opposite burst kind, BS ones with negative polarity
        add(DMR_MS_DATA_STR, 32, 1, 1, 1, 0, 0)
        add(DMR_MS_VOICE_STR, 33, 1, 1, 1, 0, 0)
        add(DMR_BS_DATA_STR, 11, 1, 1, 1, 0, 1)
        add(DMR_BS_VOICE_STR, 13, 1, 1, 1, 0, 1)
    return (OracleAcqPattern * len(pats))(*pats), keep


# ---- the UNMODIFIED dmr_data_sync() on a replayed dibit stream (oracle/ref_shim_dmr.c -> oracle/_ref/libdsdneo_ref_dmr.so) ----
REF_DMR_DTYPE = np.dtype([("handler_called", "<i4"), ("cach_called", "<i4"), ("burst", "<i4"), ("color_code", "<i4"),
                          ("color_code_ok", "<i4"), ("dmr_color_code", "<i4"), ("currentslot", "<i4"), ("live_dibits", "<i4"),
                          ("info", "u1", (196,)), ("rel98", "u1", (98,)), ("cach", "u1", (25,)), ("stereo_payload", "u1", (144,)),
                          ("_pad", "u1", (1,))])
_ref_dmr_lib = None


def ref_dmr():
    global _ref_dmr_lib
    if _ref_dmr_lib is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libdsdneo_ref_dmr.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        assert L.ref_dmr_burst_size() == REF_DMR_DTYPE.itemsize
        L.ref_dmr_data_sync.restype = C.c_long
        L.ref_dmr_data_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p]
        _ref_dmr_lib = L
    return _ref_dmr_lib


def ref_dmr_bursts(dib, rel, sync_ends, inverted_dmr, first_is_raw=True):
    """dmr_data_sync on every burst of a CONTINUOUS, polarity-corrected dibit stream (what the device slicer produces after
    acquisition).  The reference reads the 90 dibits up to the sync from its raw rolling buffer and corrects their polarity
    itself; for bursts behind the first the stream is already corrected, so it is un-corrected here before the call.  The
    colour-code confidence gate persists over the bursts, as in the reference.  Returns a REF_DMR_DTYPE array."""
    L = ref_dmr()
    L.ref_dmr_reset(1 if inverted_dmr else 0)
    out = np.zeros(len(sync_ends), REF_DMR_DTYPE)
    rel = np.ascontiguousarray(rel, np.uint8)
    for k, p in enumerate(sync_ends):
        raw = np.ascontiguousarray(dib, np.uint8).copy()
        if inverted_dmr and not (k == 0 and first_is_raw):
            raw[p - 89:p + 1] ^= 2
        rc = L.ref_dmr_data_sync(raw.ctypes.data, rel.ctypes.data, raw.size, int(p), out[k:k + 1].ctypes.data)
        assert rc >= 0
    return out
