"""ctypes binding of libdsdneo_b200.so (the C-ABI declared in include/dsdneo_b200.h).

This is plumbing for tests and bench.py: PyTorch supplies device buffers and streams, the
library supplies every kernel.  There is no Python or CPU implementation of any stage here;
if the shared library is missing, or no sm_100 device is present, calls fail loudly.

The directory name contains a hyphen, so import it through `load_package()` in
`__graft_entry__.py` (module name `dsdneo_b200`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsdneo_b200.so")

OK, EINVAL, ENODEV, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4, -5

CH_LPF_PROFILE_WIDE, CH_LPF_PROFILE_6K25, CH_LPF_PROFILE_12K5 = 0, 1, 2
CH_LPF_PROFILE_PROVOICE, CH_LPF_PROFILE_P25_C4FM, CH_LPF_PROFILE_P25_CQPSK = 3, 4, 5
FIR_ARITH_FMA, FIR_ARITH_NOFMA = 0, 1
LPF_MAX_TAPS = 144


class B200Error(RuntimeError):
    pass


class DemodBankConfig(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int),
        ("rate_out_hz", C.c_int),
        ("channel_lpf_enable", C.c_int),
        ("channel_lpf_profile", C.POINTER(C.c_int)),
        ("channel_squelch_level", C.POINTER(C.c_float)),
        ("fir_arith", C.c_int),
    ]


class DemodChanState(C.Structure):
    _fields_ = [
        ("prev_i", C.c_float),
        ("prev_q", C.c_float),
        ("have_prev", C.c_int),
        ("dc_est", C.c_float),
        ("discriminator_peak_est", C.c_float),
        ("channel_pwr", C.c_float),
        ("channel_squelched", C.c_int),
    ]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)"
        )
    L = C.CDLL(LIB_PATH)
    vp, sz, ci, cf = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    L.dsdneo_b200_abi_version.restype = ci
    L.dsdneo_b200_last_error.restype = C.c_char_p
    L.dsdneo_b200_init.argtypes = [ci]
    L.dsdneo_b200_device_sm_count.restype = ci
    L.dsdneo_b200_stream_sync.argtypes = [vp]
    L.dsdneo_b200_launch_count.restype = C.c_ulonglong
    L.dsdneo_b200_malloc_device.restype = vp
    L.dsdneo_b200_malloc_device.argtypes = [sz]
    L.dsdneo_b200_free_device.argtypes = [vp]
    L.dsdneo_b200_malloc_pinned.restype = vp
    L.dsdneo_b200_malloc_pinned.argtypes = [sz]
    L.dsdneo_b200_free_pinned.argtypes = [vp]
    L.dsdneo_b200_memcpy_h2d.argtypes = [vp, vp, sz, vp]
    L.dsdneo_b200_memcpy_d2h.argtypes = [vp, vp, sz, vp]
    L.dsdneo_b200_channel_lpf_design.argtypes = [ci, ci, C.POINTER(cf), ci]
    L.dsdneo_b200_sps_fir_design.argtypes = [ci, vp, ci, ci, cf, ci, vp, ci]
    L.dsdneo_b200_demod_bank_create.restype = vp
    L.dsdneo_b200_demod_bank_create.argtypes = [C.POINTER(DemodBankConfig)]
    L.dsdneo_b200_demod_bank_destroy.argtypes = [vp]
    L.dsdneo_b200_demod_bank_reset.argtypes = [vp, vp]
    L.dsdneo_b200_demod_bank_get_state.argtypes = [vp, ci, C.POINTER(DemodChanState)]
    L.dsdneo_b200_demod_bank_get_taps.argtypes = [vp, ci, C.POINTER(cf), ci]
    L.dsdneo_b200_full_demod_batch.argtypes = [vp, vp, sz, ci, ci, vp, sz, vp]
    L.dsdneo_b200_full_demod_batch_cu8.argtypes = [vp, vp, sz, ci, ci, vp, sz, vp]
    L.dsdneo_b200_full_demod_batch_host.argtypes = [vp, vp, sz, ci, ci, vp, sz]
    _bind_optional(L)
    _lib = L
    return L


def _bind_optional(L: C.CDLL) -> None:
    """Prototypes for entry points added after ABI v1 bring-up (channelizer, symbol side, FEC)."""
    from . import _protos  # noqa: WPS433  (kept separate so this file stays readable)

    _protos.bind(L)


def last_error() -> str:
    return lib().dsdneo_b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise B200Error(f"{what or 'libdsdneo_b200'} failed ({rc}): {last_error()}")
    return rc


def init(device: int = 0) -> None:
    check(lib().dsdneo_b200_init(device), "dsdneo_b200_init")


def launch_count() -> int:
    return int(lib().dsdneo_b200_launch_count())


def channel_lpf_design(rate_out_hz: int, profile: int):
    import numpy as np

    buf = (C.c_float * LPF_MAX_TAPS)()
    n = check(lib().dsdneo_b200_channel_lpf_design(rate_out_hz, profile, buf, LPF_MAX_TAPS), "channel_lpf_design")
    return np.frombuffer(buf, dtype=np.float32, count=n).copy()


SPS_FIR_DESIGN_INTERP, SPS_FIR_DESIGN_RRC = 0, 1


def sps_fir_design(design_kind: int, base, base_sps: int, rrc_alpha: float, sps: int):
    """design_sps_fir() (src/dsp/dsd_filters.c:94-170): normalised matched-filter taps of the reference's coefficient table
    `base` (designed at `base_sps`) for `sps` samples per symbol."""
    import numpy as np

    b = np.ascontiguousarray(base, dtype=np.float32)
    out = np.zeros(1024, np.float32)
    n = check(lib().dsdneo_b200_sps_fir_design(design_kind, b.ctypes.data, b.size, base_sps, float(rrc_alpha), sps, out.ctypes.data,
                                               out.size), "sps_fir_design")
    return out[:n].copy()


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    return int(getattr(stream, "cuda_stream", stream))


class DemodBank:
    """N-channel twin of the reference's `struct demod_state` + full_demod() (FSK discriminator kind)."""

    def __init__(
        self,
        n_channels: int,
        rate_out_hz: int = 48000,
        channel_lpf_enable: bool = True,
        profiles: Optional[Sequence[int]] = None,
        squelch_levels: Optional[Sequence[float]] = None,
        fir_arith: int = FIR_ARITH_FMA,
    ):
        cfg = DemodBankConfig()
        cfg.n_channels = n_channels
        cfg.rate_out_hz = rate_out_hz
        cfg.channel_lpf_enable = 1 if channel_lpf_enable else 0
        self._prof = (C.c_int * n_channels)(*profiles) if profiles is not None else None
        self._sq = (C.c_float * n_channels)(*squelch_levels) if squelch_levels is not None else None
        cfg.channel_lpf_profile = self._prof if self._prof is not None else None
        cfg.channel_squelch_level = self._sq if self._sq is not None else None
        cfg.fir_arith = fir_arith
        self.n_channels = n_channels
        self._h = lib().dsdneo_b200_demod_bank_create(C.byref(cfg))
        if not self._h:
            raise B200Error(f"demod_bank_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_demod_bank_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None) -> None:
        check(lib().dsdneo_b200_demod_bank_reset(self._h, _stream_ptr(stream)), "demod_bank_reset")

    def state(self, channel: int) -> DemodChanState:
        st = DemodChanState()
        check(lib().dsdneo_b200_demod_bank_get_state(self._h, channel, C.byref(st)), "demod_bank_get_state")
        return st

    def taps(self, profile: int):
        import numpy as np

        buf = (C.c_float * LPF_MAX_TAPS)()
        n = check(lib().dsdneo_b200_demod_bank_get_taps(self._h, profile, buf, LPF_MAX_TAPS), "get_taps")
        return np.frombuffer(buf, dtype=np.float32, count=n).copy()

    def full_demod(self, d_iq, block_pairs: int, n_blocks: int, d_result=None, stream=None):
        """d_iq: torch cuda float32 tensor [n_channels, pitch_pairs, 2]; returns [n_channels, result_pitch] f32."""
        import torch

        assert d_iq.is_cuda and d_iq.dtype == torch.float32 and d_iq.is_contiguous()
        assert d_iq.shape[0] == self.n_channels and d_iq.shape[-1] == 2
        pitch = d_iq.shape[1]
        if d_result is None:
            d_result = torch.empty((self.n_channels, block_pairs * n_blocks), dtype=torch.float32, device=d_iq.device)
        assert d_result.is_cuda and d_result.dtype == torch.float32 and d_result.is_contiguous()
        if stream is None:
            stream = torch.cuda.current_stream(d_iq.device)
        check(
            lib().dsdneo_b200_full_demod_batch(
                self._h, d_iq.data_ptr(), pitch, block_pairs, n_blocks, d_result.data_ptr(), d_result.shape[1],
                _stream_ptr(stream),
            ),
            "full_demod_batch",
        )
        return d_result

    def full_demod_cu8(self, d_iq_u8, block_pairs: int, n_blocks: int, d_result=None, stream=None):
        """d_iq_u8: torch cuda uint8 tensor [n_channels, pitch_pairs, 2] (widened on the device); returns [n_channels, n] f32."""
        import torch

        assert d_iq_u8.is_cuda and d_iq_u8.dtype == torch.uint8 and d_iq_u8.is_contiguous()
        assert d_iq_u8.shape[0] == self.n_channels and d_iq_u8.shape[-1] == 2
        if d_result is None:
            d_result = torch.empty((self.n_channels, block_pairs * n_blocks), dtype=torch.float32, device=d_iq_u8.device)
        if stream is None:
            stream = torch.cuda.current_stream(d_iq_u8.device)
        check(lib().dsdneo_b200_full_demod_batch_cu8(self._h, d_iq_u8.data_ptr(), d_iq_u8.shape[1], block_pairs, n_blocks,
                                                     d_result.data_ptr(), d_result.shape[1], _stream_ptr(stream)), "full_demod_batch_cu8")
        return d_result

    def full_demod_host(self, h_iq, block_pairs: int, n_blocks: int):
        """h_iq: numpy float32 [n_channels, pitch_pairs, 2] (C order). Returns numpy [n_channels, n] f32."""
        import numpy as np

        h_iq = np.ascontiguousarray(h_iq, dtype=np.float32)
        assert h_iq.shape[0] == self.n_channels and h_iq.shape[-1] == 2
        out = np.empty((self.n_channels, block_pairs * n_blocks), dtype=np.float32)
        check(
            lib().dsdneo_b200_full_demod_batch_host(
                self._h, h_iq.ctypes.data, h_iq.shape[1], block_pairs, n_blocks, out.ctypes.data, out.shape[1]
            ),
            "full_demod_batch_host",
        )
        return out


class CqpskBankConfig(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int),
        ("rate_out_hz", C.c_int),
        ("ted_sps", C.POINTER(C.c_int)),
        ("ted_gain", C.c_float),
        ("ted_gain_is_set", C.c_int),
    ]


class CqpskChanState(C.Structure):
    _fields_ = [
        ("cqpsk_agc_avg", C.c_float),
        ("fll_phase", C.c_float), ("fll_freq", C.c_float), ("fll_alpha", C.c_float), ("fll_beta", C.c_float),
        ("ted_mu", C.c_float), ("ted_omega", C.c_float), ("ted_last_r", C.c_float), ("ted_last_j", C.c_float),
        ("ted_lock_accum", C.c_float), ("ted_lock_count", C.c_int), ("ted_effective_gain", C.c_float),
        ("cqpsk_diff_prev_r", C.c_float), ("cqpsk_diff_prev_j", C.c_float),
        ("costas_phase", C.c_float), ("costas_freq", C.c_float), ("costas_error", C.c_float),
        ("costas_error_smooth", C.c_float),
        ("costas_err_avg_q14", C.c_int), ("costas_err_raw_avg_q14", C.c_int), ("costas_conf_avg_q14", C.c_int),
        ("costas_zero_conf_pct", C.c_int), ("overflow", C.c_int),
    ]


class CqpskBank:
    """N-channel twin of full_demod() for the CQPSK symbol output kind (AGC -> FLL -> Gardner -> diff -> Costas ->
    4/pi atan): owns the channel-LPF bank (profile P25_CQPSK by default) and the per-channel loop state."""

    def __init__(self, n_channels: int, rate_out_hz: int = 24000, ted_sps=None, channel_lpf_enable: bool = True,
                 profiles=None, squelch_levels=None, fir_arith: int = FIR_ARITH_FMA, ted_gain: float = 0.0,
                 ted_gain_is_set: bool = False):
        self.lpf = DemodBank(n_channels, rate_out_hz, channel_lpf_enable,
                             profiles if profiles is not None else [5] * n_channels, squelch_levels, fir_arith)
        cfg = CqpskBankConfig()
        cfg.n_channels = n_channels
        cfg.rate_out_hz = rate_out_hz
        self._sps = (C.c_int * n_channels)(*ted_sps) if ted_sps is not None else None
        cfg.ted_sps = self._sps if self._sps is not None else None
        cfg.ted_gain = ted_gain
        cfg.ted_gain_is_set = 1 if ted_gain_is_set else 0
        self.n_channels = n_channels
        self.min_sps = min(ted_sps) if ted_sps is not None else 5
        self._h = lib().dsdneo_b200_cqpsk_bank_create(C.byref(cfg))
        if not self._h:
            raise B200Error(f"cqpsk_bank_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_cqpsk_bank_destroy(self._h)
            self._h = None
        if getattr(self, "lpf", None):
            self.lpf.close()

    __del__ = close

    def reset(self, stream=None) -> None:
        self.lpf.reset(stream)
        check(lib().dsdneo_b200_cqpsk_bank_reset(self._h, _stream_ptr(stream)), "cqpsk_bank_reset")

    def state(self, channel: int) -> CqpskChanState:
        st = CqpskChanState()
        check(lib().dsdneo_b200_cqpsk_bank_get_state(self._h, channel, C.byref(st)), "cqpsk_bank_get_state")
        return st

    def fll_taps(self, channel: int):
        import numpy as np

        bufs = [np.zeros(48, np.float32) for _ in range(4)]
        n = check(lib().dsdneo_b200_cqpsk_bank_get_fll_taps(self._h, channel, *[b.ctypes.data for b in bufs], 48), "fll_taps")
        return [b[:n].copy() for b in bufs]

    def block_capacity(self, block_pairs: int) -> int:
        return check(lib().dsdneo_b200_cqpsk_block_capacity(block_pairs, self.min_sps), "cqpsk_block_capacity")

    def full_demod(self, d_iq, block_pairs: int, n_blocks: int, stream=None):
        """d_iq: torch cuda float32 [n_channels, pitch_pairs, 2].  Returns (symbols [n_channels, pitch] f32 with the
        blocks' symbols back to back, counts [n_channels, n_blocks] int32)."""
        import torch

        assert d_iq.is_cuda and d_iq.dtype == torch.float32 and d_iq.is_contiguous()
        assert d_iq.shape[0] == self.n_channels and d_iq.shape[-1] == 2
        pitch = self.block_capacity(block_pairs) * n_blocks
        sym = torch.zeros((self.n_channels, pitch), dtype=torch.float32, device=d_iq.device)
        counts = torch.zeros((self.n_channels, n_blocks), dtype=torch.int32, device=d_iq.device)
        if stream is None:
            stream = torch.cuda.current_stream(d_iq.device)
        check(
            lib().dsdneo_b200_full_demod_cqpsk_batch(
                self.lpf._h, self._h, d_iq.data_ptr(), d_iq.shape[1], block_pairs, n_blocks, sym.data_ptr(), pitch,
                counts.data_ptr(), _stream_ptr(stream),
            ),
            "full_demod_cqpsk_batch",
        )
        return sym, counts

    def full_demod_host(self, h_iq, block_pairs: int, n_blocks: int):
        import numpy as np

        h_iq = np.ascontiguousarray(h_iq, dtype=np.float32)
        assert h_iq.shape[0] == self.n_channels and h_iq.shape[-1] == 2
        pitch = self.block_capacity(block_pairs) * n_blocks
        sym = np.zeros((self.n_channels, pitch), np.float32)
        counts = np.zeros((self.n_channels, n_blocks), np.int32)
        check(
            lib().dsdneo_b200_full_demod_cqpsk_batch_host(
                self.lpf._h, self._h, h_iq.ctypes.data, h_iq.shape[1], block_pairs, n_blocks, sym.ctypes.data, pitch,
                counts.ctypes.data,
            ),
            "full_demod_cqpsk_batch_host",
        )
        return sym, counts


class CqpskSlicer:
    """Symbol-rate CQPSK sample side (tracker + CQPSK slicer + soft metrics) behind CqpskBank.full_demod."""

    def __init__(self, n_channels: int, ssize: int = 128, msize: int = 1024):
        self.n_channels = n_channels
        self._h = lib().dsdneo_b200_cqpsk_slicer_create(n_channels, ssize, msize)
        if not self._h:
            raise B200Error(f"cqpsk_slicer_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:
            lib().dsdneo_b200_cqpsk_slicer_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None) -> None:
        check(lib().dsdneo_b200_cqpsk_slicer_reset(self._h, _stream_ptr(stream)), "cqpsk_slicer_reset")

    def set_class(self, negative=None, p25_slice=None, map_idx=None, snr_cqpsk_db: float = -100.0) -> None:
        import numpy as np

        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.uint8) for a in (negative, p25_slice, map_idx)]
        check(lib().dsdneo_b200_cqpsk_slicer_set_class(self._h, *[None if a is None else a.ctypes.data for a in arrs],
                                                       snr_cqpsk_db), "cqpsk_slicer_set_class")

    def run(self, d_symbols, d_n_symbols, stream=None):
        """d_symbols cuda f32 [n_channels, pitch], d_n_symbols cuda int32 [n_channels] -> dict(dibits, reliability, llr)."""
        import torch

        assert d_symbols.is_cuda and d_symbols.dtype == torch.float32 and d_symbols.is_contiguous()
        assert d_n_symbols.dtype == torch.int32 and d_n_symbols.is_contiguous()
        pitch = d_symbols.shape[1]
        dev = d_symbols.device
        out = {"dibits": torch.zeros((self.n_channels, pitch), dtype=torch.uint8, device=dev),
               "reliability": torch.zeros((self.n_channels, pitch), dtype=torch.uint8, device=dev),
               "llr": torch.zeros((self.n_channels, pitch, 2), dtype=torch.int16, device=dev)}
        if stream is None:
            stream = torch.cuda.current_stream(dev)
        check(lib().dsdneo_b200_cqpsk_slice_batch(self._h, d_symbols.data_ptr(), pitch, d_n_symbols.data_ptr(),
                                                  out["dibits"].data_ptr(), out["reliability"].data_ptr(), out["llr"].data_ptr(),
                                                  pitch, _stream_ptr(stream)), "cqpsk_slice_batch")
        return out

    def state(self, channel: int):
        import numpy as np

        out = np.zeros(8, np.float32)
        check(lib().dsdneo_b200_cqpsk_slicer_get_state(self._h, channel, out.ctypes.data), "cqpsk_slicer_get_state")
        return out


class SyncPattern(C.Structure):
    _fields_ = [("symbols", C.c_char_p), ("sync_type", C.c_int)]


class SyncHit(C.Structure):
    _fields_ = [("position", C.c_int), ("sync_type", C.c_int)]


# include/dsd-neo/core/sync_patterns.h:33-67 with the ids of include/dsd-neo/core/synctype_ids.h (non-inverted DMR setting)
DEFAULT_SYNC_PATTERNS = [
    ("111113113311333313133333", 0),   # P25P1_SYNC            -> DSD_SYNC_P25P1_POS
    ("333331331133111131311111", 1),   # INV_P25P1_SYNC        -> DSD_SYNC_P25P1_NEG
    ("313333111331131131331131", 10),  # DMR_BS_DATA_SYNC      -> DSD_SYNC_DMR_BS_DATA_POS
    ("131111333113313313113313", 12),  # DMR_BS_VOICE_SYNC     -> DSD_SYNC_DMR_BS_VOICE_POS
    ("311131133313133331131113", 33),  # DMR_MS_DATA_SYNC      -> DSD_SYNC_DMR_MS_DATA
    ("133313311131311113313331", 32),  # DMR_MS_VOICE_SYNC     -> DSD_SYNC_DMR_MS_VOICE
    ("331313111113131113331133", 2),   # X2TDMA_BS_DATA_SYNC   -> DSD_SYNC_X2TDMA_DATA_POS
    ("113131333331313331113311", 4),   # X2TDMA_BS_VOICE_SYNC  -> DSD_SYNC_X2TDMA_VOICE_POS
    ("31111311313113131131", 30),      # FUSION_SYNC           -> DSD_SYNC_YSF_POS
    ("13333133131331313313", 31),      # INV_FUSION_SYNC       -> DSD_SYNC_YSF_NEG
]


class FrameSync:
    """Batched twin of the per-symbol sync hunt of getFrameSync() (dsd_frame_sync.c:3098-3148)."""

    def __init__(self, n_channels: int, patterns=None):
        patterns = DEFAULT_SYNC_PATTERNS if patterns is None else patterns
        self._keep = [p.encode() if isinstance(p, str) else p for p, _ in patterns]
        arr = (SyncPattern * len(patterns))(*[SyncPattern(k, t) for k, (_, t) in zip(self._keep, patterns)])
        self.n_channels = n_channels
        self._h = lib().dsdneo_b200_frame_sync_create(n_channels, arr, len(patterns))
        if not self._h:
            raise B200Error(f"frame_sync_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_frame_sync_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None) -> None:
        check(lib().dsdneo_b200_frame_sync_reset(self._h, _stream_ptr(stream)), "frame_sync_reset")

    def search(self, d_symbols, d_n_symbols, max_hits: int = 64, stream=None):
        """d_symbols: cuda f32 [n_channels, pitch]; d_n_symbols: cuda int32 [n_channels].
        Returns (hits int32 [n_channels, max_hits, 2] = (position, sync_type), n_hits int32 [n_channels])."""
        import torch

        assert d_symbols.is_cuda and d_symbols.dtype == torch.float32 and d_symbols.is_contiguous()
        assert d_n_symbols.is_cuda and d_n_symbols.dtype == torch.int32
        hits = torch.zeros((self.n_channels, max_hits, 2), dtype=torch.int32, device=d_symbols.device)
        n_hits = torch.zeros(self.n_channels, dtype=torch.int32, device=d_symbols.device)
        if stream is None:
            stream = torch.cuda.current_stream(d_symbols.device)
        check(lib().dsdneo_b200_frame_sync_search_batch(self._h, d_symbols.data_ptr(), d_symbols.shape[1], d_n_symbols.data_ptr(),
                                                        hits.data_ptr(), max_hits, n_hits.data_ptr(), _stream_ptr(stream)),
              "frame_sync_search_batch")
        return hits, n_hits


class HalfbandCascade:
    """N-channel twin of full_demod_apply_halfband_decimation (demod_pipeline.cpp:983-1001): `passes` half-band /2 stages."""

    def __init__(self, n_channels: int, passes: int, fir_arith: int = FIR_ARITH_FMA):
        self.n_channels, self.passes = n_channels, passes
        self._h = lib().dsdneo_b200_hb_cascade_create(n_channels, passes, fir_arith)
        if not self._h:
            raise B200Error(f"hb_cascade_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_hb_cascade_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None) -> None:
        check(lib().dsdneo_b200_hb_cascade_reset(self._h, _stream_ptr(stream)), "hb_cascade_reset")

    def decimate(self, d_in, block_pairs: int, n_blocks: int, d_out=None, stream=None):
        """d_in: cuda f32 [n_channels, pitch_pairs, 2]; returns [n_channels, (block_pairs >> passes) * n_blocks, 2]."""
        import torch

        assert d_in.is_cuda and d_in.dtype == torch.float32 and d_in.is_contiguous() and d_in.shape[0] == self.n_channels
        n_out = (block_pairs >> self.passes) * n_blocks
        if d_out is None:
            d_out = torch.empty((self.n_channels, n_out, 2), dtype=torch.float32, device=d_in.device)
        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(lib().dsdneo_b200_hb_cascade_decim_batch(self._h, d_in.data_ptr(), d_in.shape[1], block_pairs, n_blocks,
                                                       d_out.data_ptr(), d_out.shape[1], _stream_ptr(stream)), "hb_cascade_decim_batch")
        return d_out

    def decimate_host(self, h_in, block_pairs: int, n_blocks: int):
        import numpy as np

        h_in = np.ascontiguousarray(h_in, dtype=np.float32)
        out = np.empty((self.n_channels, (block_pairs >> self.passes) * n_blocks, 2), dtype=np.float32)
        check(lib().dsdneo_b200_hb_cascade_decim_batch_host(self._h, h_in.ctypes.data, h_in.shape[1], block_pairs, n_blocks,
                                                            out.ctypes.data, out.shape[1]), "hb_cascade_decim_batch_host")
        return out


class Channelizer:
    """Polyphase FIR channelizer (K2, optionally fused K1 cu8 widening). See include/dsdneo_b200.h."""

    def __init__(self, n_channels: int = 256, taps_per_branch: int = 8, input_is_cu8: bool = False, prototype=None):
        import numpy as np

        self.M, self.T, self.cu8 = n_channels, taps_per_branch, bool(input_is_cu8)
        proto_ptr = None
        if prototype is not None:
            self._proto = np.ascontiguousarray(prototype, dtype=np.float32)
            assert self._proto.size == n_channels * taps_per_branch
            proto_ptr = self._proto.ctypes.data_as(C.POINTER(C.c_float))
        self._h = lib().dsdneo_b200_channelizer_create(n_channels, taps_per_branch, 1 if input_is_cu8 else 0, proto_ptr)
        if not self._h:
            raise B200Error(f"channelizer_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_channelizer_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None) -> None:
        check(lib().dsdneo_b200_channelizer_reset(self._h, _stream_ptr(stream)), "channelizer_reset")

    def prototype(self):
        import numpy as np

        buf = np.empty(self.M * self.T, dtype=np.float32)
        check(lib().dsdneo_b200_channelizer_get_prototype(self._h, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size))
        return buf

    def channelize(self, d_in, d_out=None, stream=None):
        """d_in: cuda tensor, float32 [n, 2] (cf32) or uint8 [n, 2] (cu8); returns float32 [M, n/M, 2]."""
        import torch

        assert d_in.is_cuda and d_in.is_contiguous() and d_in.shape[-1] == 2
        assert d_in.dtype == (torch.uint8 if self.cu8 else torch.float32)
        n = d_in.shape[0]
        if d_out is None:
            d_out = torch.empty((self.M, n // self.M, 2), dtype=torch.float32, device=d_in.device)
        assert d_out.is_cuda and d_out.is_contiguous() and d_out.dtype == torch.float32
        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(
            lib().dsdneo_b200_channelize(self._h, d_in.data_ptr(), n, d_out.data_ptr(), d_out.shape[1], _stream_ptr(stream)),
            "channelize",
        )
        return d_out

    def prime(self, d_in, stream=None) -> None:
        """Carry the history over d_in ([n, 2] in the input format, on the GPU) without producing output."""
        import torch

        assert d_in.is_cuda and d_in.is_contiguous() and d_in.dtype == (torch.uint8 if self.cu8 else torch.float32)
        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(lib().dsdneo_b200_channelizer_prime(self._h, d_in.data_ptr(), d_in.shape[0], _stream_ptr(stream)), "channelizer_prime")

    def channelize_bins(self, d_in, bin_stride: int, bin_first: int, d_out=None, advance: bool = True, stream=None):
        """Only the channels k = bin_first (mod bin_stride): returns float32 [M / bin_stride, n/M, 2], row k' = channel
        bin_stride * k' + bin_first (the per-GPU share of one broadcast wideband tile)."""
        import torch

        assert d_in.is_cuda and d_in.is_contiguous() and d_in.shape[-1] == 2
        assert d_in.dtype == (torch.uint8 if self.cu8 else torch.float32)
        n = d_in.shape[0]
        if d_out is None:
            d_out = torch.empty((self.M // bin_stride, n // self.M, 2), dtype=torch.float32, device=d_in.device)
        assert d_out.is_cuda and d_out.is_contiguous() and d_out.dtype == torch.float32 and d_out.shape[0] == self.M // bin_stride
        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(
            lib().dsdneo_b200_channelize_bins(self._h, d_in.data_ptr(), n, bin_stride, bin_first, 1 if advance else 0,
                                              d_out.data_ptr(), d_out.shape[1], _stream_ptr(stream)),
            "channelize_bins",
        )
        return d_out

    def channelize_bins_cu8(self, d_in, bin_stride: int, bin_first: int, gain: float, d_out=None, advance: bool = True, stream=None):
        """channelize_bins with the rows re-quantised to cu8 (uint8 [M / bin_stride, n/M, 2]): the receive bank's native input."""
        import torch

        assert d_in.is_cuda and d_in.is_contiguous() and d_in.shape[-1] == 2
        assert d_in.dtype == (torch.uint8 if self.cu8 else torch.float32)
        n = d_in.shape[0]
        if d_out is None:
            d_out = torch.empty((self.M // bin_stride, n // self.M, 2), dtype=torch.uint8, device=d_in.device)
        assert d_out.is_cuda and d_out.is_contiguous() and d_out.dtype == torch.uint8 and d_out.shape[0] == self.M // bin_stride
        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(
            lib().dsdneo_b200_channelize_bins_cu8(self._h, d_in.data_ptr(), n, bin_stride, bin_first, 1 if advance else 0, float(gain),
                                                  d_out.data_ptr(), d_out.shape[1], _stream_ptr(stream)),
            "channelize_bins_cu8",
        )
        return d_out

    def channelize_host(self, h_in):
        import numpy as np

        h_in = np.ascontiguousarray(h_in, dtype=np.uint8 if self.cu8 else np.float32)
        n = h_in.shape[0]
        out = np.empty((self.M, n // self.M, 2), dtype=np.float32)
        check(lib().dsdneo_b200_channelize_host(self._h, h_in.ctypes.data, n, out.ctypes.data, out.shape[1]), "channelize_host")
        return out


class FrontendConfig(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int),
        ("taps_per_branch", C.c_int),
        ("input_is_cu8", C.c_int),
        ("wideband_rate_hz", C.c_int),
        ("block_pairs", C.c_int),
        ("prototype", C.POINTER(C.c_float)),
        ("channel_lpf_enable", C.c_int),
        ("channel_lpf_profile", C.POINTER(C.c_int)),
        ("channel_squelch_level", C.POINTER(C.c_float)),
        ("fir_arith", C.c_int),
    ]


def timing_enable(on: bool) -> None:
    check(lib().dsdneo_b200_timing_enable(1 if on else 0))


def timing_report() -> dict:
    import json

    buf = C.create_string_buffer(16384)
    check(lib().dsdneo_b200_timing_report(buf, len(buf)))
    return json.loads(buf.value.decode())


class Frontend:
    """Wideband IQ -> per-channel discriminator samples (channelizer + full_demod) behind one C-ABI call."""

    def __init__(self, n_channels=256, taps_per_branch=8, input_is_cu8=False, wideband_rate_hz=12_288_000,
                 block_pairs=8192, channel_lpf_enable=True, profiles=None, squelch_levels=None, fir_arith=FIR_ARITH_FMA):
        cfg = FrontendConfig()
        cfg.n_channels = n_channels
        cfg.taps_per_branch = taps_per_branch
        cfg.input_is_cu8 = 1 if input_is_cu8 else 0
        cfg.wideband_rate_hz = wideband_rate_hz
        cfg.block_pairs = block_pairs
        cfg.prototype = None
        cfg.channel_lpf_enable = 1 if channel_lpf_enable else 0
        self._prof = (C.c_int * n_channels)(*profiles) if profiles is not None else None
        self._sq = (C.c_float * n_channels)(*squelch_levels) if squelch_levels is not None else None
        cfg.channel_lpf_profile = self._prof if self._prof is not None else None
        cfg.channel_squelch_level = self._sq if self._sq is not None else None
        cfg.fir_arith = fir_arith
        self.M, self.cu8, self.block_pairs = n_channels, bool(input_is_cu8), block_pairs
        self._h = lib().dsdneo_b200_frontend_create(C.byref(cfg))
        if not self._h:
            raise B200Error(f"frontend_create failed: {last_error()}")

    def close(self) -> None:
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_frontend_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None) -> None:
        check(lib().dsdneo_b200_frontend_reset(self._h, _stream_ptr(stream)), "frontend_reset")

    def process(self, d_in, d_result=None, stream=None):
        import torch

        assert d_in.is_cuda and d_in.is_contiguous() and d_in.shape[-1] == 2
        assert d_in.dtype == (torch.uint8 if self.cu8 else torch.float32)
        n = d_in.shape[0]
        if d_result is None:
            d_result = torch.empty((self.M, n // self.M), dtype=torch.float32, device=d_in.device)
        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(lib().dsdneo_b200_frontend_process(self._h, d_in.data_ptr(), n, d_result.data_ptr(), d_result.shape[1],
                                                  _stream_ptr(stream)), "frontend_process")
        return d_result

    def process_async(self, d_in, d_result, stream=None):
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(d_in.device)
        check(lib().dsdneo_b200_frontend_process_async(self._h, d_in.data_ptr(), d_in.shape[0], d_result.data_ptr(),
                                                        d_result.shape[1], _stream_ptr(stream)), "frontend_process_async")
        return d_result

    def join(self, stream=None) -> None:
        import torch

        if stream is None:
            stream = torch.cuda.current_stream()
        check(lib().dsdneo_b200_frontend_join(self._h, _stream_ptr(stream)), "frontend_join")

    def process_host(self, h_in, h_out=None):
        """h_in / h_out: numpy arrays or pinned torch CPU tensors."""
        import numpy as np

        n = h_in.shape[0]
        if h_out is None:
            h_out = np.empty((self.M, n // self.M), dtype=np.float32)
        pin = h_in.data_ptr() if hasattr(h_in, "data_ptr") else h_in.ctypes.data
        pout = h_out.data_ptr() if hasattr(h_out, "data_ptr") else h_out.ctypes.data
        check(lib().dsdneo_b200_frontend_process_host(self._h, pin, n, pout, h_out.shape[1]), "frontend_process_host")
        return h_out

    def submit_host(self, h_in, h_out) -> int:
        """Streaming form: queue one tile, return its ticket; both buffers must stay untouched until wait_host(ticket)."""
        pin = h_in.data_ptr() if hasattr(h_in, "data_ptr") else h_in.ctypes.data
        pout = h_out.data_ptr() if hasattr(h_out, "data_ptr") else h_out.ctypes.data
        t = lib().dsdneo_b200_frontend_submit_host(self._h, pin, h_in.shape[0], pout, h_out.shape[1])
        if t < 0:
            check(int(t), "frontend_submit_host")
        return int(t)

    def wait_host(self, ticket: int) -> None:
        check(lib().dsdneo_b200_frontend_wait_host(self._h, ticket), "frontend_wait_host")


# ------------------------------------------------------------------------------------------ FEC (host-buffer entry points)

FEC_HAMMING_7_4, FEC_HAMMING_12_8, FEC_HAMMING_13_9, FEC_HAMMING_15_11, FEC_HAMMING_16_11_4 = 0, 1, 2, 3, 4
FEC_GOLAY_20_8, FEC_GOLAY_24_12, FEC_QR_16_7_6 = 5, 6, 7
P25_RS_36_20_17, P25_RS_24_12_13, P25_RS_24_16_9 = 0, 1, 2


class P25Candidate(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 12), ("metric", C.c_uint32)]


def fec_block_decode(code: int, bits, decoded=None):
    """bits: uint8 [n_words, n] corrected in place (numpy). Returns ok[n_words] uint8."""
    import numpy as np

    assert bits.dtype == np.uint8 and bits.flags.c_contiguous
    n_words = bits.shape[0]
    ok = np.zeros(n_words, dtype=np.uint8)
    dptr = decoded.ctypes.data if decoded is not None else None
    check(lib().dsdneo_b200_fec_block_decode_batch_host(code, bits.ctypes.data, dptr, ok.ctypes.data, n_words), "fec_block_decode")
    return ok


def bptc_196x96(bursts, interleaved: bool):
    import numpy as np

    bursts = np.ascontiguousarray(bursts, dtype=np.uint8)
    n = bursts.shape[0]
    out, r3, errs = np.zeros((n, 96), np.uint8), np.zeros((n, 3), np.uint8), np.zeros(n, np.uint32)
    check(lib().dsdneo_b200_bptc_196x96_batch_host(bursts.ctypes.data, 1 if interleaved else 0, out.ctypes.data, r3.ctypes.data,
                                                   errs.ctypes.data, n), "bptc_196x96")
    return out, r3, errs


def bptc_128x77(mats):
    """mats: uint8 [n, 128] (8 x 16 byte-per-bit matrices). Returns (out77 [n, 77], errs [n])."""
    import numpy as np

    mats = np.ascontiguousarray(mats, dtype=np.uint8).reshape(-1, 128)
    n = mats.shape[0]
    out, errs = np.zeros((n, 77), np.uint8), np.zeros(n, np.uint32)
    check(lib().dsdneo_b200_bptc_128x77_batch_host(mats.ctypes.data, out.ctypes.data, errs.ctypes.data, n), "bptc_128x77")
    return out, errs


def bptc_16x2(words, parity_odd: bool):
    """words: uint8 [n, 32] interleaved bits. Returns (out32 [n, 32], errs [n])."""
    import numpy as np

    words = np.ascontiguousarray(words, dtype=np.uint8).reshape(-1, 32)
    n = words.shape[0]
    out, errs = np.zeros((n, 32), np.uint8), np.zeros(n, np.uint32)
    check(lib().dsdneo_b200_bptc_16x2_batch_host(words.ctypes.data, out.ctypes.data, errs.ctypes.data, 1 if parity_odd else 0, n),
          "bptc_16x2")
    return out, errs


def p25_12_soft_llr(llr):
    import numpy as np

    llr = np.ascontiguousarray(llr, dtype=np.int16)
    n = llr.shape[0]
    out, met = np.zeros((n, 12), np.uint8), np.zeros(n, np.int32)
    check(lib().dsdneo_b200_p25_12_soft_llr_batch_host(llr.ctypes.data, out.ctypes.data, met.ctypes.data, n), "p25_12_soft_llr")
    return out, met


def p25_12_soft_llr_list(llr, max_candidates=8):
    import numpy as np

    llr = np.ascontiguousarray(llr, dtype=np.int16)
    n = llr.shape[0]
    cands = (P25Candidate * (8 * n))()
    cnt = np.zeros(n, np.int32)
    check(lib().dsdneo_b200_p25_12_soft_llr_list_batch_host(llr.ctypes.data, C.addressof(cands), cnt.ctypes.data, max_candidates, n),
          "p25_12_soft_llr_list")
    return cands, cnt


def p25_rs_decode(variant: int, data_bits, parity_bits):
    """data_bits uint8 [n, k*6] corrected in place; returns status[n]."""
    import numpy as np

    assert data_bits.dtype == np.uint8 and data_bits.flags.c_contiguous
    parity_bits = np.ascontiguousarray(parity_bits, dtype=np.uint8)
    n = data_bits.shape[0]
    st = np.zeros(n, np.uint8)
    check(lib().dsdneo_b200_p25_rs_decode_batch_host(variant, data_bits.ctypes.data, parity_bits.ctypes.data, st.ctypes.data, n),
          "p25_rs_decode")
    return st


# ------------------------------------------------------------------------------------------ sample side

SYM_MODE_GET_SYMBOL, SYM_MODE_GET_DIBIT_SOFT = 0, 1
SYM_FILTER_NONE, SYM_FILTER_P25, SYM_FILTER_DMR, SYM_FILTER_NXDN, SYM_FILTER_DPMR, SYM_FILTER_M17 = -1, 0, 1, 2, 3, 4


def p25_rs_decode_erasures(variant: int, data_bits, parity_bits, erasures, n_erasures):
    """check_and_fix_*_soft twin. erasures: int32 [n, pitch]; n_erasures: int32 [n]. Returns (data_bits, status)."""
    import numpy as np

    data_bits = np.ascontiguousarray(data_bits, dtype=np.uint8).copy()
    parity_bits = np.ascontiguousarray(parity_bits, dtype=np.uint8)
    erasures = np.ascontiguousarray(erasures, dtype=np.int32)
    n_erasures = np.ascontiguousarray(n_erasures, dtype=np.int32)
    n = data_bits.shape[0]
    status = np.zeros(n, np.uint8)
    check(lib().dsdneo_b200_p25_rs_decode_erasures_batch_host(variant, data_bits.ctypes.data, parity_bits.ctypes.data,
                                                              erasures.ctypes.data, erasures.shape[1], n_erasures.ctypes.data,
                                                              status.ctypes.data, n), "p25_rs_decode_erasures")
    return data_bits, status


def p25_rs_soft_reliability(variant: int, data_bits, parity_bits, data_reliab, parity_reliab, threshold: int = 64):
    """p25p1_rs_*_soft_reliability twin. Returns (data_bits, status)."""
    import numpy as np

    data_bits = np.ascontiguousarray(data_bits, dtype=np.uint8).copy()
    parity_bits = np.ascontiguousarray(parity_bits, dtype=np.uint8)
    data_reliab = np.ascontiguousarray(data_reliab, dtype=np.uint8)
    parity_reliab = np.ascontiguousarray(parity_reliab, dtype=np.uint8)
    n = data_bits.shape[0]
    status = np.zeros(n, np.uint8)
    check(lib().dsdneo_b200_p25_rs_soft_reliability_batch_host(variant, data_bits.ctypes.data, parity_bits.ctypes.data,
                                                               data_reliab.ctypes.data, parity_reliab.ctypes.data, threshold,
                                                               status.ctypes.data, n), "p25_rs_soft_reliability")
    return data_bits, status


def symbol_capture_pack(dibits, reliability, llr, symbols, with_header: bool = True) -> bytes:
    """One channel's symbolizer output -> the reference's DSDNSYM2 capture bytes (see include/dsdneo_b200.h)."""
    import numpy as np

    d = np.ascontiguousarray(dibits, np.uint8)
    r = np.ascontiguousarray(reliability, np.uint8)
    l = np.ascontiguousarray(llr, np.int16).reshape(-1, 2)
    s = np.ascontiguousarray(symbols, np.float32)
    n = d.size
    out = np.zeros(lib().dsdneo_b200_symbol_capture_size(n, 1 if with_header else 0), np.uint8)
    check(lib().dsdneo_b200_symbol_capture_pack(d.ctypes.data, r.ctypes.data, l.ctypes.data, s.ctypes.data, n, 1 if with_header else 0,
                                                out.ctypes.data), "symbol_capture_pack")
    return out.tobytes()


def symbol_capture_unpack(data: bytes):
    """DSDNSYM2 bytes (with or without header) -> (dibits u8, reliability u8, llr i16 [n,2], symbols f32)."""
    import numpy as np

    buf = np.frombuffer(data, np.uint8)
    cap = buf.size // 10 + 1
    d, r, l, s = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8), np.zeros((cap, 2), np.int16), np.zeros(cap, np.float32)
    n = lib().dsdneo_b200_symbol_capture_unpack(buf.ctypes.data, buf.size, d.ctypes.data, r.ctypes.data, l.ctypes.data, s.ctypes.data, cap)
    if n < 0:
        check(int(n), "symbol_capture_unpack")
    return d[:n], r[:n], l[:n], s[:n]


P25_WORD_GOLAY_24_6, P25_WORD_GOLAY_24_12, P25_WORD_HAMMING_10_6_3 = 0, 1, 2


def p25p1_nid_decode(code63, reliab63=None, observed_nac=None, parity=None, parity_reliab=None, threshold: int = 64):
    """Batched p25p1_nid_decode on host arrays: code63 [n, 63] uint8 bits, reliab63 [n, 63] uint8 or None, observed_nac [n]
    int32 or None, parity [n] uint8, parity_reliab [n] uint8 or None.  Returns (status int8, nac int32, duid uint8, errs int32)."""
    import numpy as np

    code63 = np.ascontiguousarray(code63, dtype=np.uint8)
    n = code63.shape[0]
    parity = np.ascontiguousarray(parity, dtype=np.uint8)
    rel = None if reliab63 is None else np.ascontiguousarray(reliab63, dtype=np.uint8)
    obs = None if observed_nac is None else np.ascontiguousarray(observed_nac, dtype=np.int32)
    prel = None if parity_reliab is None else np.ascontiguousarray(parity_reliab, dtype=np.uint8)
    st, nac, duid, errs = np.zeros(n, np.int8), np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.int32)
    check(
        lib().dsdneo_b200_p25p1_nid_decode_batch_host(
            code63.ctypes.data, None if rel is None else rel.ctypes.data, None if obs is None else obs.ctypes.data,
            parity.ctypes.data, None if prel is None else prel.ctypes.data, threshold, st.ctypes.data, nac.ctypes.data,
            duid.ctypes.data, errs.ctypes.data, n,
        ),
        "p25p1_nid_decode_batch_host",
    )
    return st, nac, duid, errs


def dmr_burst_cut(d_dibits, d_reliability, d_counts, d_hits, d_n_hits, inverted_dmr: bool = False, stream=None):
    """Device-side DMR BS data burst cutter on the symbolizer / frame-sync outputs (torch cuda tensors).  Returns a dict of cuda
    uint8 tensors indexed by slot = channel * max_hits + hit: cach24, info196, rel98, slot_type20, valid."""
    import torch

    n_ch, max_hits = d_hits.shape[0], d_hits.shape[1]
    slots = n_ch * max_hits
    dev = d_dibits.device
    u8 = lambda *shape: torch.zeros(shape, dtype=torch.uint8, device=dev)
    out = {"cach24": u8(slots, 24), "info196": u8(slots, 196), "rel98": u8(slots, 98), "slot_type20": u8(slots, 20), "valid": u8(slots)}
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    assert d_dibits.dtype == torch.uint8 and d_reliability.dtype == torch.uint8 and d_counts.dtype == torch.int32
    assert d_hits.dtype == torch.int32 and d_n_hits.dtype == torch.int32 and d_hits.is_contiguous()
    check(
        lib().dsdneo_b200_dmr_burst_cut_batch(
            d_dibits.data_ptr(), d_dibits.shape[1], d_reliability.data_ptr(), d_reliability.shape[1], d_counts.data_ptr(),
            d_hits.data_ptr(), d_n_hits.data_ptr(), n_ch, max_hits, 1 if inverted_dmr else 0, out["cach24"].data_ptr(),
            out["info196"].data_ptr(), out["rel98"].data_ptr(), out["slot_type20"].data_ptr(), out["valid"].data_ptr(),
            _stream_ptr(stream),
        ),
        "dmr_burst_cut_batch",
    )
    return out


def _p25_dtypes():
    import numpy as np

    frame = np.dtype([
        ("position", "<i8"), ("channel", "<i4"), ("voice_index", "<i4"), ("nac", "<i2"), ("nid_errs", "<i2"), ("nid_status", "i1"),
        ("duid", "u1"), ("n_tsbk", "u1"), ("tsbk_crc_ok", "u1"), ("rs_kind", "u1"), ("rs_status", "u1"), ("lsd_ok", "u1"),
        ("n_word_soft", "u1"), ("lsd", "u1", (2,)), ("reserved", "u1", (6,)), ("tsbk", "u1", (3, 12)), ("rs_data", "u1", (20,)),
        ("rs_in_data", "u1", (20,)), ("rs_in_parity", "u1", (16,))])
    voice = np.dtype([("bits", "<u4", (9, 8)), ("reliab", "u1", (9, 8, 23))])
    assert frame.itemsize == 128 and voice.itemsize == 1944
    return frame, voice


P25_FRAME_DTYPE, P25_VOICE_DTYPE = _p25_dtypes()


def p25p1_frame_cut(d_dibits, d_llr, d_counts, d_hits, d_n_hits, n_payload: int, stream=None):
    """Device-side P25p1 frame cutter on the symbolizer / frame-sync outputs (torch cuda tensors).  Returns a dict of cuda
    tensors indexed by slot = channel * max_hits + hit: nid_code63, nid_reliab63, nid_parity, nid_parity_reliab, nid_valid,
    payload_dibits [slots, n_payload], payload_llr [slots, n_payload, 2], payload_valid."""
    import torch

    n_ch, max_hits = d_hits.shape[0], d_hits.shape[1]
    slots = n_ch * max_hits
    dev = d_dibits.device
    u8 = lambda *shape: torch.zeros(shape, dtype=torch.uint8, device=dev)
    out = {"nid_code63": u8(slots, 63), "nid_reliab63": u8(slots, 63), "nid_parity": u8(slots), "nid_parity_reliab": u8(slots),
           "nid_valid": u8(slots), "payload_dibits": u8(slots, max(n_payload, 1)),
           "payload_llr": torch.zeros((slots, max(n_payload, 1), 2), dtype=torch.int16, device=dev), "payload_valid": u8(slots)}
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    assert d_dibits.dtype == torch.uint8 and d_llr.dtype == torch.int16 and d_counts.dtype == torch.int32
    assert d_hits.dtype == torch.int32 and d_n_hits.dtype == torch.int32 and d_hits.is_contiguous()
    check(
        lib().dsdneo_b200_p25p1_frame_cut_batch(
            d_dibits.data_ptr(), d_dibits.shape[1], d_llr.data_ptr(), d_llr.shape[1], d_counts.data_ptr(), d_hits.data_ptr(),
            d_n_hits.data_ptr(), n_ch, max_hits, n_payload, out["nid_code63"].data_ptr(), out["nid_reliab63"].data_ptr(),
            out["nid_parity"].data_ptr(), out["nid_parity_reliab"].data_ptr(), out["nid_valid"].data_ptr(),
            out["payload_dibits"].data_ptr(), out["payload_llr"].data_ptr(), out["payload_valid"].data_ptr(),
            _stream_ptr(stream),
        ),
        "p25p1_frame_cut_batch",
    )
    return out


def p25p1_frames_decode(d_dibits, d_llr, d_counts, d_hits, d_n_hits, observed_nac=None, threshold=64, hard_override=True,
                        region_offset=0, stream=None):
    """Sync hits -> frame records, all on the device: NID cut + p25p1_nid_decode + the frame decoder (TSBK / HDU / LDU1 / LDU2).
    Returns (frames, voices): numpy structured arrays (dsdneo_b200_p25p1_frame / _voice layouts), trimmed to the counts."""
    import numpy as np
    import torch

    n_ch, max_hits = d_hits.shape[0], d_hits.shape[1]
    slots = n_ch * max_hits
    dev = d_dibits.device
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    assert region_offset == 0, "the python wrapper cuts from buffer index 0"
    cut = p25p1_frame_cut(d_dibits, d_llr, d_counts, d_hits, d_n_hits, 0, stream)
    st = torch.zeros(slots, dtype=torch.int8, device=dev)
    nac = torch.zeros(slots, dtype=torch.int32, device=dev)
    duid = torch.zeros(slots, dtype=torch.uint8, device=dev)
    errs = torch.zeros(slots, dtype=torch.int32, device=dev)
    obs = None
    if observed_nac is not None:
        obs = torch.as_tensor(np.repeat(np.asarray(observed_nac, np.int32), max_hits), device=dev).contiguous()
    check(lib().dsdneo_b200_p25p1_nid_decode_batch(cut["nid_code63"].data_ptr(), cut["nid_reliab63"].data_ptr(),
                                                   obs.data_ptr() if obs is not None else None, cut["nid_parity"].data_ptr(),
                                                   cut["nid_parity_reliab"].data_ptr(), threshold, st.data_ptr(), nac.data_ptr(),
                                                   duid.data_ptr(), errs.data_ptr(), slots, _stream_ptr(stream)), "p25p1_nid_decode_batch")
    f_off = torch.zeros(n_ch, dtype=torch.int32, device=dev)
    v_off = torch.zeros(n_ch, dtype=torch.int32, device=dev)
    totals = torch.zeros(2, dtype=torch.int32, device=dev)
    frames = torch.zeros((slots, 128), dtype=torch.uint8, device=dev)
    voices = torch.zeros((slots, 1944), dtype=torch.uint8, device=dev)
    check(lib().dsdneo_b200_p25p1_frames_decode_batch(
        d_dibits.data_ptr(), d_dibits.shape[1], d_llr.data_ptr(), d_llr.shape[1], d_counts.data_ptr(), d_hits.data_ptr(),
        d_n_hits.data_ptr(), n_ch, max_hits, region_offset, None, st.data_ptr(), cut["nid_valid"].data_ptr(), nac.data_ptr(),
        duid.data_ptr(), errs.data_ptr(),
        threshold, 1 if hard_override else 0, f_off.data_ptr(), v_off.data_ptr(), totals.data_ptr(), frames.data_ptr(), slots,
        voices.data_ptr(), slots, _stream_ptr(stream)), "p25p1_frames_decode_batch")
    torch.cuda.synchronize()
    nf, nv = (int(x) for x in totals.cpu().numpy())
    fr = frames[:nf].cpu().numpy().reshape(-1).view(P25_FRAME_DTYPE)
    vo = voices[:nv].cpu().numpy().reshape(-1).view(P25_VOICE_DTYPE)
    return fr, vo


class P25p1RxConfig(C.Structure):
    _fields_ = [("n_channels", C.c_int), ("rate_hz", C.c_int), ("block_pairs", C.c_int), ("max_pairs_per_call", C.c_int),
                ("input_cu8", C.c_int), ("fir_arith", C.c_int), ("max_hits", C.c_int), ("erasure_threshold", C.c_int),
                ("hard_override_disabled", C.c_int), ("track_nac", C.c_int), ("channel_squelch_level", C.POINTER(C.c_float)),
                ("p25_filter_taps", C.POINTER(C.c_float)), ("p25_filter_len", C.c_int), ("acquire_tiles", C.c_int),
                ("auto_reacquire_tiles", C.c_int)]


class P25p1RxOut(C.Structure):
    _fields_ = [("d_frames", C.c_void_p), ("frame_capacity", C.c_int), ("d_voices", C.c_void_p), ("voice_capacity", C.c_int),
                ("d_totals", C.c_void_p), ("d_dibits", C.c_void_p), ("dibit_pitch", C.c_size_t), ("d_counts", C.c_void_p)]


class P25p1RxHostOut(C.Structure):
    _fields_ = [("h_frames", C.c_void_p), ("frame_capacity", C.c_int), ("h_voices", C.c_void_p), ("voice_capacity", C.c_int),
                ("h_totals", C.c_void_p), ("h_dibits", C.c_void_p), ("dibit_pitch", C.c_size_t), ("h_counts", C.c_void_p)]


class P25p1Rx:
    """P25 Phase 1 C4FM receiver bank (dsdneo_b200_p25p1_rx_*): per-channel IQ -> frames / voice records / dibits."""

    def __init__(self, n_channels, p25_taps, rate_hz=48000, block_pairs=8192, max_pairs_per_call=49152, input_cu8=True,
                 fir_arith=FIR_ARITH_FMA, max_hits=32, track_nac=False, acquire_tiles=0, auto_reacquire_tiles=0):
        import numpy as np

        self._taps = np.ascontiguousarray(p25_taps, dtype=np.float32)
        cfg = P25p1RxConfig()
        cfg.n_channels, cfg.rate_hz, cfg.block_pairs, cfg.max_pairs_per_call = n_channels, rate_hz, block_pairs, max_pairs_per_call
        cfg.input_cu8, cfg.fir_arith, cfg.max_hits, cfg.track_nac = 1 if input_cu8 else 0, fir_arith, max_hits, 1 if track_nac else 0
        cfg.p25_filter_taps = self._taps.ctypes.data_as(C.POINTER(C.c_float))
        cfg.p25_filter_len = self._taps.size
        cfg.acquire_tiles = acquire_tiles
        cfg.auto_reacquire_tiles = auto_reacquire_tiles
        self.n_channels, self.input_cu8 = n_channels, input_cu8
        self._h = lib().dsdneo_b200_p25p1_rx_create(C.byref(cfg))
        if not self._h:
            raise B200Error(f"p25p1_rx_create failed: {last_error()}")
        self.frame_capacity = lib().dsdneo_b200_p25p1_rx_frame_capacity(self._h)
        self.voice_capacity = lib().dsdneo_b200_p25p1_rx_voice_capacity(self._h)
        self.dibit_pitch = lib().dsdneo_b200_p25p1_rx_dibit_pitch(self._h)

    def close(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().dsdneo_b200_p25p1_rx_destroy(self._h)
            self._h = None

    __del__ = close

    def alloc_device_out(self, device, with_dibits=True):
        import torch

        o = {"frames": torch.zeros((self.frame_capacity, 128), dtype=torch.uint8, device=device),
             "voices": torch.zeros((self.voice_capacity, 1944), dtype=torch.uint8, device=device),
             "totals": torch.zeros(2, dtype=torch.int32, device=device)}
        if with_dibits:
            o["dibits"] = torch.zeros((self.n_channels, self.dibit_pitch), dtype=torch.uint8, device=device)
            o["counts"] = torch.zeros(self.n_channels, dtype=torch.int32, device=device)
        return o

    def _dev_out(self, d_iq, out, stream):
        import torch

        assert d_iq.is_cuda and d_iq.is_contiguous() and d_iq.shape[0] == self.n_channels
        assert d_iq.dtype == (torch.uint8 if self.input_cu8 else torch.float32)
        if stream is None:
            stream = torch.cuda.current_stream(d_iq.device)
        o = P25p1RxOut(out["frames"].data_ptr(), out["frames"].shape[0], out["voices"].data_ptr(), out["voices"].shape[0],
                       out["totals"].data_ptr(), out["dibits"].data_ptr() if "dibits" in out else None,
                       out["dibits"].shape[1] if "dibits" in out else 0, out["counts"].data_ptr() if "counts" in out else None)
        return o, stream

    def process(self, d_iq, n_pairs, out, stream=None):
        """d_iq: cuda uint8 [n_ch, pitch, 2] (cu8) or float32 [n_ch, pitch, 2]; out from alloc_device_out."""
        o, stream = self._dev_out(d_iq, out, stream)
        check(lib().dsdneo_b200_p25p1_rx_process(self._h, d_iq.data_ptr(), d_iq.shape[1], n_pairs, C.byref(o), _stream_ptr(stream)),
              "p25p1_rx_process")

    def submit(self, d_iq, n_pairs, out, stream=None):
        """Pipelined form of process(): queues the tile and returns a ticket; consecutive tiles overlap on the device."""
        o, stream = self._dev_out(d_iq, out, stream)
        t = lib().dsdneo_b200_p25p1_rx_submit(self._h, d_iq.data_ptr(), d_iq.shape[1], n_pairs, C.byref(o), _stream_ptr(stream))
        if t < 0:
            check(int(t), "p25p1_rx_submit")
        return t

    def wait(self, ticket, stream=None):
        """Orders `stream` (default: torch's current stream) behind the tile's outputs; the host does not block."""
        import torch

        if stream is None:
            stream = torch.cuda.current_stream()
        check(lib().dsdneo_b200_p25p1_rx_wait(self._h, ticket, _stream_ptr(stream)), "p25p1_rx_wait")

    def reacquire(self, synchronised=None, tiles=1):
        """Sends the channels whose flag is 0 (None: all) back to the sync hunt for the next `tiles` tiles."""
        import numpy as np

        if synchronised is None:
            ptr = None
        else:
            flags = np.ascontiguousarray(synchronised, dtype=np.int32)
            assert flags.size == self.n_channels
            ptr = flags.ctypes.data
        check(lib().dsdneo_b200_p25p1_rx_reacquire(self._h, ptr, int(tiles)), "p25p1_rx_reacquire")

    def channel_status(self):
        """(synchronised flags, idle tiles) per channel as numpy int32 arrays."""
        import numpy as np

        a, b = np.zeros(self.n_channels, np.int32), np.zeros(self.n_channels, np.int32)
        check(lib().dsdneo_b200_p25p1_rx_channel_status(self._h, a.ctypes.data, b.ctypes.data), "p25p1_rx_channel_status")
        return a, b

    def input_consumed(self, ticket, stream=None):
        import torch

        if stream is None:
            stream = torch.cuda.current_stream()
        check(lib().dsdneo_b200_p25p1_rx_input_consumed(self._h, ticket, _stream_ptr(stream)), "p25p1_rx_input_consumed")

    @staticmethod
    def records(out):
        """(frames, voices) numpy structured arrays of a finished device call (synchronises)."""
        import torch

        torch.cuda.synchronize()
        nf, nv = (int(x) for x in out["totals"].cpu().numpy())
        fr = out["frames"][:nf].cpu().numpy().reshape(-1).view(P25_FRAME_DTYPE)
        vo = out["voices"][:nv].cpu().numpy().reshape(-1).view(P25_VOICE_DTYPE)
        return fr, vo

    def alloc_host_out(self, with_dibits=True):
        import torch

        o = {"frames": torch.zeros((self.frame_capacity, 128), dtype=torch.uint8).pin_memory(),
             "voices": torch.zeros((self.voice_capacity, 1944), dtype=torch.uint8).pin_memory(),
             "totals": torch.zeros(2, dtype=torch.int32).pin_memory()}
        if with_dibits:
            o["dibits"] = torch.zeros((self.n_channels, self.dibit_pitch), dtype=torch.uint8).pin_memory()
            o["counts"] = torch.zeros(self.n_channels, dtype=torch.int32).pin_memory()
        return o

    def _host_out(self, out):
        return P25p1RxHostOut(out["frames"].data_ptr(), out["frames"].shape[0], out["voices"].data_ptr(), out["voices"].shape[0],
                              out["totals"].data_ptr(), out["dibits"].data_ptr() if "dibits" in out else None,
                              out["dibits"].shape[1] if "dibits" in out else 0, out["counts"].data_ptr() if "counts" in out else None)

    def submit_host(self, h_iq, n_pairs, out):
        """h_iq: host tensor [n_ch, pitch, 2] (pinned for overlap).  Returns a ticket."""
        o = self._host_out(out)
        t = lib().dsdneo_b200_p25p1_rx_submit_host(self._h, h_iq.data_ptr(), h_iq.shape[1], n_pairs, C.byref(o))
        if t < 0:
            check(int(t), "p25p1_rx_submit_host")
        return t

    def wait_host(self, ticket):
        check(lib().dsdneo_b200_p25p1_rx_wait_host(self._h, ticket), "p25p1_rx_wait_host")

    @staticmethod
    def host_records(out):
        nf, nv = (int(x) for x in out["totals"].numpy())
        return (out["frames"][:nf].numpy().reshape(-1).view(P25_FRAME_DTYPE).copy(),
                out["voices"][:nv].numpy().reshape(-1).view(P25_VOICE_DTYPE).copy())


def ambe3600x2450_decode(ambe_fr):
    """ambe_fr uint8 [n, 4, 24] -> (ambe_d [n, 49], c0_errors [n], total_errors [n]); host arrays."""
    import numpy as np

    fr = np.ascontiguousarray(ambe_fr, dtype=np.uint8).reshape(-1, 96)
    n = fr.shape[0]
    d, c0, tot = np.zeros((n, 49), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
    check(lib().dsdneo_b200_ambe3600x2450_decode_batch_host(fr.ctypes.data, d.ctypes.data, c0.ctypes.data, tot.ctypes.data, n),
          "ambe3600x2450_decode")
    return d, c0, tot


def imbe7200x4400_decode(imbe_fr):
    """imbe_fr uint8 [n, 8, 23] -> (imbe_d [n, 88], c0_errors [n], total_errors [n]); host arrays."""
    import numpy as np

    fr = np.ascontiguousarray(imbe_fr, dtype=np.uint8).reshape(-1, 184)
    n = fr.shape[0]
    d, c0, tot = np.zeros((n, 88), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
    check(lib().dsdneo_b200_imbe7200x4400_decode_batch_host(fr.ctypes.data, d.ctypes.data, c0.ctypes.data, tot.ctypes.data, n),
          "imbe7200x4400_decode")
    return d, c0, tot


def p25p1_voice_imbe_decode(d_voices, n_records: int, stream=None):
    """d_voices: device uint8 [>= n_records, 1944] (the bank's voice records).  Returns device tensors
    (imbe_d uint8 [n, 9, 88], c0_errors int32 [n, 9], total_errors int32 [n, 9])."""
    import torch

    dev = d_voices.device
    d = torch.zeros((n_records, 9, 88), dtype=torch.uint8, device=dev)
    c0 = torch.zeros((n_records, 9), dtype=torch.int32, device=dev)
    tot = torch.zeros((n_records, 9), dtype=torch.int32, device=dev)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    check(lib().dsdneo_b200_p25p1_voice_imbe_decode_batch(d_voices.data_ptr(), n_records, d.data_ptr(), c0.data_ptr(), tot.data_ptr(),
                                                          _stream_ptr(stream)), "p25p1_voice_imbe_decode")
    return d, c0, tot


def ambe_2450_dibit_map():
    import numpy as np

    m = np.zeros((36, 4), np.uint8)
    check(lib().dsdneo_b200_ambe_2450_dibit_map(m.ctypes.data), "ambe_2450_dibit_map")
    return m


def dmr_voice_cut(d_dibits, d_counts, d_hits, d_n_hits, max_hits: int, n_bursts: int, inverted_dmr: bool = False, stream=None):
    """Device tensors in (dibits uint8 [n_ch, pitch], counts int32 [n_ch], hits int32 [n_ch, max_hits, 2], n_hits int32 [n_ch]);
    returns device tensors (cach24 [R, 24], ambe_fr [R, 3, 4, 24], sync48 [R, 48], valid [R]), R = n_ch * max_hits * n_bursts."""
    import torch

    n_ch = d_dibits.shape[0]
    R = n_ch * max_hits * n_bursts
    dev = d_dibits.device
    cach = torch.zeros((R, 24), dtype=torch.uint8, device=dev)
    fr = torch.zeros((R, 3, 4, 24), dtype=torch.uint8, device=dev)
    sync = torch.zeros((R, 48), dtype=torch.uint8, device=dev)
    valid = torch.zeros(R, dtype=torch.uint8, device=dev)
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    check(lib().dsdneo_b200_dmr_voice_cut_batch(d_dibits.data_ptr(), d_dibits.shape[1], d_counts.data_ptr(), d_hits.data_ptr(),
                                                d_n_hits.data_ptr(), n_ch, max_hits, n_bursts, 1 if inverted_dmr else 0,
                                                cach.data_ptr(), fr.data_ptr(), sync.data_ptr(), valid.data_ptr(), _stream_ptr(stream)),
          "dmr_voice_cut")
    return cach, fr, sync, valid


def p25_word_decode(code: int, data_bits, parity_bits):
    """Golay(24,6)/(24,12)/Hamming(10,6,3) P25 words. Returns (data_bits corrected, status u8 [n], fixed i32 [n])."""
    import numpy as np

    data_bits = np.ascontiguousarray(data_bits, dtype=np.uint8).copy()
    parity_bits = np.ascontiguousarray(parity_bits, dtype=np.uint8)
    n = data_bits.shape[0]
    st, fx = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    check(lib().dsdneo_b200_p25_word_decode_batch_host(code, data_bits.ctypes.data, parity_bits.ctypes.data, st.ctypes.data,
                                                       fx.ctypes.data, n), "p25_word_decode")
    return data_bits, st, fx


def bch_63_16_decode(in63, out16_init=None):
    """P25 NID BCH(63,16,11). Returns (out16 [n,16], ok u8 [n], err_count i32 [n])."""
    import numpy as np

    in63 = np.ascontiguousarray(in63, dtype=np.uint8).reshape(-1, 63)
    n = in63.shape[0]
    out = np.zeros((n, 16), np.uint8) if out16_init is None else np.ascontiguousarray(out16_init, dtype=np.uint8).copy()
    ok, ec = np.zeros(n, np.uint8), np.zeros(n, np.int32)
    check(lib().dsdneo_b200_bch_63_16_decode_batch_host(in63.ctypes.data, out.ctypes.data, ok.ctypes.data, ec.ctypes.data, n),
          "bch_63_16_decode")
    return out, ok, ec


class MbeParms(C.Structure):
    """struct mbe_parameters of mbelib 1.3.0 (see include/dsdneo_b200.h: parity unpinned)."""

    _fields_ = [("w0", C.c_float), ("L", C.c_int), ("K", C.c_int), ("Vl", C.c_int * 57), ("Ml", C.c_float * 57),
                ("log2Ml", C.c_float * 57), ("PHIl", C.c_float * 57), ("PSIl", C.c_float * 57), ("gamma", C.c_float),
                ("un", C.c_int), ("repeat", C.c_int)]


def mbe_synth(cur, prev_enhanced, keys=None, uvquality: int = 3, want_int16: bool = True):
    """cur / prev_enhanced: ctypes arrays (MbeParms * n), updated in place. Returns (pcm_f [n,160] f32, pcm_s [n,160] i16)."""
    import numpy as np

    n = len(cur)
    pcm_f = np.zeros((n, 160), np.float32)
    pcm_s = np.zeros((n, 160), np.int16) if want_int16 else None
    k = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
    check(lib().dsdneo_b200_mbe_synth_batch_host(C.byref(cur), C.byref(prev_enhanced), None if k is None else k.ctypes.data, uvquality,
                                                 pcm_f.ctypes.data, None if pcm_s is None else pcm_s.ctypes.data, n), "mbe_synth")
    return pcm_f, pcm_s


class SymClass(C.Structure):
    _fields_ = [("filter", C.c_int), ("window_l", C.c_int), ("track_minmax", C.c_int), ("negative", C.c_int), ("rf_mod", C.c_int)]


class SymbolizerConfig(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int), ("output_rate_hz", C.c_int), ("symbol_rate_hz", C.c_int), ("ssize", C.c_int), ("msize", C.c_int),
        ("use_cosine_filter", C.c_int), ("n_filters", C.c_int),
        ("filter_taps", C.POINTER(C.c_float) * 8), ("filter_len", C.c_int * 8),
    ]


class AcqPattern(C.Structure):
    """dsdneo_b200_acq_pattern: one 24-symbol sync pattern of getFrameSync and the decoder class behind it."""
    _fields_ = [("symbols", C.c_char_p), ("sync_type", C.c_int), ("kind", C.c_int), ("cls", SymClass)]


ACQ_INFO_BYTES = 84


def acq_info_dtype():
    import numpy as np

    return np.dtype([("acquired", "<i4"), ("sync_type", "<i4"), ("hit_index", "<i4"), ("hunt_symbols", "<i4"), ("warm_start", "u1"),
                     ("resample_ok", "u1"), ("resampled", "u1", (66,))])


class SymbolOut(C.Structure):
    _fields_ = [("d_symbols", C.c_void_p), ("d_dibits", C.c_void_p), ("d_reliability", C.c_void_p), ("d_llr", C.c_void_p),
                ("d_count", C.c_void_p), ("pitch", C.c_size_t)]


def sym_class_from_synctype(synctype: int, lastsynctype: int, use_cosine_filter: bool = True) -> SymClass:
    out = SymClass()
    check(lib().dsdneo_b200_sym_class_from_synctype(synctype, lastsynctype, 1 if use_cosine_filter else 0, C.byref(out)),
          "sym_class_from_synctype")
    return out


class Symbolizer:
    """Batched twin of getSymbol()/getDibitSoft() for n_channels discriminator streams."""

    def __init__(self, n_channels, output_rate_hz=48000, symbol_rate_hz=4800, filters=None, ssize=0, msize=0, use_cosine_filter=True):
        import numpy as np

        cfg = SymbolizerConfig()
        cfg.n_channels, cfg.output_rate_hz, cfg.symbol_rate_hz = n_channels, output_rate_hz, symbol_rate_hz
        cfg.ssize, cfg.msize, cfg.use_cosine_filter = ssize, msize, 1 if use_cosine_filter else 0
        self._taps = {}
        filters = filters or {}
        cfg.n_filters = (max(filters) + 1) if filters else 0
        for idx, taps in filters.items():
            a = np.ascontiguousarray(taps, dtype=np.float32)
            self._taps[idx] = a
            cfg.filter_taps[idx] = a.ctypes.data_as(C.POINTER(C.c_float))
            cfg.filter_len[idx] = a.size
        self.n_channels = n_channels
        self.sps_floor = max(2, min(64, output_rate_hz // symbol_rate_hz))
        self._h = lib().dsdneo_b200_symbolizer_create(C.byref(cfg))
        if not self._h:
            raise B200Error(f"symbolizer_create failed: {last_error()}")

    def close(self):
        if getattr(self, "_h", None) and lib is not None:  # lib is None while the interpreter shuts down
            lib().dsdneo_b200_symbolizer_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None):
        check(lib().dsdneo_b200_symbolizer_reset(self._h, _stream_ptr(stream)), "symbolizer_reset")

    def set_class(self, classes):
        arr = (SymClass * self.n_channels)(*classes)
        check(lib().dsdneo_b200_symbolizer_set_class(self._h, arr), "symbolizer_set_class")

    def set_snr(self, snr_db=None):
        """Per-channel C4FM SNR (dB) as the metrics hook reports it; None = no hook installed (the default)."""
        import numpy as np

        if snr_db is None:
            check(lib().dsdneo_b200_symbolizer_set_snr(self._h, None), "symbolizer_set_snr")
            return
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(snr_db, dtype=np.float64), (self.n_channels,)))
        check(lib().dsdneo_b200_symbolizer_set_snr(self._h, a.ctypes.data), "symbolizer_set_snr")

    def out_pitch(self, n_samples):
        return (n_samples + 256) // (self.sps_floor - 1) + 2

    def set_acquire_patterns(self, patterns):
        """patterns: [(symbols '1'/'3' x 24, sync_type, kind 0 P25p1 / 1 DMR, SymClass)], compared in this order."""
        self._pat_keep = [p[0].encode() if isinstance(p[0], str) else p[0] for p in patterns]
        arr = (AcqPattern * len(patterns))()
        for k, (sym, st, kind, cls) in enumerate(patterns):
            arr[k].symbols, arr[k].sync_type, arr[k].kind, arr[k].cls = self._pat_keep[k], st, kind, cls
        check(lib().dsdneo_b200_symbolizer_set_acquire_patterns(self._h, arr, len(patterns)), "symbolizer_set_acquire_patterns")

    def set_acquired(self, flags=None):
        import numpy as np

        if flags is None:
            check(lib().dsdneo_b200_symbolizer_set_acquired(self._h, None), "symbolizer_set_acquired")
            return
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(flags, dtype=np.int32), (self.n_channels,)))
        check(lib().dsdneo_b200_symbolizer_set_acquired(self._h, a.ctypes.data), "symbolizer_set_acquired")

    def _alloc_out(self, dev, n_samples):
        import torch

        pitch = self.out_pitch(n_samples)
        res = {
            "symbols": torch.zeros((self.n_channels, pitch), dtype=torch.float32, device=dev),
            "dibits": torch.zeros((self.n_channels, pitch), dtype=torch.uint8, device=dev),
            "reliability": torch.zeros((self.n_channels, pitch), dtype=torch.uint8, device=dev),
            "llr": torch.zeros((self.n_channels, pitch, 2), dtype=torch.int16, device=dev),
            "count": torch.zeros((self.n_channels,), dtype=torch.int32, device=dev),
        }
        out = SymbolOut(res["symbols"].data_ptr(), res["dibits"].data_ptr(), res["reliability"].data_ptr(), res["llr"].data_ptr(),
                        res["count"].data_ptr(), pitch)
        return res, out

    def run_acquire(self, d_disc, n_samples, stream=None, filtered=False):
        """getFrameSync + getDibitSoft over one launch of RAW discriminator samples; adds res['info'] uint8 [n_channels, 84]
        (view the host copy with acq_info_dtype())."""
        import torch

        assert d_disc.is_cuda and d_disc.dtype == torch.float32 and d_disc.is_contiguous() and d_disc.shape[0] == self.n_channels
        res, out = self._alloc_out(d_disc.device, n_samples)
        res["info"] = torch.zeros((self.n_channels, ACQ_INFO_BYTES), dtype=torch.uint8, device=d_disc.device)
        if stream is None:
            stream = torch.cuda.current_stream(d_disc.device)
        fn = lib().dsdneo_b200_symbolize_reacquire_batch if filtered else lib().dsdneo_b200_symbolize_acquire_batch
        check(fn(self._h, d_disc.data_ptr(), d_disc.shape[1], n_samples, C.byref(out), res["info"].data_ptr(), _stream_ptr(stream)),
              "symbolize_acquire_batch")
        return res

    def run(self, d_disc, n_samples, mode=SYM_MODE_GET_DIBIT_SOFT, have_sync=1, stream=None):
        """d_disc: cuda float32 [n_channels, pitch]. Returns dict of cuda tensors + per-channel counts."""
        import torch

        assert d_disc.is_cuda and d_disc.dtype == torch.float32 and d_disc.is_contiguous() and d_disc.shape[0] == self.n_channels
        pitch = self.out_pitch(n_samples)
        dev = d_disc.device
        res = {
            "symbols": torch.zeros((self.n_channels, pitch), dtype=torch.float32, device=dev),
            "dibits": torch.zeros((self.n_channels, pitch), dtype=torch.uint8, device=dev),
            "reliability": torch.zeros((self.n_channels, pitch), dtype=torch.uint8, device=dev),
            "llr": torch.zeros((self.n_channels, pitch, 2), dtype=torch.int16, device=dev),
            "count": torch.zeros((self.n_channels,), dtype=torch.int32, device=dev),
        }
        out = SymbolOut(res["symbols"].data_ptr(), res["dibits"].data_ptr(), res["reliability"].data_ptr(), res["llr"].data_ptr(),
                        res["count"].data_ptr(), pitch)
        if stream is None:
            stream = torch.cuda.current_stream(dev)
        check(lib().dsdneo_b200_symbolize_batch(self._h, d_disc.data_ptr(), d_disc.shape[1], n_samples, mode, have_sync, C.byref(out),
                                                _stream_ptr(stream)), "symbolize_batch")
        return res


def sync_hits_select(d_hits, d_n_hits, sync_type: int, out_max_hits: int, stream=None):
    """The hits of one sync type per channel, in stream order (device tensors in and out): (hits [n_ch, out_max_hits, 2], n [n_ch])."""
    import torch

    n_ch, max_hits = d_hits.shape[0], d_hits.shape[1]
    out = torch.zeros((n_ch, out_max_hits, 2), dtype=torch.int32, device=d_hits.device)
    n_out = torch.zeros(n_ch, dtype=torch.int32, device=d_hits.device)
    if stream is None:
        stream = torch.cuda.current_stream(d_hits.device)
    check(lib().dsdneo_b200_sync_hits_select(d_hits.data_ptr(), d_n_hits.data_ptr(), n_ch, max_hits, sync_type, out.data_ptr(), out_max_hits,
                                             n_out.data_ptr(), _stream_ptr(stream)), "sync_hits_select")
    return out, n_out


class SymbolStreamView(C.Structure):
    _fields_ = [("d_symbols", C.c_void_p), ("d_dibits", C.c_void_p), ("d_reliability", C.c_void_p), ("d_llr", C.c_void_p),
                ("pitch", C.c_size_t), ("d_valid", C.c_void_p), ("d_new", C.c_void_p), ("d_stream_base", C.c_void_p), ("keep", C.c_int)]


class SymbolStream:
    """dsdneo_b200_symbol_stream: the launches of a Symbolizer joined into one stream per channel (history kept on the device)."""

    def __init__(self, n_channels: int, keep: int, max_new: int):
        self.n_channels, self.keep, self.max_new = n_channels, keep, max_new
        self._h = lib().dsdneo_b200_symbol_stream_create(n_channels, keep, max_new)
        if not self._h:
            raise B200Error("symbol_stream_create: " + last_error())

    def close(self):
        if self._h:
            lib().dsdneo_b200_symbol_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def run(self, symbolizer, d_disc, n_samples, mode=SYM_MODE_GET_DIBIT_SOFT, have_sync=1, stream=None):
        """begin -> dsdneo_b200_symbolize_batch into the stream's rows -> commit; returns the view (device pointers)."""
        import torch

        assert d_disc.is_cuda and d_disc.dtype == torch.float32 and d_disc.is_contiguous()
        assert symbolizer.out_pitch(n_samples) <= self.max_new
        if stream is None:
            stream = torch.cuda.current_stream(d_disc.device)
        out = SymbolOut()
        check(lib().dsdneo_b200_symbol_stream_begin(self._h, C.byref(out)), "symbol_stream_begin")
        check(lib().dsdneo_b200_symbolize_batch(symbolizer._h, d_disc.data_ptr(), d_disc.shape[1], n_samples, mode, have_sync,
                                                C.byref(out), _stream_ptr(stream)), "symbolize_batch")
        view = SymbolStreamView()
        check(lib().dsdneo_b200_symbol_stream_commit(self._h, C.byref(view), _stream_ptr(stream)), "symbol_stream_commit")
        return view

    def fetch(self, view):
        """Host copies of the joined rows (test helper): dict of numpy arrays."""
        import numpy as np
        import torch

        torch.cuda.synchronize()
        n, p = self.n_channels, int(view.pitch)
        res = {"symbols": np.zeros((n, p), np.float32), "dibits": np.zeros((n, p), np.uint8), "reliability": np.zeros((n, p), np.uint8),
               "llr": np.zeros((n, p, 2), np.int16), "valid": np.zeros(n, np.int32), "new": np.zeros(n, np.int32),
               "stream_base": np.zeros(n, np.int64)}
        for key, ptr in (("symbols", view.d_symbols), ("dibits", view.d_dibits), ("reliability", view.d_reliability), ("llr", view.d_llr),
                         ("valid", view.d_valid), ("new", view.d_new), ("stream_base", view.d_stream_base)):
            check(lib().dsdneo_b200_memcpy_d2h(res[key].ctypes.data, ptr, res[key].nbytes, None), "memcpy_d2h")
        torch.cuda.synchronize()
        return res


def viterbi_k5_decode(cost, in_len, punct=None, out_pitch=None, out_init=None):
    """cost: uint16 [n, pitch]; returns (out uint8 [n, out_pitch], metric uint32 [n])."""
    import numpy as np

    cost = np.ascontiguousarray(cost, dtype=np.uint16)
    n = cost.shape[0]
    if out_pitch is None:
        out_pitch = 80
    out = np.zeros((n, out_pitch), np.uint8) if out_init is None else np.ascontiguousarray(out_init, dtype=np.uint8).copy()
    met = np.zeros(n, np.uint32)
    pp, pl = (None, 0)
    if punct is not None:
        punct = np.ascontiguousarray(punct, dtype=np.uint8)
        pp, pl = punct.ctypes.data, punct.size
    check(lib().dsdneo_b200_viterbi_k5_decode_batch_host(cost.ctypes.data, cost.shape[1], in_len, pp, pl, out.ctypes.data, out.shape[1],
                                                         met.ctypes.data, n), "viterbi_k5_decode")
    return out, met


def nxdn_conv_decode(sym, rel, n_steps, n_bits_out, metrics, out_pitch=40, out_init=None):
    """sym/rel: uint8 [n, pitch]; metrics: uint16 [n, 32] updated in place. Returns out uint8 [n, out_pitch]."""
    import numpy as np

    sym = np.ascontiguousarray(sym, dtype=np.uint8)
    n = sym.shape[0]
    out = np.zeros((n, out_pitch), np.uint8) if out_init is None else np.ascontiguousarray(out_init, dtype=np.uint8).copy()
    rp = None
    if rel is not None:
        rel = np.ascontiguousarray(rel, dtype=np.uint8)
        rp = rel.ctypes.data
    assert metrics.dtype == np.uint16 and metrics.shape == (n, 32) and metrics.flags.c_contiguous
    check(lib().dsdneo_b200_nxdn_conv_decode_batch_host(sym.ctypes.data, rp, sym.shape[1], n_steps, n_bits_out, metrics.ctypes.data,
                                                        out.ctypes.data, out.shape[1], n), "nxdn_conv_decode")
    return out
