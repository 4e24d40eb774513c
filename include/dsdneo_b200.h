/* SPDX-License-Identifier: GPL-3.0-or-later */
/**
 * @file dsdneo_b200.h
 * @brief C-ABI of libdsdneo_b200.so: the B200 (sm_100a) many-channel twin of dsd-neo's
 *        demodulation hot path.
 *
 * Plain C, plain pointers and sizes.  Every entry point names the reference interface
 * (file:line under arancormonk/dsd-neo @ 4d06905) that it replaces or batches.
 *
 * Conventions
 *  - All functions return 0 on success and a negative DSDNEO_B200_E* code on failure unless
 *    documented otherwise; the text of the last failure on the calling thread is available
 *    from dsdneo_b200_last_error().  There is NO CPU fallback: without a CUDA device every
 *    compute entry point fails with DSDNEO_B200_ENODEV.
 *  - `stream` parameters are `cudaStream_t` passed as `void*` (NULL = default stream), so this
 *    header needs no CUDA headers.
 *  - Pointers prefixed `d_` are device pointers, `h_` host pointers.  `*_host` variants copy
 *    host->device, run the same kernels, and copy the results back (synchronous on return).
 *  - Carried per-channel DSP state lives in device arenas owned by the bank objects
 *    (struct-of-arrays: one value per channel per field).
 */
#ifndef DSDNEO_B200_H_
#define DSDNEO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSDNEO_B200_ABI_VERSION 2 /* 2: dsdneo_b200_p25p1_rx_config grew auto_reacquire_tiles */

enum {
    DSDNEO_B200_OK = 0,
    DSDNEO_B200_EINVAL = -1,  /* bad argument (same cases where the reference early-returns) */
    DSDNEO_B200_ENODEV = -2,  /* no usable CUDA device / wrong architecture */
    DSDNEO_B200_ECUDA = -3,   /* CUDA runtime error (sticky text in last_error) */
    DSDNEO_B200_ENOMEM = -4,
    DSDNEO_B200_EUNSUPPORTED = -5,
};

/* ---- library / device ------------------------------------------------------------------ */

int dsdneo_b200_abi_version(void);
const char* dsdneo_b200_last_error(void);
/** Select the CUDA device used by the calling thread (cudaSetDevice) and verify it is sm_100. */
int dsdneo_b200_init(int device_ordinal);
int dsdneo_b200_device_sm_count(void);
/** Blocks until all work queued on `stream` has finished. */
int dsdneo_b200_stream_sync(void* stream);
/** Number of CUDA kernels this library has launched in this process (bench.py reports it). */
unsigned long long dsdneo_b200_launch_count(void);

/**
 * Per-kernel device timing for bench.py's roofline: when enabled, every kernel launch of this library is
 * bracketed by CUDA events on its own stream and accumulated per kernel name.  `timing_report` synchronises
 * the outstanding events and writes {"kernel": {"launches": n, "ms": total}, ...} as JSON text.
 */
int dsdneo_b200_timing_enable(int on);
int dsdneo_b200_timing_report(char* buf, size_t cap);

/* Raw device/pinned memory so a pure-C host needs no CUDA runtime of its own. */
void* dsdneo_b200_malloc_device(size_t bytes);
void dsdneo_b200_free_device(void* d_ptr);
void* dsdneo_b200_malloc_pinned(size_t bytes);
void dsdneo_b200_free_pinned(void* h_ptr);
int dsdneo_b200_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int dsdneo_b200_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);

/* ---- K4 host helper: channel low-pass design ------------------------------------------- */

/* Channel LPF profile ids: include/dsd-neo/dsp/demod_state.h:36-43 (same numeric values). */
enum {
    DSDNEO_CH_LPF_PROFILE_WIDE = 0,
    DSDNEO_CH_LPF_PROFILE_6K25 = 1,
    DSDNEO_CH_LPF_PROFILE_12K5 = 2,
    DSDNEO_CH_LPF_PROFILE_PROVOICE = 3,
    DSDNEO_CH_LPF_PROFILE_P25_C4FM = 4,
    DSDNEO_CH_LPF_PROFILE_P25_CQPSK = 5,
    DSDNEO_CH_LPF_PROFILE_COUNT = 6,
};
#define DSDNEO_B200_LPF_MAX_TAPS 144

/**
 * Host-side twin of channel_lpf_design_low_pass() + dsd_firdes_low_pass()
 * (src/dsp/demod_pipeline.cpp:443-460,478-489; src/dsp/firdes.cpp low_pass/Blackman):
 * Blackman-windowed sinc, 1200 Hz transition, profile cutoff clamped to [100, 0.45*fs], unit DC gain.
 * @return number of taps written (odd), or a negative error when the design does not fit
 *         `max_taps` (the reference then falls back to fixed 63-tap tables; this library
 *         reports DSDNEO_B200_EUNSUPPORTED instead).
 */
int dsdneo_b200_channel_lpf_design(int rate_out_hz, int profile, float* taps_out, int max_taps);

enum {
    DSDNEO_SPS_FIR_DESIGN_INTERP = 0, /* base table re-sampled linearly (p25_filter, nxdn_filter) */
    DSDNEO_SPS_FIR_DESIGN_RRC = 1,    /* closed-form root-raised cosine (dmr 0.7, dpmr 0.2, m17 0.5) */
};
#define DSDNEO_B200_SPS_FIR_MAX_TAPS 1024

/**
 * Host-side twin of design_sps_fir() (src/dsp/dsd_filters.c:94-170), the lazy per-sps redesign behind p25_filter /
 * dmr_filter / nxdn_filter / dpmr_filter / m17_filter (:347-370): the base table as it is at its own samples per symbol,
 * otherwise a filter of the same span in symbols (odd length, <= 1023 taps) re-sampled from the table or evaluated from the
 * RRC closed form, then scaled to unit DC gain.  `base` is the reference's coefficient table of the filter
 * (dsd_filters.c:204-323), `design_kind` / `base_sps` / `rrc_alpha` the fields of its sps_fir descriptor (:325-345).
 * The result is what dsdneo_b200_symbolizer_config::filter_taps and dsdneo_b200_p25p1_rx_config::p25_filter_taps expect,
 * bit-identical to the taps the reference's filter holds after its first sample at `sps`.
 * @return the tap count, DSDNEO_B200_EINVAL for arguments the reference would leave the filter unready on (sps <= 1, empty
 *         table), DSDNEO_B200_EUNSUPPORTED when the taps do not fit `max_taps`.
 */
int dsdneo_b200_sps_fir_design(int design_kind, const float* base, int base_len, int base_sps, float rrc_alpha, int sps,
                               float* taps_out, int max_taps);

/* ---- block side: batched full_demod() for the FSK-discriminator output kind ------------- */

/**
 * A bank of N independent channels, each the twin of one reference `struct demod_state`
 * (include/dsd-neo/dsp/demod_state.h:67-258) configured as
 *   output_kind = DSD_DEMOD_OUTPUT_FSK_DISCRIMINATOR, cqpsk_enable = 0, downsample_passes = 0,
 *   iq_dc_block_enable = 0, iqbal_enable = 0 (the reference defaults for 4FSK modes,
 *   src/io/radio/rtl_demod_config.cpp:189-228,481,491-509).
 * Carried per channel: channel-LPF history (last taps-1 inputs), discriminator state
 * {prev_i, prev_q, have_prev, dc_est, discriminator_peak_est} (include/dsd-neo/dsp/fsk_modem.h:29-36),
 * channel_pwr, channel_squelched.
 */
typedef struct dsdneo_b200_demod_bank dsdneo_b200_demod_bank;

/* Which of the reference's FIR kernels the channel LPF reproduces bit-for-bit. */
enum {
    /* acc = fma(tap, x[-d] + x[+d], acc): simd_fir_complex_apply_avx2 (src/dsp/simd_fir_avx2.cpp:120-141),
     * the kernel the reference dispatches to on AVX2+FMA x86-64 hosts (src/dsp/simd_fir.cpp:322-330). Default. */
    DSDNEO_FIR_ARITH_FMA = 0,
    /* acc = acc + tap * (x[-d] + x[+d]) with separate roundings: the scalar and SSE2 kernels
     * (src/dsp/simd_fir.cpp:96-110, src/dsp/simd_fir_sse2.cpp:283-308). */
    DSDNEO_FIR_ARITH_NOFMA = 1,
};

typedef struct dsdneo_b200_demod_bank_config {
    int n_channels;
    int rate_out_hz;        /* demod_state.rate_out; channel sample rate (e.g. 48000) */
    int channel_lpf_enable; /* demod_state.channel_lpf_enable */
    /* Per-channel arrays (length n_channels) or NULL for the default in parentheses. */
    const int* channel_lpf_profile;     /* DSDNEO_CH_LPF_PROFILE_* (P25_C4FM) */
    const float* channel_squelch_level; /* demod_state.channel_squelch_level (0 = squelch off) */
    int fir_arith;                      /* DSDNEO_FIR_ARITH_* */
} dsdneo_b200_demod_bank_config;

/** Snapshot of one channel's carried state, for parity dumps ("sync state to host"). */
typedef struct dsdneo_b200_demod_chan_state {
    float prev_i, prev_q;
    int have_prev;
    float dc_est;
    float discriminator_peak_est;
    float channel_pwr;
    int channel_squelched;
} dsdneo_b200_demod_chan_state;

dsdneo_b200_demod_bank* dsdneo_b200_demod_bank_create(const dsdneo_b200_demod_bank_config* cfg);
void dsdneo_b200_demod_bank_destroy(dsdneo_b200_demod_bank* bank);
/** Zero all carried state (twin of a fresh demod_state + dsd_fsk_modem_init, src/dsp/fsk_modem.c:75-82). */
int dsdneo_b200_demod_bank_reset(dsdneo_b200_demod_bank* bank, void* stream);
int dsdneo_b200_demod_bank_get_state(dsdneo_b200_demod_bank* bank, int channel, dsdneo_b200_demod_chan_state* out);
int dsdneo_b200_demod_bank_get_taps(dsdneo_b200_demod_bank* bank, int profile, float* taps_out, int max_taps);

/**
 * Batched twin of `void full_demod(struct demod_state*)` (src/dsp/demod_pipeline.cpp:1330-1350;
 * declared include/dsd-neo/dsp/demod_pipeline.h:106) for every channel of the bank:
 *   channel_lpf_apply -> simd_fir_complex_apply   (demod_pipeline.cpp:526-555, simd_fir.cpp:55-133)
 *   mean_power over the first <=512 floats + channel squelch (demod_pipeline.cpp:926-945,1003-1020)
 *   dsd_fsk_modem_discriminator_process           (src/dsp/fsk_modem.c:135-164)
 *
 * Each channel receives `n_blocks` consecutive reference blocks of `block_pairs` complex samples;
 * one reference block == one full_demod() call with lp_len = 2*block_pairs (so the FIR's
 * "pad the right edge with the block's last sample" behaviour, simd_fir.cpp:65-84, is reproduced
 * at every block boundary, and squelch is evaluated once per block).
 *
 * @param d_iq      [n_channels][iq_pitch_pairs] interleaved cf32 (float2); pitch >= n_blocks*block_pairs
 * @param d_result  [n_channels][result_pitch] f32 discriminator samples (demod_state.result),
 *                  result_len == block_pairs per block
 * Results are bit-identical to the reference's float output for the selected `fir_arith`, with the rest of the
 * chain compiled without -ffast-math / FMA contraction (the reference's default build type).
 */
int dsdneo_b200_full_demod_batch(dsdneo_b200_demod_bank* bank, const float* d_iq, size_t iq_pitch_pairs,
                                 int block_pairs, int n_blocks, float* d_result, size_t result_pitch, void* stream);
/** The same on cu8 IQ ([n_channels][iq_pitch_pairs] uchar2): widen_u8_to_f32_bias127 (src/dsp/simd_widen.cpp:139-147) fused into
 *  the channel filter's loads, bit-identical to widening first (2 B per pair through HBM instead of 8). */
int dsdneo_b200_full_demod_batch_cu8(dsdneo_b200_demod_bank* bank, const uint8_t* d_iq_u8, size_t iq_pitch_pairs, int block_pairs,
                                     int n_blocks, float* d_result, size_t result_pitch, void* stream);

/** Same, with host buffers: H2D copy, kernels, D2H copy, synchronous. */
int dsdneo_b200_full_demod_batch_host(dsdneo_b200_demod_bank* bank, const float* h_iq, size_t iq_pitch_pairs,
                                      int block_pairs, int n_blocks, float* h_result, size_t result_pitch);

/* ---- CQPSK symbol output kind of full_demod ------------------------ */

/**
 * The reference's OP25-style CQPSK chain for P25 LSM / simulcast and P25 Phase 2 channels, batched over a bank of
 * channels.  Per channel and per block this is what full_demod() runs when `cqpsk_enable = 1` and
 * `output_kind = DSD_DEMOD_OUTPUT_SYMBOL_CQPSK` (src/dsp/demod_pipeline.cpp:1100-1117,1330-1350):
 *   cqpsk_rms_agc            src/dsp/demod_pipeline.cpp:796-842      (alpha 0.45, reference 0.85)
 *   op25_fll_band_edge_cc    src/dsp/costas.cpp:1176-1224            (2 sps + 1 band-edge taps, loop_bw 2 pi / sps / 350)
 *   op25_gardner_cc          src/dsp/costas.cpp:804-858              (8-tap MMSE interpolator, src/dsp/mmse_interp.cpp)
 *   op25_diff_phasor_cc      src/dsp/costas.cpp:872-902
 *   op25_costas_loop_cc      src/dsp/costas.cpp:935-961              (symbol rate, loop_bw 0.008, phase clamped to +-pi/2)
 *   qpsk_differential_demod  src/dsp/demod_pipeline.cpp:742-764      (4/pi atan approximation -> {-3,-1,+1,+3})
 * A squelched block yields ceil(block_pairs / sps) zero symbols and leaves every loop untouched (:1022-1040).
 *
 * Carried per channel: AGC average, FLL {phase, freq, delay line}, ted_state_t {mu, omega, last sample, lock
 * accumulator, delay line}, cqpsk_diff_prev, dsd_costas_loop_state_t {phase, freq, error, error_smooth}; the FLL and
 * Gardner delay lines are one ring of the FLL's outputs on the device.  Results (symbols, counts, state) are
 * bit-identical to the reference built without -ffast-math / FMA contraction.
 * Out of contract: non-finite input samples, blocks shorter than 4 pairs (the reference then skips timing recovery,
 * costas.cpp:811-813), a change of sps on a live channel (create a new bank), sps outside 2..10.
 */
typedef struct dsdneo_b200_cqpsk_bank dsdneo_b200_cqpsk_bank;

typedef struct dsdneo_b200_cqpsk_bank_config {
    int n_channels;
    int rate_out_hz;     /* demod_state.rate_out (24000 or 48000 in the reference's own configurations) */
    const int* ted_sps;  /* per channel demod_state.ted_sps (NULL: 5 everywhere = P25p1 at 24 kHz) */
    float ted_gain;      /* demod_state.ted_gain; <= 0 selects the OP25 default 0.025 (costas.cpp:145) */
    int ted_gain_is_set; /* demod_state.ted_gain_is_set: 1 disables the locked-loop gain 0.018 at >= 5500 sym/s */
} dsdneo_b200_cqpsk_bank_config;

/** Snapshot of one channel's carried CQPSK state (field names follow the reference structs). */
typedef struct dsdneo_b200_cqpsk_chan_state {
    float cqpsk_agc_avg;
    float fll_phase, fll_freq, fll_alpha, fll_beta;
    float ted_mu, ted_omega, ted_last_r, ted_last_j, ted_lock_accum;
    int ted_lock_count;
    float ted_effective_gain;
    float cqpsk_diff_prev_r, cqpsk_diff_prev_j;
    float costas_phase, costas_freq, costas_error, costas_error_smooth;
    int costas_err_avg_q14, costas_err_raw_avg_q14, costas_conf_avg_q14, costas_zero_conf_pct;
    int overflow; /* 1 if a block produced more symbols than dsdneo_b200_cqpsk_block_capacity() allows (never on finite input) */
} dsdneo_b200_cqpsk_chan_state;

dsdneo_b200_cqpsk_bank* dsdneo_b200_cqpsk_bank_create(const dsdneo_b200_cqpsk_bank_config* cfg);
void dsdneo_b200_cqpsk_bank_destroy(dsdneo_b200_cqpsk_bank* q);
/** Fresh-channel state: ted_init_state + first-call initialisation of the FLL / Gardner / Costas blocks. */
int dsdneo_b200_cqpsk_bank_reset(dsdneo_b200_cqpsk_bank* q, void* stream);
int dsdneo_b200_cqpsk_bank_get_state(dsdneo_b200_cqpsk_bank* q, int channel, dsdneo_b200_cqpsk_chan_state* out);
/** Band-edge filter taps of one channel (fll_band_edge_design_filter, costas.cpp:936-1024); returns n_taps. */
int dsdneo_b200_cqpsk_bank_get_fll_taps(dsdneo_b200_cqpsk_bank* q, int channel, float* lower_r, float* lower_i,
                                        float* upper_r, float* upper_i, int max_taps);
/** Host-side band-edge filter design (same arithmetic as the reference, libm sinf/cosf); returns n_taps or <0. */
int dsdneo_b200_fll_band_edge_design(int sps, float* lower_r, float* lower_i, float* upper_r, float* upper_i,
                                     int max_taps);
/** Symbols one block can produce per channel; `symbols_pitch` must be >= n_blocks * this value. */
int dsdneo_b200_cqpsk_block_capacity(int block_pairs, int min_sps);

/**
 * Batched twin of full_demod() for the CQPSK symbol output kind: channel LPF (profile P25_CQPSK unless configured
 * otherwise) + power / squelch from `bank`, then the chain above with the per-channel loop state of `q`.
 * @param d_iq       [n_channels][iq_pitch_pairs] cf32, n_blocks consecutive blocks of block_pairs samples per channel
 * @param d_symbols  [n_channels][symbols_pitch] f32: the blocks' symbols back to back (demod_state.result)
 * @param d_counts   [n_channels][n_blocks] int: result_len of every block
 */
int dsdneo_b200_full_demod_cqpsk_batch(dsdneo_b200_demod_bank* bank, dsdneo_b200_cqpsk_bank* q, const float* d_iq,
                                       size_t iq_pitch_pairs, int block_pairs, int n_blocks, float* d_symbols,
                                       size_t symbols_pitch, int* d_counts, void* stream);
int dsdneo_b200_full_demod_cqpsk_batch_host(dsdneo_b200_demod_bank* bank, dsdneo_b200_cqpsk_bank* q, const float* h_iq,
                                            size_t iq_pitch_pairs, int block_pairs, int n_blocks, float* h_symbols,
                                            size_t symbols_pitch, int* h_counts);

/**
 * Sample side behind the CQPSK chain: the reference's symbol-rate input path (dsd_rtl_stream_metrics_hooks.output_kind == 2),
 * i.e. what getDibitSoft() does per symbol when the stream already carries one float per symbol:
 *   symbol_try_rtl_symbol_rate_fast_path   src/dsp/dsd_symbol.c:1583-1625
 *   use_symbol with rf_mod == 1            src/core/frames/dsd_dibit.c:243-299 (rolling min / max tracker, ssize x msize windows)
 *   digitize                               :1018-1041: cqpsk_slice(symbol - center) + OP25 dibit orientation map
 *                                           (include/dsd-neo/core/p25_cqpsk_dibit.h) when `p25_slice`, else threshold regions
 *   soft metric                            :685-721 (CQPSK or standard ideals), cqpsk_reliability_raw x CQPSK SNR weight
 * Per channel: `negative` = is_four_level_neg_synctype(synctype), `p25_slice` = is_cqpsk_active() && P25 sync type
 * (dsd_dibit.c:950-961), `map_idx` = state->p25_cqpsk_dibit_map_idx (0..4).  `snr_cqpsk_db` = the SNR hook's value for the
 * weight (<= -50: none).  The DSD_NEO_CQPSK_SYNC_INV / _NEG debug switches are not supported.
 * Outputs use the layout of dsdneo_b200_symbolize_batch (dibits, reliability, LLR pairs); bit-exact.
 */
typedef struct dsdneo_b200_cqpsk_slicer dsdneo_b200_cqpsk_slicer;
dsdneo_b200_cqpsk_slicer* dsdneo_b200_cqpsk_slicer_create(int n_channels, int ssize, int msize);
void dsdneo_b200_cqpsk_slicer_destroy(dsdneo_b200_cqpsk_slicer* q);
int dsdneo_b200_cqpsk_slicer_reset(dsdneo_b200_cqpsk_slicer* q, void* stream);
int dsdneo_b200_cqpsk_slicer_set_class(dsdneo_b200_cqpsk_slicer* q, const uint8_t* h_negative, const uint8_t* h_p25_slice,
                                       const uint8_t* h_map_idx, double snr_cqpsk_db);
/** @param d_symbols [n_channels][symbols_pitch] f32 (dsdneo_b200_full_demod_cqpsk_batch output), d_n_symbols [n_channels] */
int dsdneo_b200_cqpsk_slice_batch(dsdneo_b200_cqpsk_slicer* q, const float* d_symbols, size_t symbols_pitch,
                                  const int* d_n_symbols, uint8_t* d_dibits, uint8_t* d_reliability, int16_t* d_llr,
                                  size_t out_pitch, void* stream);
/** {min, max, center, umid, lmid, minref, maxref, lastsample} of one channel */
int dsdneo_b200_cqpsk_slicer_get_state(dsdneo_b200_cqpsk_slicer* q, int channel, float* out8);

/* ---- IQ capture sidecar ("dsd-neo-iq" metadata of --iq-capture / --iq-replay): harness I/O for the block side ---- */

enum { DSDNEO_B200_IQ_CU8 = 1, DSDNEO_B200_IQ_CF32 = 2 }; /* dsd_iq_sample_format, include/dsd-neo/io/iq_types.h:37-41 */
/** The fields of dsd_iq_replay_config (include/dsd-neo/io/iq_replay.h:26-64) an ingest needs. */
typedef struct dsdneo_b200_iq_info {
    uint32_t version;                    /* 1: single segment, 2: with an "events" replay timeline */
    int32_t sample_format;               /* DSDNEO_B200_IQ_CU8 / _CF32 */
    uint32_t sample_rate_hz;
    uint64_t center_frequency_hz, capture_center_frequency_hz, data_bytes;
    uint32_t base_decimation, post_downsample, demod_rate_hz;
    int32_t offset_tuning_enabled, fs4_shift_enabled, historical_cu8_two_pass /* combine_rotate_enabled == false */;
    int32_t muted_bytes_excluded, contains_retunes, size_limit_reached;
    uint32_t capture_retune_count, event_count;
    char data_file[256];
    char capture_stage[64];
} dsdneo_b200_iq_info;
/**
 * Parses sidecar JSON text the way dsd_iq_replay_read_metadata does (src/io/iq/iq_replay.c:1520-1790: one flat object, the
 * required fields, the value checks of :675-720; the v2 "events" array is counted, not interpreted).  0 on success,
 * DSDNEO_B200_EINVAL for malformed / incomplete / inconsistent metadata, DSDNEO_B200_EUNSUPPORTED for a sample format no
 * kernel here takes (cs16).  Host code; no device needed.
 */
int dsdneo_b200_iq_sidecar_parse(const char* json, size_t len, dsdneo_b200_iq_info* out);
/** dsd_iq_replay_compute_effective_bytes (iq_replay.c:1859-1882): the replayable byte count, whole samples only. */
long long dsdneo_b200_iq_effective_bytes(const dsdneo_b200_iq_info* info, uint64_t actual_file_size, int* size_mismatch);

/* ---- K2 (+K1): polyphase FIR channelizer -------------------------------------------------------- */

/**
 * Wideband complex IQ -> n_channels narrowband channels (critically sampled: every channel runs at
 * fs_in / n_channels).  The reference has no channelizer (it tunes ONE channel with a half-band cascade,
 * src/dsp/demod_pipeline.cpp:983-1001); this is the many-channel front end the north star adds.  Definition
 * (also the float64 oracle, oracle/oracle_dsp.c:oracle_pfb_direct):
 *     y_k[n] = sum_{m<L} h[m] x[t_n - m] exp(-j 2 pi k (t_n - m) / M),  t_n = n M + M - 1,  L = M * taps_per_branch
 * Channel k is centred at k * fs_in / M (k > M/2 are the negative frequencies).
 * With input_is_cu8 the input is unsigned 8-bit I/Q pairs widened on load exactly like
 * widen_u8_to_f32_bias127 (src/dsp/simd_widen.cpp:139-147): (u8 - 127.5f) * (1/127.5f).
 * Carried state: the last (taps_per_branch-1)*M input samples.
 * n_channels: a power of two, 256 ... 8192 (4096 x 12.5 kHz = 51.2 MHz, 8192 x 6.25 kHz = 51.2 MHz: BASELINE configs
 * C4 / C5); taps_per_branch 4, 8, 12 or 16.
 */
typedef struct dsdneo_b200_channelizer dsdneo_b200_channelizer;

/** Default prototype: Blackman-windowed sinc, cutoff = cutoff_rel * fs_in/(2M), unit DC gain. Returns L. */
int dsdneo_b200_channelizer_design_prototype(int n_channels, int taps_per_branch, double cutoff_rel, float* h_out);
/** prototype == NULL selects the default design with cutoff_rel = 1.0. */
dsdneo_b200_channelizer* dsdneo_b200_channelizer_create(int n_channels, int taps_per_branch, int input_is_cu8,
                                                        const float* prototype);
void dsdneo_b200_channelizer_destroy(dsdneo_b200_channelizer* c);
int dsdneo_b200_channelizer_reset(dsdneo_b200_channelizer* c, void* stream);
int dsdneo_b200_channelizer_get_prototype(dsdneo_b200_channelizer* c, float* h_out, int max_taps);
/** Moves the carried history over n_in_samples input samples (device, the channelizer's input format) without producing
 *  output: resuming a stream whose preceding samples are known, so that the first outputs carry no filter start-up. */
int dsdneo_b200_channelizer_prime(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, void* stream);
/**
 * @param d_in   n_in_samples wideband samples (cf32 pairs, or cu8 pairs), n_in_samples % n_channels == 0
 * @param d_out  [n_channels][out_pitch_pairs] cf32; n_in_samples / n_channels outputs per channel -- the layout
 *               dsdneo_b200_full_demod_batch() consumes directly.
 */
int dsdneo_b200_channelize(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, float* d_out,
                           size_t out_pitch_pairs, void* stream);
int dsdneo_b200_channelize_host(dsdneo_b200_channelizer* c, const void* h_in, size_t n_in_samples, float* h_out,
                                size_t out_pitch_pairs);
/**
 * Bin-pruned form for channel sharding (one wideband stream, the raw tile broadcast to every GPU, each
 * GPU demodulates its own channels): computes only the channels k = bin_first (mod bin_stride), bin_stride a power of
 * two <= 32 with n_channels / bin_stride >= 256.  Row k' of d_out ([n_channels / bin_stride][out_pitch_pairs]) is channel
 * bin_stride * k' + bin_first.  The branch filters still see the whole tile; the transform and the output shrink by
 * bin_stride.  advance != 0 moves the carried history past this tile (pass 0 for all but the last call when one GPU
 * computes several bin classes of the same tile).  dsdneo_b200_channelize() == bin_stride 1, bin_first 0, advance 1.
 */
int dsdneo_b200_channelize_bins(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, int bin_stride,
                                int bin_first, int advance, float* d_out, size_t out_pitch_pairs, void* stream);
/**
 * The same with the channels re-quantised to cu8 on the way out -- u8 = clamp(round(y * gain * 127.5 + 127.5), 0, 255), the
 * inverse of widen_u8_to_f32_bias127 -- i.e. in the reference's native per-channel capture format: d_out is
 * [n_channels / bin_stride][out_pitch_pairs] uchar2, the input dsdneo_b200_p25p1_rx_* / dsdneo_b200_full_demod_* take with
 * input_cu8 = 1 (2 B per sample through HBM instead of 8).  `gain` scales a channel into the 8-bit range (a channel of a
 * band of M equally loaded channels sits 1/sqrt(M) below the wideband level).  At most 4096 output channels,
 * out_pitch_pairs a multiple of 4.
 */
int dsdneo_b200_channelize_bins_cu8(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, int bin_stride,
                                    int bin_first, int advance, float gain, uint8_t* d_out, size_t out_pitch_pairs, void* stream);

/* ---- FSK front end: channelizer -> full_demod, one call ----------------------------------------- */

/**
 * The whole block side for N channels: what the reference's demod thread does per process
 * (src/io/radio/rtl_sdr_fm.cpp:3458-3512: take a block, call full_demod(), publish result[]), fed from one
 * shared wideband stream instead of one tuner per channel.
 */
typedef struct dsdneo_b200_frontend dsdneo_b200_frontend;

typedef struct dsdneo_b200_frontend_config {
    int n_channels;         /* M */
    int taps_per_branch;    /* channelizer prototype length / M (4, 8, 12 or 16) */
    int input_is_cu8;       /* 0: cf32 wideband input, 1: cu8 (RTL-SDR native), widened on load */
    int wideband_rate_hz;   /* channel rate = wideband_rate_hz / n_channels (demod_state.rate_out) */
    int block_pairs;        /* complex samples per channel per reference block (one full_demod() call) */
    const float* prototype; /* optional M*taps_per_branch taps, NULL = default design */
    int channel_lpf_enable;
    const int* channel_lpf_profile;     /* per channel or NULL */
    const float* channel_squelch_level; /* per channel or NULL */
    int fir_arith;                      /* DSDNEO_FIR_ARITH_* */
} dsdneo_b200_frontend_config;

dsdneo_b200_frontend* dsdneo_b200_frontend_create(const dsdneo_b200_frontend_config* cfg);
void dsdneo_b200_frontend_destroy(dsdneo_b200_frontend* fe);
int dsdneo_b200_frontend_reset(dsdneo_b200_frontend* fe, void* stream);
/** Borrow the demod bank (for dsdneo_b200_demod_bank_get_state). */
dsdneo_b200_demod_bank* dsdneo_b200_frontend_bank(dsdneo_b200_frontend* fe);
/**
 * @param d_wideband n_in_samples wideband samples on the device; n_in_samples % (n_channels*block_pairs) == 0
 * @param d_result   [n_channels][result_pitch] f32 discriminator samples, n_in_samples/n_channels per channel
 */
int dsdneo_b200_frontend_process(dsdneo_b200_frontend* fe, const void* d_wideband, size_t n_in_samples, float* d_result,
                                 size_t result_pitch, void* stream);
/**
 * Pipelined form for back-to-back tiles: the time-parallel stages (channelizer, channel LPF + phase
 * discriminator) of call i+1 overlap the time-serial recurrence stage of call i on internal streams.
 * Inputs must be ready on `stream` when called; results are complete after dsdneo_b200_frontend_join()
 * has been ordered into a stream (call it once after a run of process_async calls).  Bit-identical output.
 */
int dsdneo_b200_frontend_process_async(dsdneo_b200_frontend* fe, const void* d_wideband, size_t n_in_samples,
                                       float* d_result, size_t result_pitch, void* stream);
int dsdneo_b200_frontend_join(dsdneo_b200_frontend* fe, void* stream);
/** Host buffers (pinned recommended): per-block H2D / kernels / D2H pipelined on three streams; synchronous. */
int dsdneo_b200_frontend_process_host(dsdneo_b200_frontend* fe, const void* h_wideband, size_t n_in_samples,
                                      float* h_result, size_t result_pitch);
/**
 * Streaming form of process_host, the shape of the reference's demod thread (src/io/radio/rtl_sdr_fm.cpp:3458-3512:
 * blocks stream from the input ring through full_demod into the output ring): queues one tile and returns a ticket
 * (>= 0, or a negative DSDNEO_B200_E* code) without waiting, so the PCIe transfers and kernels of consecutive tiles
 * overlap.  Both host buffers must stay valid and untouched until wait_host(ticket) returns.  At most four tickets may
 * be outstanding.  Bit-identical to process_host.
 */
long long dsdneo_b200_frontend_submit_host(dsdneo_b200_frontend* fe, const void* h_wideband, size_t n_in_samples,
                                           float* h_result, size_t result_pitch);
int dsdneo_b200_frontend_wait_host(dsdneo_b200_frontend* fe, long long ticket);

/* ---- K12: frame-sync correlator, batched over channels ------------------------------------------------------- */
/*
 * The per-symbol part of getFrameSync() (src/dsp/dsd_frame_sync.c:3098-3148) for FSK inputs: hunt-time slice
 * symbol > 0 -> '1' else '3' (:2110-2127), rolling window, strcmp against the sync strings of
 * include/dsd-neo/core/sync_patterns.h:33-67.  Patterns are given as the reference's own strings ('1'/'3', 8..32 symbols,
 * oldest first) with the DSD_SYNC_* id (include/dsd-neo/core/synctype_ids.h) to report; at one position the first
 * matching pattern in table order wins, as in the reference's ordered matcher list.  A pattern can only match once as
 * many symbols as its length have been seen since create/reset (history carries across launches).
 * Everything else getFrameSync does (protocol gating by opts, threshold warm start, timeouts) is host control flow.
 */
#define DSDNEO_B200_SYNC_MAX_PATTERNS 32
typedef struct {
    const char* symbols; /* e.g. "111113113311333313133333" (P25P1_SYNC) */
    int sync_type;       /* e.g. 0 (DSD_SYNC_P25P1_POS) */
} dsdneo_b200_sync_pattern;
typedef struct {
    int position;  /* index (within the launch) of the symbol that completes the pattern */
    int sync_type;
} dsdneo_b200_sync_hit;
typedef struct dsdneo_b200_frame_sync dsdneo_b200_frame_sync;
dsdneo_b200_frame_sync* dsdneo_b200_frame_sync_create(int n_channels, const dsdneo_b200_sync_pattern* patterns, int n_patterns);
void dsdneo_b200_frame_sync_destroy(dsdneo_b200_frame_sync* fs);
int dsdneo_b200_frame_sync_reset(dsdneo_b200_frame_sync* fs, void* stream);
/**
 * @param d_symbols   [n_channels][pitch] symbol-rate floats (dsdneo_b200_symbolizer output, getSymbol() values)
 * @param d_n_symbols [n_channels] valid symbols per channel in this launch
 * @param d_hits      [n_channels][max_hits] hits in stream order; @param d_n_hits [n_channels] hits found (may exceed
 *                    max_hits, in which case only the first max_hits were stored)
 */
int dsdneo_b200_frame_sync_search_batch(dsdneo_b200_frame_sync* fs, const float* d_symbols, size_t pitch, const int* d_n_symbols,
                                        dsdneo_b200_sync_hit* d_hits, int max_hits, int* d_n_hits, void* stream);

/* ---- K3: complex half-band decimator cascade, batched over channels ------------------------------------ */
/*
 * Replaces full_demod_apply_halfband_decimation (src/dsp/demod_pipeline.cpp:983-1001) = `passes` calls of
 * simd_hb_decim2_complex (src/dsp/simd_fir.cpp:363-373; stage 0 uses hb31_q15_taps, later stages hb_q15_taps,
 * src/dsp/halfband.cpp:35-74), with the per-stage histories of struct demod_state (hb_hist_i/q[10][30]) kept on the device.
 * Input  [n_channels][in_pitch_pairs] cf32, n_blocks reference blocks of block_pairs each (the reference pads the right
 * edge of every block with its last sample, so block boundaries are part of the result);
 * output [n_channels][out_pitch_pairs] cf32, block_pairs >> passes pairs per block.
 * fir_arith selects the reference kernel whose arithmetic is reproduced bit for bit (DSDNEO_FIR_ARITH_*); as in the
 * reference, blocks shorter than the tap count always take the unfused kernel.  block_pairs must be a multiple of 2^passes.
 */
typedef struct dsdneo_b200_hb_cascade dsdneo_b200_hb_cascade;
dsdneo_b200_hb_cascade* dsdneo_b200_hb_cascade_create(int n_channels, int passes, int fir_arith);
void dsdneo_b200_hb_cascade_destroy(dsdneo_b200_hb_cascade* h);
int dsdneo_b200_hb_cascade_reset(dsdneo_b200_hb_cascade* h, void* stream);
int dsdneo_b200_hb_cascade_decim_batch(dsdneo_b200_hb_cascade* h, const float* d_in, size_t in_pitch_pairs, int block_pairs,
                                       int n_blocks, float* d_out, size_t out_pitch_pairs, void* stream);
int dsdneo_b200_hb_cascade_decim_batch_host(dsdneo_b200_hb_cascade* h, const float* h_in, size_t in_pitch_pairs,
                                            int block_pairs, int n_blocks, float* h_out, size_t out_pitch_pairs);

/* ---- sample side: matched filter + getSymbol + use_symbol + digitize, batched over channels (K9-K11) ----- */

#define DSDNEO_B200_SYM_MAX_TAPS 256
#define DSDNEO_B200_SYM_MAX_FILTERS 8
/* Slot convention for the matched filters (the reference's five sps_fir objects, src/dsp/dsd_filters.c:325-345). */
enum {
    DSDNEO_SYM_FILTER_NONE = -1,
    DSDNEO_SYM_FILTER_P25 = 0,  /* p25_filter  */
    DSDNEO_SYM_FILTER_DMR = 1,  /* dmr_filter (also YSF, NXDN96) */
    DSDNEO_SYM_FILTER_NXDN = 2, /* nxdn_filter */
    DSDNEO_SYM_FILTER_DPMR = 3, /* dpmr_filter */
    DSDNEO_SYM_FILTER_M17 = 4,  /* m17_filter  */
};
enum {
    DSDNEO_SYM_MODE_GET_SYMBOL = 0,     /* twin of getSymbol(opts, state, have_sync): float symbols only */
    DSDNEO_SYM_MODE_GET_DIBIT_SOFT = 1, /* twin of getDibitSoft(): have_sync = 1, + use_symbol + digitize + soft metrics */
};

/** What the reference derives from state->synctype / state->lastsynctype for each symbol. */
typedef struct dsdneo_b200_sym_class {
    int filter;       /* DSDNEO_SYM_FILTER_*: symbol_apply_matched_filter, src/dsp/dsd_symbol.c:301-337 */
    int window_l;     /* left edge of the averaging window, 2 or 1: select_window_c4fm, dsd_symbol.c:197-211 (right edge is 2) */
    int track_minmax; /* use_symbol() threshold tracking (P25p1): src/core/frames/dsd_dibit.c:264 */
    int negative;     /* is_four_level_neg_synctype(synctype): dsd_dibit.c:915-935 */
    int rf_mod;       /* state->rf_mod: 0 C4FM (default), 2 GFSK (two-sample window, GFSK timing nudge, no sync clip: dsd_symbol.c:213-225,
                       * 347-358,428-434,480-487; what the reference's DMR / NXDN96 / M17 presets select, src/runtime/decode_mode.c) */
} dsdneo_b200_sym_class;
/** The reference's rules for P25p1 / DMR / YSF / M17 / X2-TDMA / none (ids: include/dsd-neo/core/synctype_ids.h). */
int dsdneo_b200_sym_class_from_synctype(int synctype, int lastsynctype, int use_cosine_filter, dsdneo_b200_sym_class* out);

typedef struct dsdneo_b200_symbolizer_config {
    int n_channels;
    int output_rate_hz;    /* discriminator sample rate (dsd_rtl_stream_metrics_hooks.output_rate_hz), e.g. 48000 */
    int symbol_rate_hz;    /* symbol_profile rate, e.g. 4800; samples per symbol = rate/symbol rate with remainder accumulator */
    int ssize, msize;      /* opts->ssize / opts->msize (0 = reference defaults 128 / 1024) */
    int use_cosine_filter; /* opts->use_cosine_filter */
    int n_filters;
    /* NORMALISED taps exactly as the reference's design_sps_fir() leaves them for this samples-per-symbol
     * (src/dsd_filters.c:94-170): dsd-neo passes its own coefficient tables; oldest tap first. */
    const float* filter_taps[DSDNEO_B200_SYM_MAX_FILTERS];
    int filter_len[DSDNEO_B200_SYM_MAX_FILTERS];
} dsdneo_b200_symbolizer_config;

typedef struct dsdneo_b200_symbol_out {
    float* d_symbols;       /* [n_channels][pitch] */
    uint8_t* d_dibits;      /* [n_channels][pitch]     (GET_DIBIT_SOFT) value returned by getDibitSoft */
    uint8_t* d_reliability; /* [n_channels][pitch]     dsd_dibit_soft_t.reliability (include/dsd-neo/core/dibit.h:24-27) */
    int16_t* d_llr;         /* [n_channels][pitch][2]  dsd_dibit_soft_t.llr */
    int32_t* d_count;       /* [n_channels] symbols produced by this call */
    size_t pitch;           /* capacity per channel; must be >= (n_samples + 256) / (samples_per_symbol - 1) + 2 */
} dsdneo_b200_symbol_out;

typedef struct dsdneo_b200_symbolizer dsdneo_b200_symbolizer;
dsdneo_b200_symbolizer* dsdneo_b200_symbolizer_create(const dsdneo_b200_symbolizer_config* cfg);
void dsdneo_b200_symbolizer_destroy(dsdneo_b200_symbolizer* y);
/** Back to initState() values (src/core/util/dsd_init.c:519-592) and empty matched-filter history. */
int dsdneo_b200_symbolizer_reset(dsdneo_b200_symbolizer* y, void* stream);
/** Per-channel class (host array of n_channels).  Synchronises the device. */
int dsdneo_b200_symbolizer_set_class(dsdneo_b200_symbolizer* y, const dsdneo_b200_sym_class* per_channel);
/**
 * Per-channel C4FM SNR in dB as dsd_rtl_stream_metrics_hooks.snr_c4fm_db reports it (host array of n_channels, or NULL
 * for the "no hook installed" sentinel of -100 dB on every channel, the default).  Selects the reliability weight
 * (204 + (w256 >> 2)) / 256 of apply_c4fm_snr_weight (src/core/frames/dsd_dibit.c:504-546).  Synchronises the device.
 */
int dsdneo_b200_symbolizer_set_snr(dsdneo_b200_symbolizer* y, const double* h_snr_c4fm_db);
/**
 * Consume n_samples discriminator samples per channel ([n_channels][disc_pitch] f32, the output layout of
 * dsdneo_b200_full_demod_batch) and emit every complete symbol; an unfinished symbol's samples are carried to the next
 * call.  With have_sync == 0 (GET_SYMBOL mode) the reference's +-1 sample jitter nudge is active, so channels may emit
 * different symbol counts.  Bit-identical to the reference for 4-level C4FM-family modes (rf_mod == 0).
 */
int dsdneo_b200_symbolize_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int mode,
                                int have_sync, const dsdneo_b200_symbol_out* out, void* stream);

/*
 * Acquisition on the device: getFrameSync() from the never-synchronised state (src/dsp/dsd_frame_sync.c:3098-3148) for the
 * FSK discriminator path, per channel and bit-exact.  A hunting channel takes RAW discriminator samples (no matched filter
 * before the first sync, dsd_symbol.c:301-337) through getSymbol(have_sync = 0) with the timing nudges, keeps the 24-symbol
 * level ring and the rolling payload dibits / reliabilities (:1747-1764, :2161-2190), slices the hunt window ('1' when the
 * symbol is positive, else '3', :2110-2127) and compares it with the configured 24-symbol sync patterns from the 8th symbol on
 * (:2638-2676).  On a match: frame_sync_set_basic_lock (:386-392); the sync warm start of the slicer thresholds
 * (dsd_sync_warm_start_thresholds_outer_only(24), src/dsp/sync_calibration.c:155-226) for P25 Phase 1 when rf_mod == 0 and
 * for DMR always; for DMR (kind 1) the 66 dibits in front of the sync re-sliced with the new thresholds
 * (dmr_resample_on_sync, src/dsp/dmr_sync.c:60-126); then the channel switches to the pattern's decoder class and continues,
 * inside the same launch, as a synchronised channel (getDibitSoft), with the matched filter starting from an all-zero delay
 * line exactly where the reference turns it on.  A hunt that saw 1800 symbols without sync restarts with an empty window
 * (:3039); the no-carrier housekeeping of that moment is the host's.
 * Out of scope here: the other sync families of getFrameSync (NXDN, YSF, dPMR, M17, D-STAR, ProVoice, EDACS, P25p2), the
 * CQPSK sync path, and re-acquisition policy after a lost sync (the host drops a channel back to hunting with _set_acquired).
 */
typedef struct dsdneo_b200_acq_pattern {
    const char* symbols;       /* 24 characters '1' / '3' (include/dsd-neo/core/sync_patterns.h) */
    int sync_type;             /* what getFrameSync returns for it (include/dsd-neo/core/synctype_ids.h) */
    int kind;                  /* 0 P25 Phase 1 rules, 1 DMR rules (warm start always + resample-on-sync) */
    dsdneo_b200_sym_class cls; /* decoder class in force after the sync (rf_mod is a channel property and is not switched) */
} dsdneo_b200_acq_pattern;

typedef struct dsdneo_b200_acq_info {
    int32_t acquired;     /* 1 = synchronised (now or earlier) */
    int32_t sync_type;    /* >= 0 only in the launch that found the sync */
    int32_t hit_index;    /* output index (this launch) of the last sync symbol; the frame body starts at hit_index + 1 */
    int32_t hunt_symbols; /* symbols this launch spent hunting */
    uint8_t warm_start;   /* the sync warm start replaced the thresholds (DSD_WARM_START_OK) */
    uint8_t resample_ok;  /* resampled[] is valid (DMR sync with >= 90 symbols of history) */
    uint8_t resampled[66]; /* DMR: the 66 dibits in front of the sync, re-sliced (dmr_resample_cach); those of them that lie
                            * in this launch's output are also rewritten in d_dibits */
} dsdneo_b200_acq_info;

/** Up to 8 patterns, compared in the given order (the reference tests P25p1 before DMR).  Allocates the hunt state. */
int dsdneo_b200_symbolizer_set_acquire_patterns(dsdneo_b200_symbolizer* y, const dsdneo_b200_acq_pattern* patterns, int n_patterns);
/** Per-channel hunting (0) / synchronised (1) flags, host array of n_channels or NULL for all hunting.  Synchronises. */
int dsdneo_b200_symbolizer_set_acquired(dsdneo_b200_symbolizer* y, const int* h_acquired);
/**
 * dsdneo_b200_symbolize_batch(GET_DIBIT_SOFT) for a bank in which some channels still hunt for sync.  d_disc are the RAW
 * discriminator samples; hunting channels emit their hunt symbols with the payload dibits / reliabilities of that moment,
 * synchronised channels what getDibitSoft returns.  d_info: [n_channels] device records for this launch (may be NULL).
 */
int dsdneo_b200_symbolize_acquire_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples,
                                        const dsdneo_b200_symbol_out* out, dsdneo_b200_acq_info* d_info, void* stream);
/**
 * The same for a hunt AFTER a channel's first sync: the reference keeps the sample-side matched filter of the last sync type
 * running while it hunts again (dsd_symbol.c:301-337 selects the filter from lastsynctype), so hunting channels read the
 * matched filter's output and nothing is restarted at the sync.  This is the form a stream should be acquired with in
 * practice: a cold hunt on raw samples locks onto symbol centres that move by the filter's group delay -- 4.5 symbols for the
 * 91-tap p25_filter at 10 samples per symbol -- the moment the filter switches on, and the reference itself only recovers from
 * that by losing the first frame and hunting again with the filter on.  Built from the hunting and hand-over rules pinned by
 * the cold form; not separately pinned.
 */
int dsdneo_b200_symbolize_reacquire_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples,
                                          const dsdneo_b200_symbol_out* out, dsdneo_b200_acq_info* d_info, void* stream);

/* ---- batched FEC leaves (K13, K16, K17, K19) ------------------------------------------------------ */

/*
 * Layouts are the reference's own: one bit per unsigned char (only the LSB is significant on input; in-place
 * correction flips the LSB and leaves the other bits of the byte alone, as src/fec/fec.c does), P25 hex words as six
 * bytes MSB first, LLRs as int16 (positive = bit 1).  Every item is decoded independently -- the reference's
 * nbCodewords argument is always 1 at its call sites.  No table initialisation call is
 * needed (the twin of InitAllFecFunction(), src/fec/fec.c:827, runs on first use).
 */
enum {
    DSDNEO_FEC_HAMMING_7_4 = 0, /* Hamming_7_4_decode      src/fec/fec.c:145-168 */
    DSDNEO_FEC_HAMMING_12_8,    /* Hamming_12_8_decode     src/fec/fec.c:188-230 (keeps going after a bad word) */
    DSDNEO_FEC_HAMMING_13_9,    /* Hamming_13_9_decode     src/fec/fec.c:253-302 */
    DSDNEO_FEC_HAMMING_15_11,   /* Hamming_15_11_decode    src/fec/fec.c:327-376 */
    DSDNEO_FEC_HAMMING_16_11_4, /* Hamming_16_11_4_decode  src/fec/fec.c:402-450 */
    DSDNEO_FEC_GOLAY_20_8,      /* Golay_20_8_decode       src/fec/fec.c:543-587 (3 flips are applied but reported false) */
    DSDNEO_FEC_GOLAY_24_12,     /* Golay_24_12_decode      src/fec/fec.c:682-733 */
    DSDNEO_FEC_QR_16_7_6,       /* QR_16_7_6_decode        src/fec/fec.c:787-824 */
    DSDNEO_FEC_BLOCK_CODE_COUNT
};
int dsdneo_b200_fec_block_code_len(int code); /* n: bytes per codeword */
int dsdneo_b200_fec_block_code_k(int code);   /* k: information bits */
/**
 * @param d_bits    [n_words][n] codewords, corrected in place
 * @param d_decoded [n_words][k] information bits (Hamming codes other than 7_4; may be NULL).  Not written for a word
 *                  the reference reports uncorrectable before its memcpy (13_9, 15_11, 16_11_4).
 * @param d_ok      [n_words] the reference's bool return (1 = no error or corrected)
 */
int dsdneo_b200_fec_block_decode_batch(int code, uint8_t* d_bits, uint8_t* d_decoded, uint8_t* d_ok, int n_words, void* stream);
int dsdneo_b200_fec_block_decode_batch_host(int code, uint8_t* h_bits, uint8_t* h_decoded, uint8_t* h_ok, int n_words);
/** Golay_24_12_encode (src/fec/fec.c:670-680): [n][12] data bits -> [n][24] codeword bits. */
int dsdneo_b200_fec_golay_24_12_encode_batch(const uint8_t* d_data, uint8_t* d_out, int n_words, void* stream);

/**
 * Batched twins of `int dmr_r34_viterbi_decode(const uint8_t* dibits98, uint8_t out_bytes18[18])` and
 * `dmr_r34_viterbi_decode_soft(dibits98, reliab98, out_bytes18)` (include/dsd-neo/protocol/dmr/r34_viterbi.h:33,46;
 * src/protocol/dmr/dmr_34_viterbi.c:402-474): DMR rate 3/4 trellis, 98 dibits -> 18 payload bytes.
 * d_dibits98 [n][98] (0..3, as received); d_reliab98 [n][98] per-dibit reliabilities or NULL for the hard-decision decoder;
 * d_out18 [n][18].  Bit-exact including the tie-break (lowest previous state).
 */
int dsdneo_b200_dmr_r34_decode_batch(const uint8_t* d_dibits98, const uint8_t* d_reliab98, uint8_t* d_out18, int n_blocks,
                                     void* stream);
int dsdneo_b200_dmr_r34_decode_batch_host(const uint8_t* h_dibits98, const uint8_t* h_reliab98, uint8_t* h_out18, int n_blocks);

/**
 * Batched twin of rs_12_9_calc_syndrome + rs_12_9_check_syndrome + rs_12_9_correct_errors (include/dsd-neo/fec/rs_12_9.h:41-44;
 * src/fec/rs-12-9.c:237-323) as the reference's callers chain them: RS(12,9) over GF(2^8), 9 data + 3 checksum bytes.
 * d_codewords [n][12] corrected in place; d_syndrome3 [n][3] (may be NULL); d_result[i]: 0 = syndrome zero (corrector not run),
 * 1 = RS_12_9_CORRECT_ERRORS_RESULT_NO_ERRORS_FOUND, 2 = _ERRORS_CORRECTED, 3 = _ERRORS_CANT_BE_CORRECTED;
 * d_errors_found[i] = the reference's *errors_found.
 */
int dsdneo_b200_rs_12_9_decode_batch(uint8_t* d_codewords, uint8_t* d_syndrome3, uint8_t* d_result, uint8_t* d_errors_found,
                                     int n_words, void* stream);
int dsdneo_b200_rs_12_9_decode_batch_host(uint8_t* h_codewords, uint8_t* h_syndrome3, uint8_t* h_result, uint8_t* h_errors_found,
                                          int n_words);

/**
 * BPTCDeInterleaveDMRData + BPTC_196x96_Extract_Data (src/fec/bptc.c:51-59,136-149; include/dsd-neo/fec/bptc.h).
 * @param d_in        [n][196] burst bits; interleaved != 0: as received (de-interleave fused), else already de-interleaved
 * @param d_out96     [n][96] payload bits;  d_r3 [n][3] reserved bits R(2..0) (may be NULL)
 * @param d_errs      [n] number of uncorrectable Hamming lines in the second pass (the reference's return value)
 * The reference copies a stale callee buffer when a Hamming(13,9) column is uncorrectable: the previous good column
 * of the same pass is reproduced; when there is none (first column) the reference reads uninitialised stack and this
 * library leaves the column unchanged.
 */
int dsdneo_b200_bptc_196x96_batch(const uint8_t* d_in, int interleaved, uint8_t* d_out96, uint8_t* d_r3, uint32_t* d_errs,
                                  int n_bursts, void* stream);
int dsdneo_b200_bptc_196x96_batch_host(const uint8_t* h_in, int interleaved, uint8_t* h_out96, uint8_t* h_r3, uint32_t* h_errs,
                                       int n_bursts);

/**
 * BPTC_128x77_Extract_Data (src/fec/bptc.c:167-252): in = the 8 x 16 byte-per-bit matrix (row-major, only the LSB of
 * each byte is used), out = 77 bytes (72 data + 5 CRC bits), errs = uncorrectable Hamming(16,11,4) rows + column-parity
 * failures, exactly the reference's return value.  An uncorrectable row takes the information bits of the most recent
 * correctable row, as the reference's stale callee buffer does; where the reference reads an uninitialised buffer (no
 * correctable row before it) the row is left unchanged.
 */
int dsdneo_b200_bptc_128x77_batch(const uint8_t* d_in128, uint8_t* d_out77, uint32_t* d_errs, int n_items, void* stream);
int dsdneo_b200_bptc_128x77_batch_host(const uint8_t* h_in128, uint8_t* h_out77, uint32_t* h_errs, int n_items);
/**
 * BPTC_16x2_Extract_Data (src/fec/bptc.c:272-333): in = 32 interleaved bits, out = 32 de-interleaved bits with the
 * first 11 Hamming(16,11,4)-corrected, errs = (uncorrectable ? 1 : 0) + parity mismatches (odd or even rule).
 */
int dsdneo_b200_bptc_16x2_batch(const uint8_t* d_in32, uint8_t* d_out32, uint32_t* d_errs, int parity_odd, int n_items, void* stream);
int dsdneo_b200_bptc_16x2_batch_host(const uint8_t* h_in32, uint8_t* h_out32, uint32_t* h_errs, int parity_odd, int n_items);

/** p25_12_candidate_t (include/dsd-neo/protocol/p25/p25_12.h:13-17), same layout. */
typedef struct dsdneo_b200_p25_12_candidate {
    uint8_t bytes[12];
    uint32_t metric;
} dsdneo_b200_p25_12_candidate;
/** p25_12_soft_llr (src/protocol/p25/p25_12.c:204-283): [n][196] LLRs -> [n][12] bytes, [n] return values (metric >> 8). */
int dsdneo_b200_p25_12_soft_llr_batch(const int16_t* d_llr196, uint8_t* d_out12, int32_t* d_metric, int n_blocks, void* stream);
int dsdneo_b200_p25_12_soft_llr_batch_host(const int16_t* h_llr196, uint8_t* h_out12, int32_t* h_metric, int n_blocks);
/** p25_12_soft_llr_list (src/protocol/p25/p25_12.c:144-202): d_cands is [n][8] (P25_12_MAX_CANDIDATES slots per block,
 *  the first d_count[i] valid, sorted by metric); max_candidates is clamped to 8 like the reference. */
int dsdneo_b200_p25_12_soft_llr_list_batch(const int16_t* d_llr196, dsdneo_b200_p25_12_candidate* d_cands, int32_t* d_count,
                                           int max_candidates, int n_blocks, void* stream);
int dsdneo_b200_p25_12_soft_llr_list_batch_host(const int16_t* h_llr196, dsdneo_b200_p25_12_candidate* h_cands,
                                                int32_t* h_count, int max_candidates, int n_blocks);

enum {
    DSDNEO_P25_RS_36_20_17 = 0, /* check_and_fix_redsolomon_36_20_17   phase1/p25p1_check_hdu.cpp:38-45 */
    DSDNEO_P25_RS_24_12_13 = 1, /* check_and_fix_reedsolomon_24_12_13  phase1/p25p1_check_ldu.cpp:37-44 */
    DSDNEO_P25_RS_24_16_9 = 2,  /* check_and_fix_reedsolomon_24_16_9   phase1/p25p1_check_ldu.cpp:55-62 */
};
/**
 * RS(63,k) over GF(64) shortened to the P25 sizes (engine include/dsd-neo/fec/ReedSolomon.hpp:61-816).
 * @param d_data_bits   [n][k][6]  data hex words, corrected in place
 * @param d_parity_bits [n][n-k][6]
 * @param d_status      [n] 0 = ok / corrected, 1 = irrecoverable (data left as received)
 */
int dsdneo_b200_p25_rs_decode_batch(int variant, uint8_t* d_data_bits, const uint8_t* d_parity_bits, uint8_t* d_status,
                                    int n_words, void* stream);
int dsdneo_b200_p25_rs_decode_batch_host(int variant, uint8_t* h_data_bits, const uint8_t* h_parity_bits, uint8_t* h_status,
                                         int n_words);

#define DSDNEO_P25P1_SOFT_ERASURE_THRESHOLD 64 /* P25P1_SOFT_ERASURE_THRESHOLD, phase1/p25p1_soft.cpp:20 */
/**
 * check_and_fix_redsolomon_36_20_17_soft / check_and_fix_reedsolomon_24_12_13_soft / _24_16_9_soft
 * (phase1/p25p1_check_hdu.cpp:47-54, p25p1_check_ldu.cpp:46-71 -> DSDReedSolomon_*::decode_soft, ReedSolomon.hpp:879-913):
 * hard decode first, then one bounded errors-and-erasures decode (ReedSolomon_63::decode_with_erasures, :621-683,773-795)
 * with the caller's erasure positions in codeword space (parity symbols 0..n-k-1, data after them).
 * @param d_erasures   [n][erasure_pitch] positions, @param d_n_erasures [n] how many of them are valid (1..2t)
 * @param d_status     [n] 0 = ok / corrected, 1 = irrecoverable; data bits are rewritten as 0/1 either way, as the
 *                     reference's hard decoder does
 */
int dsdneo_b200_p25_rs_decode_erasures_batch(int variant, uint8_t* d_data_bits, const uint8_t* d_parity_bits,
                                             const int32_t* d_erasures, int erasure_pitch, const int32_t* d_n_erasures,
                                             uint8_t* d_status, int n_words, void* stream);
int dsdneo_b200_p25_rs_decode_erasures_batch_host(int variant, uint8_t* h_data_bits, const uint8_t* h_parity_bits,
                                                  const int32_t* h_erasures, int erasure_pitch, const int32_t* h_n_erasures,
                                                  uint8_t* h_status, int n_words);
/**
 * p25p1_rs_36_20_17_soft_reliability / p25p1_rs_24_16_9_soft_reliability (phase1/p25p1_check_hdu.cpp:56-77,
 * p25p1_check_ldu.cpp:73-94; the same rule is offered for the (24,12,13) shape): every symbol is ranked by
 * (reliability, position) with parity positions numbered first (p25p1_build_rs_ranked_erasures, p25p1_soft.cpp:140-170),
 * the count is max(#symbols below erasure_threshold, t) capped at 2t, and n = 1..count weakest symbols are tried as
 * erasures in order; the first success wins.  Reliabilities are the per-symbol minima of |LLR| clamped to 0..255
 * (p25p1_llr_reliability, p25p1_soft.cpp:64-79).  On failure the data bits are left untouched, as in the reference.
 * erasure_threshold: pass DSDNEO_P25P1_SOFT_ERASURE_THRESHOLD unless DSD_NEO_P25P1_SOFT_ERASURE_THRESHOLD overrides it.
 */
int dsdneo_b200_p25_rs_soft_reliability_batch(int variant, uint8_t* d_data_bits, const uint8_t* d_parity_bits,
                                              const uint8_t* d_data_reliab, const uint8_t* d_parity_reliab, int erasure_threshold,
                                              uint8_t* d_status, int n_words, void* stream);
int dsdneo_b200_p25_rs_soft_reliability_batch_host(int variant, uint8_t* h_data_bits, const uint8_t* h_parity_bits,
                                                   const uint8_t* h_data_reliab, const uint8_t* h_parity_reliab,
                                                   int erasure_threshold, uint8_t* h_status, int n_words);

/* ---- P25 Phase 1 word codes around the RS decoders: Golay(24,6) / (24,12), Hamming(10,6,3), NID BCH(63,16,11) ---- */
enum {
    DSDNEO_P25_WORD_GOLAY_24_6 = 0,     /* check_and_fix_golay_24_6   phase1/p25p1_check_hdu.cpp:26-30, Golay24.hpp:336-368 */
    DSDNEO_P25_WORD_GOLAY_24_12 = 1,    /* check_and_fix_golay_24_12  phase1/p25p1_check_hdu.cpp:32-36, Golay24.hpp:370-404 */
    DSDNEO_P25_WORD_HAMMING_10_6_3 = 2, /* hamming_10_6_3_decode      src/fec/hamming_10_6_3.cpp:90-105 */
};
/**
 * @param d_data_bits   [n][6 | 12 | 6] byte-per-bit words (MSB first), corrected in place exactly where the reference does
 * @param d_parity_bits [n][12 | 12 | 4]
 * @param d_status      Golay: 0 ok / 1 uncorrectable or non-binary input (word untouched); Hamming: 0 clean / 1 corrected /
 *                      2 uncorrectable or non-binary input
 * @param d_fixed       optional [n]: Golay = the reference's *fixed_errors (weight of the last syndrome examined, also on
 *                      failure); Hamming = 1 when a bit was corrected
 */
int dsdneo_b200_p25_word_decode_batch(int code, uint8_t* d_data_bits, const uint8_t* d_parity_bits, uint8_t* d_status,
                                      int32_t* d_fixed, int n_words, void* stream);
int dsdneo_b200_p25_word_decode_batch_host(int code, uint8_t* h_data_bits, const uint8_t* h_parity_bits, uint8_t* h_status,
                                           int32_t* h_fixed, int n_words);
/**
 * BCH_63_16_11::decode_with_result (include/dsd-neo/fec/BCH_63_16.hpp:288-329; caller p25p1_nid_decode,
 * phase1/p25p1_check_nid.cpp:246-305): 63 byte-per-bit inputs (16 data bits MSB first, then 47 parity bits) -> 16 corrected
 * data bits, ok flag (1 = success, up to 11 bit errors) and the number of corrected bits (0 on failure; output untouched).
 */
int dsdneo_b200_bch_63_16_decode_batch(const uint8_t* d_in63, uint8_t* d_out16, uint8_t* d_ok, int32_t* d_err_count, int n_words,
                                       void* stream);
int dsdneo_b200_bch_63_16_decode_batch_host(const uint8_t* h_in63, uint8_t* h_out16, uint8_t* h_ok, int32_t* h_err_count,
                                            int n_words);

/**
 * Batched twin of `struct p25p1_nid_result p25p1_nid_decode(const char bch_code[63], const uint8_t reliab63[63],
 * int observed_nac, unsigned char parity, uint8_t parity_reliab)` (include/dsd-neo/protocol/p25/p25p1_check_nid.h:76,
 * src/protocol/p25/phase1/p25p1_check_nid.cpp:322-354): hard BCH(63,16,11) decode + DUID / parity validation, one retry
 * with the known NAC written over the received one after a BCH failure, then the bounded Chase search (at most 3 flips among
 * the <= 8 least reliable positions, from the received word and from the NAC-rewritten word).
 * d_reliab63 / d_observed_nac / d_parity_reliab may be NULL (hard decode only / no known NAC / reliability 0).
 * `erasure_threshold` = p25p1_get_erasure_threshold() (64 unless configured, p25p1_soft.cpp:20-39).
 * d_status[i] is enum NidResult (0 fail, 1 ok, 2 parity override); nac / duid / error_count as in the reference struct.
 */
int dsdneo_b200_p25p1_nid_decode_batch(const uint8_t* d_code63, const uint8_t* d_reliab63, const int32_t* d_observed_nac,
                                       const uint8_t* d_parity, const uint8_t* d_parity_reliab, int erasure_threshold,
                                       int8_t* d_status, int32_t* d_nac, uint8_t* d_duid, int32_t* d_error_count,
                                       int n_words, void* stream);
int dsdneo_b200_p25p1_nid_decode_batch_host(const uint8_t* h_code63, const uint8_t* h_reliab63,
                                            const int32_t* h_observed_nac, const uint8_t* h_parity,
                                            const uint8_t* h_parity_reliab, int erasure_threshold, int8_t* h_status,
                                            int32_t* h_nac, uint8_t* h_duid, int32_t* h_error_count, int n_words);

/* ---- symbol stream: the launches of one channel joined into one stream --------------------------------------- */
/*
 * The reference's frame readers work on ONE sequential dibit stream per channel: they look back into dibits already taken
 * (DMR: 90 dibits from a rolling buffer at every sync, src/protocol/dmr/dmr_data.c:118-157, src/protocol/dmr/dmr_bs.c:137-148)
 * and keep reading past the end of whatever block of samples arrived.  A symbol stream keeps the last `keep` symbols, dibits,
 * reliabilities and LLRs of every channel on the device in front of the next launch's outputs, so that the sync hunt and the
 * frame cutters see frames that straddle launches whole, once.  Per launch:
 *     dsdneo_b200_symbol_stream_begin(ss, &out);               // rows the slicer writes this launch's outputs into
 *     dsdneo_b200_symbolize_batch(y, d_disc, pitch, n, mode, have_sync, &out, stream);
 *     dsdneo_b200_symbol_stream_commit(ss, &view, stream);     // history moved in front of them, positions counted
 *     // sync hunt `delay` symbols behind the slicer (delay >= the symbols a frame needs after its sync, keep >= delay + the
 *     // symbols it needs before it): every stream position is searched exactly once, with its whole frame present
 *     dsdneo_b200_frame_sync_search_batch(fs, view.d_symbols + (view.keep - delay), view.pitch, view.d_new, d_hits, max_hits, d_n_hits, stream);
 *     dsdneo_b200_sync_hits_rebase(d_hits, d_n_hits, n_channels, max_hits, view.keep - delay, stream);   // hit positions -> row indices
 *     dsdneo_b200_dmr_burst_cut_batch(view.d_dibits, view.pitch, view.d_reliability, view.pitch, view.d_valid, d_hits, ...);
 * Stream position of row index i of channel c = view.d_stream_base[c] + i (negative before the stream's first symbol; the rows
 * hold zeros there).  `max_new` must cover the symbols one launch can add per channel (the pitch rule of
 * dsdneo_b200_symbol_out for the longest launch).  DMR data bursts: delay >= 54, keep >= delay + 90 + 24 (e.g. 64 / 256);
 * DMR voice superframes of n_bursts bursts (dsdneo_b200_dmr_voice_cut_batch): delay >= (n_bursts - 1) * 144 + 54, e.g.
 * 800 / 1024 for six bursts.  The P25 Phase 1 receive bank has the same layout built in (1024 kept, 864 behind).
 */
typedef struct dsdneo_b200_symbol_stream dsdneo_b200_symbol_stream;
typedef struct dsdneo_b200_symbol_stream_view {
    const float* d_symbols;        /* [n_channels][pitch]: `keep` symbols of history, then this launch's */
    const uint8_t* d_dibits;       /* [n_channels][pitch] */
    const uint8_t* d_reliability;  /* [n_channels][pitch] */
    const int16_t* d_llr;          /* [n_channels][pitch][2] */
    size_t pitch;
    const int32_t* d_valid;        /* [n_channels] keep + symbols of this launch = valid row length */
    const int32_t* d_new;          /* [n_channels] symbols of this launch */
    const long long* d_stream_base; /* [n_channels] stream position of row index 0 */
    int keep;
} dsdneo_b200_symbol_stream_view;
/** `keep`: multiple of 32.  Rows are zero until written. */
dsdneo_b200_symbol_stream* dsdneo_b200_symbol_stream_create(int n_channels, int keep, int max_new);
void dsdneo_b200_symbol_stream_destroy(dsdneo_b200_symbol_stream* ss);
int dsdneo_b200_symbol_stream_reset(dsdneo_b200_symbol_stream* ss, void* stream);
int dsdneo_b200_symbol_stream_begin(dsdneo_b200_symbol_stream* ss, dsdneo_b200_symbol_out* out);
int dsdneo_b200_symbol_stream_commit(dsdneo_b200_symbol_stream* ss, dsdneo_b200_symbol_stream_view* view, void* stream);
/** Adds `offset` to the position of every reported hit (dsdneo_b200_sync_hit[n_channels][max_hits]). */
int dsdneo_b200_sync_hits_rebase(void* d_hits, const int32_t* d_n_hits, int n_channels, int max_hits, int offset, void* stream);
/** The hits of one sync type per channel, in stream order: what the reference's dispatch does when it hands a sync to the
 * handler of its type (src/engine/dispatch: BS DATA syncs -> dmr_data_sync, BS VOICE syncs -> dmrBSBootstrap).  d_hits_out is
 * dsdneo_b200_sync_hit[n_channels][out_max_hits]; hits beyond out_max_hits are dropped, d_n_out[c] <= out_max_hits. */
int dsdneo_b200_sync_hits_select(const void* d_hits, const int32_t* d_n_hits, int n_channels, int max_hits, int sync_type,
                                 void* d_hits_out, int out_max_hits, int32_t* d_n_out, void* stream);

/**
 * DMR base-station data burst cutter: the collection phase of `dmr_data_sync` (src/protocol/dmr/dmr_data.c:54-65,118-157,
 * 159-179,218-226,261-268) for every BS DATA sync hit of every channel -- 90 dibits back from the dibit after the sync
 * (12 CACH dibits de-interleaved with dmr_cach_interleave, 49 info dibits, 5 slot-type dibits, the sync) and 5 slot-type +
 * 49 info dibits after it.  Outputs per slot (= channel * max_hits + hit): CACH bits [24] (bits 0..6 = the TACT word for
 * Hamming(7,4)), info bits [196] in transmitted (interleaved) order = the input of dsdneo_b200_bptc_196x96_batch, per-dibit
 * reliabilities of the 98 info dibits, slot-type bits [20] = the input of Golay(20,8), and whether the channel's stream
 * (d_counts dibits) holds the whole burst.  `inverted_dmr` = opts->inverted_dmr (XOR 2 on the part before the sync's end).
 * The cutters see the dibit rows they are given: on one launch's rows a burst that straddles two launches is reported
 * valid = 0 by both (it needs 90 dibits before the end of its sync and 54 after).  A streaming caller hands them the joined
 * rows of a symbol stream (dsdneo_b200_symbol_stream_*, above: history kept on the device in front of every launch, sync hunt
 * 64 symbols behind the slicer), which cuts every burst once and whole however the stream is split into launches.
 */
int dsdneo_b200_dmr_burst_cut_batch(const uint8_t* d_dibits, size_t dibit_pitch, const uint8_t* d_reliability,
                                    size_t reliability_pitch, const int32_t* d_counts, const void* d_hits,
                                    const int32_t* d_n_hits, int n_channels, int max_hits, int inverted_dmr,
                                    uint8_t* d_cach24, uint8_t* d_info196, uint8_t* d_rel98, uint8_t* d_slot_type20,
                                    uint8_t* d_valid, void* stream);

/**
 * DMR base-station VOICE burst cutter: the collection phase of dmrBSBootstrap / dmrBS (src/protocol/dmr/dmr_bs.c:137-148,
 * 150-170, 182-187, 711-722, 745-746, 838-848) for every BS VOICE sync hit of every channel.  Burst j = 0 is the hit's own
 * burst (90 dibits back from the dibit after the sync: 12 CACH, 36 + 18 vocoder dibits, the sync, then 18 + 36 vocoder dibits
 * after it); bursts j = 1 .. n_bursts-1 are the following 144-dibit bursts of the stream (the two TDMA slots alternate; the
 * TACT word in the CACH names the slot).  Per record (= (channel * max_hits + hit) * n_bursts + j): CACH bits [24] (bits 0..6 =
 * the TACT word for Hamming(7,4)), ambe_fr[3][4][24] (the three `char ambe_fr[4][24]` the reference hands to
 * processMbeFrame, de-interleaved with dsd_ambe_2450_dibit_map, unreached cells 0), the 48 sync / EMB bits, and whether the
 * channel's stream (d_counts dibits) holds the whole burst.  `inverted_dmr` = opts->inverted_dmr: XOR 2 on the first 90
 * dibits of burst 0 (the part the reference takes from its rolling buffer).
 */
int dsdneo_b200_dmr_voice_cut_batch(const uint8_t* d_dibits, size_t dibit_pitch, const int32_t* d_counts, const void* d_hits,
                                    const int32_t* d_n_hits, int n_channels, int max_hits, int n_bursts, int inverted_dmr,
                                    uint8_t* d_cach24, uint8_t* d_ambe_fr, uint8_t* d_sync48, uint8_t* d_valid, void* stream);
/** dsd_ambe_2450_dibit_map (include/dsd-neo/core/ambe_interleave.h:25-32) as this library generates it:
 *  out[i] = {high_row, high_col, low_row, low_col}, i < 36. */
int dsdneo_b200_ambe_2450_dibit_map(uint8_t* out36x4);

/**
 * Vocoder frame ECC, batched: what `int mbe_decodeAmbe3600x2450Frame(const char ambe_fr[4][24], char ambe_d[49],
 * mbe_process_result*)` and `int mbe_decodeImbe7200x4400Frame(const char imbe_fr[8][23], char imbe_d[88],
 * mbe_process_result*)` do for the reference (mbelib-neo API, contract CMakeLists.txt:622-655; call sites
 * src/core/vocoder/dsd_mbe.c:168,188): [23,12] Golay on C0, pseudo-random demodulation seeded with C0's data, Golay /
 * [15,11] Hamming on the protected words, output bits in priority order.  d_c0_errors / d_total_errors = the `errs` / `errs2`
 * the reference stores in its state (result.c0_errors / result.total_errors).  PARITY UNPINNED: mbelib-neo is not in the
 * reference tree; this follows the published mbelib 1.3.0 / TIA-102.BABA algorithm.
 * Frames are the reference's own arrays as bytes: ambe_fr [n][4][24], imbe_fr [n][8][23], values 0 / 1.
 */
int dsdneo_b200_ambe3600x2450_decode_batch(const uint8_t* d_ambe_fr, uint8_t* d_ambe_d, int32_t* d_c0_errors,
                                           int32_t* d_total_errors, int n_frames, void* stream);
int dsdneo_b200_ambe3600x2450_decode_batch_host(const uint8_t* h_ambe_fr, uint8_t* h_ambe_d, int32_t* h_c0_errors,
                                                int32_t* h_total_errors, int n_frames);
int dsdneo_b200_imbe7200x4400_decode_batch(const uint8_t* d_imbe_fr, uint8_t* d_imbe_d, int32_t* d_c0_errors,
                                           int32_t* d_total_errors, int n_frames, void* stream);
int dsdneo_b200_imbe7200x4400_decode_batch_host(const uint8_t* h_imbe_fr, uint8_t* h_imbe_d, int32_t* h_c0_errors,
                                                int32_t* h_total_errors, int n_frames);
/** The same straight from the receive bank's voice records (below): nine IMBE frames per record -> d_imbe_d [n_records][9][88],
 *  d_c0_errors / d_total_errors [n_records][9]: what processMbeFrame -> mbe_decodeImbe7200x4400Frame sees for every LDU. */
struct dsdneo_b200_p25p1_voice;
int dsdneo_b200_p25p1_voice_imbe_decode_batch(const struct dsdneo_b200_p25p1_voice* d_voices, int n_records, uint8_t* d_imbe_d,
                                              int32_t* d_c0_errors, int32_t* d_total_errors, void* stream);

/**
 * Batched twins of `int check_and_fix_golay_24_6_soft(char* data, const char* parity, const int* reliab, int* fixed)` and
 * `check_and_fix_golay_24_12_soft` (include/dsd-neo/protocol/p25/p25p1_soft.h, src/protocol/p25/phase1/p25p1_soft.cpp:477-593):
 * Golay(24,6) / (24,12) with a bounded search over the 8 least reliable bits (at most 4 flips).
 * code = DSDNEO_P25_WORD_GOLAY_24_6 / _24_12; d_data_bits [n][6 | 12] corrected in place, d_parity_bits [n][12],
 * d_reliab [n][18 | 24] int (data bits first, then parity; clamped to 0..255 like the reference);
 * d_status[i]: 0 ok / 1 no valid candidate (data untouched); d_fixed[i] as the reference's *fixed.
 */
int dsdneo_b200_p25_golay_soft_batch(int code, uint8_t* d_data_bits, const uint8_t* d_parity_bits, const int32_t* d_reliab,
                                     int hard_override_enabled, int erasure_threshold, uint8_t* d_status, int32_t* d_fixed,
                                     int n_words, void* stream);
int dsdneo_b200_p25_golay_soft_batch_host(int code, uint8_t* h_data_bits, const uint8_t* h_parity_bits,
                                          const int32_t* h_reliab, int hard_override_enabled, int erasure_threshold,
                                          uint8_t* h_status, int32_t* h_fixed, int n_words);

/**
 * Batched twin of `int hamming_10_6_3_soft(const char* bits, const int* reliab, char* out_bits)`
 * (include/dsd-neo/protocol/p25/p25p1_soft.h:40, src/protocol/p25/phase1/p25p1_soft.cpp:444-475): Hamming(10,6,3) with a
 * bounded search over the 5 least reliable bits (at most 2 flips).  `hard_override_enabled` = p25_soft_hard_override_enabled()
 * (1 unless configured, p25p1_soft.cpp:41-52), `erasure_threshold` = p25p1_get_erasure_threshold() (64 unless configured).
 * d_bits10 / d_out10: [n][10] byte-per-bit (6 data + 4 parity); d_reliab10: [n][10] int (clamped to 0..255 like the reference);
 * d_status[i]: 0 unchanged / 1 corrected / 2 failed (out = in).
 */
int dsdneo_b200_hamming_10_6_3_soft_batch(const uint8_t* d_bits10, const int32_t* d_reliab10, int hard_override_enabled,
                                          int erasure_threshold, uint8_t* d_out10, uint8_t* d_status, int n_words,
                                          void* stream);
int dsdneo_b200_hamming_10_6_3_soft_batch_host(const uint8_t* h_bits10, const int32_t* h_reliab10,
                                               int hard_override_enabled, int erasure_threshold, uint8_t* h_out10,
                                               uint8_t* h_status, int n_words);

/**
 * P25 Phase 1 frame cutter: what the reference does dibit by dibit between frame sync and the FEC leaves, for every sync
 * hit of every channel at once, so that frames go from the slicer to the FEC kernels without a host round trip:
 *   - NID fields: the 32 dibits after the sync with the status symbol at frame offset 35 dropped, as 63 BCH bits,
 *     reliabilities min(|llr|, 255) and the final parity bit (p25p1_read_nid_fields + p25p1_append_bch_bits,
 *     src/engine/dispatch/dispatch_p25p1.c:59-82,121-143) -- the inputs of dsdneo_b200_p25p1_nid_decode_batch;
 *   - payload: `n_payload` dibits from frame offset 57 on with every 36th dibit of the frame (the status symbols) removed
 *     (tsbk_read_repetition_samples, src/protocol/p25/phase1/p25p1_tsbk.c:135-152, skipdibit = 36 - 14 at :1054), with
 *     their LLR pairs -- 98 dibits per half-rate trellis block, the input of dsdneo_b200_p25_12_soft_llr[_list]_batch.
 * Slot s = channel * max_hits + hit.  d_hits is the hit array of dsdneo_b200_frame_sync_search_batch
 * ({position of the last sync dibit, sync type} per hit).  nid_valid / payload_valid tell whether the stream of that channel
 * (d_counts dibits) still holds the whole NID / payload; invalid slots are zero-filled.
 */
int dsdneo_b200_p25p1_frame_cut_batch(const uint8_t* d_dibits, size_t dibit_pitch, const int16_t* d_llr, size_t llr_pitch,
                                      const int32_t* d_counts, const void* d_hits, const int32_t* d_n_hits, int n_channels,
                                      int max_hits, int n_payload, uint8_t* d_nid_code63, uint8_t* d_nid_reliab63,
                                      uint8_t* d_nid_parity, uint8_t* d_nid_parity_reliab, uint8_t* d_nid_valid,
                                      uint8_t* d_payload_dibits, int16_t* d_payload_llr, uint8_t* d_payload_valid,
                                      void* stream);

/**
 * P25 Phase 1 frame decoder: everything the reference's frame handlers do between the frame sync and the vocoder / message
 * parsers, for every sync hit of every channel at once, straight from the slicer's dibit + LLR streams:
 *   processTSBK  (src/protocol/p25/phase1/p25p1_tsbk.c:108-161,1051-1081)  1..3 half-rate trellis blocks, list-8 + CRC-16 pick
 *   processHDU   (p25p1_hdu.c:108-303)                                     36 Golay(24,6) words hard + soft, RS(36,20,17)
 *   processLDU1 / processLDU2 (p25p1_ldu.c:89-222, p25p1_ldu1.c:54-245, p25p1_ldu2.c:54-280)
 *                                                                          9 IMBE de-interleaves (the imbe_fr[8][23] +
 *                                                                          reliabilities handed to processMbeFrameSoft), 24
 *                                                                          Hamming(10,6,3) words hard + soft, RS(24,12,13) /
 *                                                                          (24,16,9) hard + ranked erasures, LSD (16,8) x 2
 *   processTDULC (p25p1_tdulc.c:75-238,284-300)                            12 Golay(24,12) dodeca words hard + soft, each two
 *                                                                          RS symbols (bits 6..11 first, swap_hex_words),
 *                                                                          RS(24,12,13) hard + ranked erasures; the link
 *                                                                          control word is dodeca 5..0 = rs_data[11], [10], ...
 * NID results come from dsdneo_b200_p25p1_nid_decode_batch on the slots of dsdneo_b200_p25p1_frame_cut_batch (slot = channel
 * * max_hits + hit; d_nid_valid = the cutter's nid_valid, may be NULL).  One frame record per hit, ordered by (channel, stream order): record index d_frame_off[channel] + hit;
 * LDUs additionally get a voice record (frame.voice_index).  d_totals = {frames, voice records} written by the call.
 * Hit positions are relative to buffer index region_offset of each channel row; frames that do not fit inside d_counts
 * dibits keep their NID fields and are flagged (reserved[0] = 1).  TDU has no payload; MPDU payloads are not decoded (NID only).
 * Bit-exact with the reference handlers (tests/test_gpu_p25p1_frames.py; golden records from the unmodified handlers).
 */
typedef struct dsdneo_b200_p25p1_frame {
    int64_t position;      /* stream index of the LAST sync dibit (d_stream_base[channel] + buffer index) */
    int32_t channel;
    int32_t voice_index;   /* LDU1 / LDU2: index into the voice records, else -1 */
    int16_t nac;
    int16_t nid_errs;
    int8_t nid_status;     /* enum NidResult of p25p1_nid_decode (> 0 = decoded) */
    uint8_t duid;          /* 0xFF when the NID failed */
    uint8_t n_tsbk;        /* TSDU: blocks read (stops after the block flagged last) */
    uint8_t tsbk_crc_ok;   /* bit b: block b passed crc16_lb_bridge */
    uint8_t rs_kind;       /* 0 none, 1 RS(36,20,17) HDU, 2 RS(24,12,13) LDU1 and TDULC, 3 RS(24,16,9) LDU2 */
    uint8_t rs_status;     /* 0 hard decode ok, 1 recovered by ranked erasures, 2 irrecoverable */
    uint8_t lsd_ok;        /* bit k: p25_lsd_fec_16x8_soft accepted LSD word k */
    uint8_t n_word_soft;   /* words whose soft decode changed the outcome (p25_p1_soft_hamming_ok / _golay_ok) */
    uint8_t lsd[2];        /* corrected low speed data octets */
    uint8_t reserved[6];   /* [0] = 1: the stream ended inside the frame */
    uint8_t tsbk[3][12];
    uint8_t rs_data[20];      /* hex_data[i] after Reed-Solomon (6-bit values) */
    uint8_t rs_in_data[20];   /* hex_data[i] / hex_parity[i] after the word-level FEC, as handed to the RS decoder */
    uint8_t rs_in_parity[16];
} dsdneo_b200_p25p1_frame;
typedef struct dsdneo_b200_p25p1_voice {
    uint32_t bits[9][8];        /* imbe_fr[row][col] of voice frame v = bit col of bits[v][row] */
    uint8_t reliab[9][8][23];   /* dsd_vocoder_soft_bit.reliability */
} dsdneo_b200_p25p1_voice;
int dsdneo_b200_p25p1_frames_decode_batch(const uint8_t* d_dibits, size_t dibit_pitch, const int16_t* d_llr, size_t llr_pitch,
                                          const int32_t* d_counts, const dsdneo_b200_sync_hit* d_hits, const int32_t* d_n_hits,
                                          int n_channels, int max_hits, int region_offset, const long long* d_stream_base,
                                          const int8_t* d_nid_status, const uint8_t* d_nid_valid, const int32_t* d_nid_nac,
                                          const uint8_t* d_nid_duid, const int32_t* d_nid_errs, int erasure_threshold, int hard_override_enabled,
                                          int32_t* d_frame_off, int32_t* d_voice_off, int32_t* d_totals,
                                          dsdneo_b200_p25p1_frame* d_frames, int frame_capacity, dsdneo_b200_p25p1_voice* d_voices,
                                          int voice_capacity, void* stream);

/**
 * viterbi_decode / viterbi_decode_punctured (src/core/util/dsd_misc.c:118-182; include/dsd-neo/fec/viterbi.h:23-29), the
 * K = 5 soft decoder used by M17 and YSF.  Costs are uint16 "probability of a 1" (0 / 0xFFFF strong, 0x7FFF erased).
 * @param d_cost   [n][cost_pitch] received soft bits, in_len used per frame (after de-puncturing at most 488)
 * @param d_punct  puncture pattern (1 = transmitted) of p_len entries on the device, or NULL for the unpunctured decoder
 * @param d_out    [n][out_pitch] decoded bytes; like the reference only the first (bits-1)/8+1 bytes are cleared and later
 *                 bytes are OR-ed into (the decoded message starts at bit 8); out_pitch >= (bits+4+7)/8
 * @param d_metric [n] the reference's return value (minimum final metric, neutral puncture cost removed)
 */
int dsdneo_b200_viterbi_k5_decode_batch(const uint16_t* d_cost, size_t cost_pitch, int in_len, const uint8_t* d_punct, int p_len,
                                        uint8_t* d_out, size_t out_pitch, uint32_t* d_metric, int n_frames, void* stream);
int dsdneo_b200_viterbi_k5_decode_batch_host(const uint16_t* h_cost, size_t cost_pitch, int in_len, const uint8_t* h_punct,
                                             int p_len, uint8_t* h_out, size_t out_pitch, uint32_t* h_metric, int n_frames);
/**
 * CNXDNConvolution_start + n_steps x CNXDNConvolution_decode[_soft] + CNXDNConvolution_chainback(out, n_bits_out)
 * (src/protocol/nxdn/nxdn_convolution.c:58-164).
 * @param d_sym     [n][pitch] symbol pairs s0,s1 (values 0 / 2, 1 = erased), 2*n_steps used; d_rel: reliabilities r0,r1 for the
 *                  soft decoder or NULL for the hard one
 * @param d_metrics [n][32] uint16: the reference's two ping-pong metric arrays (m_metrics1 | m_metrics2), which it zeroes once
 *                  at start-up and then carries from frame to frame -- pass zeros for a fresh decoder, keep them per channel
 * @param d_out     [n][out_pitch] n_bits_out decoded bits MSB first (other bits of the buffer untouched), n_steps <= 300
 */
int dsdneo_b200_nxdn_conv_decode_batch(const uint8_t* d_sym, const uint8_t* d_rel, size_t pitch, int n_steps, int n_bits_out,
                                       uint16_t* d_metrics, uint8_t* d_out, size_t out_pitch, int n_frames, void* stream);
int dsdneo_b200_nxdn_conv_decode_batch_host(const uint8_t* h_sym, const uint8_t* h_rel, size_t pitch, int n_steps, int n_bits_out,
                                            uint16_t* h_metrics, uint8_t* h_out, size_t out_pitch, int n_frames);

/* ---- the reference's soft symbol-capture format ("DSDNSYM2") as harness I/O (host side, no device work) ---------- */
/*
 * Written by `dsd-neo -c file.bin`, replayed by `dsd-neo -i file.bin`: header src/core/file/dsd_file.c:876-888, record
 * writer src/core/frames/dsd_dibit.c:794-818, reader src/dsp/dsd_symbol.c:120-173, constants include/dsd-neo/core/dibit.h:35-37.
 * Packing one channel's symbolizer output (dibits, reliability, llr[n][2], symbols) into this format lets the unmodified
 * reference CLI decode what the GPU demodulated.  unpack accepts data with or without the 16-byte header and returns the
 * number of records (negative DSDNEO_B200_E* on a bad header).
 */
#define DSDNEO_B200_SYMCAP_HEADER_SIZE 16
#define DSDNEO_B200_SYMCAP_RECORD_SIZE 10
size_t dsdneo_b200_symbol_capture_size(size_t n_records, int with_header);
int dsdneo_b200_symbol_capture_pack(const uint8_t* dibits, const uint8_t* reliability, const int16_t* llr, const float* symbols,
                                    size_t n_records, int with_header, uint8_t* out);
long long dsdneo_b200_symbol_capture_unpack(const uint8_t* in, size_t len, uint8_t* dibits, uint8_t* reliability, int16_t* llr,
                                            float* symbols, size_t max_records);
int dsdneo_b200_symbol_capture_write_file(const char* path, int append, const uint8_t* dibits, const uint8_t* reliability,
                                          const int16_t* llr, const float* symbols, size_t n_records);

/* ---- stream server for the reference's runtime hook seam (host side, no device work) -------------------------- */
/*
 * The unmodified reference decoder reads discriminator floats through dsd_rtl_stream_io_hooks {read, return_pwr}
 * (include/dsd-neo/runtime/rtl_stream_io_hooks.h:25-28) and asks dsd_rtl_stream_metrics_hooks for output_rate_hz,
 * output_kind, symbol_profile and stream_generation (rtl_stream_metrics_hooks.h:28-47).  The dsdneo_b200_stream_hook_*
 * functions have exactly those signatures, so a maintainer installs them in the two tables
 * (src/engine/rtl_stream_io_hooks_install.c:26-34) and passes the server handle as rtl_ctx; the ingest loop pushes one
 * channel's rows of the front end's output.  read blocks until data or close, returns 0 with *out_got floats, < 0 once
 * the stream is closed and drained.  The context-free metrics hooks answer for the server made current.
 */
typedef struct dsdneo_b200_stream_server dsdneo_b200_stream_server;
dsdneo_b200_stream_server* dsdneo_b200_stream_server_create(size_t ring_floats, unsigned int output_rate_hz, int symbol_rate_hz,
                                                            int levels, int channel_profile);
void dsdneo_b200_stream_server_destroy(dsdneo_b200_stream_server* s);
void dsdneo_b200_stream_server_make_current(dsdneo_b200_stream_server* s);
size_t dsdneo_b200_stream_server_push(dsdneo_b200_stream_server* s, const float* samples, size_t n, int block);
void dsdneo_b200_stream_server_close(dsdneo_b200_stream_server* s);
void dsdneo_b200_stream_server_bump_generation(dsdneo_b200_stream_server* s);
/** output_kind 1 (default): FSK discriminator samples; 2: symbol-rate CQPSK symbols (the output of
 *  dsdneo_b200_full_demod_cqpsk_batch), for which the decoder takes its symbol-rate fast path (src/dsp/dsd_symbol.c:1583-1625)
 *  and, with `cqpsk_active`, slices with cqpsk_slice(); `snr_cqpsk_db` feeds the reliability weight (<= -50: none). */
void dsdneo_b200_stream_server_set_output_kind(dsdneo_b200_stream_server* s, int output_kind, int cqpsk_active,
                                               double snr_cqpsk_db);
void dsdneo_b200_stream_server_set_power(dsdneo_b200_stream_server* s, double pwr);
int dsdneo_b200_stream_hook_read(void* rtl_ctx, float* out, size_t count, int* out_got);
double dsdneo_b200_stream_hook_return_pwr(const void* rtl_ctx);
unsigned int dsdneo_b200_stream_hook_output_rate_hz(void);
int dsdneo_b200_stream_hook_output_kind(void);
int dsdneo_b200_stream_hook_symbol_profile(int* out_symbol_rate_hz, int* out_levels, int* out_channel_profile);
uint32_t dsdneo_b200_stream_hook_stream_generation(void);
/* dsd_rtl_stream_metrics_hooks.cqpsk_status / .snr_cqpsk_db (rtl_stream_metrics_hooks.h:33,39) */
int dsdneo_b200_stream_hook_cqpsk_status(int* out_cqpsk_enable, int* out_cqpsk_timing_active);
double dsdneo_b200_stream_hook_snr_cqpsk_db(void);


/* ---- P25 Phase 1 C4FM receiver bank: IQ -> frames in one object (the metric's chain, BASELINE.json configs[2]) ---------- */
/*
 * N already-channelised streams (48 kS/s cu8 or cf32 IQ, the reference's RTL ingest / --iq-replay format) through
 *   widen_u8_to_f32_bias127 -> full_demod -> p25_filter + getDibitSoft -> frame sync -> NID -> TSBK / HDU / LDU1 / LDU2
 * on the device, per channel what the reference runs as one process (dispatch: src/engine/dispatch/dispatch_p25p1.c:401-426).
 * Every channel is configured as the reference is once it has seen a +P25p1 sync (synctype = lastsynctype = 0: p25_filter,
 * window 2/2, min / max tracker).  Frames are decoded 864 symbols behind the slicer from a per-channel stream history kept
 * on the device, so frames that straddle two calls are decoded once, from a contiguous stream.
 */
typedef struct dsdneo_b200_p25p1_rx_config {
    int n_channels;
    int rate_hz;             /* per-channel sample rate (48000) */
    int block_pairs;         /* one full_demod() block (the reference's DEFAULT_BUF_LENGTH / 2 = 8192) */
    int max_pairs_per_call;  /* capacity of one process call per channel; multiple of block_pairs */
    int input_cu8;           /* 1 = cu8 IQ widened on the device, 0 = cf32 */
    int fir_arith;           /* DSDNEO_FIR_ARITH_* */
    int max_hits;            /* frame syncs kept per channel per call (<= 32; 0 = 32) */
    int erasure_threshold;   /* p25p1_get_erasure_threshold() (0 = 64) */
    int hard_override_disabled; /* !p25_soft_hard_override_enabled() */
    int track_nac;           /* 1: the channel's last decoded NAC feeds the known-NAC retry of p25p1_nid_decode (state->nac) */
    const float* channel_squelch_level; /* per channel or NULL (squelch off) */
    const float* p25_filter_taps;       /* NORMALISED p25_filter taps for rate_hz as the reference's design_sps_fir leaves them */
    int p25_filter_len;
    int acquire_tiles;       /* 0: every channel starts synchronised (the stream must start symbol-aligned); k > 0: every channel
                              * starts never-synchronised and the first k tiles run getFrameSync()'s acquisition on the device
                              * (dsdneo_b200_symbolize_reacquire_batch: hunt on the matched filter's output, timing nudges, sync
                              * warm start), one tile at a time; a channel that has not found the P25 Phase 1 sync by then continues
                              * with the synchronised rules (with auto_reacquire_tiles it keeps hunting, inside the pipeline) */
    int auto_reacquire_tiles; /* 0: off; k > 0: loss-of-sync watch on the device -- a channel none of whose sync hits of the
                               * last k tiles decoded to a valid NID goes back to getFrameSync()'s hunt (warm form, on the matched
                               * filter's output) two tiles later, inside the tile pipeline (no serialised tiles, no host step) and
                               * in stream order, so a stream decodes the same way however it is cut into calls of the same size;
                               * it hunts until it finds the sync again.  The reference takes this decision per frame (NID failure
                               * -> getFrameSync); k * tile length should exceed the longest gap between frames of a live channel */
} dsdneo_b200_p25p1_rx_config;
typedef struct dsdneo_b200_p25p1_rx_out { /* device buffers */
    dsdneo_b200_p25p1_frame* d_frames;
    int frame_capacity;
    dsdneo_b200_p25p1_voice* d_voices;
    int voice_capacity;
    int32_t* d_totals;      /* {frame records, voice records} of this call */
    uint8_t* d_dibits;      /* optional: the call's new dibits [n_channels][dibit_pitch] */
    size_t dibit_pitch;
    int32_t* d_counts;      /* optional: new dibits per channel */
} dsdneo_b200_p25p1_rx_out;
typedef struct dsdneo_b200_p25p1_rx_host_out { /* host buffers (pinned for overlap) */
    dsdneo_b200_p25p1_frame* h_frames;
    int frame_capacity;
    dsdneo_b200_p25p1_voice* h_voices;
    int voice_capacity;
    int32_t* h_totals;
    uint8_t* h_dibits;
    size_t dibit_pitch;
    int32_t* h_counts;
} dsdneo_b200_p25p1_rx_host_out;
typedef struct dsdneo_b200_p25p1_rx dsdneo_b200_p25p1_rx;
dsdneo_b200_p25p1_rx* dsdneo_b200_p25p1_rx_create(const dsdneo_b200_p25p1_rx_config* cfg);
void dsdneo_b200_p25p1_rx_destroy(dsdneo_b200_p25p1_rx* rx);
int dsdneo_b200_p25p1_rx_frame_capacity(const dsdneo_b200_p25p1_rx* rx); /* records one call can produce at most */
int dsdneo_b200_p25p1_rx_voice_capacity(const dsdneo_b200_p25p1_rx* rx);
size_t dsdneo_b200_p25p1_rx_dibit_pitch(const dsdneo_b200_p25p1_rx* rx); /* dibits one call can add per channel at most */
/** d_iq: [n_channels][iq_pitch_pairs] cu8 (uchar2) or cf32 (float2) per config; n_pairs a multiple of block_pairs. */
int dsdneo_b200_p25p1_rx_process(dsdneo_b200_p25p1_rx* rx, const void* d_iq, size_t iq_pitch_pairs, int n_pairs,
                                 const dsdneo_b200_p25p1_rx_out* out, void* stream);
/**
 * Pipelined form.  Inside the bank a tile passes three stages on three streams (channel filter | discriminator recurrences +
 * matched filter | slicer + frames), and consecutive tiles overlap on the device: the latency-bound per-channel recurrences of
 * one tile run under the throughput-bound filters of the next.  _submit queues one tile (its input must be complete on
 * `stream` at call time) and returns a ticket without waiting; _wait makes `stream` wait (on the device, the host does not
 * block) until that tile's outputs are complete; _input_consumed does the same for "d_iq may be overwritten".  Tiles complete
 * in submission order.  _process == _submit + _wait on the same stream (no overlap between consecutive calls, because the
 * next tile's input is only known to be ready after the wait).  Outputs of tile i must be consumed before they are handed to
 * a later _submit again.
 */
long long dsdneo_b200_p25p1_rx_submit(dsdneo_b200_p25p1_rx* rx, const void* d_iq, size_t iq_pitch_pairs, int n_pairs,
                                      const dsdneo_b200_p25p1_rx_out* out, void* stream);
int dsdneo_b200_p25p1_rx_wait(dsdneo_b200_p25p1_rx* rx, long long ticket, void* stream);
int dsdneo_b200_p25p1_rx_input_consumed(dsdneo_b200_p25p1_rx* rx, long long ticket, void* stream);
/**
 * Mid-stream loss of sync: sends channels back to getFrameSync()'s hunt (src/dsp/dsd_frame_sync.c, the warm form the
 * reference runs after a channel's first sync: matched filter kept, hunting rules with the +-1-sample timing nudges, basic
 * lock, sync warm start) for the next `tiles` tiles.  `h_synchronised`: host array of n_channels flags, 1 = the channel
 * keeps its lock and runs the synchronised rules as before (its outputs are bit-identical to a stream without this call),
 * 0 = the channel hunts from an empty window; NULL = every channel hunts.  The reference takes this decision per frame
 * inside its decode loop (no valid NID -> back to getFrameSync); here the caller takes it from the frame records of the
 * tiles it has seen (e.g. no frame with nid_status > 0 for a few tiles).  Drains the pipeline (device synchronise); the
 * acquiring tiles run their stages one after the other, as at stream start (cfg.acquire_tiles).
 */
int dsdneo_b200_p25p1_rx_reacquire(dsdneo_b200_p25p1_rx* rx, const int* h_synchronised, int tiles);
/** Monitoring: per channel whether it is synchronised (1) or hunting (0) -- always 1 when acquisition was never configured --
 * and, with cfg.auto_reacquire_tiles, the watch's count of tiles without a valid NID (negative while a drop is taking effect).
 * Either array may be NULL.  Host arrays of n_channels; drains the pipeline (device synchronise). */
int dsdneo_b200_p25p1_rx_channel_status(dsdneo_b200_p25p1_rx* rx, int* h_synchronised, int* h_idle_tiles);
/** Host buffers, streaming: returns a ticket >= 0; results are in the caller's buffers once wait_host(ticket) returned.
 * Up to six tickets may be outstanding (H2D, the four pipeline stages and D2H of consecutive tiles overlap); a seventh submit
 * first completes the oldest one.  Input and output buffers of a ticket must stay untouched until its wait_host returned. */
long long dsdneo_b200_p25p1_rx_submit_host(dsdneo_b200_p25p1_rx* rx, const void* h_iq, size_t iq_pitch_pairs, int n_pairs,
                                           const dsdneo_b200_p25p1_rx_host_out* out);
int dsdneo_b200_p25p1_rx_wait_host(dsdneo_b200_p25p1_rx* rx, long long ticket);
int dsdneo_b200_p25p1_rx_process_host(dsdneo_b200_p25p1_rx* rx, const void* h_iq, size_t iq_pitch_pairs, int n_pairs,
                                      const dsdneo_b200_p25p1_rx_host_out* out);

/* ---- K21: MBE speech synthesis stage, batched over frames -- PARITY UNPINNED ---------------------------------- */
/*
 * dsd-neo obtains PCM from mbelib-neo 2.x (mbe_processImbe4400Dataf / mbe_processAmbe2450Dataf, call sites
 * src/core/vocoder/dsd_mbe.c:268,296,581,617,685), an un-vendored dependency that is absent here, so this stage cannot be
 * compared with the reference.  It follows the published algorithm of mbelib 1.3.0
 * (mbe_spectralAmpEnhance + mbe_synthesizeSpeechf + mbe_floattoshort, then mbe_moveMbeParms(cur, prev_enhanced)) with one
 * documented difference: random phases / noise come from a counter-based hash of (key, band, sample, index) instead of
 * libc rand().  The struct is mbelib 1.3.0's `struct mbe_parameters` (mbelib.h); whether mbelib-neo 2.x kept that
 * layout is unverified.  The bit-level parameter decode (quantiser tables) is not part of this stage.
 * d_cur[i] is enhanced and phase-updated in place and copied to d_prev_enhanced[i]; d_keys may be NULL (key = index).
 */
typedef struct {
    float w0;
    int L;
    int K;
    int Vl[57];
    float Ml[57];
    float log2Ml[57];
    float PHIl[57];
    float PSIl[57];
    float gamma;
    int un;
    int repeat;
} dsdneo_b200_mbe_parms;
int dsdneo_b200_mbe_synth_batch(dsdneo_b200_mbe_parms* d_cur, dsdneo_b200_mbe_parms* d_prev_enhanced, const uint64_t* d_keys,
                                int uvquality, float* d_pcm_f, int16_t* d_pcm_s, int n_frames, void* stream);
int dsdneo_b200_mbe_synth_batch_host(dsdneo_b200_mbe_parms* h_cur, dsdneo_b200_mbe_parms* h_prev_enhanced, const uint64_t* h_keys,
                                     int uvquality, float* h_pcm_f, int16_t* h_pcm_s, int n_frames);

/** Self-test hook: the device atan2f used by the discriminator's large-angle branch (fsk_modem.c:34),
 *  evaluated on caller-supplied inputs so tests can compare it with the host libm bit for bit. */
int dsdneo_b200_selftest_atan2f(const float* d_y, const float* d_x, float* d_out, int n, void* stream);

/** Exhaustive check (all 2^32 operands) of the symbolizer's x / 5 sequence against the IEEE operator: writes the number of
 * mismatching operands and the smallest mismatching bit pattern (0xffffffff if none). */
int dsdneo_b200_selftest_div5(unsigned long long* d_n_mismatch, unsigned* d_first_bad, void* stream);

/** Self-test hook: the discriminator's output scale 30000.0f / peak (fsk_modem.c:127-129) as the recurrence kernel
 *  evaluates it (d_fast) next to the device's IEEE division (d_ieee), so tests can check they agree bit for bit. */
int dsdneo_b200_selftest_scale(const float* d_pk, float* d_fast, float* d_ieee, int n, void* stream);

/** Self-test hook: the CQPSK chain's branch-free division a / b and square root sqrt(a) (csrc/cqpsk.cu) next to the
 *  device's IEEE operators; d_flags bit 0 / bit 1 = the sequence declared the division / square root operands inside its
 *  safe range (outside it the kernel re-runs the symbol with the plain operators). */
int dsdneo_b200_selftest_divsqrt(const float* d_a, const float* d_b, float* d_q_fast, float* d_q_ieee, float* d_s_fast,
                                 float* d_s_ieee, unsigned char* d_flags, int n, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* DSDNEO_B200_H_ */
