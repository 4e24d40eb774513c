// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Thin extern "C" driver around the UNMODIFIED reference sources compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libdsdneo_ref*.so.  It
 * only allocates the reference's own structs and calls the reference's own
 * entry points so that Python (ctypes) can run them:
 *   - full_demod()                    include/dsd-neo/dsp/demod_pipeline.h:106
 *   - simd_fir_complex_apply[_scalar] src/dsp/simd_fir.cpp:55,350
 *   - ReedSolomon_63 wrappers / Golay24 / BCH  (C++ headers -> C symbols)
 * Nothing here restates an algorithm; the restatement lives in oracle/*.c.
 */
#include <cstdlib>
#include <cstring>
#include <new>

#include <dsd-neo/dsp/demod_pipeline.h>
#include <dsd-neo/dsp/demod_state.h>
#include <dsd-neo/dsp/fsk_modem.h>
#include <dsd-neo/dsp/simd_fir.h>
#include <dsd-neo/dsp/ted.h>

#include "dsp/simd_fir_internal.h"

extern "C" {

/* ---- block side: one reference demod_state per channel ----------------- */

void*
ref_demod_create(int sample_rate, int symbol_rate, int lpf_profile, int lpf_enable, float squelch_level) {
    void* mem = NULL;
    if (posix_memalign(&mem, 64, sizeof(demod_state)) != 0 || !mem) {
        return NULL;
    }
    memset(mem, 0, sizeof(demod_state));
    demod_state* s = (demod_state*)mem;
    /* Same field set as the reference bench's FSK configuration
     * (tests/dsp/bench_dsp.cpp:1022-1046). */
    s->rate_in = sample_rate;
    s->rate_out = sample_rate;
    s->rate_out2 = 0;
    s->lowpassed = s->input_cb_buf;
    s->mode_demod = &dsd_fm_demod;
    s->output_kind = DSD_DEMOD_OUTPUT_FSK_DISCRIMINATOR;
    s->symbol_rate_hz = symbol_rate;
    s->symbol_levels = 4;
    s->ted_sps = sample_rate / symbol_rate;
    s->sps_is_integer = (s->ted_sps * symbol_rate == sample_rate) ? 1 : 0;
    s->channel_lpf_enable = lpf_enable;
    s->channel_lpf_profile = lpf_profile;
    s->channel_squelch_level = squelch_level;
    s->squelch_env = 1.0f;
    s->squelch_env_attack = 0.125f;
    s->squelch_env_release = 0.03125f;
    dsd_fsk_modem_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.sample_rate_hz = s->rate_out;
    cfg.symbol_rate_hz = s->symbol_rate_hz;
    cfg.levels = s->symbol_levels;
    cfg.channel_profile = s->channel_lpf_profile;
    dsd_fsk_modem_init(&s->fsk_modem_state, &cfg);
    return s;
}

/* Wideband variant: rate_in -> rate_out through `passes` half-band stages (demod_pipeline.cpp:983-1001),
 * the reference's own way of selecting one channel from a wide capture. */
void*
ref_demod_create_wideband(int rate_in, int passes, int symbol_rate, int lpf_profile, int lpf_enable, float squelch_level) {
    demod_state* s = (demod_state*)ref_demod_create(rate_in >> passes, symbol_rate, lpf_profile, lpf_enable, squelch_level);
    if (!s) {
        return NULL;
    }
    s->rate_in = rate_in;
    s->downsample_passes = passes;
    return s;
}

void
ref_demod_destroy(void* h) {
    free(h);
}

/* Feed one block (n_floats interleaved I/Q floats) exactly like the demod
 * thread does (src/io/radio/rtl_sdr_fm.cpp:3419-3420) and copy result[]. */
int
ref_demod_block(void* h, const float* iq, int n_floats, float* out, int out_cap) {
    demod_state* s = (demod_state*)h;
    if (!s || n_floats > MAXIMUM_BUF_LENGTH) {
        return -1;
    }
    memcpy(s->input_cb_buf, iq, (size_t)n_floats * sizeof(float));
    s->lowpassed = s->input_cb_buf;
    s->lp_len = n_floats;
    full_demod(s);
    int n = s->result_len < out_cap ? s->result_len : out_cap;
    memcpy(out, s->result, (size_t)n * sizeof(float));
    return s->result_len;
}

/* Carried state, for parity dumps: {prev_i, prev_q, have_prev, dc_est, peak, channel_pwr, squelched} */
void
ref_demod_get_state(void* h, float* out7) {
    demod_state* s = (demod_state*)h;
    out7[0] = s->fsk_modem_state.prev_i;
    out7[1] = s->fsk_modem_state.prev_q;
    out7[2] = (float)s->fsk_modem_state.have_prev;
    out7[3] = s->fsk_modem_state.dc_est;
    out7[4] = s->fsk_modem_state.discriminator_peak_est;
    out7[5] = s->channel_pwr;
    out7[6] = (float)s->channel_squelched;
}

int
ref_demod_get_lpf_taps(void* h, float* taps_out, int cap) {
    demod_state* s = (demod_state*)h;
    int n = s->channel_lpf_plan_taps_len;
    if (n > cap) {
        n = cap;
    }
    memcpy(taps_out, s->channel_lpf_plan_taps, (size_t)n * sizeof(float));
    return s->channel_lpf_plan_taps_len;
}

/* CQPSK symbol output kind, configured like the reference bench's configure_common_cqpsk_state
 * (tests/dsp/bench_dsp.cpp:1073-1091) and the runtime defaults (src/io/radio/rtl_demod_config.cpp:364-366). */
void*
ref_demod_create_cqpsk(int sample_rate, int symbol_rate, int sps, int lpf_enable, float squelch_level, float ted_gain,
                       int ted_gain_is_set) {
    demod_state* s = (demod_state*)ref_demod_create(sample_rate, symbol_rate, DSD_CH_LPF_PROFILE_P25_CQPSK, lpf_enable,
                                                    squelch_level);
    if (!s) {
        return NULL;
    }
    s->output_kind = DSD_DEMOD_OUTPUT_SYMBOL_CQPSK;
    s->cqpsk_enable = 1;
    s->ted_sps = sps;
    s->sps_is_integer = 1;
    s->ted_gain = ted_gain;
    s->ted_gain_is_set = ted_gain_is_set;
    s->cqpsk_diff_prev_r = 1.0f;
    s->cqpsk_diff_prev_j = 0.0f;
    s->cqpsk_agc_avg = 1.0f;
    ted_init_state(&s->ted_state);
    return s;
}

/* Carried CQPSK loop state for parity dumps (24 floats, ints converted):
 * {agc_avg, fll.phase, fll.freq, fll.alpha, fll.beta, ted.mu, ted.omega, ted.last_r, ted.last_j, ted.lock_accum,
 *  ted.lock_count, ted_effective_gain, diff_prev_r, diff_prev_j, costas.phase, costas.freq, costas.error,
 *  costas.error_smooth, err_avg_q14, err_raw_avg_q14, conf_avg_q14, zero_conf_pct, channel_pwr, channel_squelched} */
void
ref_demod_get_cqpsk_state(void* h, float* out24) {
    demod_state* s = (demod_state*)h;
    out24[0] = s->cqpsk_agc_avg;
    out24[1] = s->fll_band_edge_state.phase;
    out24[2] = s->fll_band_edge_state.freq;
    out24[3] = s->fll_band_edge_state.alpha;
    out24[4] = s->fll_band_edge_state.beta;
    out24[5] = s->ted_state.mu;
    out24[6] = s->ted_state.omega;
    out24[7] = s->ted_state.last_r;
    out24[8] = s->ted_state.last_j;
    out24[9] = s->ted_state.lock_accum;
    out24[10] = (float)s->ted_state.lock_count;
    out24[11] = s->ted_effective_gain;
    out24[12] = s->cqpsk_diff_prev_r;
    out24[13] = s->cqpsk_diff_prev_j;
    out24[14] = s->costas_state.phase;
    out24[15] = s->costas_state.freq;
    out24[16] = s->costas_state.error;
    out24[17] = s->costas_state.error_smooth;
    out24[18] = (float)s->costas_err_avg_q14;
    out24[19] = (float)s->costas_err_raw_avg_q14;
    out24[20] = (float)s->costas_conf_avg_q14;
    out24[21] = (float)s->costas_zero_conf_pct;
    out24[22] = s->channel_pwr;
    out24[23] = (float)s->channel_squelched;
}

int
ref_demod_get_fll_taps(void* h, float* lower_r, float* lower_i, float* upper_r, float* upper_i, int cap) {
    demod_state* s = (demod_state*)h;
    int n = s->fll_band_edge_state.n_taps;
    if (n > cap) {
        n = cap;
    }
    memcpy(lower_r, s->fll_band_edge_state.taps_lower_r, (size_t)n * sizeof(float));
    memcpy(lower_i, s->fll_band_edge_state.taps_lower_i, (size_t)n * sizeof(float));
    memcpy(upper_r, s->fll_band_edge_state.taps_upper_r, (size_t)n * sizeof(float));
    memcpy(upper_i, s->fll_band_edge_state.taps_upper_i, (size_t)n * sizeof(float));
    return s->fll_band_edge_state.n_taps;
}

void
ref_fir_complex_scalar(const float* in, int in_len, float* out, float* hist_i, float* hist_q, const float* taps,
                       int taps_len) {
    simd_fir_complex_apply_scalar(in, in_len, out, hist_i, hist_q, taps, taps_len);
}

} /* extern "C" */
