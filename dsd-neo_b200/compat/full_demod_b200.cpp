/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Drop-in `full_demod()` for dsd-neo (compiled INSIDE the dsd-neo tree as C++14, like src/dsp/demod_pipeline.cpp,
 * against its own headers).
 *
 * Same signature and caller contract as src/dsp/demod_pipeline.cpp:1330-1350: the demod thread sets d->lowpassed /
 * d->lp_len (interleaved I/Q floats) and reads d->result[0..d->result_len).  Covered configurations, both with no
 * half-band passes and IQ DC block / IQ balance off:
 *   - output_kind == DSD_DEMOD_OUTPUT_FSK_DISCRIMINATOR, cqpsk off (the defaults of every 4FSK mode,
 *     src/io/radio/rtl_demod_config.cpp:189-228);
 *   - output_kind == DSD_DEMOD_OUTPUT_SYMBOL_CQPSK with cqpsk_enable and the OP25 chain (mode_demod != raw_demod),
 *     blocks of at least 4 pairs, ted_sps 2..10 (P25 LSM / Phase 2).
 * Anything else is handed back to the reference's own implementation, which the integrator renames to full_demod_cpu
 * (one line in demod_pipeline.cpp).
 *
 * One demod_state == one bank of 1 channel: correctness shim, not the fast path (the fast path is
 * dsdneo_b200_frontend_process* over hundreds of channels).
 */
#include <dsd-neo/dsp/demod_pipeline.h>
#include <dsd-neo/dsp/demod_state.h>
#include <stdio.h>
#include <string.h>

#include "dsdneo_b200.h"

extern "C" void full_demod_cpu(struct demod_state* d); /* the reference body, renamed */

#define B200_MAX_STATES 64
static struct {
    const struct demod_state* key;
    dsdneo_b200_demod_bank* bank;
    dsdneo_b200_cqpsk_bank* cqpsk; /* created on first CQPSK block; its loop state lives on the device from then on */
    int rate_out, profile, lpf_enable;
    float squelch;
    int cqpsk_sps;
    int mirrored_have_prev; /* fsk_modem_state.have_prev as this shim last wrote it: a 0 found there since is a host-side reset */
} g_banks[B200_MAX_STATES];

/* Drop the device state kept for `d` (call where the reference frees or re-initialises a demod_state, so that a later state
 * at the same address does not inherit it). */
extern "C" void
full_demod_b200_forget(const struct demod_state* d) {
    for (int i = 0; i < B200_MAX_STATES; i++) {
        if (g_banks[i].key == d) {
            dsdneo_b200_demod_bank_destroy(g_banks[i].bank);
            dsdneo_b200_cqpsk_bank_destroy(g_banks[i].cqpsk);
            memset(&g_banks[i], 0, sizeof(g_banks[i]));
        }
    }
}

static int
slot_of(const struct demod_state* d) {
    for (int i = 0; i < B200_MAX_STATES; i++) {
        if (g_banks[i].key == d) {
            return i;
        }
    }
    return -1;
}

static dsdneo_b200_demod_bank*
bank_for(const struct demod_state* d) {
    int free_slot = -1;
    for (int i = 0; i < B200_MAX_STATES; i++) {
        if (g_banks[i].key == d) {
            if (g_banks[i].rate_out == d->rate_out && g_banks[i].profile == d->channel_lpf_profile
                && g_banks[i].lpf_enable == d->channel_lpf_enable && g_banks[i].squelch == d->channel_squelch_level) {
                return g_banks[i].bank;
            }
            dsdneo_b200_demod_bank_destroy(g_banks[i].bank); /* plan changed: same as channel_lpf_ensure_plan re-design */
            dsdneo_b200_cqpsk_bank_destroy(g_banks[i].cqpsk);
            g_banks[i].cqpsk = NULL;
            g_banks[i].key = NULL;
            free_slot = i;
            break;
        }
        if (!g_banks[i].key && free_slot < 0) {
            free_slot = i;
        }
    }
    if (free_slot < 0) {
        return NULL;
    }
    int profile = d->channel_lpf_profile;
    float squelch = d->channel_squelch_level;
    dsdneo_b200_demod_bank_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_channels = 1;
    cfg.rate_out_hz = d->rate_out;
    cfg.channel_lpf_enable = d->channel_lpf_enable;
    cfg.channel_lpf_profile = &profile;
    cfg.channel_squelch_level = &squelch;
    cfg.fir_arith = DSDNEO_FIR_ARITH_FMA;
    dsdneo_b200_demod_bank* b = dsdneo_b200_demod_bank_create(&cfg);
    if (!b) {
        return NULL;
    }
    g_banks[free_slot].key = d;
    g_banks[free_slot].bank = b;
    g_banks[free_slot].rate_out = d->rate_out;
    g_banks[free_slot].profile = d->channel_lpf_profile;
    g_banks[free_slot].lpf_enable = d->channel_lpf_enable;
    g_banks[free_slot].squelch = d->channel_squelch_level;
    return b;
}

/* The CQPSK loop state of `d`: one bank of 1 channel, re-created when ted_sps changes (the reference re-initialises the
 * FLL and Gardner blocks on an sps change, costas.cpp:352-398,635-683; the Costas loop and AGC restart with it here). */
static dsdneo_b200_cqpsk_bank*
cqpsk_for(const struct demod_state* d) {
    const int sps = d->ted_sps > 0 ? d->ted_sps : 5;
    for (int i = 0; i < B200_MAX_STATES; i++) {
        if (g_banks[i].key != d) {
            continue;
        }
        if (g_banks[i].cqpsk && g_banks[i].cqpsk_sps == sps) {
            return g_banks[i].cqpsk;
        }
        dsdneo_b200_cqpsk_bank_destroy(g_banks[i].cqpsk);
        dsdneo_b200_cqpsk_bank_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.n_channels = 1;
        cfg.rate_out_hz = d->rate_out;
        cfg.ted_sps = &sps;
        cfg.ted_gain = d->ted_gain;
        cfg.ted_gain_is_set = d->ted_gain_is_set;
        g_banks[i].cqpsk = dsdneo_b200_cqpsk_bank_create(&cfg);
        g_banks[i].cqpsk_sps = sps;
        return g_banks[i].cqpsk;
    }
    return NULL;
}

static int
full_demod_cqpsk_b200(struct demod_state* d) {
    dsdneo_b200_demod_bank* bank = bank_for(d);
    dsdneo_b200_cqpsk_bank* q = bank ? cqpsk_for(d) : NULL;
    if (!q) {
        return 0;
    }
    const int pairs = d->lp_len >> 1;
    const int sps = d->ted_sps > 0 ? d->ted_sps : 5;
    const int cap = dsdneo_b200_cqpsk_block_capacity(pairs, sps);
    int count = 0;
    if (cap <= 0 || cap > MAXIMUM_BUF_LENGTH
        || dsdneo_b200_full_demod_cqpsk_batch_host(bank, q, d->lowpassed, (size_t)pairs, pairs, 1, d->result, (size_t)cap,
                                                   &count) != 0) {
        d->result_len = 0; /* no silent fallback on a GPU failure */
        return 1;
    }
    d->result_len = count;
    dsdneo_b200_demod_chan_state st;
    dsdneo_b200_cqpsk_chan_state cs;
    if (dsdneo_b200_demod_bank_get_state(bank, 0, &st) == 0 && dsdneo_b200_cqpsk_bank_get_state(q, 0, &cs) == 0) {
        d->channel_pwr = st.channel_pwr;
        d->channel_squelched = st.channel_squelched;
        d->squelch_gate_open = !st.channel_squelched;
        d->cqpsk_agc_avg = cs.cqpsk_agc_avg;
        d->fll_band_edge_state.phase = cs.fll_phase;
        d->fll_band_edge_state.freq = cs.fll_freq;
        d->ted_state.mu = cs.ted_mu;
        d->ted_state.omega = cs.ted_omega;
        d->ted_state.last_r = cs.ted_last_r;
        d->ted_state.last_j = cs.ted_last_j;
        d->ted_state.lock_accum = cs.ted_lock_accum;
        d->ted_state.lock_count = cs.ted_lock_count;
        d->ted_effective_gain = cs.ted_effective_gain;
        d->cqpsk_diff_prev_r = cs.cqpsk_diff_prev_r;
        d->cqpsk_diff_prev_j = cs.cqpsk_diff_prev_j;
        d->costas_state.phase = cs.costas_phase;
        d->costas_state.freq = cs.costas_freq;
        d->costas_state.error = cs.costas_error;
        d->costas_state.error_smooth = cs.costas_error_smooth;
        d->costas_err_avg_q14 = cs.costas_err_avg_q14;
        d->costas_err_raw_avg_q14 = cs.costas_err_raw_avg_q14;
        d->costas_conf_avg_q14 = cs.costas_conf_avg_q14;
        d->costas_zero_conf_pct = cs.costas_zero_conf_pct;
    }
    return 1;
}

extern "C" void
full_demod(struct demod_state* d) {
    const int cqpsk_covered = d && d->output_kind == DSD_DEMOD_OUTPUT_SYMBOL_CQPSK && d->cqpsk_enable
                              && d->mode_demod != &raw_demod && d->downsample_passes <= 0 && !d->iq_dc_block_enable
                              && !d->iqbal_enable && d->lowpassed && d->lp_len >= 8 && d->lp_len <= MAXIMUM_BUF_LENGTH
                              && (d->ted_sps <= 0 || (d->ted_sps >= 2 && d->ted_sps <= 10));
    if (cqpsk_covered && full_demod_cqpsk_b200(d)) {
        return;
    }
    const int covered = d && d->output_kind == DSD_DEMOD_OUTPUT_FSK_DISCRIMINATOR && !d->cqpsk_enable
                        && d->downsample_passes <= 0 && !d->iq_dc_block_enable && !d->iqbal_enable && d->lowpassed
                        && d->lp_len >= 2 && d->lp_len <= MAXIMUM_BUF_LENGTH;
    if (!covered) {
        full_demod_cpu(d); /* configurations this library does not build stay on the reference's own code */
        return;
    }
    dsdneo_b200_demod_bank* bank = bank_for(d);
    if (!bank) {
        /* no silent fallback for a covered configuration: the table is full (full_demod_b200_forget was never called) or the
         * device refused the bank */
        static int warned = 0;
        if (!warned++) {
            fprintf(stderr, "full_demod (b200): no device bank for demod_state %p: %s\n", (const void*)d, dsdneo_b200_last_error());
        }
        d->result_len = 0;
        return;
    }
    const int slot = slot_of(d);
    /* the reference resets the modem on retune (dsd_fsk_modem_reset clears have_prev, dc and peak estimates): mirror it */
    if (slot >= 0 && g_banks[slot].mirrored_have_prev && !d->fsk_modem_state.have_prev) {
        (void)dsdneo_b200_demod_bank_reset(bank, NULL);
    }
    const int pairs = d->lp_len >> 1;
    if (dsdneo_b200_full_demod_batch_host(bank, d->lowpassed, (size_t)pairs, pairs, 1, d->result, (size_t)pairs) != 0) {
        /* No silent fallback on a GPU failure: surface it the way the reference surfaces a dead stream. */
        d->result_len = 0;
        return;
    }
    d->result_len = pairs;
    dsdneo_b200_demod_chan_state st;
    if (dsdneo_b200_demod_bank_get_state(bank, 0, &st) == 0) {
        d->channel_pwr = st.channel_pwr;
        d->channel_squelched = st.channel_squelched;
        d->squelch_gate_open = !st.channel_squelched;
        d->fsk_modem_state.prev_i = st.prev_i;
        d->fsk_modem_state.prev_q = st.prev_q;
        d->fsk_modem_state.have_prev = st.have_prev;
        if (slot >= 0) {
            g_banks[slot].mirrored_have_prev = st.have_prev;
        }
        d->fsk_modem_state.dc_est = st.dc_est;
        d->fsk_modem_state.discriminator_peak_est = st.discriminator_peak_est;
    }
}
