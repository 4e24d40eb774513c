/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's CQPSK block side
 * (SURVEY.md section 8f rank 3): what full_demod() does for output_kind == SYMBOL_CQPSK after the
 * channel LPF / squelch step (arancormonk/dsd-neo @ 4d06905):
 *
 *   cqpsk_rms_agc                src/dsp/demod_pipeline.cpp:796-842
 *   op25_fll_band_edge_cc        src/dsp/costas.cpp:1176-1224 (+ helpers :635-765, NCO :80-133,
 *                                 band-edge design :936-1024)
 *   op25_gardner_cc              src/dsp/costas.cpp:804-858 (+ helpers :352-534),
 *                                 MMSE 8-tap interpolator src/dsp/mmse_interp.cpp:9-99
 *   op25_diff_phasor_cc          src/dsp/costas.cpp:872-902
 *   op25_costas_loop_cc          src/dsp/costas.cpp:935-961 (+ helpers :179-259, :536-633)
 *   qpsk_differential_demod      src/dsp/demod_pipeline.cpp:742-764 (atan approximation :73-98)
 *   squelched block              src/dsp/demod_pipeline.cpp:1022-1040
 *
 * Parity status: PINNED against the unmodified reference full_demod() compiled into
 * oracle/_ref/libdsdneo_ref.so (tests/test_oracle_cqpsk.py: symbols and carried loop state, bit for bit,
 * over multi-block streams with frequency offset, noise, squelch transitions and odd block sizes).
 *
 * Differences in form (not in arithmetic): the reference keeps two "doubled" circular delay lines (one in
 * the FLL, one in the Gardner block).  Both hold the same sample stream -- the FLL's own output -- so the
 * restatement keeps ONE history ring of FLL outputs per channel; the band-edge filters read its newest
 * n_taps entries and the interpolator reads entries (pushed - T + j).  The Gardner loop is restated in
 * push form: before sample i is consumed, symbols are emitted while mu <= 1.
 *
 * Out of contract (documented in DESIGN.md): non-finite samples, blocks shorter than 4 pairs (the
 * reference then skips timing recovery and emits sample-rate garbage, costas.cpp:811-813), sps changes on
 * a live channel (re-create the channel instead), sps > 10.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

static const float kTwoPiF = 6.28318530717958647692f;
static const float kPiF = 3.14159265358979323846f;

/* GNU Radio MMSE interpolator taps, every eighth row of interpolator_taps.h (mmse_interp.cpp:17-50). */
static const float kMmse[17][8] = {
    {0.00000e+00f, 0.00000e+00f, 0.00000e+00f, 0.00000e+00f, 1.00000e+00f, 0.00000e+00f, 0.00000e+00f, 0.00000e+00f},
    {-1.23337e-03f, 6.84261e-03f, -2.24178e-02f, 6.57852e-02f, 9.83392e-01f, -4.04519e-02f, 9.56876e-03f, -1.54221e-03f},
    {-2.43121e-03f, 1.35716e-02f, -4.49929e-02f, 1.36968e-01f, 9.55956e-01f, -7.43154e-02f, 1.80759e-02f, -2.94361e-03f},
    {-3.55283e-03f, 1.99599e-02f, -6.70018e-02f, 2.12443e-01f, 9.18329e-01f, -1.01501e-01f, 2.53295e-02f, -4.16581e-03f},
    {-4.55932e-03f, 2.57844e-02f, -8.77011e-02f, 2.91006e-01f, 8.71305e-01f, -1.22047e-01f, 3.11866e-02f, -5.17776e-03f},
    {-5.41467e-03f, 3.08323e-02f, -1.06342e-01f, 3.71376e-01f, 8.15826e-01f, -1.36111e-01f, 3.55525e-02f, -5.95620e-03f},
    {-6.08674e-03f, 3.49066e-02f, -1.22185e-01f, 4.52218e-01f, 7.52958e-01f, -1.43968e-01f, 3.83800e-02f, -6.48585e-03f},
    {-6.54823e-03f, 3.78315e-02f, -1.34515e-01f, 5.32164e-01f, 6.83875e-01f, -1.45993e-01f, 3.96678e-02f, -6.75943e-03f},
    {-6.77751e-03f, 3.94578e-02f, -1.42658e-01f, 6.09836e-01f, 6.09836e-01f, -1.42658e-01f, 3.94578e-02f, -6.77751e-03f},
    {-6.73929e-03f, 3.95900e-02f, -1.46043e-01f, 6.92808e-01f, 5.22267e-01f, -1.33190e-01f, 3.75341e-02f, -6.50285e-03f},
    {-6.48585e-03f, 3.83800e-02f, -1.43968e-01f, 7.52958e-01f, 4.52218e-01f, -1.22185e-01f, 3.49066e-02f, -6.08674e-03f},
    {-5.95620e-03f, 3.55525e-02f, -1.36111e-01f, 8.15826e-01f, 3.71376e-01f, -1.06342e-01f, 3.08323e-02f, -5.41467e-03f},
    {-5.17776e-03f, 3.11866e-02f, -1.22047e-01f, 8.71305e-01f, 2.91006e-01f, -8.77011e-02f, 2.57844e-02f, -4.55932e-03f},
    {-4.16581e-03f, 2.53295e-02f, -1.01501e-01f, 9.18329e-01f, 2.12443e-01f, -6.70018e-02f, 1.99599e-02f, -3.55283e-03f},
    {-2.94361e-03f, 1.80759e-02f, -7.43154e-02f, 9.55956e-01f, 1.36968e-01f, -4.49929e-02f, 1.35716e-02f, -2.43121e-03f},
    {-1.54221e-03f, 9.56876e-03f, -4.04519e-02f, 9.83392e-01f, 6.57852e-02f, -2.24178e-02f, 6.84261e-03f, -1.23337e-03f},
    {0.00000e+00f, 0.00000e+00f, 0.00000e+00f, 1.00000e+00f, 0.00000e+00f, 0.00000e+00f, 0.00000e+00f, 0.00000e+00f},
};

const float*
oracle_cqpsk_mmse_table(void) {
    return &kMmse[0][0];
}

static float
clip_sym(float x, float lim) {
    return x > lim ? lim : (x < -lim ? -lim : x);
}

static float
clamp_rng(float x, float lo, float hi) {
    return x < lo ? lo : (x > hi ? hi : x);
}

/* second-order loop gains, GNU Radio control_loop::update_gains as used at costas.cpp:541-545, :648-655 */
static void
loop_gains(float loop_bw, float* alpha, float* beta) {
    const float damping = 0.70710678118654752440f;
    const float denom = 1.0f + 2.0f * damping * loop_bw + loop_bw * loop_bw;
    *alpha = (4.0f * damping * loop_bw) / denom;
    *beta = (4.0f * loop_bw * loop_bw) / denom;
}

/* costas.cpp:936-1024: half-sine band-edge prototype (sum of two sincs), divided by its POWER (not its
 * root), modulated to -/+ (1 + rolloff) / (2 sps) and stored time-reversed.  Returns n_taps. */
int
oracle_fll_band_edge_design(int sps, float* lower_r, float* lower_i, float* upper_r, float* upper_i, int max_taps) {
    const float rolloff = 0.2f;
    int n_taps = 2 * sps + 1; /* costas.cpp:662 */
    if (n_taps > ORACLE_FLL_MAX_TAPS) {
        n_taps = ORACLE_FLL_MAX_TAPS;
    }
    if (n_taps < 3) {
        n_taps = 3;
    }
    if (n_taps > max_taps) {
        return -1;
    }
    const float m_span = roundf((float)n_taps / (float)sps);
    const int half = (n_taps - 1) / 2;
    float proto[ORACLE_FLL_MAX_TAPS];
    float power = 0.0f;
    for (int i = 0; i < n_taps; i++) {
        const float k = -m_span + (float)i * 2.0f / (float)sps;
        const float a_lo = rolloff * k - 0.5f;
        const float a_hi = rolloff * k + 0.5f;
        const float s_lo = (fabsf(a_lo) < 1e-6f) ? 1.0f : sinf(kPiF * a_lo) / (kPiF * a_lo);
        const float s_hi = (fabsf(a_hi) < 1e-6f) ? 1.0f : sinf(kPiF * a_hi) / (kPiF * a_hi);
        proto[i] = s_lo + s_hi;
        power += proto[i] * proto[i];
    }
    if (power > 0.0f) {
        const float norm = 1.0f / power;
        for (int i = 0; i < n_taps; i++) {
            proto[i] *= norm;
        }
    }
    for (int i = 0; i < n_taps; i++) {
        const float f = (float)(i - half) / (2.0f * (float)sps);
        const float ph = kTwoPiF * (1.0f + rolloff) * f;
        const int r = n_taps - 1 - i;
        lower_r[r] = proto[i] * cosf(-ph);
        lower_i[r] = proto[i] * sinf(-ph);
        upper_r[r] = proto[i] * cosf(ph);
        upper_i[r] = proto[i] * sinf(ph);
    }
    return n_taps;
}

int
oracle_cqpsk_chan_init(oracle_cqpsk_chan* q, int rate_out_hz, int sps, float ted_gain, int ted_gain_is_set) {
    memset(q, 0, sizeof(*q));
    if (sps < 2 || sps > 10) {
        return -1;
    }
    q->rate_out_hz = rate_out_hz;
    q->sps = sps;
    q->ted_gain = ted_gain;
    q->ted_gain_is_set = ted_gain_is_set;
    q->agc_avg = 1.0f;      /* rtl_demod_config.cpp:366 */
    q->diff_prev_r = 1.0f;  /* rtl_demod_config.cpp:364 */
    q->diff_prev_j = 0.0f;
    /* FLL, costas.cpp:660-668 (first init: freq = 0, phase = 0, delay cleared) */
    q->fll_ntaps = oracle_fll_band_edge_design(sps, q->fll_lower_r, q->fll_lower_i, q->fll_upper_r, q->fll_upper_i,
                                               ORACLE_FLL_MAX_TAPS);
    loop_gains(kTwoPiF / (float)sps / 350.0f, &q->fll_alpha, &q->fll_beta);
    /* Gardner first init, costas.cpp:365-398 */
    q->mu = (float)sps;
    q->omega = (float)sps;
    q->omega_rel = 0.002f;
    q->omega_mid = q->omega;
    {
        const float omega_max = q->omega * (1.0f + q->omega_rel);
        const int t_op25 = 2 * (int)ceilf(omega_max);
        const int t_mmse = (int)ceilf(omega_max / 2.0f) + 8 + 1;
        q->ted_span = t_op25 > t_mmse ? t_op25 : t_mmse;
    }
    /* Costas, costas.cpp:536-551 */
    loop_gains(0.008f, &q->costas_alpha, &q->costas_beta);
    return 0;
}

/* costas.cpp:80-100: Maclaurin sine / cosine through x^11 / x^10, Horner form */
static void
sincos_poly(float x, float* s, float* c) {
    const float x2 = x * x;
    *s = x
         * (1.0f
            + x2
                  * (-0.16666666666666666667f
                     + x2
                           * (0.00833333333333333333f
                              + x2 * (-0.00019841269841269841f + x2 * (0.00000275573192239859f + x2 * -0.00000002505210838544f)))));
    *c = 1.0f
         + x2
               * (-0.5f
                  + x2
                        * (0.04166666666666666667f
                           + x2 * (-0.00138888888888888889f + x2 * (0.00002480158730158730f + x2 * -0.00000027557319223986f))));
}

/* costas.cpp:102-133 */
static void
sincos_wrapped(float ph, float* s, float* c) {
    if (!isfinite(ph) || ph < -kTwoPiF || ph > kTwoPiF) {
        *s = sinf(ph);
        *c = cosf(ph);
        return;
    }
    if (ph > kPiF) {
        ph -= kTwoPiF;
    } else if (ph < -kPiF) {
        ph += kTwoPiF;
    }
    if (ph > (kPiF / 2.0f)) {
        float cc;
        sincos_poly(kPiF - ph, s, &cc);
        *c = -cc;
    } else if (ph < (-kPiF / 2.0f)) {
        float cc;
        sincos_poly(-kPiF - ph, s, &cc);
        *c = -cc;
    } else {
        sincos_poly(ph, s, c);
    }
}

static void
ring_get(const oracle_cqpsk_chan* q, long n, float* r, float* j) {
    if (n < 0) { /* before the stream: cleared delay lines */
        *r = 0.0f;
        *j = 0.0f;
        return;
    }
    *r = q->ring_r[n % ORACLE_CQPSK_RING];
    *j = q->ring_j[n % ORACLE_CQPSK_RING];
}

/* mmse_interp.cpp:52-82: taps blended linearly between rows floor(16 mu) and floor(16 mu) + 1, applied back to front to
 * 8 consecutive history samples starting at sample number `base` */
static void
mmse8(const oracle_cqpsk_chan* q, long base, float mu, float* out_r, float* out_j) {
    float pos = mu * 16.0f;
    int lo = (int)pos;
    float frac = pos - (float)lo;
    if (lo < 0) {
        lo = 0;
        frac = 0.0f;
    }
    if (lo >= 16) {
        lo = 15;
        frac = 1.0f;
    }
    const float w_lo = 1.0f - frac;
    float acc_r = 0.0f, acc_j = 0.0f;
    float sr[8], sj[8];
    for (int i = 0; i < 8; i++) {
        ring_get(q, base + i, &sr[i], &sj[i]);
    }
    for (int i = 0; i < 8; i++) {
        const float tap = w_lo * kMmse[lo][i] + frac * kMmse[lo + 1][i];
        acc_r += tap * sr[7 - i];
    }
    for (int i = 0; i < 8; i++) {
        const float tap = w_lo * kMmse[lo][i] + frac * kMmse[lo + 1][i];
        acc_j += tap * sj[7 - i];
    }
    *out_r = acc_r;
    *out_j = acc_j;
}

static float
smoothstep_f(float e0, float e1, float x) {
    if (x <= e0) {
        return 0.0f;
    }
    if (x >= e1) {
        return 1.0f;
    }
    const float t = (x - e0) / (e1 - e0);
    return t * t * (3.0f - 2.0f * t);
}

/* demod_pipeline.cpp:73-98 */
static float
atan2_qpsk(float y, float x) {
    if (x == 0.0f && y == 0.0f) {
        return 0.0f;
    }
    const float ax = fabsf(x), ay = fabsf(y);
    if (ax >= ay) {
        const float r = y / x;
        float ang = r * (0.78539816339744830962f - (fabsf(r) - 1.0f) * (0.2447f + 0.0663f * fabsf(r)));
        if (x < 0.0f) {
            ang += (y < 0.0f) ? -3.14159265358979323846f : 3.14159265358979323846f;
        }
        return ang;
    }
    const float r = x / y;
    const float ang = r * (0.78539816339744830962f - (fabsf(r) - 1.0f) * (0.2447f + 0.0663f * fabsf(r)));
    return (y > 0.0f) ? (1.57079632679489661923f - ang) : (-1.57079632679489661923f - ang);
}

/* One symbol through op25_diff_phasor_cc, the Costas loop and the output phase extractor. */
static float
symbol_back_end(oracle_cqpsk_chan* q, float sym_r, float sym_j) {
    /* costas.cpp:884-898: y = x conj(prev) */
    const float d_r = sym_r * q->diff_prev_r + sym_j * q->diff_prev_j;
    const float d_j = sym_j * q->diff_prev_r - sym_r * q->diff_prev_j;
    q->diff_prev_r = sym_r;
    q->diff_prev_j = sym_j;

    /* costas.cpp:572-608 */
    float nco_r, nco_j;
    sincos_poly(-q->costas_phase, &nco_j, &nco_r);
    const float rot_r = d_r * nco_r - d_j * nco_j;
    const float rot_j = d_r * nco_j + d_j * nco_r;

    /* normalize_costas_detector_sample, costas.cpp:229-259 */
    float det_r, det_j, conf;
    const float mag2 = rot_r * rot_r + rot_j * rot_j;
    if (!isfinite(mag2)) {
        det_r = det_j = 0.0f;
        conf = 0.0f;
    } else if (mag2 <= 0.10f * 0.10f) {
        det_r = rot_r;
        det_j = rot_j;
        conf = 0.0f;
    } else {
        const float mag = sqrtf(mag2);
        conf = (mag2 >= 0.35f * 0.35f) ? 1.0f : (isfinite(mag) ? smoothstep_f(0.10f, 0.35f, mag) : 0.0f);
        const float scale = (0.85f * 0.85f) / mag;
        if (!isfinite(scale)) {
            det_r = det_j = 0.0f;
            conf = 0.0f;
        } else {
            det_r = rot_r * scale;
            det_j = rot_j * scale;
        }
    }

    float err = 0.0f, err_raw = 0.0f;
    if (conf <= 0.0f || !isfinite(conf)) {
        q->costas_err_smooth = 0.0f;
        q->m_zero_conf++;
    } else {
        /* phase_detector_4, costas.cpp:179-182 */
        const float pd = (det_r > 0.0f ? 1.0f : -1.0f) * det_j - (det_j > 0.0f ? 1.0f : -1.0f) * det_r;
        err_raw = clip_sym(pd * conf, 1.0f);
        /* cqpsk_costas_error_smooth_alpha, costas.cpp:215-227 */
        float a = 0.25f;
        if (isfinite(err_raw) && isfinite(q->costas_err_smooth) && !(fabsf(q->costas_err_smooth) <= 1.0e-6f)) {
            const float kick = smoothstep_f(0.02f, 0.18f, fabsf(err_raw - q->costas_err_smooth));
            a = 0.25f + (0.10f - 0.25f) * kick;
        }
        q->costas_err_smooth += a * (err_raw - q->costas_err_smooth);
        err = clip_sym(q->costas_err_smooth, 1.0f);
        q->m_conf_acc += conf;
    }
    q->costas_error = err;
    q->m_err_abs += fabsf(err);
    q->m_err_raw_abs += fabsf(err_raw);
    q->costas_freq += q->costas_beta * err;
    q->costas_phase += q->costas_freq + q->costas_alpha * err;
    q->costas_phase = clamp_rng(q->costas_phase, -(kPiF / 2.0f), kPiF / 2.0f);
    q->costas_freq = clamp_rng(q->costas_freq, -1.0f, 1.0f);

    /* demod_pipeline.cpp:755-761 */
    const float k4_over_pi = 4.0f / 3.14159265358979323846f;
    return atan2_qpsk(det_j, det_r) * k4_over_pi;
}

/* Gardner gain schedule, costas.cpp:143-168 (no env override, i.e. cfg->ted_gain_is_set == 0) */
static float
gardner_gain_mu(const oracle_cqpsk_chan* q) {
    const float requested = (q->ted_gain > 0.0f) ? q->ted_gain : 0.025f;
    if (q->ted_gain_is_set) {
        return requested;
    }
    const int sym_rate = (q->rate_out_hz <= 0) ? 4800 : (q->rate_out_hz + q->sps / 2) / q->sps;
    if (sym_rate < 5500 || q->lock_count < 240) {
        return requested;
    }
    if (q->lock_accum / (float)q->lock_count < 0.05f) {
        return requested;
    }
    return 0.018f;
}

/* One un-squelched block of channel-filtered samples -> symbols.  Returns the number of symbols. */
int
oracle_cqpsk_block(oracle_cqpsk_chan* q, const float* lp, int n_floats, float* out) {
    const int pairs = n_floats / 2;
    if (pairs < 4) {
        return -1;
    }
    const float gain_mu = gardner_gain_mu(q);
    const float gain_omega = 0.1f * gain_mu * gain_mu;
    q->ted_effective_gain = gain_mu;
    /* per-block Costas context, costas.cpp:553-569 */
    q->costas_phase = isfinite(q->costas_phase) ? clamp_rng(q->costas_phase, -(kPiF / 2.0f), kPiF / 2.0f) : 0.0f;
    if (!isfinite(q->costas_err_smooth)) {
        q->costas_err_smooth = 0.0f;
    }
    q->m_err_abs = q->m_err_raw_abs = q->m_conf_acc = 0.0f;
    q->m_zero_conf = 0;

    float avg = q->agc_avg;
    if (avg <= 0.0f) {
        avg = 1.0f;
    }
    const float agc_alpha = 0.45f, agc_beta = 1.0f - agc_alpha, agc_ref = 0.85f;
    int n_sym = 0;
    int timing_open = 1; /* cleared if the output buffer bound of costas.cpp:830 (o < buf_len) is hit */
    for (int i = 0; i < pairs; i++) {
        /* --- AGC, demod_pipeline.cpp:819-838 --- */
        float xr = lp[2 * i], xj = lp[2 * i + 1];
        const float mag2 = xr * xr + xj * xj;
        avg = agc_beta * avg + agc_alpha * mag2;
        if (avg > 0.0f) {
            const float sc = agc_ref / sqrtf(avg);
            xr = xr * sc;
            xj = xj * sc;
        }
        /* --- FLL, costas.cpp:742-765 --- */
        float s, c;
        sincos_wrapped(q->fll_phase, &s, &c);
        const float yr = xr * c - xj * s;
        const float yj = xr * s + xj * c;
        q->ring_r[q->pushed % ORACLE_CQPSK_RING] = yr;
        q->ring_j[q->pushed % ORACLE_CQPSK_RING] = yj;
        float lo_r = 0.0f, lo_j = 0.0f, up_r = 0.0f, up_j = 0.0f;
        for (int k = 0; k < q->fll_ntaps; k++) { /* newest first, costas.cpp:709-718 */
            float dr, dj;
            ring_get(q, q->pushed - k, &dr, &dj);
            lo_r += dr * q->fll_lower_r[k] - dj * q->fll_lower_i[k];
            lo_j += dr * q->fll_lower_i[k] + dj * q->fll_lower_r[k];
            up_r += dr * q->fll_upper_r[k] - dj * q->fll_upper_i[k];
            up_j += dr * q->fll_upper_i[k] + dj * q->fll_upper_r[k];
        }
        const float lo_p = lo_r * lo_r + lo_j * lo_j;
        const float up_p = up_r * up_r + up_j * up_j;
        const float ferr = clip_sym(up_p - lo_p, 1.0f);
        q->fll_freq += q->fll_beta * ferr;
        q->fll_freq = clamp_rng(q->fll_freq, -1.0f, 1.0f);
        q->fll_phase += q->fll_freq + q->fll_alpha * ferr;
        while (q->fll_phase > kTwoPiF) {
            q->fll_phase -= kTwoPiF;
        }
        while (q->fll_phase < -kTwoPiF) {
            q->fll_phase += kTwoPiF;
        }

        /* --- Gardner, costas.cpp:830-855, push form: emit while mu <= 1, then consume this sample --- */
        if (timing_open) {
            while (!(q->mu > 1.0f)) {
                if (n_sym >= pairs) {
                    timing_open = 0;
                    break;
                }
                /* gardner_compute_half_timing, costas.cpp:475-489 */
                const float half_omega = q->omega / 2.0f;
                int hs = (int)floorf(half_omega);
                float hmu = q->mu + half_omega - (float)hs;
                if (hmu > 1.0f) {
                    hmu -= 1.0f;
                    hs += 1;
                }
                if (hs < 0) {
                    hs = 0;
                }
                /* delay line = the last ted_span consumed samples; its oldest entry is sample (consumed - span) */
                const long oldest = q->consumed - q->ted_span;
                const int widx = (int)(q->consumed % q->ted_span); /* the reference's dl_index */
                if (widx + 7 >= 2 * q->ted_span || widx + hs + 7 >= 2 * q->ted_span) { /* costas.cpp:494-498 */
                    q->mu += q->omega;
                    continue;
                }
                float mid_r, mid_j, sym_r, sym_j;
                mmse8(q, oldest, q->mu, &mid_r, &mid_j);
                mmse8(q, oldest + hs, hmu, &sym_r, &sym_j);
                /* costas.cpp:505-513 */
                float terr = (q->last_r - sym_r) * mid_r + (q->last_j - sym_j) * mid_j;
                if (terr != terr) {
                    terr = 0.0f;
                }
                terr = clip_sym(terr, 1.0f);
                /* lock detector, costas.cpp:516-526 */
                {
                    const float ie2 = sym_r * sym_r, io2 = mid_r * mid_r, qe2 = sym_j * sym_j, qo2 = mid_j * mid_j;
                    const float yi = ((ie2 + io2) != 0.0f) ? (ie2 - io2) / (ie2 + io2) : 0.0f;
                    const float yq = ((qe2 + qo2) != 0.0f) ? (qe2 - qo2) / (qe2 + qo2) : 0.0f;
                    q->lock_accum += yi + yq;
                    q->lock_count++;
                }
                /* loop update, costas.cpp:528-534 */
                const float smag = sqrtf(sym_r * sym_r + sym_j * sym_j);
                q->omega += gain_omega * terr * smag;
                q->omega = q->omega_mid + clip_sym(q->omega - q->omega_mid, q->omega_rel);
                q->mu += q->omega + gain_mu * terr;
                q->last_r = sym_r;
                q->last_j = sym_j;
                out[n_sym++] = symbol_back_end(q, sym_r, sym_j);
            }
            if (timing_open) {
                q->mu -= 1.0f;
                q->consumed++;
            }
        }
        q->pushed++;
    }
    q->agc_avg = avg;
    /* costas_store_metrics, costas.cpp:610-625 (the Costas block sees n_sym pairs; below 1 symbol it does not run) */
    if (n_sym >= 1) {
        const float inv = 1.0f / (float)n_sym;
        long v;
        v = lrintf(q->m_err_abs * inv * 16384.0f);
        q->costas_err_avg_q14 = (int)(v < 0 ? 0 : (v > 32767 ? 32767 : v));
        v = lrintf(q->m_err_raw_abs * inv * 16384.0f);
        q->costas_err_raw_avg_q14 = (int)(v < 0 ? 0 : (v > 32767 ? 32767 : v));
        v = lrintf(q->m_conf_acc * inv * 16384.0f);
        q->costas_conf_avg_q14 = (int)(v < 0 ? 0 : (v > 16384 ? 16384 : v));
        v = lrint((100.0 * (double)q->m_zero_conf) / (double)n_sym);
        q->costas_zero_conf_pct = (int)(v < 0 ? 0 : (v > 100 ? 100 : v));
    }
    return n_sym;
}

/* full_demod() for the CQPSK symbol output kind (demod_pipeline.cpp:1330-1350): channel LPF, power / squelch, then either
 * ceil(pairs / sps) zero symbols (squelched, :1022-1040) or the chain above. */
int
oracle_full_demod_cqpsk_block(oracle_demod_chan* c, oracle_cqpsk_chan* q, const float* iq, int n_floats, float* scratch,
                              float* out) {
    const float* lp = iq;
    if (c->lpf_enable && n_floats >= 2 && c->taps_len >= 3) {
        oracle_fir_complex(iq, n_floats, scratch, c->hist_i, c->hist_q, c->taps, c->taps_len, c->fir_fma);
        lp = scratch;
    }
    const int pairs = n_floats / 2;
    if (n_floats >= 2) {
        c->channel_pwr = oracle_mean_power(lp, n_floats > 512 ? 512 : n_floats, 1);
    }
    if (n_floats > 0 && c->squelch_level > 0.0f && c->channel_pwr < c->squelch_level) {
        c->channel_squelched = 1;
        int n = (pairs + q->sps - 1) / q->sps;
        if (n < 1) {
            n = 1;
        }
        for (int i = 0; i < n; i++) {
            out[i] = 0.0f;
        }
        return n;
    }
    c->channel_squelched = 0;
    return oracle_cqpsk_block(q, lp, n_floats, out);
}
