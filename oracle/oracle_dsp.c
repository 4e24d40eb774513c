/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the block side of the hot path.
 * See oracle.h for the rules.  Each function cites the reference code it restates
 * (paths relative to arancormonk/dsd-neo @ 4d06905).
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

#define ORACLE_PI 3.14159265358979323846

/* ---- channel LPF design ------------------------------------------------------------- */
/* src/dsp/demod_pipeline.cpp:133-149 (cutoffs), :443-460 (clamp + call), src/dsp/firdes.cpp
 * (dsd_firdes_compute_ntaps, dsd_window_blackman, dsd_firdes_low_pass). */
int
oracle_channel_lpf_design(int rate_out_hz, int profile, float* taps_out, int max_taps) {
    static const double cutoff_by_profile[6] = {8600.0, 3725.0, 6850.0, 6850.0, 6850.0, 7250.0};
    if (rate_out_hz <= 0 || !taps_out || max_taps <= 0) {
        return -1;
    }
    double fs = (double)rate_out_hz;
    double cutoff = (profile >= 0 && profile < 6) ? cutoff_by_profile[profile] : cutoff_by_profile[0];
    double hi = fs * 0.5 * 0.90;
    if (cutoff < 100.0) {
        cutoff = 100.0;
    }
    if (cutoff > hi) {
        cutoff = hi;
    }
    if (cutoff <= 0.0 || cutoff > fs / 2.0) {
        return -1;
    }
    int ntaps = (int)(74.0 * fs / (22.0 * 1200.0));
    ntaps |= 1; /* even -> next odd */
    if (ntaps > max_taps || ntaps > 1024) {
        return -1;
    }
    float w[1024];
    float Mf = (float)(ntaps - 1);
    for (int n = 0; n < ntaps; n++) {
        w[n] = 0.42f - 0.5f * cosf((2.0f * (float)ORACLE_PI * (float)n) / Mf)
               + 0.08f * cosf((4.0f * (float)ORACLE_PI * (float)n) / Mf);
    }
    int M = (ntaps - 1) / 2;
    double fwT0 = 2.0 * ORACLE_PI * cutoff / fs;
    for (int n = -M; n <= M; n++) {
        if (n == 0) {
            taps_out[M] = (float)((fwT0 / ORACLE_PI) * w[M]);
        } else {
            taps_out[n + M] = (float)((sin(n * fwT0) / (n * ORACLE_PI)) * w[n + M]);
        }
    }
    double fmax = taps_out[M];
    for (int n = 1; n <= M; n++) {
        fmax += 2.0 * taps_out[n + M];
    }
    double gain = 1.0 / fmax;
    for (int i = 0; i < ntaps; i++) {
        taps_out[i] *= (float)gain;
    }
    return ntaps;
}

/* ---- symmetric complex FIR ------------------------------------------------------------ */
/* fma == 0: src/dsp/simd_fir.cpp:55-133 (simd_fir_complex_apply_scalar; the SSE2 kernel
 *           src/dsp/simd_fir_sse2.cpp:262-345 performs the same per-output operations).
 * fma == 1: src/dsp/simd_fir_avx2.cpp:120-141,399-450 (fused multiply-add per tap pair). */
void
oracle_fir_complex(const float* in, int in_len, float* out, float* hist_i, float* hist_q, const float* taps,
                   int taps_len, int fma) {
    if (taps_len < 3 || !(taps_len & 1) || in_len < 2) {
        return;
    }
    if (in_len < taps_len * 2) {
        fma = 0; /* simd_fir_prefer_scalar_for_block, src/dsp/simd_fir.cpp:302-305,353-356: short blocks take the scalar kernel */
    }
    int N = in_len / 2;
    int H = taps_len - 1;
    int c = H / 2;
#define SAMPLE_I(idx) ((idx) < H ? hist_i[(idx)] : ((idx) - H < N ? in[2 * ((idx) - H)] : in[2 * (N - 1)]))
#define SAMPLE_Q(idx) ((idx) < H ? hist_q[(idx)] : ((idx) - H < N ? in[2 * ((idx) - H) + 1] : in[2 * (N - 1) + 1]))
    for (int n = 0; n < N; n++) {
        int mid = H + n;
        float ai, aq;
        if (fma) {
            ai = fmaf(taps[c], SAMPLE_I(mid), 0.0f);
            aq = fmaf(taps[c], SAMPLE_Q(mid), 0.0f);
        } else {
            ai = 0.0f;
            aq = 0.0f;
            ai += taps[c] * SAMPLE_I(mid);
            aq += taps[c] * SAMPLE_Q(mid);
        }
        for (int k = 0; k < c; k++) {
            float t = taps[k];
            if (t == 0.0f) {
                continue;
            }
            int d = c - k;
            float si = SAMPLE_I(mid - d) + SAMPLE_I(mid + d);
            float sq = SAMPLE_Q(mid - d) + SAMPLE_Q(mid + d);
            if (fma) {
                ai = fmaf(t, si, ai);
                aq = fmaf(t, sq, aq);
            } else {
                ai += t * si;
                aq += t * sq;
            }
        }
        out[2 * n] = ai;
        out[2 * n + 1] = aq;
    }
#undef SAMPLE_I
#undef SAMPLE_Q
    /* history := last H inputs (simd_fir.cpp:117-132) */
    if (N >= H) {
        for (int k = 0; k < H; k++) {
            hist_i[k] = in[2 * (N - H + k)];
            hist_q[k] = in[2 * (N - H + k) + 1];
        }
    } else {
        int keep = H - N;
        memmove(hist_i, hist_i + N, (size_t)keep * sizeof(float));
        memmove(hist_q, hist_q + N, (size_t)keep * sizeof(float));
        for (int k = 0; k < N; k++) {
            hist_i[keep + k] = in[2 * k];
            hist_q[keep + k] = in[2 * k + 1];
        }
    }
}

/* ---- complex half-band decimator by 2 and its cascade --------------------------------------- */
/* Taps: src/dsp/halfband.cpp:35-74 (Q15-normalised; odd taps zero except the centre). */
const float oracle_hb15_taps[15] = {-108.0f / 32768.0f, 0.0f, 1800.0f / 32768.0f, 0.0f, -500.0f / 32768.0f, 0.0f,
                                    7000.0f / 32768.0f, 0.5f, 7000.0f / 32768.0f, 0.0f, -500.0f / 32768.0f, 0.0f,
                                    1800.0f / 32768.0f, 0.0f, -108.0f / 32768.0f};
const float oracle_hb31_taps[31] = {0.0f, 0.0f, 13.0f / 32768.0f, 0.0f, -73.0f / 32768.0f, 0.0f, 233.0f / 32768.0f, 0.0f,
                                    -587.0f / 32768.0f, 0.0f, 1314.0f / 32768.0f, 0.0f, -2953.0f / 32768.0f, 0.0f,
                                    10244.0f / 32768.0f, 16386.0f / 32768.0f, 10244.0f / 32768.0f, 0.0f,
                                    -2953.0f / 32768.0f, 0.0f, 1314.0f / 32768.0f, 0.0f, -587.0f / 32768.0f, 0.0f,
                                    233.0f / 32768.0f, 0.0f, -73.0f / 32768.0f, 0.0f, 13.0f / 32768.0f, 0.0f, 0.0f};

/* fma == 0: simd_hb_decim2_complex_scalar, src/dsp/simd_fir.cpp:139-222 (the SSE2 kernel does the same per-output
 *           operations): acc = 0; acc += cc*x[centre]; for even e: acc += taps[e] * (x[-d] + x[+d]), zero taps skipped.
 * fma == 1: src/dsp/simd_fir_avx2.cpp:312-390: centre product, then acc = fma(tap, x[-d] + x[+d], acc).
 * Input beyond the block is the block's last sample; input before it is the carried history (taps_len - 1 samples).
 * Returns the number of floats written (2 * (pairs / 2)). */
int
oracle_hb_decim2_complex(const float* in, int in_len, float* out, float* hist_i, float* hist_q, const float* taps,
                         int taps_len, int fma) {
    if (taps_len < 3 || !(taps_len & 1)) {
        return 0;
    }
    int N = in_len / 2;
    if (N <= 0) {
        return 0;
    }
    if (in_len < taps_len * 2) {
        fma = 0; /* simd_fir_prefer_scalar_for_block, src/dsp/simd_fir.cpp:302-305,366-368 */
    }
    int n_out = N / 2;
    int H = taps_len - 1;
    int c = H / 2;
#define HB_I(rel) ((rel) < 0 ? hist_i[H + (rel)] : ((rel) < N ? in[2 * (rel)] : in[2 * (N - 1)]))
#define HB_Q(rel) ((rel) < 0 ? hist_q[H + (rel)] : ((rel) < N ? in[2 * (rel) + 1] : in[2 * (N - 1) + 1]))
    for (int n = 0; n < n_out; n++) {
        const int mid = 2 * n; /* the reference centres on src index (taps_len-1) + 2n of [hist | block | pad] */
        float ai, aq;
        if (fma) {
            ai = taps[c] * HB_I(mid);
            aq = taps[c] * HB_Q(mid);
        } else {
            ai = 0.0f;
            aq = 0.0f;
            ai += taps[c] * HB_I(mid);
            aq += taps[c] * HB_Q(mid);
        }
        for (int e = 0; e < c; e += 2) {
            float t = taps[e];
            if (t == 0.0f) {
                continue;
            }
            int d = c - e;
            float si = HB_I(mid - d) + HB_I(mid + d);
            float sq = HB_Q(mid - d) + HB_Q(mid + d);
            if (fma) {
                ai = fmaf(t, si, ai);
                aq = fmaf(t, sq, aq);
            } else {
                ai += t * si;
                aq += t * sq;
            }
        }
        out[2 * n] = ai;
        out[2 * n + 1] = aq;
    }
#undef HB_I
#undef HB_Q
    if (N >= H) {
        for (int k = 0; k < H; k++) {
            hist_i[k] = in[2 * (N - H + k)];
            hist_q[k] = in[2 * (N - H + k) + 1];
        }
    } else {
        int keep = H - N;
        memmove(hist_i, hist_i + N, (size_t)keep * sizeof(float));
        memmove(hist_q, hist_q + N, (size_t)keep * sizeof(float));
        for (int k = 0; k < N; k++) {
            hist_i[keep + k] = in[2 * k];
            hist_q[keep + k] = in[2 * k + 1];
        }
    }
    return 2 * n_out;
}

/* full_demod_apply_halfband_decimation, src/dsp/demod_pipeline.cpp:983-1001: `passes` stages, stage 0 = 31 taps, the rest
 * 15; hist is [passes][2][30].  `work` must hold in_len floats.  Result (in_len >> passes floats) is written to `out`. */
int
oracle_hb_cascade(const float* in, int in_len, int passes, float* hist /* [passes][2][30] */, float* work, float* out,
                  int fma) {
    const float* src = in;
    int len = in_len;
    for (int i = 0; i < passes; i++) {
        float* dst = (i == passes - 1) ? out : ((i & 1) ? work + in_len / 2 : work);
        len = oracle_hb_decim2_complex(src, len, dst, hist + (size_t)i * 60, hist + (size_t)i * 60 + 30,
                                       i == 0 ? oracle_hb31_taps : oracle_hb15_taps, i == 0 ? 31 : 15, fma);
        src = dst;
    }
    if (passes == 0) {
        memcpy(out, in, (size_t)in_len * sizeof(float));
    }
    return len;
}

/* ---- mean power ----------------------------------------------------------------------- */
/* src/dsp/demod_pipeline.cpp:926-945 */
float
oracle_mean_power(const float* samples, int len, int step) {
    double sum = 0.0, sumsq = 0.0;
    for (int i = 0; i < len; i += step) {
        double s = (double)samples[i];
        sum += s;
        sumsq += s * s;
    }
    double corr = len > 0 ? (sum * sum) / (double)len : 0.0;
    double e = sumsq - corr;
    if (e < 0.0) {
        e = 0.0;
    }
    return (float)(e / (double)(len > 0 ? len : 1));
}

/* ---- FSK discriminator ------------------------------------------------------------------ */
/* src/dsp/fsk_modem.c:23-35 (phase), :84-89 (conjugate product), :96-103 (dc), :116-133 (peak/scale/clip),
 * :135-164 (loop, first-sample rule) */
static float
disc_phase(float im, float re) {
    if (re > 1.0e-7f && fabsf(im) <= (0.35f * re)) {
        float x = im / re;
        float x2 = x * x;
        return x * (1.0f + x2 * (-0.3333333333333333f + x2 * 0.2f));
    }
    return atan2f(im, re);
}

static int
disc_process(oracle_demod_chan* c, const float* iq, int len, float* out) {
    int pairs = len / 2;
    for (int n = 0; n < pairs; n++) {
        float ci = iq[2 * n], cq = iq[2 * n + 1];
        if (!c->have_prev) {
            c->prev_i = ci;
            c->prev_q = cq;
            c->have_prev = 1;
            out[n] = 0.0f;
            continue;
        }
        float re = ci * c->prev_i + cq * c->prev_q;
        float im = cq * c->prev_i - ci * c->prev_q;
        float f = disc_phase(im, re);
        c->dc_est += 0.00025f * (f - c->dc_est);
        float cen = f - c->dc_est;
        float mag = fabsf(cen);
        if (mag > 1.0e-7f) {
            if (c->peak_est <= 1.0e-7f) {
                c->peak_est = mag;
            } else if (mag > c->peak_est) {
                c->peak_est += 0.125f * (mag - c->peak_est);
            } else {
                c->peak_est += 0.00005f * (mag - c->peak_est);
            }
        }
        float pk = c->peak_est <= 1.0e-7f ? 1.0f : c->peak_est;
        float o = cen * (30000.0f / pk);
        if (o > 32767.0f) {
            o = 32767.0f;
        }
        if (o < -32768.0f) {
            o = -32768.0f;
        }
        out[n] = o;
        c->prev_i = ci;
        c->prev_q = cq;
    }
    return pairs;
}

int
oracle_demod_chan_init(oracle_demod_chan* c, int rate_out_hz, int profile, int lpf_enable, float squelch_level,
                       int fir_fma) {
    memset(c, 0, sizeof(*c));
    c->rate_out_hz = rate_out_hz;
    c->lpf_profile = profile;
    c->lpf_enable = lpf_enable;
    c->squelch_level = squelch_level;
    c->fir_fma = fir_fma;
    c->taps_len = oracle_channel_lpf_design(rate_out_hz, profile, c->taps, ORACLE_LPF_MAX_TAPS);
    return c->taps_len > 0 ? 0 : -1;
}

/* src/dsp/demod_pipeline.cpp:1330-1350 restricted to output_kind == FSK discriminator, no half-band
 * passes, IQ DC block and IQ balance off: channel_lpf_apply (:526-555) -> mean_power/squelch
 * (:1003-1020) -> full_demod_handle_fsk_output (:1173-1190). */
int
oracle_full_demod_block(oracle_demod_chan* c, const float* iq, int n_floats, float* scratch, float* out) {
    const float* lp = iq;
    if (c->lpf_enable && n_floats >= 2 && c->taps_len >= 3) {
        oracle_fir_complex(iq, n_floats, scratch, c->hist_i, c->hist_q, c->taps, c->taps_len, c->fir_fma);
        lp = scratch;
    }
    int pairs = n_floats / 2;
    if (n_floats >= 2) {
        c->channel_pwr = oracle_mean_power(lp, n_floats > 512 ? 512 : n_floats, 1);
    }
    if (n_floats > 0 && c->squelch_level > 0.0f && c->channel_pwr < c->squelch_level) {
        c->channel_squelched = 1;
        /* dsd_fsk_modem_reset (fsk_modem.c:50-58) + zero result */
        c->prev_i = c->prev_q = 0.0f;
        c->have_prev = 0;
        c->dc_est = 0.0f;
        c->peak_est = 0.0f;
        for (int i = 0; i < pairs; i++) {
            out[i] = 0.0f;
        }
        return pairs;
    }
    c->channel_squelched = 0;
    return disc_process(c, lp, n_floats, out);
}

/* ---- polyphase channelizer: float64 direct form (no reference implementation exists) ------ */
void
oracle_pfb_direct(const float* xh, int hist_len, int n_in, const float* h, int L, int M, int D, const int* channels,
                  int n_sel, double* out, int n_out) {
    (void)n_in;
    for (int s = 0; s < n_sel; s++) {
        int k = channels[s];
        for (int n = 0; n < n_out; n++) {
            long t = (long)n * D + (M - 1); /* newest input index used by output n (relative to x[0]) */
            double ar = 0.0, ai = 0.0;
            for (int m = 0; m < L; m++) {
                long idx = t - m + hist_len;
                if (idx < 0) {
                    continue;
                }
                double xr = xh[2 * idx], xi = xh[2 * idx + 1];
                /* exp(-j 2 pi k (t-m)/M); reduce the integer phase index exactly */
                long ph = ((long)k * (t - m)) % M;
                if (ph < 0) {
                    ph += M;
                }
                double ang = -2.0 * ORACLE_PI * (double)ph / (double)M;
                double cr = cos(ang), ci = sin(ang);
                double hr = (double)h[m];
                ar += hr * (xr * cr - xi * ci);
                ai += hr * (xr * ci + xi * cr);
            }
            out[((long)s * n_out + n) * 2] = ar;
            out[((long)s * n_out + n) * 2 + 1] = ai;
        }
    }
}

/* The reference's large-angle branch is a plain libm call (src/dsp/fsk_modem.c:34); expose it over arrays
 * so device results can be compared with this host's libm without going through numpy's SIMD loops. */
void
oracle_libm_atan2f_array(const float* y, const float* x, float* out, long n) {
    for (long i = 0; i < n; i++) {
        out[i] = atan2f(y[i], x[i]);
    }
}
