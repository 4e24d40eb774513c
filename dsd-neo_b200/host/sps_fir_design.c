/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Host-side design of the sample-side matched filters (K9 taps: p25 / dmr / nxdn / dpmr / m17 filter at any
 * samples-per-symbol).
 *
 * Twin of the reference's per-sps redesign, which runs lazily inside the first filtered sample after a rate change:
 *   design_sps_fir            src/dsp/dsd_filters.c:153-170   (length rule, taps, normalisation)
 *   sps_fir_compute_taps_len  src/dsp/dsd_filters.c:94-116
 *   sps_fir_design_taps       src/dsp/dsd_filters.c:118-135   (base table copied, linearly re-sampled, or closed-form RRC)
 *   rrc_impulse / interp_base src/dsp/dsd_filters.c:43-92
 *   sps_fir_normalize_and_clear :137-151                      (f64 sum in tap order, f32 reciprocal, f32 products)
 * The symbolizer and the receive bank take NORMALISED taps (dsdneo_b200_symbolizer_config::filter_taps); this function is how
 * a caller gets them from the reference's own coefficient table for the channel's samples per symbol, bit-identical to what
 * the reference's filter would hold (same libm calls, same float / double staging; compile with -fno-fast-math
 * -ffp-contract=off).  The coefficient tables stay where they are (static data of dsd_filters.c): the caller passes one in.
 */
#include <math.h>

#include "../../include/dsdneo_b200.h"

#define SPS_FIR_CAP 1024 /* FIR_MAX_TAPS, dsd_filters.c:12 */

/* base table read at a fractional index; zero outside it, the last interval closes on the last tap */
static float
base_at(const float* base, int len, float idx) {
    if (idx < 0.0f || idx > (float)(len - 1)) {
        return 0.0f;
    }
    const int lo = (int)idx;
    const int hi = lo + 1 >= len ? len - 1 : lo + 1;
    const float frac = idx - (float)lo;
    const float step = base[hi] - base[lo];
    const float prod = frac * step;
    return base[lo] + prod;
}

/* root-raised-cosine impulse response at t symbols, T = 1, with the reference's guards and evaluation order */
static float
rrc_at(float t, float alpha) {
    const float pi = 3.14159265358979323846f;
    const float tiny = 1e-6f;
    if (alpha <= 0.0f || alpha > 1.0f) { /* degenerate roll-off: plain sinc */
        if (fabsf(t) < tiny) {
            return 1.0f;
        }
        const float x = pi * t;
        return sinf(x) / x;
    }
    if (fabsf(t) < tiny) {
        const float k = (4.0f / pi) - 1.0f;
        const float ak = alpha * k;
        return 1.0f + ak;
    }
    const float at4 = 4.0f * alpha * t;
    if (fabsf(fabsf(at4) - 1.0f) < 1e-4f) { /* t = +-1 / (4 alpha) */
        const float a = pi / (4.0f * alpha);
        const float s_part = (1.0f + (2.0f / pi)) * sinf(a);
        const float c_part = (1.0f - (2.0f / pi)) * cosf(a);
        return (alpha / 1.41421356237309504880f) * (s_part + c_part);
    }
    const float arg_s = pi * t * (1.0f - alpha);
    const float arg_c = pi * t * (1.0f + alpha);
    const float c_term = at4 * cosf(arg_c);
    const float num = sinf(arg_s) + c_term;
    const float sq = at4 * at4;
    const float den = pi * t * (1.0f - sq);
    if (fabsf(den) < tiny) {
        return 0.0f;
    }
    return num / den;
}

int
dsdneo_b200_sps_fir_design(int design_kind, const float* base, int base_len, int base_sps, float rrc_alpha, int sps, float* taps_out,
                           int max_taps) {
    if (!base || base_len <= 0 || base_sps <= 0 || sps <= 1 || !taps_out || max_taps <= 0) {
        return DSDNEO_B200_EINVAL; /* the reference leaves the filter not ready and passes samples through */
    }
    if (design_kind != DSDNEO_SPS_FIR_DESIGN_INTERP && design_kind != DSDNEO_SPS_FIR_DESIGN_RRC) {
        return DSDNEO_B200_EINVAL;
    }

    /* length: the base table as it is at its own sps, else the same span in symbols (+ centre tap), odd, capped */
    int n_taps;
    const int exact = (sps == base_sps && base_len <= SPS_FIR_CAP);
    if (exact) {
        n_taps = base_len;
    } else {
        const double span_sym = (double)(base_len - 1) / (double)base_sps;
        n_taps = (int)(span_sym * (double)sps + 0.5) + 1;
    }
    if (n_taps < 3) {
        n_taps = 3;
    }
    if ((n_taps & 1) == 0) {
        n_taps++;
    }
    if (n_taps > SPS_FIR_CAP) {
        n_taps = SPS_FIR_CAP - 1;
    }
    if (n_taps > max_taps) {
        return DSDNEO_B200_EUNSUPPORTED;
    }

    if (sps == base_sps && n_taps == base_len) {
        for (int n = 0; n < n_taps; n++) {
            taps_out[n] = base[n];
        }
    } else {
        const float c_new = 0.5f * (float)(n_taps - 1);
        const float c_base = 0.5f * (float)(base_len - 1);
        for (int n = 0; n < n_taps; n++) {
            const float t = ((float)n - c_new) / (float)sps;
            if (design_kind == DSDNEO_SPS_FIR_DESIGN_RRC) {
                taps_out[n] = rrc_at(t, rrc_alpha);
            } else {
                const float scaled = t * (float)base_sps;
                taps_out[n] = base_at(base, base_len, scaled + c_base);
            }
        }
    }

    /* unit DC gain */
    double dc = 0.0;
    for (int n = 0; n < n_taps; n++) {
        dc += taps_out[n];
    }
    if (fabs(dc) < 1e-12) {
        dc = 1.0;
    }
    const float scale = (float)(1.0 / dc);
    for (int n = 0; n < n_taps; n++) {
        taps_out[n] *= scale;
    }
    return n_taps;
}
