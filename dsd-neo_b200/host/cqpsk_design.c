/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Host-side design for the CQPSK chain (cqpsk.cu): band-edge FLL filters and second-order loop gains.
 *
 * Twin of the reference's
 *   fll_band_edge_design_filter     src/dsp/costas.cpp:936-1024  (GNU Radio fll_band_edge_cc::design_filter)
 *   fll_configure_loop_params       src/dsp/costas.cpp:647-657
 *   costas_init_if_needed           src/dsp/costas.cpp:536-551
 * The taps depend on sps only; they are computed once per bank on the host with the same libm calls (sinf / cosf /
 * roundf) and the same float staging as the reference, so they come out bit-identical, then uploaded.
 * Compile with -fno-fast-math -ffp-contract=off.
 */
#include <math.h>

#include "../../include/dsdneo_b200.h"

#define FLL_MAX_TAPS 48

int
dsdneo_b200_fll_band_edge_design(int sps, float* lower_r, float* lower_i, float* upper_r, float* upper_i, int max_taps) {
    static const float two_pi = 6.28318530717958647692f;
    static const float pi = 3.14159265358979323846f;
    const float excess_bw = 0.2f;
    if (sps < 1 || !lower_r || !lower_i || !upper_r || !upper_i) {
        return DSDNEO_B200_EINVAL;
    }
    int n = 2 * sps + 1;
    if (n > FLL_MAX_TAPS) {
        n = FLL_MAX_TAPS;
    }
    if (n < 3) {
        n = 3;
    }
    if (n > max_taps) {
        return DSDNEO_B200_EINVAL;
    }
    float base[FLL_MAX_TAPS];
    const float span = roundf((float)n / (float)sps);
    float energy = 0.0f;
    for (int i = 0; i < n; i++) {
        /* half-sine spectrum edge == sum of two sincs in time */
        const float k = -span + (float)i * 2.0f / (float)sps;
        const float u = excess_bw * k - 0.5f;
        const float v = excess_bw * k + 0.5f;
        const float su = (fabsf(u) < 1e-6f) ? 1.0f : sinf(pi * u) / (pi * u);
        const float sv = (fabsf(v) < 1e-6f) ? 1.0f : sinf(pi * v) / (pi * v);
        base[i] = su + sv;
        energy += base[i] * base[i];
    }
    if (energy > 0.0f) {
        const float g = 1.0f / energy; /* divided by the power itself, as GNU Radio does */
        for (int i = 0; i < n; i++) {
            base[i] *= g;
        }
    }
    const int mid = (n - 1) / 2;
    for (int i = 0; i < n; i++) {
        const float fr = (float)(-mid + i) / (2.0f * (float)sps);
        const float ang = two_pi * (1.0f + excess_bw) * fr;
        const int j = n - 1 - i; /* stored time-reversed */
        lower_r[j] = base[i] * cosf(-ang);
        lower_i[j] = base[i] * sinf(-ang);
        upper_r[j] = base[i] * cosf(ang);
        upper_i[j] = base[i] * sinf(ang);
    }
    return n;
}

/* control_loop::update_gains with damping sqrt(2)/2 */
void
dsdneo_b200_loop_gains(float loop_bw, float* alpha, float* beta) {
    const float damping = 0.70710678118654752440f;
    const float denom = 1.0f + 2.0f * damping * loop_bw + loop_bw * loop_bw;
    *alpha = (4.0f * damping * loop_bw) / denom;
    *beta = (4.0f * loop_bw * loop_bw) / denom;
}

/* FLL loop bandwidth 2 pi / sps / 350 (costas.cpp:649) */
void
dsdneo_b200_fll_loop_gains(int sps, float* alpha, float* beta) {
    static const float two_pi = 6.28318530717958647692f;
    dsdneo_b200_loop_gains(two_pi / (float)sps / 350.0f, alpha, beta);
}

/* Gardner delay-line span for a fresh channel (costas.cpp:377-392): max(2 ceil(omega_max), ceil(omega_max / 2) + 9) */
int
dsdneo_b200_gardner_span(int sps) {
    const float omega = (float)sps;
    const float omega_max = omega * (1.0f + 0.002f);
    const int a = 2 * (int)ceilf(omega_max);
    const int b = (int)ceilf(omega_max / 2.0f) + 8 + 1;
    return a > b ? a : b;
}
