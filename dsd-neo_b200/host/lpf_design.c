/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Host-side channel low-pass design for the demod bank (K4 taps).
 *
 * Twin of the reference's per-rate channel filter plan:
 *   channel_lpf_design_low_pass / channel_lpf_cutoff_for_profile   src/dsp/demod_pipeline.cpp:443-489
 *   dsd_firdes_low_pass + Blackman window + ntaps rule              src/dsp/firdes.cpp (GNU Radio firdes port)
 * The taps are a pure function of (rate, profile); they are designed once on the host with the same
 * libm calls and the same float/double staging as the reference so that they come out bit-identical,
 * then uploaded to the device.  Compile with -fno-fast-math -ffp-contract=off.
 */
#include <math.h>

#include "../../include/dsdneo_b200.h"

#define PI_D 3.14159265358979323846

/* Profile -> cutoff: protected channel edge + half of the 1200 Hz transition (demod_pipeline.cpp:134-149). */
static double
profile_cutoff_hz(int profile) {
    const double guard = 1200.0 * 0.5;
    switch (profile) {
        case DSDNEO_CH_LPF_PROFILE_6K25: return 3125.0 + guard;
        case DSDNEO_CH_LPF_PROFILE_12K5:
        case DSDNEO_CH_LPF_PROFILE_PROVOICE:
        case DSDNEO_CH_LPF_PROFILE_P25_C4FM: return 6250.0 + guard;
        case DSDNEO_CH_LPF_PROFILE_P25_CQPSK: return 7250.0;
        case DSDNEO_CH_LPF_PROFILE_WIDE:
        default: return 8000.0 + guard;
    }
}

int
dsdneo_b200_channel_lpf_design(int rate_out_hz, int profile, float* taps_out, int max_taps) {
    if (rate_out_hz <= 0 || !taps_out || max_taps <= 0) {
        return DSDNEO_B200_EINVAL;
    }
    const double fs = (double)rate_out_hz;
    const double transition = 1200.0;

    double fc = profile_cutoff_hz(profile);
    const double fc_max = (fs * 0.5) * 0.90;
    if (fc < 100.0) {
        fc = 100.0;
    }
    if (fc > fc_max) {
        fc = fc_max;
    }
    if (fc <= 0.0 || fc > fs / 2.0) {
        return DSDNEO_B200_EINVAL;
    }

    /* Blackman: 74 dB; ntaps = (int)(A*fs/(22*tw)), forced odd. */
    int ntaps = (int)(74.0 * fs / (22.0 * transition));
    if ((ntaps & 1) == 0) {
        ntaps++;
    }
    if (ntaps > max_taps || ntaps > 1024 || ntaps < 3) {
        return DSDNEO_B200_EUNSUPPORTED;
    }

    float win[1024];
    {
        const float span = (float)(ntaps - 1);
        for (int i = 0; i < ntaps; i++) {
            win[i] = 0.42f - 0.5f * cosf((2.0f * (float)PI_D * (float)i) / span)
                     + 0.08f * cosf((4.0f * (float)PI_D * (float)i) / span);
        }
    }

    const int half = (ntaps - 1) / 2;
    const double w0 = 2.0 * PI_D * fc / fs;
    for (int n = -half; n <= half; n++) {
        if (n == 0) {
            taps_out[half] = (float)((w0 / PI_D) * win[half]);
        } else {
            taps_out[n + half] = (float)((sin(n * w0) / (n * PI_D)) * win[n + half]);
        }
    }

    /* Unit gain at DC, exploiting symmetry exactly as the reference sums it. */
    double dc = taps_out[half];
    for (int n = 1; n <= half; n++) {
        dc += 2.0 * taps_out[n + half];
    }
    double g = 1.0;
    g /= dc;
    for (int i = 0; i < ntaps; i++) {
        taps_out[i] *= (float)g;
    }
    return ntaps;
}
