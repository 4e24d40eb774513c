// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the MBE synthesis stage.
 *
 * PARITY UNPINNED.  dsd-neo does not contain this code: it calls the un-vendored dependency mbelib-neo 2.x
 * (arancormonk/mbelib-neo @ 6138cce7091d90e4be9e889ac166006265d3e8fb, vcpkg-ports/mbe-neo/portfile.cmake:3-7) through
 * mbe_processImbe4400Dataf / mbe_processAmbe2450Dataf (call sites src/core/vocoder/dsd_mbe.c:268,296,581,617,685), and
 * that library is absent from /root/reference and from this machine.  What follows restates the published algorithm of
 * the library it descends from (szechyjs/mbelib 1.3.0, mbelib.c: mbe_spectralAmpEnhance, mbe_synthesizeSpeechf,
 * mbe_floattoshort; TIA-102.BABA eq. 105-111 and 127-141), from its public description:
 *   - spectral amplitude enhancement of the current frame's Ml[],
 *   - per band l: voiced bands as windowed cosines of the previous and current frame (eq. 131-133), unvoiced bands as a
 *     `uvquality`-tone multisine with random phases plus band-limited noise above 2700 Hz, overlap-added with the
 *     211-tap trapezoid synthesis window ws(n),
 *   - phase tracking PSIl/PHIl (eq. 139-140), then float -> int16 with gain 7 and clipping at +-32760.
 * The bit-level parameter decode (mbe_decodeImbe4400Parms / mbe_decodeAmbe2450Parms) needs the codec's quantiser tables
 * and is NOT restated: the stage boundary is the decoded parameter set (w0, L, Vl, Ml) -> 160 PCM samples.
 * One deliberate difference: mbelib draws its random phases / noise from libc rand(), whose stream is shared process
 * state; here every draw is a counter-based hash of (frame key, band, sample, index), so batches are reproducible and
 * order independent.  No golden vector for this stage exists in the reference tree.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"

/* ws(n), n = -160..160 stored at index n + 160: 0 outside +-105, 1 inside +-55, linear ramps of 0.02 in between */
static float
mbe_ws(int idx) {
    int n = idx - 160;
    if (n < 0) {
        n = -n;
    }
    if (n > 105) {
        return 0.0f;
    }
    if (n <= 55) {
        return 1.0f;
    }
    return (float)(105 - n) * 0.02f;
}

/* counter-based uniform [0,1): splitmix64 finaliser over (key, band, sample, index, stream) */
float
oracle_mbe_uniform(uint64_t key, int band, int sample, int index, int stream) {
    uint64_t z = key + 0x9E3779B97F4A7C15ull * (uint64_t)(1 + band) + 0xBF58476D1CE4E5B9ull * (uint64_t)(1 + sample)
                 + 0x94D049BB133111EBull * (uint64_t)(1 + index) + 0xD6E8FEB86659FD93ull * (uint64_t)(1 + stream);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

static float
rand_phase(uint64_t key, int band, int index, int stream) { /* mbe_rand_phase: uniform in [-pi, pi) */
    return oracle_mbe_uniform(key, band, -1, index, stream) * 6.2831853071795864769f - 3.14159265358979323846f;
}

/* mbe_spectralAmpEnhance (mbelib.c; TIA-102.BABA eq. 105-111) */
void
oracle_mbe_spectral_amp_enhance(oracle_mbe_parms* cur) {
    float Rm0 = 0.0f, Rm1 = 0.0f, Wl[57];
    for (int l = 1; l <= cur->L; l++) {
        Rm0 = Rm0 + (cur->Ml[l] * cur->Ml[l]);
        Rm1 = Rm1 + ((cur->Ml[l] * cur->Ml[l]) * cosf(cur->w0 * (float)l));
    }
    const float R2m0 = Rm0 * Rm0, R2m1 = Rm1 * Rm1;
    for (int l = 1; l <= cur->L; l++) {
        if (cur->Ml[l] != 0.0f) {
            Wl[l] = sqrtf(cur->Ml[l])
                    * powf(((0.96f * 3.14159265358979323846f * ((R2m0 + R2m1) - (2.0f * Rm0 * Rm1 * cosf(cur->w0 * (float)l))))
                            / (cur->w0 * Rm0 * (R2m0 - R2m1))),
                           0.25f);
            if ((8 * l) <= cur->L) {
                /* low bands are left alone */
            } else if (Wl[l] > 1.2f) {
                cur->Ml[l] = 1.2f * cur->Ml[l];
            } else if (Wl[l] < 0.5f) {
                cur->Ml[l] = 0.5f * cur->Ml[l];
            } else {
                cur->Ml[l] = Wl[l] * cur->Ml[l];
            }
        }
    }
    float sum = 0.0f;
    for (int l = 1; l <= cur->L; l++) {
        float M = cur->Ml[l];
        if (M < 0.0f) {
            M = -M;
        }
        sum += M * M;
    }
    const float gamma = (sum == 0.0f) ? 1.0f : sqrtf(Rm0 / sum);
    for (int l = 1; l <= cur->L; l++) {
        cur->Ml[l] = gamma * cur->Ml[l];
    }
}

/* mbe_synthesizeSpeechf (mbelib.c; eq. 127-141).  Updates cur->PSIl/PHIl and pads Ml/Vl above min(L) like the original. */
void
oracle_mbe_synthesize_speechf(float* aout, oracle_mbe_parms* cur, oracle_mbe_parms* prev, int uvquality, uint64_t key) {
    const int N = 160;
    const float uvthreshold = (2700.0f * 3.14159265358979323846f) / 4000.0f;
    const float uvsine = 1.3591409f * 2.7182818284590452354f, uvrand = 2.0f;
    if (uvquality < 1 || uvquality > 64) {
        uvquality = 3;
    }
    const float qfactor = (uvquality == 1) ? (1.0f / 2.7182818284590452354f) : (logf((float)uvquality) / (float)uvquality);
    const float uvstep = 1.0f / (float)uvquality;
    const float uvoffset = (uvstep * (float)(uvquality - 1)) / 2.0f;
    int numUv = 0;
    for (int l = 1; l <= cur->L; l++) {
        numUv += cur->Vl[l] == 0;
    }
    const float cw0 = cur->w0, pw0 = prev->w0;
    for (int n = 0; n < N; n++) {
        aout[n] = 0.0f;
    }
    int maxl;
    if (cur->L > prev->L) { /* eq. 128, 129 */
        maxl = cur->L;
        for (int l = prev->L + 1; l <= maxl; l++) {
            prev->Ml[l] = 0.0f;
            prev->Vl[l] = 1;
        }
    } else {
        maxl = prev->L;
        for (int l = cur->L + 1; l <= maxl; l++) {
            cur->Ml[l] = 0.0f;
            cur->Vl[l] = 1;
        }
    }
    for (int l = 1; l <= 56; l++) { /* eq. 139, 140 */
        cur->PSIl[l] = prev->PSIl[l] + ((pw0 + cw0) * ((float)(l * N) / 2.0f));
        if (l <= (int)(cur->L / 4)) {
            cur->PHIl[l] = cur->PSIl[l];
        } else {
            cur->PHIl[l] = cur->PSIl[l] + (((float)numUv * rand_phase(key, l, 0, 0)) / (float)cur->L);
        }
    }
    for (int l = 1; l <= maxl; l++) {
        const float cw0l = cw0 * (float)l, pw0l = pw0 * (float)l;
        const int cv = cur->Vl[l], pv = prev->Vl[l];
        if (cv == 0 && pv == 1) {
            for (int n = 0; n < N; n++) {
                const float C1 = mbe_ws(n + N) * prev->Ml[l] * cosf((pw0l * (float)n) + prev->PHIl[l]); /* eq. 131 */
                float C3 = 0.0f;
                for (int i = 0; i < uvquality; i++) {
                    C3 = C3 + cosf((cw0 * (float)n * ((float)l + ((float)i * uvstep) - uvoffset)) + rand_phase(key, l, i, 1));
                    if (cw0l > uvthreshold) {
                        C3 = C3 + ((cw0l - uvthreshold) * uvrand * oracle_mbe_uniform(key, l, n, i, 3));
                    }
                }
                C3 = C3 * uvsine * mbe_ws(n) * cur->Ml[l] * qfactor;
                aout[n] = aout[n] + C1 + C3;
            }
        } else if (cv == 1 && pv == 0) {
            for (int n = 0; n < N; n++) {
                const float C1 = mbe_ws(n) * cur->Ml[l] * cosf((cw0l * (float)(n - N)) + cur->PHIl[l]); /* eq. 132 */
                float C3 = 0.0f;
                for (int i = 0; i < uvquality; i++) {
                    C3 = C3 + cosf((pw0 * (float)n * ((float)l + ((float)i * uvstep) - uvoffset)) + rand_phase(key, l, i, 1));
                    if (pw0l > uvthreshold) {
                        C3 = C3 + ((pw0l - uvthreshold) * uvrand * oracle_mbe_uniform(key, l, n, i, 3));
                    }
                }
                C3 = C3 * uvsine * mbe_ws(n + N) * prev->Ml[l] * qfactor;
                aout[n] = aout[n] + C1 + C3;
            }
        } else if (cv == 1 || pv == 1) {
            for (int n = 0; n < N; n++) { /* eq. 133 */
                const float C1 = mbe_ws(n + N) * prev->Ml[l] * cosf((pw0l * (float)n) + prev->PHIl[l]);
                const float C2 = mbe_ws(n) * cur->Ml[l] * cosf((cw0l * (float)(n - N)) + cur->PHIl[l]);
                aout[n] = aout[n] + C1 + C2;
            }
        } else {
            for (int n = 0; n < N; n++) {
                float C3 = 0.0f, C4 = 0.0f;
                for (int i = 0; i < uvquality; i++) {
                    C3 = C3 + cosf((pw0 * (float)n * ((float)l + ((float)i * uvstep) - uvoffset)) + rand_phase(key, l, i, 1));
                    if (pw0l > uvthreshold) {
                        C3 = C3 + ((pw0l - uvthreshold) * uvrand * oracle_mbe_uniform(key, l, n, i, 3));
                    }
                }
                C3 = C3 * uvsine * mbe_ws(n + N) * prev->Ml[l] * qfactor;
                for (int i = 0; i < uvquality; i++) {
                    C4 = C4 + cosf((cw0 * (float)n * ((float)l + ((float)i * uvstep) - uvoffset)) + rand_phase(key, l, i, 2));
                    if (cw0l > uvthreshold) {
                        C4 = C4 + ((cw0l - uvthreshold) * uvrand * oracle_mbe_uniform(key, l, n, i, 4));
                    }
                }
                C4 = C4 * uvsine * mbe_ws(n) * cur->Ml[l] * qfactor;
                aout[n] = aout[n] + C3 + C4;
            }
        }
    }
}

/* mbe_floattoshort: gain 7, clip to +-32760, truncate */
void
oracle_mbe_floattoshort(const float* in, int16_t* out) {
    for (int i = 0; i < 160; i++) {
        float v = in[i] * 7.0f;
        if (v > 32760.0f) {
            v = 32760.0f;
        } else if (v < -32760.0f) {
            v = -32760.0f;
        }
        out[i] = (int16_t)v;
    }
}

/* The tail of mbe_process*Dataf for a good frame: enhance, synthesise against the previous enhanced frame, then
 * mbe_moveMbeParms(cur, prev_enhanced). */
void
oracle_mbe_synth_frame(float* aout, int16_t* pcm, oracle_mbe_parms* cur, oracle_mbe_parms* prev_enhanced, int uvquality, uint64_t key) {
    oracle_mbe_spectral_amp_enhance(cur);
    oracle_mbe_synthesize_speechf(aout, cur, prev_enhanced, uvquality, key);
    if (pcm) {
        oracle_mbe_floattoshort(aout, pcm);
    }
    memcpy(prev_enhanced, cur, sizeof(*cur));
}
