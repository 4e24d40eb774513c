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

/* ==================================================================================================================
 * Vocoder frame ECC: (imbe_fr[8][23] | ambe_fr[4][24]) -> (imbe_d[88] | ambe_d[49]) + error counts.
 *
 * PARITY UNPINNED (same reason as above).  Call sites in the reference: mbe_decodeImbe7200x4400Frame /
 * mbe_decodeAmbe3600x2450Frame, src/core/vocoder/dsd_mbe.c:168,188 (contract: CMakeLists.txt:622-655).  Restated from the
 * published algorithm of mbelib 1.3.0 (ecc.c: mbe_checkGolayBlock, mbe_golay2312, mbe_hamming1511; imbe7200x4400.c:
 * mbe_eccImbe7200x4400C0 / mbe_demodulateImbe7200x4400Data / mbe_eccImbe7200x4400Data; ambe3600x2450.c:
 * mbe_eccAmbe3600x2450C0 / mbe_demodulateAmbe3600x2450Data / mbe_eccAmbe3600x2450Data) and of TIA-102.BABA section 7
 * (bit prioritisation, [23,12] Golay and [15,11] Hamming codes, pseudo-random modulation p_r(n) = (173 p_r(n-1) + 13849)
 * mod 65536 seeded with 16 u0).  The code tables are not recalled but DERIVED here: the Golay parity rows are x^(11+i) mod
 * g(x), g = 0xC75, the Golay correction table is the coset-leader table of that perfect code (every syndrome has exactly one
 * error pattern of weight <= 3), the Hamming correction table is the column table of its four check masks; the tests check
 * minimum distance 7 / single-error correction exhaustively, so any implementation of these codes gives the same result.
 * ================================================================================================================== */

static int g_ecc_ready = 0;
static uint16_t g_golay_gen[12];      /* parity of data bit (22 - i) */
static uint16_t g_golay_fix[2048];    /* syndrome -> 12-bit data error pattern */
static const uint16_t kHammingCheck[4] = {0x7f08, 0x78e4, 0x66d2, 0x55b1};
static uint16_t g_hamming_fix[16];    /* syndrome -> 15-bit error pattern */

static int
parity16(unsigned v) {
    v ^= v >> 8;
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return (int)(v & 1u);
}

static unsigned
golay_syndrome(unsigned block23) {
    unsigned ecc = 0;
    for (int i = 0; i < 12; i++) {
        if (block23 & (0x400000u >> i)) {
            ecc ^= g_golay_gen[i];
        }
    }
    return ecc ^ (block23 & 0x7ffu);
}

static unsigned
hamming_syndrome(unsigned block15) {
    unsigned s = 0;
    for (int i = 0; i < 4; i++) {
        s = (s << 1) | (unsigned)parity16(block15 & kHammingCheck[i]);
    }
    return s;
}

static void
ecc_tables(void) {
    if (g_ecc_ready) {
        return;
    }
    unsigned r = 0x475; /* x^11 mod g */
    for (int i = 11; i >= 0; i--) {
        g_golay_gen[i] = (uint16_t)r;
        r <<= 1;
        if (r & 0x800u) {
            r ^= 0xC75u;
        }
    }
    memset(g_golay_fix, 0, sizeof(g_golay_fix));
    g_golay_fix[golay_syndrome(0)] = 0;
    for (int a = 0; a < 23; a++) {
        unsigned ea = 1u << a;
        g_golay_fix[golay_syndrome(ea)] = (uint16_t)(ea >> 11);
        for (int b = a + 1; b < 23; b++) {
            unsigned eb = ea | (1u << b);
            g_golay_fix[golay_syndrome(eb)] = (uint16_t)(eb >> 11);
            for (int c = b + 1; c < 23; c++) {
                unsigned ec = eb | (1u << c);
                g_golay_fix[golay_syndrome(ec)] = (uint16_t)(ec >> 11);
            }
        }
    }
    memset(g_hamming_fix, 0, sizeof(g_hamming_fix));
    for (int b = 0; b < 15; b++) {
        g_hamming_fix[hamming_syndrome(1u << b)] = (uint16_t)(1u << b);
    }
    g_ecc_ready = 1;
}

/* mbe_golay2312: in/out[23], index 22 = first transmitted bit; data = bits 22..11; returns the number of data bits changed */
int
oracle_mbe_golay2312(const uint8_t* in, uint8_t* out) {
    ecc_tables();
    unsigned block = 0;
    for (int i = 22; i >= 0; i--) {
        block = (block << 1) | (in[i] & 1u);
    }
    unsigned data = (block >> 11) ^ g_golay_fix[golay_syndrome(block)];
    int errs = 0;
    for (int i = 22; i >= 11; i--) {
        out[i] = (uint8_t)((data >> (i - 11)) & 1u);
        errs += out[i] != (in[i] & 1u);
    }
    for (int i = 10; i >= 0; i--) {
        out[i] = in[i] & 1u;
    }
    return errs;
}

/* encoder twin for the tests: 12 data bits (bit 11 = index 22) -> 23-bit codeword */
unsigned
oracle_mbe_golay2312_encode(unsigned data12) {
    ecc_tables();
    unsigned ecc = 0;
    for (int i = 0; i < 12; i++) {
        if (data12 & (0x800u >> i)) {
            ecc ^= g_golay_gen[i];
        }
    }
    return (data12 << 11) | ecc;
}

/* mbe_hamming1511: in/out[15], data = bits 14..4; returns 1 when the syndrome was non-zero */
int
oracle_mbe_hamming1511(const uint8_t* in, uint8_t* out) {
    ecc_tables();
    unsigned block = 0;
    for (int i = 14; i >= 0; i--) {
        block = (block << 1) | (in[i] & 1u);
    }
    unsigned s = hamming_syndrome(block);
    int errs = 0;
    if (s) {
        errs = 1;
        block ^= g_hamming_fix[s];
    }
    for (int i = 14; i >= 0; i--) {
        out[i] = (uint8_t)((block >> i) & 1u);
    }
    return errs;
}

unsigned
oracle_mbe_hamming1511_encode(unsigned data11) {
    unsigned block = data11 << 4;
    for (int i = 0; i < 4; i++) {
        if (parity16(block & kHammingCheck[i] & 0x7ff0u)) {
            block |= 8u >> i;
        }
    }
    return block;
}

static void
mbe_pr_bits(unsigned seed12, int n, uint8_t* bits) { /* bits[1..n-1] */
    unsigned pr = 16u * seed12;
    bits[0] = 0;
    for (int i = 1; i < n; i++) {
        pr = (173u * pr + 13849u) & 0xffffu;
        bits[i] = (uint8_t)(pr >> 15);
    }
}

/* ambe_fr[4][24] (bytes 0/1) -> ambe_d[49]; *c0_errs = errs, *total_errs = errs2 of mbe_processAmbe3600x2450Framef */
void
oracle_ambe3600x2450_decode(const uint8_t* ambe_fr_in, uint8_t* ambe_d, int* c0_errs, int* total_errs) {
    uint8_t fr[4][24], gout[23], pr[24];
    memcpy(fr, ambe_fr_in, sizeof(fr));
    int errs = oracle_mbe_golay2312(&fr[0][1], gout); /* C0 is the [24,12] word; column 0 (overall parity) is not used */
    memcpy(&fr[0][1], gout, 23);
    unsigned seed = 0;
    for (int i = 23; i >= 12; i--) {
        seed = (seed << 1) | fr[0][i];
    }
    mbe_pr_bits(seed, 24, pr);
    int k = 1;
    for (int j = 22; j >= 0; j--) {
        fr[1][j] ^= pr[k++];
    }
    uint8_t* o = ambe_d;
    for (int j = 23; j > 11; j--) {
        *o++ = fr[0][j];
    }
    int errs2 = errs + oracle_mbe_golay2312(&fr[1][0], gout);
    for (int j = 22; j > 10; j--) {
        *o++ = gout[j];
    }
    for (int j = 10; j >= 0; j--) {
        *o++ = fr[2][j];
    }
    for (int j = 13; j >= 0; j--) {
        *o++ = fr[3][j];
    }
    *c0_errs = errs;
    *total_errs = errs2;
}

/* imbe_fr[8][23] -> imbe_d[88]; errs (C0) and errs2 (all words) of mbe_processImbe7200x4400Framef */
void
oracle_imbe7200x4400_decode(const uint8_t* imbe_fr_in, uint8_t* imbe_d, int* c0_errs, int* total_errs) {
    uint8_t fr[8][23], gout[23], hout[15], pr[115];
    memcpy(fr, imbe_fr_in, sizeof(fr));
    int errs = oracle_mbe_golay2312(&fr[0][0], gout);
    memcpy(&fr[0][0], gout, 23);
    unsigned seed = 0;
    for (int i = 22; i >= 11; i--) {
        seed = (seed << 1) | fr[0][i];
    }
    mbe_pr_bits(seed, 115, pr);
    int k = 1;
    for (int i = 1; i < 4; i++) {
        for (int j = 22; j >= 0; j--) {
            fr[i][j] ^= pr[k++];
        }
    }
    for (int i = 4; i < 7; i++) {
        for (int j = 14; j >= 0; j--) {
            fr[i][j] ^= pr[k++];
        }
    }
    uint8_t* o = imbe_d;
    int errs2 = errs;
    for (int j = 22; j > 10; j--) {
        *o++ = fr[0][j];
    }
    for (int i = 1; i < 4; i++) {
        errs2 += oracle_mbe_golay2312(&fr[i][0], gout);
        for (int j = 22; j > 10; j--) {
            *o++ = gout[j];
        }
    }
    for (int i = 4; i < 7; i++) {
        errs2 += oracle_mbe_hamming1511(&fr[i][0], hout);
        for (int j = 14; j >= 4; j--) {
            *o++ = hout[j];
        }
    }
    for (int j = 6; j >= 0; j--) {
        *o++ = fr[7][j];
    }
    *c0_errs = errs;
    *total_errs = errs2;
}

/* ---- DMR BS voice burst cutter: src/protocol/dmr/dmr_bs.c:137-148 (unpack through dsd_ambe_2450_dibit_map,
 * include/dsd-neo/core/ambe_interleave.h:25-32), :150-170 (sync segment -> 48 bits), :182-187 (CACH), :711-722 (the 90
 * buffered dibits of the first burst are inverted when opts->inverted_dmr), :745-746 / :838-848 (segment offsets).
 * burst144: the 144 dibits of one burst.  Outputs: cach24, ambe_fr[3][4][24], sync48. */
static const uint8_t kAmbeMap[36][4] = {
    {0, 23, 0, 5},  {1, 10, 2, 3}, {0, 22, 0, 4},  {1, 9, 2, 2},  {0, 21, 0, 3},  {1, 8, 2, 1},  {0, 20, 0, 2},  {1, 7, 2, 0},
    {0, 19, 0, 1},  {1, 6, 3, 13}, {0, 18, 0, 0},  {1, 5, 3, 12}, {0, 17, 1, 22}, {1, 4, 3, 11}, {0, 16, 1, 21}, {1, 3, 3, 10},
    {0, 15, 1, 20}, {1, 2, 3, 9},  {0, 14, 1, 19}, {1, 1, 3, 8},  {0, 13, 1, 18}, {1, 0, 3, 7},  {0, 12, 1, 17}, {2, 10, 3, 6},
    {0, 11, 1, 16}, {2, 9, 3, 5},  {0, 10, 1, 15}, {2, 8, 3, 4},  {0, 9, 1, 14},  {2, 7, 3, 3},  {0, 8, 1, 13},  {2, 6, 3, 2},
    {0, 7, 1, 12},  {2, 5, 3, 1},  {0, 6, 1, 11},  {2, 4, 3, 0},
};

void
oracle_dmr_voice_cut(const uint8_t* burst144, int invert_first90, uint8_t* cach24, uint8_t* ambe_fr3, uint8_t* sync48) {
    static const uint8_t cach_il[24] = {0, 7, 8, 9, 1, 10, 11, 12, 2, 13, 14, 15, 3, 16, 4, 17, 18, 19, 5, 20, 21, 22, 6, 23};
    uint8_t d[144];
    for (int i = 0; i < 144; i++) {
        d[i] = burst144[i] & 3;
        if (invert_first90 && i < 90) {
            d[i] = (uint8_t)((d[i] ^ 2) & 3);
        }
    }
    memset(cach24, 0, 24);
    memset(ambe_fr3, 0, 3 * 96);
    for (int i = 0; i < 12; i++) {
        cach24[cach_il[2 * i]] = (d[i] >> 1) & 1;
        cach24[cach_il[2 * i + 1]] = d[i] & 1;
    }
    uint8_t(*f)[4][24] = (uint8_t(*)[4][24])ambe_fr3;
    const int seg[4][4] = {{0, 12, 36, 0}, {1, 48, 18, 0}, {1, 90, 18, 18}, {2, 108, 36, 0}}; /* frame, offset, count, map offset */
    for (int s = 0; s < 4; s++) {
        for (int i = 0; i < seg[s][2]; i++) {
            const uint8_t* m = kAmbeMap[seg[s][3] + i];
            int dib = d[seg[s][1] + i];
            f[seg[s][0]][m[0]][m[1]] = (dib >> 1) & 1;
            f[seg[s][0]][m[2]][m[3]] = dib & 1;
        }
    }
    for (int i = 0; i < 24; i++) {
        sync48[2 * i] = (d[66 + i] >> 1) & 1;
        sync48[2 * i + 1] = d[66 + i] & 1;
    }
}
