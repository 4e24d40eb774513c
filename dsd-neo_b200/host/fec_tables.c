/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Host-side construction of the FEC decode tables uploaded to the device by csrc/fec.cu.
 *
 * Twin of InitAllFecFunction() (src/fec/fec.c:827-836): the reference fills its syndrome->correction tables at
 * start-up by enumerating error patterns; the SAME enumeration order is used here because, for the shortened
 * Golay(20,8) code, distinct weight-3 patterns share syndromes and the last writer wins
 * (src/fec/fec.c:452-528, 591-667, 737-783).  The parity-check matrices themselves (src/fec/fec.c:25-126) are
 * regenerated from the codes' generator polynomials:
 *   Hamming(7,4): x^3+x+1;  Hamming(15,11)/(13,9)/(12,8)/(16,11,4): x^4+x+1 (shortened / extended);
 *   Golay(24,12) and its shortening (20,8): x^11+x^10+x^6+x^5+x^4+x^2+1 plus overall parity;
 *   QR(16,7,6): x^8+x^5+x^4+x^3+1 plus overall parity.
 */
#include <string.h>

#include "fec_tables.h"

static unsigned
poly_xpow(int e, unsigned g, int deg) {
    unsigned v = 1;
    while (e-- > 0) {
        v <<= 1;
        if (v >> deg) {
            v ^= g;
        }
    }
    return v;
}

static int
popcnt(unsigned v) {
    int n = 0;
    for (; v; v &= v - 1) {
        n++;
    }
    return n;
}

static void
build_hamming(dsdneo_hamming_table* t, int n, int r, unsigned g, int deg, int extended, int stop_on_fail) {
    memset(t, 0, sizeof(*t));
    t->n = n;
    t->r = r;
    t->k = n - r;
    t->stop_on_fail = stop_on_fail;
    memset(t->pos_of, 0xFF, sizeof(t->pos_of));
    const int ncyc = extended ? n - 1 : n;
    unsigned col[16];
    for (int j = 0; j < ncyc; j++) {
        unsigned s = poly_xpow(ncyc - 1 - j, g, deg);
        if (extended) {
            s = (s << 1) | (unsigned)((popcnt(s) & 1) ^ 1);
        }
        col[j] = s;
    }
    if (extended) {
        col[n - 1] = 1;
    }
    for (int j = 0; j < n; j++) {
        t->pos_of[col[j]] = (unsigned char)j;
        for (int row = 0; row < r; row++) { /* row 0 = most significant syndrome bit */
            if ((col[j] >> (r - 1 - row)) & 1u) {
                t->row_mask[row] |= 1u << j; /* bit j of the packed word = codeword position j */
            }
        }
    }
}

static void
build_gq(dsdneo_gq_table* t, int k, int r, int maxw, unsigned g, int deg, int full_k) {
    memset(t, 0, sizeof(*t));
    t->k = k;
    t->r = r;
    t->n = k + r;
    t->maxw = maxw;
    memset(t->corr, 0xFF, sizeof(t->corr));
    unsigned dsyn[12]; /* syndrome of a single error in data bit j */
    for (int j = 0; j < k; j++) {
        const int i = j + (full_k - k);
        const unsigned rem = poly_xpow(deg + full_k - 1 - i, g, deg);
        dsyn[j] = (rem << 1) | (unsigned)((popcnt(rem) + 1) & 1);
    }
    for (int row = 0; row < r; row++) {
        unsigned m = 1u << (k + row); /* identity part */
        for (int j = 0; j < k; j++) {
            if ((dsyn[j] >> (r - 1 - row)) & 1u) {
                m |= 1u << j;
            }
        }
        t->row_mask[row] = m;
    }
#define PAR(ip) (1u << (r - 1 - (ip)))
#define SET1(s, a) (t->corr[(s)][0] = (unsigned char)(a))
#define SET2(s, a, b) (SET1(s, a), t->corr[(s)][1] = (unsigned char)(b))
#define SET3(s, a, b, c) (SET2(s, a, b), t->corr[(s)][2] = (unsigned char)(c))
    for (int a = 0; a < k; a++) {
        for (int b = a + 1; b < k; b++) {
            const unsigned sab = dsyn[a] ^ dsyn[b];
            if (maxw >= 3) {
                for (int c = b + 1; c < k; c++) {
                    SET3(sab ^ dsyn[c], a, b, c);
                }
            }
            SET2(sab, a, b);
            if (maxw >= 3) {
                for (int p = 0; p < r; p++) {
                    SET3(sab ^ PAR(p), a, b, k + p);
                }
            }
        }
        SET1(dsyn[a], a);
        for (int p = 0; p < r; p++) {
            SET2(dsyn[a] ^ PAR(p), a, k + p);
            if (maxw >= 3) {
                for (int q = p + 1; q < r; q++) {
                    SET3(dsyn[a] ^ PAR(p) ^ PAR(q), a, k + p, k + q);
                }
            }
        }
    }
    for (int p = 0; p < r; p++) {
        SET1(PAR(p), k + p);
        for (int q = p + 1; q < r; q++) {
            SET2(PAR(p) ^ PAR(q), k + p, k + q);
            if (maxw >= 3) {
                for (int w = q + 1; w < r; w++) {
                    SET3(PAR(p) ^ PAR(q) ^ PAR(w), k + p, k + q, k + w);
                }
            }
        }
    }
#undef PAR
#undef SET1
#undef SET2
#undef SET3
}

void
dsdneo_fec_build_tables(dsdneo_fec_tables* t) {
    build_hamming(&t->ham[DSDNEO_FEC_HAMMING_7_4], 7, 3, 0xB, 3, 0, 1);
    build_hamming(&t->ham[DSDNEO_FEC_HAMMING_12_8], 12, 4, 0x13, 4, 0, 0);
    build_hamming(&t->ham[DSDNEO_FEC_HAMMING_13_9], 13, 4, 0x13, 4, 0, 1);
    build_hamming(&t->ham[DSDNEO_FEC_HAMMING_15_11], 15, 4, 0x13, 4, 0, 1);
    build_hamming(&t->ham[DSDNEO_FEC_HAMMING_16_11_4], 16, 5, 0x13, 4, 1, 1);
    build_gq(&t->gq[0], 8, 12, 3, 0xC75, 11, 12);  /* Golay(20,8)  */
    build_gq(&t->gq[1], 12, 12, 3, 0xC75, 11, 12); /* Golay(24,12) */
    build_gq(&t->gq[2], 7, 9, 2, 0x139, 8, 7);     /* QR(16,7,6)   */
    /* GF(64), primitive polynomial x^6+x+1 (include/dsd-neo/fec/ReedSolomon.hpp:685-727) */
    int v = 1;
    for (int i = 0; i < 63; i++) {
        t->gf_exp[i] = (signed char)v;
        t->gf_log[v] = (signed char)i;
        v <<= 1;
        if (v & 0x40) {
            v ^= 0x43;
        }
    }
    t->gf_exp[63] = 0;
    t->gf_log[0] = -1;
}
