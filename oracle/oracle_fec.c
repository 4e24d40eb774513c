/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the FEC leaves of the hot path (see oracle.h).
 * Each decoder cites the reference code it restates (arancormonk/dsd-neo @ 4d06905) and is pinned against
 * the reference's known-answer vectors and the compiled reference by tests/test_oracle_fec.py.
 *
 * Code definitions are generated from their generator polynomials instead of being stored as matrices:
 *   Hamming(15,11) and its shortenings (13,9) (12,8): column j of H = x^(n-1-j) mod (x^4+x+1)
 *   Hamming(7,4):   x^(6-j) mod (x^3+x+1);   Hamming(16,11,4): the (15,11) columns with an odd-weight
 *   completion bit and a 16th column 00001     (src/fec/fec.c:25-69 are exactly these matrices)
 *   Golay(24,12): parity row i = [x^(22-i) mod g23 , overall parity], g23 = x^11+x^10+x^6+x^5+x^4+x^2+1;
 *   Golay(20,8) = the same shortened by 4 data bits; QR(16,7,6): g = x^8+x^5+x^4+x^3+1, same construction
 *   (src/fec/fec.c:73-126).
 */
#include "oracle.h"

#include <string.h>

/* ------------------------------------------------------------------ Hamming family */

typedef struct {
    int n, r;              /* codeword bits, syndrome bits */
    uint8_t col[16];       /* syndrome of a single error at position j */
    uint8_t pos_of[32];    /* syndrome -> position, 0xFF = not a single-bit syndrome */
    int stop_on_fail;      /* reference `break`s before copying info bits (all but 12_8) */
} ham_code;

static ham_code g_ham[5];
static int g_ready = 0;

static unsigned
xpow_mod(int e, unsigned g, int deg) {
    unsigned r = 1;
    for (int i = 0; i < e; i++) {
        r <<= 1;
        if (r & (1u << deg)) {
            r ^= g;
        }
    }
    return r;
}

static void
ham_build(ham_code* c, int n, int r, unsigned g, int deg, int extended) {
    memset(c, 0, sizeof(*c));
    c->n = n;
    c->r = r;
    memset(c->pos_of, 0xFF, sizeof(c->pos_of));
    int ncyc = extended ? n - 1 : n;
    for (int j = 0; j < ncyc; j++) {
        unsigned s = xpow_mod(ncyc - 1 - j, g, deg);
        if (extended) {
            s = (s << 1) | ((__builtin_popcount(s) & 1) ? 0u : 1u);
        }
        c->col[j] = (uint8_t)s;
    }
    if (extended) {
        c->col[n - 1] = 1;
    }
    for (int j = 0; j < n; j++) {
        c->pos_of[c->col[j]] = (uint8_t)j;
    }
}

/* ------------------------------------------------------------------ Golay / QR family */

typedef struct {
    int k, r, maxw;           /* data bits, parity bits, patterns enumerated up to this weight */
    uint16_t pt[12];          /* row s of P^T: bit (k-1-j) set when data bit j feeds parity s */
    uint8_t corr[4096][3];
} gq_code;

static gq_code g_golay24, g_golay20, g_qr16;

static unsigned
gq_data_syndrome(const gq_code* c, int j) { /* syndrome (r bits, row 0 = MSB) of a single data-bit error */
    unsigned s = 0;
    for (int row = 0; row < c->r; row++) {
        s |= (unsigned)((c->pt[row] >> (c->k - 1 - j)) & 1u) << (c->r - 1 - row);
    }
    return s;
}

/* Reproduces the write order of Golay_20_8_init / Golay_24_12_init / QR_16_7_6_init (src/fec/fec.c:452-528,
 * 591-667, 737-783) including the partial writes that leave stale higher slots untouched. */
static void
gq_build_table(gq_code* c) {
    const int k = c->k, r = c->r;
    memset(c->corr, 0xFF, sizeof(c->corr));
#define PB(ip) (1u << (r - 1 - (ip)))
    for (int i1 = 0; i1 < k; i1++) {
        unsigned s1 = gq_data_syndrome(c, i1);
        for (int i2 = i1 + 1; i2 < k; i2++) {
            unsigned s2 = s1 ^ gq_data_syndrome(c, i2);
            if (c->maxw >= 3) {
                for (int i3 = i2 + 1; i3 < k; i3++) {
                    unsigned s3 = s2 ^ gq_data_syndrome(c, i3);
                    c->corr[s3][0] = (uint8_t)i1;
                    c->corr[s3][1] = (uint8_t)i2;
                    c->corr[s3][2] = (uint8_t)i3;
                }
            }
            c->corr[s2][0] = (uint8_t)i1;
            c->corr[s2][1] = (uint8_t)i2;
            if (c->maxw >= 3) {
                for (int ip = 0; ip < r; ip++) {
                    unsigned s = s2 ^ PB(ip);
                    c->corr[s][0] = (uint8_t)i1;
                    c->corr[s][1] = (uint8_t)i2;
                    c->corr[s][2] = (uint8_t)(k + ip);
                }
            }
        }
        c->corr[s1][0] = (uint8_t)i1;
        for (int ip1 = 0; ip1 < r; ip1++) {
            unsigned sa = s1 ^ PB(ip1);
            c->corr[sa][0] = (uint8_t)i1;
            c->corr[sa][1] = (uint8_t)(k + ip1);
            if (c->maxw >= 3) {
                for (int ip2 = ip1 + 1; ip2 < r; ip2++) {
                    unsigned sb = sa ^ PB(ip2);
                    c->corr[sb][0] = (uint8_t)i1;
                    c->corr[sb][1] = (uint8_t)(k + ip1);
                    c->corr[sb][2] = (uint8_t)(k + ip2);
                }
            }
        }
    }
    for (int ip1 = 0; ip1 < r; ip1++) {
        unsigned sa = PB(ip1);
        c->corr[sa][0] = (uint8_t)(k + ip1);
        for (int ip2 = ip1 + 1; ip2 < r; ip2++) {
            unsigned sb = sa ^ PB(ip2);
            c->corr[sb][0] = (uint8_t)(k + ip1);
            c->corr[sb][1] = (uint8_t)(k + ip2);
            if (c->maxw >= 3) {
                for (int ip3 = ip2 + 1; ip3 < r; ip3++) {
                    unsigned sc = sb ^ PB(ip3);
                    c->corr[sc][0] = (uint8_t)(k + ip1);
                    c->corr[sc][1] = (uint8_t)(k + ip2);
                    c->corr[sc][2] = (uint8_t)(k + ip3);
                }
            }
        }
    }
#undef PB
}

static void
gq_build(gq_code* c, int k, int r, int maxw, unsigned g, int deg, int full_k) {
    /* parity word of data bit i (of the unshortened code with full_k data bits):
     * [x^(deg + full_k - 1 - i) mod g, overall parity]; shortening drops the first full_k - k data bits */
    memset(c, 0, sizeof(*c));
    c->k = k;
    c->r = r;
    c->maxw = maxw;
    for (int j = 0; j < k; j++) {
        int i = j + (full_k - k);
        unsigned rem = xpow_mod(deg + full_k - 1 - i, g, deg);
        unsigned word = (rem << 1) | ((__builtin_popcount(rem) + 1) & 1u);
        for (int row = 0; row < r; row++) {
            if ((word >> (r - 1 - row)) & 1u) {
                c->pt[row] |= (uint16_t)(1u << (k - 1 - j));
            }
        }
    }
    gq_build_table(c);
}

void
oracle_fec_init(void) {
    if (g_ready) {
        return;
    }
    ham_build(&g_ham[ORACLE_HAMMING_7_4], 7, 3, 0xB, 3, 0);
    ham_build(&g_ham[ORACLE_HAMMING_12_8], 12, 4, 0x13, 4, 0);
    ham_build(&g_ham[ORACLE_HAMMING_13_9], 13, 4, 0x13, 4, 0);
    ham_build(&g_ham[ORACLE_HAMMING_15_11], 15, 4, 0x13, 4, 0);
    ham_build(&g_ham[ORACLE_HAMMING_16_11_4], 16, 5, 0x13, 4, 1);
    g_ham[ORACLE_HAMMING_13_9].stop_on_fail = 1;
    g_ham[ORACLE_HAMMING_15_11].stop_on_fail = 1;
    g_ham[ORACLE_HAMMING_16_11_4].stop_on_fail = 1;
    gq_build(&g_golay24, 12, 12, 3, 0xC75, 11, 12);
    gq_build(&g_golay20, 8, 12, 3, 0xC75, 11, 12);
    gq_build(&g_qr16, 7, 9, 2, 0x139, 8, 7);
    g_ready = 1;
}

static unsigned
ham_syndrome(const ham_code* c, const uint8_t* bits) {
    unsigned s = 0;
    for (int j = 0; j < c->n; j++) {
        if (bits[j] & 1) { /* the reference sums raw bytes mod 2: only the LSB matters (fec.c:150-157) */
            s ^= c->col[j];
        }
    }
    return s;
}

/* Hamming_7_4_decode (fec.c:145-168) and Hamming_{12_8,13_9,15_11,16_11_4}_decode with nbCodewords = 1
 * (fec.c:188-450).  Returns the reference's bool.  `decoded` may be NULL; for the codes that stop on failure the
 * info bits are NOT copied when the word is uncorrectable (the reference breaks out before the memcpy). */
int
oracle_hamming_decode(int code, uint8_t* bits, uint8_t* decoded) {
    oracle_fec_init();
    const ham_code* c = &g_ham[code];
    unsigned s = ham_syndrome(c, bits);
    int ok = 1;
    if (s) {
        if (c->pos_of[s] == 0xFF) {
            ok = 0;
            if (c->stop_on_fail || code == ORACLE_HAMMING_7_4) {
                return 0;
            }
        } else {
            bits[c->pos_of[s]] ^= 1;
        }
    }
    if (decoded && code != ORACLE_HAMMING_7_4) {
        memcpy(decoded, bits, (size_t)(c->n - c->r));
    }
    return ok;
}

static unsigned
gq_syndrome(const gq_code* c, const uint8_t* bits) {
    unsigned s = 0;
    for (int row = 0; row < c->r; row++) {
        unsigned par = bits[c->k + row] & 1u;
        for (int j = 0; j < c->k; j++) {
            par ^= (bits[j] & 1u) & ((c->pt[row] >> (c->k - 1 - j)) & 1u);
        }
        s |= par << (c->r - 1 - row);
    }
    return s;
}

/* Golay_24_12_decode (fec.c:682-733), Golay_20_8_decode (fec.c:543-587; more than 2 flips => false AFTER flipping),
 * QR_16_7_6_decode (fec.c:787-824). */
int
oracle_golay_24_12_decode(uint8_t* bits) {
    oracle_fec_init();
    unsigned s = gq_syndrome(&g_golay24, bits);
    if (!s) {
        return 1;
    }
    int i = 0;
    for (; i < 3 && g_golay24.corr[s][i] != 0xFF; i++) {
        bits[g_golay24.corr[s][i]] ^= 1;
    }
    return i != 0;
}

void
oracle_golay_24_12_encode(const uint8_t* data12, uint8_t* out24) {
    oracle_fec_init();
    for (int j = 0; j < 12; j++) {
        out24[j] = data12[j] & 1;
    }
    for (int row = 0; row < 12; row++) {
        unsigned par = 0;
        for (int j = 0; j < 12; j++) {
            par ^= (data12[j] & 1u) & ((g_golay24.pt[row] >> (11 - j)) & 1u);
        }
        out24[12 + row] = (uint8_t)par;
    }
}

int
oracle_golay_20_8_decode(uint8_t* bits) {
    oracle_fec_init();
    unsigned s = gq_syndrome(&g_golay20, bits);
    if (!s) {
        return 1;
    }
    int i = 0;
    for (; i < 3 && g_golay20.corr[s][i] != 0xFF; i++) {
        bits[g_golay20.corr[s][i]] ^= 1;
    }
    return (i != 0) && (i <= 2);
}

int
oracle_qr_16_7_6_decode(uint8_t* bits) {
    oracle_fec_init();
    unsigned s = gq_syndrome(&g_qr16, bits);
    if (!s) {
        return 1;
    }
    int i = 0;
    for (; i < 2 && g_qr16.corr[s][i] != 0xFF; i++) {
        bits[g_qr16.corr[s][i]] ^= 1;
    }
    return i != 0;
}

/* ------------------------------------------------------------------ BPTC(196,96) */

/* BPTCDeInterleaveDMRData (src/fec/bptc.c:22-59): out[(13 i) mod 196] = in[i] & 1 */
void
oracle_bptc_deinterleave(const uint8_t* in196, uint8_t* out196) {
    for (int i = 0; i < 196; i++) {
        out196[(13 * i) % 196] = in196[i] & 1;
    }
}

/* BPTC_196x96_Extract_Data (src/fec/bptc.c:61-149).  Two row/column passes, only the second counted.
 * Stale-buffer behaviour of the reference: when a Hamming(13,9) column is uncorrectable the callee returns before
 * writing its output buffer, so the caller copies whatever that buffer held -- the previous successfully decoded
 * column of the same pass.  For the first column of a pass the reference reads an uninitialised stack buffer;
 * here the column is left unchanged and *undefined_out is set so tests can exclude the case.
 * (Hamming(15,11) is perfect: a row can never be uncorrectable.) */
unsigned
oracle_bptc_196x96_extract(const uint8_t* in196, uint8_t* out96, uint8_t* r3, int* undefined_out) {
    oracle_fec_init();
    uint8_t m[13][15];
    int k = 1;
    if (undefined_out) {
        *undefined_out = 0;
    }
    for (int i = 0; i < 13; i++) {
        for (int j = 0; j < 15; j++) {
            m[i][j] = in196[k++] & 1;
        }
    }
    unsigned errs = 0;
    for (int pass = 0; pass < 2; pass++) {
        unsigned e = 0;
        for (int i = 0; i < 9; i++) {
            uint8_t line[15], dec[11];
            memcpy(line, m[i], 15);
            (void)oracle_hamming_decode(ORACLE_HAMMING_15_11, line, dec); /* always correctable */
            memcpy(m[i], dec, 11);
        }
        uint8_t last[9];
        int have_last = 0;
        for (int j = 0; j < 15; j++) {
            uint8_t col[13], dec[9];
            for (int i = 0; i < 13; i++) {
                col[i] = m[i][j];
            }
            if (oracle_hamming_decode(ORACLE_HAMMING_13_9, col, dec)) {
                memcpy(last, dec, 9);
                have_last = 1;
                for (int i = 0; i < 9; i++) {
                    m[i][j] = dec[i];
                }
            } else {
                e++;
                if (have_last) {
                    for (int i = 0; i < 9; i++) {
                        m[i][j] = last[i];
                    }
                } else if (undefined_out) {
                    *undefined_out = 1;
                }
            }
        }
        if (pass == 1) {
            errs = e;
        }
    }
    k = 0;
    for (int j = 3; j < 11; j++) {
        out96[k++] = m[0][j];
    }
    for (int i = 1; i < 9; i++) {
        for (int j = 0; j < 11; j++) {
            out96[k++] = m[i][j];
        }
    }
    r3[0] = m[0][2];
    r3[1] = m[0][1];
    r3[2] = m[0][0];
    return errs;
}

/* BPTC_128x77_Extract_Data (src/fec/bptc.c:167-252): 8 x 16 matrix, rows 0-6 Hamming(16,11,4), row 7 = column
 * parity.  An uncorrectable row is overwritten with the callee's stale output buffer = the information bits of the
 * most recent correctable row of this call; if there is none yet the reference reads an uninitialised stack buffer
 * (undefined: row left unchanged here, *undefined_out set).  Return = uncorrectable rows + column-parity failures. */
unsigned
oracle_bptc_128x77_extract(const uint8_t* in128, uint8_t* out77, int* undefined_out) {
    uint8_t m[8][16];
    if (undefined_out) {
        *undefined_out = 0;
    }
    for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 16; j++) {
            m[i][j] = in128[16 * i + j] & 1u;
        }
    }
    unsigned ham_err = 0, par_err = 0;
    uint8_t last[11];
    int have_last = 0;
    for (int i = 0; i < 7; i++) {
        uint8_t line[16], dec[11];
        memcpy(line, m[i], 16);
        if (oracle_hamming_decode(ORACLE_HAMMING_16_11_4, line, dec)) {
            memcpy(last, dec, 11);
            have_last = 1;
            memcpy(m[i], dec, 11);
        } else {
            ham_err++;
            if (have_last) {
                memcpy(m[i], last, 11);
            } else if (undefined_out) {
                *undefined_out = 1;
            }
        }
    }
    int k = 0;
    for (int i = 0; i < 2; i++) {
        for (int j = 0; j < 11; j++) {
            out77[k++] = m[i][j];
        }
    }
    for (int i = 2; i < 7; i++) {
        for (int j = 0; j < 10; j++) {
            out77[k++] = m[i][j];
        }
    }
    for (int i = 2; i < 7; i++) {
        out77[k++] = m[i][10];
    }
    for (int j = 0; j < 16; j++) {
        unsigned ones = 0;
        for (int i = 0; i < 7; i++) {
            ones += m[i][j];
        }
        if ((ones % 2) != m[7][j]) {
            par_err++;
        }
    }
    return ham_err + par_err;
}

/* BPTC_16x2_Extract_Data (src/fec/bptc.c:272-333): reverse-channel de-interleave (tables :33-38), first 16 bits
 * Hamming(16,11,4), last 16 bits = parity of the first 16 (odd or even).  When the Hamming word is uncorrectable the
 * reference copies an uninitialised buffer over out[0..10] (undefined: left as de-interleaved here, flag set). */
unsigned
oracle_bptc_16x2_extract(const uint8_t* in32, uint8_t* out32, unsigned parity_odd, int* undefined_out) {
    static const uint8_t dei[32] = {0,  17, 2,  19, 4,  21, 6,  23, 8,  25, 10, 27, 12, 29, 14, 31,
                                    16, 1,  18, 3,  20, 5,  22, 7,  24, 9,  26, 11, 28, 13, 30, 15};
    static const uint8_t place[32] = {0,  16, 1,  17, 2,  18, 3,  19, 4,  20, 5,  21, 6,  22, 7,  23,
                                      8,  24, 9,  25, 10, 26, 11, 27, 12, 28, 13, 29, 14, 30, 15, 31};
    uint8_t m[32], line[16], dec[11];
    if (undefined_out) {
        *undefined_out = 0;
    }
    for (int i = 0; i < 32; i++) {
        m[place[dei[i]]] = in32[i] & 1u;
    }
    memcpy(out32, m, 32);
    memcpy(line, m, 16);
    unsigned ham_err = 0, odd_err = 0, even_err = 0;
    if (oracle_hamming_decode(ORACLE_HAMMING_16_11_4, line, dec)) {
        memcpy(out32, dec, 11);
    } else {
        ham_err = 1;
        if (undefined_out) {
            *undefined_out = 1;
        }
    }
    for (int i = 0; i < 16; i++) {
        if (out32[i] == out32[i + 16]) {
            odd_err++;
        } else {
            even_err++;
        }
    }
    return ham_err + (parity_odd ? odd_err : even_err);
}

/* ------------------------------------------------------------------ P25 half-rate trellis */

/* dibit-pair nibble expected on the transition prev -> next (src/protocol/p25/p25_12.c:19) */
static const uint64_t k_p25_dtm_packed = 0x86B54A793D0EF1C2ull; /* nibble i = entry i, entry 0 in the low nibble */

static inline unsigned
p25_dtm(int prev, int next) {
    return (unsigned)((k_p25_dtm_packed >> (4 * ((prev << 2) | next))) & 0xF);
}

/* position of received dibit i in the de-interleaved stream (src/fec/trellis34.c:8-13):
 * four groups g = 0..3, each walking dibit pairs (j, j+1) for j = 2g, 2g+8, 2g+16, ... < 98 */
static void
p25_interleave_98(uint8_t* t) {
    int n = 0;
    for (int g = 0; g < 4; g++) {
        for (int j = 2 * g; j < 98; j += 8) {
            t[n++] = (uint8_t)j;
            t[n++] = (uint8_t)(j + 1);
        }
    }
}

static inline uint32_t
llr_cost(int16_t llr, unsigned bit) {
    if (bit) {
        return llr < 0 ? (uint32_t)(-(int)llr) : 0u;
    }
    return llr > 0 ? (uint32_t)llr : 0u;
}

static void
p25_deinterleave_llr(const int16_t* llr196, int16_t* dei) {
    uint8_t t[98];
    p25_interleave_98(t);
    memset(dei, 0, 196 * sizeof(int16_t));
    for (int i = 0; i < 98; i++) {
        dei[2 * t[i]] = llr196[2 * i];
        dei[2 * t[i] + 1] = llr196[2 * i + 1];
    }
}

static inline uint32_t
p25_branch_cost(const int16_t* dei, int sym, int prev, int next) {
    unsigned e = p25_dtm(prev, next);
    const int16_t* l = dei + 4 * sym;
    return llr_cost(l[0], (e >> 3) & 1) + llr_cost(l[1], (e >> 2) & 1) + llr_cost(l[2], (e >> 1) & 1) + llr_cost(l[3], e & 1);
}

/* p25_12_soft_llr (src/protocol/p25/p25_12.c:204-283) */
int
oracle_p25_12_soft_llr(const int16_t* llr196, uint8_t out12[12]) {
    int16_t dei[196];
    p25_deinterleave_llr(llr196, dei);
    uint32_t pm[4] = {0, 256, 256, 256}, cm[4] = {0, 0, 0, 0};
    uint8_t bp[49][4];
    for (int i = 0; i < 49; i++) {
        for (int nx = 0; nx < 4; nx++) {
            uint32_t best = 0xFFFFFFFFu;
            uint8_t bprev = 0;
            for (int pv = 0; pv < 4; pv++) {
                uint32_t m = pm[pv] + p25_branch_cost(dei, i, pv, nx);
                if (m < best) {
                    best = m;
                    bprev = (uint8_t)pv;
                }
            }
            cm[nx] = best;
            bp[i][nx] = bprev;
        }
        memcpy(pm, cm, sizeof(pm));
    }
    int st = 0;
    uint32_t bf = cm[0];
    for (int j = 1; j < 4; j++) {
        if (cm[j] < bf) {
            bf = cm[j];
            st = j;
        }
    }
    uint8_t td[49];
    for (int i = 49; i-- > 0;) {
        td[i] = (uint8_t)st;
        st = bp[i][st];
    }
    for (int i = 0; i < 12; i++) {
        out12[i] = (uint8_t)((td[4 * i] << 6) | (td[4 * i + 1] << 4) | (td[4 * i + 2] << 2) | td[4 * i + 3]);
    }
    return (int)(bf >> 8);
}

/* p25_12_soft_llr_list (src/protocol/p25/p25_12.c:31-202): list Viterbi, <= 8 survivors per state */
int
oracle_p25_12_soft_llr_list(const int16_t* llr196, uint8_t* cand_bytes /*[max][12]*/, uint32_t* cand_metric, int max_candidates) {
    enum { K = 8 };
    if (!llr196 || !cand_bytes || !cand_metric || max_candidates <= 0) {
        return 0;
    }
    if (max_candidates > K) {
        max_candidates = K;
    }
    int16_t dei[196];
    p25_deinterleave_llr(llr196, dei);
    static uint32_t ma[4][K], mb[4][K];
    static uint8_t bp[49][4][K];
    uint32_t (*pm)[K] = ma;
    uint32_t (*cm)[K] = mb;
    memset(ma, 0xFF, sizeof(ma));
    memset(mb, 0xFF, sizeof(mb));
    memset(bp, 0, sizeof(bp));
    for (int s = 0; s < 4; s++) {
        pm[s][0] = s == 0 ? 0u : 256u;
    }
    for (int i = 0; i < 49; i++) {
        memset(cm, 0xFF, sizeof(ma));
        for (int pv = 0; pv < 4; pv++) {
            for (int nx = 0; nx < 4; nx++) {
                uint32_t cost = p25_branch_cost(dei, i, pv, nx);
                for (int rk = 0; rk < K; rk++) {
                    if (pm[pv][rk] == 0xFFFFFFFFu) {
                        continue;
                    }
                    uint32_t m = pm[pv][rk] + cost;
                    int at = -1;
                    for (int q = 0; q < K; q++) {
                        if (m < cm[nx][q]) {
                            at = q;
                            break;
                        }
                    }
                    if (at < 0) {
                        continue;
                    }
                    for (int q = K - 1; q > at; q--) {
                        cm[nx][q] = cm[nx][q - 1];
                        bp[i][nx][q] = bp[i][nx][q - 1];
                    }
                    cm[nx][at] = m;
                    bp[i][nx][at] = (uint8_t)((pv << 3) | rk);
                }
            }
        }
        uint32_t (*t)[K] = pm;
        pm = cm;
        cm = t;
    }
    int count = 0;
    for (int s = 0; s < 4; s++) {
        for (int rk = 0; rk < K; rk++) {
            if (pm[s][rk] == 0xFFFFFFFFu) {
                continue;
            }
            uint8_t td[49], bytes[12];
            int st = s, r = rk;
            for (int i = 49; i-- > 0;) {
                td[i] = (uint8_t)st;
                uint8_t p = bp[i][st][r];
                st = (p >> 3) & 3;
                r = p & 7;
            }
            for (int i = 0; i < 12; i++) {
                bytes[i] = (uint8_t)((td[4 * i] << 6) | (td[4 * i + 1] << 4) | (td[4 * i + 2] << 2) | td[4 * i + 3]);
            }
            int dup = 0;
            for (int c = 0; c < count; c++) {
                if (memcmp(cand_bytes + 12 * c, bytes, 12) == 0) {
                    dup = 1;
                }
            }
            if (dup) {
                continue;
            }
            uint32_t metric = pm[s][rk];
            int at = count;
            for (int c = 0; c < count; c++) {
                if (metric < cand_metric[c]) {
                    at = c;
                    break;
                }
            }
            if (count < max_candidates) {
                count++;
            } else if (at >= max_candidates) {
                continue;
            }
            for (int c = count - 1; c > at; c--) {
                memcpy(cand_bytes + 12 * c, cand_bytes + 12 * (c - 1), 12);
                cand_metric[c] = cand_metric[c - 1];
            }
            memcpy(cand_bytes + 12 * at, bytes, 12);
            cand_metric[at] = metric;
        }
    }
    return count;
}

/* ------------------------------------------------------------------ Reed-Solomon over GF(64) */

static int gf_exp[64], gf_log[64];
static int gf_ready = 0;

static void
gf_init(void) { /* GF(2^6), primitive polynomial x^6 + x + 1 (ReedSolomon.hpp:685-727) */
    if (gf_ready) {
        return;
    }
    int v = 1;
    for (int i = 0; i < 63; i++) {
        gf_exp[i] = v;
        gf_log[v] = i;
        v <<= 1;
        if (v & 0x40) {
            v ^= 0x43;
        }
    }
    gf_exp[63] = 0;
    gf_log[0] = -1;
    gf_ready = 1;
}

/* ReedSolomon_63<TT>::decode (ReedSolomon.hpp:738-771 with :353-582): Berlekamp iteration in the index-form
 * bookkeeping of Rockliff's rs.c.  in/out: 63 symbols in polynomial form.  Returns 0 = ok / corrected, 1 = irrecoverable
 * (out == in). */
int
oracle_rs63_decode(int tt, const int* in63, int* out63) {
    enum { NN = 63 };
    gf_init();
    const int n2t = 2 * tt;
    int recd[NN], s[2 * 8 + 2];
    for (int i = 0; i < NN; i++) {
        recd[i] = gf_log[in63[i]];
    }
    int syn_err = 0;
    s[0] = 0;
    for (int i = 1; i <= n2t; i++) {
        int acc = 0;
        for (int j = 0; j < NN; j++) {
            if (recd[j] != -1) {
                acc ^= gf_exp[(recd[j] + i * j) % NN];
            }
        }
        if (acc) {
            syn_err = 1;
        }
        s[i] = gf_log[acc];
    }
    if (!syn_err) {
        memcpy(out63, in63, NN * sizeof(int));
        return 0;
    }
    int elp[2 * 8 + 2][2 * 8], d[2 * 8 + 2], l[2 * 8 + 2], u_lu[2 * 8 + 2];
    d[0] = 0;
    d[1] = s[1];
    elp[0][0] = 0;
    elp[1][0] = 1;
    for (int i = 1; i < n2t; i++) {
        elp[0][i] = -1;
        elp[1][i] = 0;
    }
    l[0] = l[1] = 0;
    u_lu[0] = -1;
    u_lu[1] = 0;
    int u = 0;
    do {
        u++;
        if (d[u] == -1) {
            l[u + 1] = l[u];
            for (int i = 0; i <= l[u]; i++) {
                elp[u + 1][i] = elp[u][i];
                elp[u][i] = gf_log[elp[u][i]];
            }
        } else {
            int q = u - 1;
            while (q > 0 && d[q] == -1) {
                q--;
            }
            if (q > 0) {
                for (int j = q - 1; j > 0; j--) {
                    if (d[j] != -1 && u_lu[q] < u_lu[j]) {
                        q = j;
                    }
                }
            }
            l[u + 1] = (l[u] > l[q] + u - q) ? l[u] : l[q] + u - q;
            for (int i = 0; i < n2t; i++) {
                elp[u + 1][i] = 0;
            }
            for (int i = 0; i <= l[q]; i++) {
                if (elp[q][i] != -1) {
                    elp[u + 1][i + u - q] = gf_exp[(d[u] + NN - d[q] + elp[q][i]) % NN];
                }
            }
            for (int i = 0; i <= l[u]; i++) {
                elp[u + 1][i] ^= elp[u][i];
                elp[u][i] = gf_log[elp[u][i]];
            }
        }
        u_lu[u + 1] = u - l[u + 1];
        if (u < n2t) {
            int dd = (s[u + 1] != -1) ? gf_exp[s[u + 1]] : 0;
            for (int i = 1; i <= l[u + 1]; i++) {
                if (s[u + 1 - i] != -1 && elp[u + 1][i] != 0) {
                    dd ^= gf_exp[(s[u + 1 - i] + gf_log[elp[u + 1][i]]) % NN];
                }
            }
            d[u + 1] = gf_log[dd];
        }
    } while (u < n2t && l[u + 1] <= tt);
    u++;
    if (l[u] > tt) {
        memcpy(out63, in63, NN * sizeof(int));
        return 1;
    }
    int deg = l[u];
    int loc_poly[9];
    for (int i = 0; i <= deg; i++) {
        loc_poly[i] = gf_log[elp[u][i]];
    }
    int reg[9], root[8], loc[8], count = 0;
    for (int i = 1; i <= deg; i++) {
        reg[i] = loc_poly[i];
    }
    for (int i = 1; i <= NN; i++) {
        int q = 1;
        for (int j = 1; j <= deg; j++) {
            if (reg[j] != -1) {
                reg[j] = (reg[j] + j) % NN;
                q ^= gf_exp[reg[j]];
            }
        }
        if (!q) {
            if (count < 8) {
                root[count] = i;
                loc[count] = NN - i;
            }
            count++;
        }
    }
    if (count != deg) {
        memcpy(out63, in63, NN * sizeof(int));
        return 1;
    }
    int z[9];
    for (int i = 1; i <= deg; i++) {
        int zi;
        if (s[i] != -1 && loc_poly[i] != -1) {
            zi = gf_exp[s[i]] ^ gf_exp[loc_poly[i]];
        } else if (s[i] != -1) {
            zi = gf_exp[s[i]];
        } else if (loc_poly[i] != -1) {
            zi = gf_exp[loc_poly[i]];
        } else {
            zi = 0;
        }
        for (int j = 1; j < i; j++) {
            if (s[j] != -1 && loc_poly[i - j] != -1) {
                zi ^= gf_exp[(loc_poly[i - j] + s[j]) % NN];
            }
        }
        z[i] = gf_log[zi];
    }
    memcpy(out63, in63, NN * sizeof(int));
    for (int i = 0; i < deg; i++) {
        int num = 1;
        for (int j = 1; j <= deg; j++) {
            if (z[j] != -1) {
                num ^= gf_exp[(z[j] + j * root[i]) % NN];
            }
        }
        if (num != 0) {
            int den = 0;
            for (int j = 0; j < deg; j++) {
                if (j != i) {
                    den += gf_log[1 ^ gf_exp[(loc[j] + root[i]) % NN]];
                }
            }
            den %= NN;
            int e = gf_exp[(gf_log[num] - den + NN) % NN];
            out63[loc[i]] ^= e;
        }
    }
    return 0;
}

/* Systematic encoder for tests (the reference tree has no encoder for this code): parity symbols occupy positions
 * 0..2t-1, data the following kk positions -- the layout the reference decoder expects (ReedSolomon.hpp:838-851).
 * Generator g(x) = prod_{i=1..2t} (x + alpha^i); codeword polynomial c(x) = sum c[j] x^j has c(alpha^i) = 0. */
void
oracle_rs63_encode(int tt, const int* data /*[63-2tt]*/, int* cw63) {
    gf_init();
    const int n2t = 2 * tt, kk = 63 - n2t;
    int g[17];
    g[0] = 1;
    for (int i = 1; i <= n2t; i++) {
        g[i] = 0;
    }
    for (int i = 1; i <= n2t; i++) { /* multiply by (x + alpha^i) */
        for (int j = i; j > 0; j--) {
            int t = g[j - 1];
            if (g[j]) {
                t ^= gf_exp[(gf_log[g[j]] + i) % 63];
            }
            g[j] = t;
        }
        g[0] = gf_exp[(gf_log[g[0]] + i) % 63];
    }
    int par[16];
    for (int i = 0; i < n2t; i++) {
        par[i] = 0;
    }
    for (int i = kk - 1; i >= 0; i--) { /* LFSR division of data(x) x^{2t} by g(x) */
        int fb = data[i] ^ par[n2t - 1];
        for (int j = n2t - 1; j > 0; j--) {
            par[j] = par[j - 1] ^ (fb && g[j] ? gf_exp[(gf_log[fb] + gf_log[g[j]]) % 63] : 0);
        }
        par[0] = (fb && g[0]) ? gf_exp[(gf_log[fb] + gf_log[g[0]]) % 63] : 0;
    }
    for (int i = 0; i < n2t; i++) {
        cw63[i] = par[i];
    }
    for (int i = 0; i < kk; i++) {
        cw63[n2t + i] = data[i];
    }
}

/* check_and_fix_redsolomon_36_20_17 / _24_12_13 / _24_16_9 (phase1/p25p1_check_hdu.cpp:38-45, p25p1_check_ldu.cpp:37-71;
 * wrappers ReedSolomon.hpp:821-877,918-955,1013-1050): byte-per-bit hex words, parity first in the codeword. */
int
oracle_p25_rs_decode(int n_total, int n_data, uint8_t* data_bits /*[n_data][6]*/, const uint8_t* parity_bits /*[n_par][6]*/) {
    const int n_par = n_total - n_data, tt = n_par / 2;
    int in[63], out[63];
    for (int i = 0; i < 63; i++) {
        in[i] = 0;
    }
    for (int i = 0; i < n_par; i++) {
        int v = 0;
        for (int b = 0; b < 6; b++) {
            v = (v << 1) | (parity_bits[6 * i + b] != 0);
        }
        in[i] = v;
    }
    for (int i = 0; i < n_data; i++) {
        int v = 0;
        for (int b = 0; b < 6; b++) {
            v = (v << 1) | (data_bits[6 * i + b] != 0);
        }
        in[n_par + i] = v;
    }
    int rc = oracle_rs63_decode(tt, in, out);
    for (int i = 0; i < n_data; i++) {
        for (int b = 0; b < 6; b++) {
            data_bits[6 * i + b] = (uint8_t)((out[n_par + i] >> (5 - b)) & 1);
        }
    }
    return rc;
}

/* ---- bounded errors-and-erasures decoding + the P25p1 ranked-erasure soft wrappers ---------------------------- */

static int
gfm(int a, int b) { /* ReedSolomon.hpp:72-78 */
    return (a == 0 || b == 0) ? 0 : gf_exp[(gf_log[a] + gf_log[b]) % 63];
}

static int
gfd(int a, int b) { /* ReedSolomon.hpp:80-89 (division by zero yields 0) */
    return (a == 0 || b == 0) ? 0 : gf_exp[(gf_log[a] - gf_log[b] + 63) % 63];
}

static int
gfa(int e) { /* ReedSolomon.hpp:91-98 */
    e %= 63;
    if (e < 0) {
        e += 63;
    }
    return gf_exp[e];
}

static int
rs_syndromes(const int* w, int n2t, int* syn) { /* ReedSolomon.hpp:102-119; syn[1..2t], polynomial form */
    int err = 0;
    syn[0] = 0;
    for (int i = 1; i <= n2t; i++) {
        int acc = 0;
        for (int j = 0; j < 63; j++) {
            if (w[j]) {
                acc ^= gfm(w[j], gfa(i * j));
            }
        }
        syn[i] = acc;
        err |= acc != 0;
    }
    return err;
}

/* ReedSolomon_63<TT>::decode_with_erasures (ReedSolomon.hpp:773-795) = run_decode_with_erasures (:621-683): erasure
 * locator, modified syndromes, Berlekamp-Massey on the tail, combined locator, root search over all 63 positions,
 * error values by Gauss-Jordan on the Vandermonde system (first non-zero pivot), re-check of the syndromes.
 * Returns 0 ok (out corrected) / 1 failure (out == in). */
int
oracle_rs63_decode_with_erasures(int tt, const int* in63, int* out63, const int* erasures, int n_er) {
    gf_init();
    const int n2t = 2 * tt;
    if (!in63 || !out63 || n_er < 0 || n_er > n2t) {
        return 1;
    }
    memcpy(out63, in63, 63 * sizeof(int));
    int syn[17];
    if (n_er == 0) {
        return rs_syndromes(out63, n2t, syn) ? 1 : 0;
    }
    if (!erasures) {
        return 1;
    }
    {
        int seen[63] = {0};
        for (int i = 0; i < n_er; i++) {
            if (erasures[i] < 0 || erasures[i] >= 63 || seen[erasures[i]]) {
                return 1;
            }
            seen[erasures[i]] = 1;
        }
    }
    int status = 1;
    do {
        if (!rs_syndromes(out63, n2t, syn)) {
            status = 0;
            break;
        }
        int el[17] = {0}; /* erasure locator */
        el[0] = 1;
        for (int e = 0, deg = 0; e < n_er; e++, deg++) {
            int f = gfa(erasures[e]);
            for (int i = deg; i >= 0; i--) {
                el[i + 1] ^= gfm(el[i], f);
            }
        }
        int ms[16]; /* modified syndromes */
        for (int i = 0; i < n2t; i++) {
            int v = 0;
            for (int j = 0; j <= n_er && j <= i; j++) {
                v ^= gfm(el[j], syn[(i - j) + 1]);
            }
            ms[i] = v;
        }
        /* Berlekamp-Massey on ms[n_er .. 2t) (ReedSolomon.hpp:204-262) */
        int c[17] = {0}, b[17] = {0}, t[17];
        c[0] = b[0] = 1;
        int l = 0, m = 1, bb = 1, fail = 0;
        const int* sy = ms + n_er;
        const int ns = n2t - n_er;
        for (int n = 0; n < ns; n++) {
            int disc = sy[n];
            for (int i = 1; i <= l; i++) {
                disc ^= gfm(c[i], sy[n - i]);
            }
            if (disc == 0) {
                m++;
                continue;
            }
            memcpy(t, c, sizeof(t));
            if (bb == 0) {
                fail = 1;
                break;
            }
            int coef = gfd(disc, bb);
            for (int i = 0; i + m <= n2t; i++) {
                if (b[i]) {
                    c[i + m] ^= gfm(coef, b[i]);
                }
            }
            if (2 * l <= n) {
                l = n + 1 - l;
                memcpy(b, t, sizeof(b));
                bb = disc;
                m = 1;
            } else {
                m++;
            }
        }
        if (fail) {
            break;
        }
        int udeg = 0;
        for (int i = n2t; i >= 0; i--) {
            if (c[i]) {
                udeg = i;
                break;
            }
        }
        if (2 * udeg + n_er > n2t) {
            break;
        }
        int edeg = 0;
        for (int i = n2t; i >= 0; i--) {
            if (el[i]) {
                edeg = i;
                break;
            }
        }
        if (edeg + udeg > n2t) {
            break;
        }
        int comb[17] = {0};
        for (int i = 0; i <= edeg; i++) {
            for (int j = 0; j <= udeg; j++) {
                comb[i + j] ^= gfm(el[i], c[j]);
            }
        }
        int cdeg = 0;
        for (int i = n2t; i >= 0; i--) {
            if (comb[i]) {
                cdeg = i;
                break;
            }
        }
        int locs[16], n_loc = 0;
        if (cdeg != 0) { /* find_error_locations (:280-303) */
            for (int pos = 0; pos < 63; pos++) {
                int x = gfa(63 - pos), v = 0, xp = 1;
                for (int i = 0; i <= cdeg; i++) {
                    v ^= gfm(comb[i], xp);
                    xp = gfm(xp, x);
                }
                if (v == 0) {
                    if (n_loc >= n2t) {
                        n_loc++;
                        break;
                    }
                    locs[n_loc++] = pos;
                }
            }
        }
        if (n_loc != cdeg || n_loc > n2t) {
            break;
        }
        int ok = 1;
        for (int i = 0; i < n_er && ok; i++) {
            int found = 0;
            for (int k = 0; k < n_loc; k++) {
                found |= locs[k] == erasures[i];
            }
            ok = found;
        }
        if (!ok) {
            break;
        }
        int mat[16][17];
        memset(mat, 0, sizeof(mat));
        for (int r = 0; r < n_loc; r++) {
            for (int k = 0; k < n_loc; k++) {
                mat[r][k] = gfa((r + 1) * locs[k]);
            }
            mat[r][n_loc] = syn[r + 1];
        }
        int singular = 0;
        for (int col = 0; col < n_loc && !singular; col++) { /* solve_gf_linear_system (:121-161) */
            int piv = -1;
            for (int r = col; r < n_loc; r++) {
                if (mat[r][col]) {
                    piv = r;
                    break;
                }
            }
            if (piv < 0) {
                singular = 1;
                break;
            }
            if (piv != col) {
                for (int k = col; k <= n_loc; k++) {
                    int tmp = mat[col][k];
                    mat[col][k] = mat[piv][k];
                    mat[piv][k] = tmp;
                }
            }
            int pv = mat[col][col];
            for (int k = col; k <= n_loc; k++) {
                mat[col][k] = gfd(mat[col][k], pv);
            }
            for (int r = 0; r < n_loc; r++) {
                if (r == col || mat[r][col] == 0) {
                    continue;
                }
                int f = mat[r][col];
                for (int k = col; k <= n_loc; k++) {
                    mat[r][k] ^= gfm(f, mat[col][k]);
                }
            }
        }
        if (singular) {
            break;
        }
        for (int i = 0; i < n_loc; i++) {
            out63[locs[i]] ^= mat[i][n_loc];
        }
        if (rs_syndromes(out63, n2t, syn)) {
            break;
        }
        status = 0;
    } while (0);
    if (status) {
        memcpy(out63, in63, 63 * sizeof(int));
    }
    return status;
}

/* p25p1_build_rs_ranked_erasures (src/protocol/p25/phase1/p25p1_soft.cpp:140-170 with :83-136): every symbol is a
 * candidate (parity positions first), sorted ascending by (reliability, position); count = max(#below threshold,
 * min_erasures) capped at max_erasures. */
int
oracle_p25_rs_ranked_erasures(const uint8_t* data_rel, int n_data, const uint8_t* par_rel, int n_par, int min_er, int threshold,
                              int* erasures, int max_er) {
    uint8_t rel[64];
    int pos[64], n = 0, hits = 0;
    for (int i = 0; i < n_par && n < 64; i++) {
        hits += par_rel[i] < threshold;
        rel[n] = par_rel[i];
        pos[n++] = i;
    }
    for (int i = 0; i < n_data && n < 64; i++) {
        hits += data_rel[i] < threshold;
        rel[n] = data_rel[i];
        pos[n++] = n_par + i;
    }
    for (int i = 0; i < n; i++) {
        for (int j = i + 1; j < n; j++) {
            if (rel[j] < rel[i] || (rel[j] == rel[i] && pos[j] < pos[i])) {
                uint8_t tr = rel[i];
                rel[i] = rel[j];
                rel[j] = tr;
                int tp = pos[i];
                pos[i] = pos[j];
                pos[j] = tp;
            }
        }
    }
    int cnt = hits > min_er ? hits : min_er;
    if (cnt > n) {
        cnt = n;
    }
    if (cnt > max_er) {
        cnt = max_er;
    }
    for (int i = 0; i < cnt; i++) {
        erasures[i] = pos[i];
    }
    return cnt;
}

/* p25p1_rs_{36_20_17,24_12_13,24_16_9}_soft_reliability (phase1/p25p1_check_hdu.cpp:56-77, p25p1_check_ldu.cpp:73-94 and
 * the (24,12,13) twin) over DSDReedSolomon_*::decode_soft (ReedSolomon.hpp:879-913 ...): hard decode first; then the n
 * weakest symbols as erasures for n = 1..ranked, first success wins.  Ranking: min_erasures = t, max = 2t.
 * On success data_bits holds the corrected 0/1 bits; on failure data_bits is untouched.  Returns 0 / 1. */
int
oracle_p25_rs_soft_reliability(int n_total, int n_data, uint8_t* data_bits, const uint8_t* parity_bits, const uint8_t* data_rel,
                               const uint8_t* par_rel, int threshold) {
    gf_init();
    const int n_par = n_total - n_data, tt = n_par / 2;
    int er[16];
    int n_ranked = oracle_p25_rs_ranked_erasures(data_rel, n_data, par_rel, n_par, tt, threshold, er, 2 * tt);
    uint8_t cand[36 * 6];
    int in[63], out[63];
    for (int i = 0; i < 63; i++) {
        in[i] = 0;
    }
    for (int i = 0; i < n_par; i++) {
        int v = 0;
        for (int b = 0; b < 6; b++) {
            v = (v << 1) | (parity_bits[6 * i + b] != 0);
        }
        in[i] = v;
    }
    for (int i = 0; i < n_data; i++) {
        int v = 0;
        for (int b = 0; b < 6; b++) {
            v = (v << 1) | (data_bits[6 * i + b] != 0);
        }
        in[n_par + i] = v;
    }
    for (int n = 1; n <= n_ranked; n++) {
        memcpy(cand, data_bits, (size_t)n_data * 6);
        if (oracle_p25_rs_decode(n_total, n_data, cand, parity_bits) == 0) { /* decode_soft tries the hard decoder first */
            memcpy(data_bits, cand, (size_t)n_data * 6);
            return 0;
        }
        if (oracle_rs63_decode_with_erasures(tt, in, out, er, n) == 0) {
            for (int i = 0; i < n_data; i++) {
                for (int b = 0; b < 6; b++) {
                    data_bits[6 * i + b] = (uint8_t)((out[n_par + i] >> (5 - b)) & 1);
                }
            }
            return 0;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------ P25 word codes: Golay(24,6/12), Hamming(10,6,3), BCH(63,16,11) */

/* Golay24 (include/dsd-neo/fec/Golay24.hpp:17-222): Hank Wallace's (23,12) decoder + overall parity. */
#define GOLAY_POLY 0xAE3u

static unsigned
g23_syndrome(unsigned cw) { /* :58-72 */
    cw &= 0x7fffffu;
    for (int i = 1; i <= 12; i++) {
        if (cw & 1u) {
            cw ^= GOLAY_POLY;
        }
        cw >>= 1;
    }
    return cw << 12;
}

static unsigned
g23_rotl(unsigned cw) {
    cw = (cw & 0x400000u) ? ((cw << 1) | 1u) : (cw << 1);
    return cw & 0x7fffffu;
}

static unsigned
g23_rotr(unsigned cw, int n) {
    for (int i = 0; i < n; i++) {
        cw = (cw & 1u) ? ((cw >> 1) | 0x400000u) : (cw >> 1);
    }
    return cw & 0x7fffffu;
}

static int
parity32(unsigned cw) { /* Golay24::parity looks at the low 24 bits only */
    unsigned p = cw ^ (cw >> 8) ^ (cw >> 16);
    p ^= p >> 4;
    p ^= p >> 2;
    p ^= p >> 1;
    return (int)(p & 1u);
}

/* Golay24::correct (:108-169): weight <= 3 syndrome over the 23 cyclic shifts, then 23 trial flips with threshold 2.
 * *errs is whatever the last examined syndrome weighed, also when nothing could be corrected (the reference reports it). */
static unsigned
g23_correct(unsigned cw, int* errs) {
    const unsigned saver = cw;
    unsigned mask = 1;
    int w = 3, j = -1;
    *errs = 0;
    while (j < 23) {
        if (j != -1) {
            if (j > 0) {
                mask += mask;
            }
            cw = saver ^ mask;
            w = 2;
        }
        unsigned s = g23_syndrome(cw);
        if (!s) {
            return cw;
        }
        for (int i = 0; i < 23; i++) {
            *errs = __builtin_popcount(s & 0x7fffffu);
            if (*errs <= w) {
                return g23_rotr(cw ^ s, i);
            }
            cw = g23_rotl(cw);
            s = g23_syndrome(cw);
        }
        j++;
    }
    return saver;
}

/* DSDGolay24::decode_6 / decode_12 (:336-405) = check_and_fix_golay_24_6 / _24_12 (phase1/p25p1_check_hdu.cpp:26-36).
 * length = 6 or 12 data bits (MSB first) + 12 parity bits; returns 0 ok / 1 uncorrectable (data untouched). */
int
oracle_p25_golay24_decode(int length, uint8_t* word, const uint8_t* parity, int* fixed_errors) {
    *fixed_errors = 0;
    for (int i = 0; i < length; i++) {
        if (word[i] > 1) {
            return 1;
        }
    }
    for (int i = 0; i < 12; i++) {
        if (parity[i] > 1) {
            return 1;
        }
    }
    unsigned cw = 0;
    for (int i = 0; i < 12; i++) {
        cw = (cw << 1) | parity[11 - i];
    }
    for (int i = 0; i < length; i++) {
        cw = (cw << 1) | word[length - 1 - i];
    }
    cw <<= (12 - length);
    const unsigned pbit = cw & 0x800000u;
    cw = g23_correct(cw & ~0x800000u, fixed_errors) | pbit;
    int bad = parity32(cw);
    if (bad && (cw & 0x3fu) != 0) {
        return 1;
    }
    unsigned mask = 1u << (12 - length);
    for (int i = 0; i < length; i++, mask <<= 1) {
        word[i] = (cw & mask) ? 1 : 0;
    }
    return 0;
}

/* hamming_10_6_3_decode (src/fec/hamming_10_6_3.cpp:14-105): 0 clean, 1 corrected (data bits only), 2 uncorrectable or
 * non-binary input. */
int
oracle_hamming_10_6_3_decode(uint8_t* data6, const uint8_t* parity4) {
    static const int bad_bit[16] = {-2, 0, 1, 5, 2, -1, -1, 6, 3, -1, -1, 7, 4, 8, 9, -1};
    unsigned v = 0;
    for (int i = 0; i < 6; i++) {
        if (data6[i] > 1) {
            return 2;
        }
        v = (v << 1) | data6[i];
    }
    for (int i = 0; i < 4; i++) {
        if (parity4[i] > 1) {
            return 2;
        }
        v = (v << 1) | parity4[i];
    }
    const int syn = (__builtin_parity(v & 0x398u) << 3) | (__builtin_parity(v & 0x354u) << 2) | (__builtin_parity(v & 0x2E2u) << 1)
                    | __builtin_parity(v & 0x1E1u);
    if (!syn) {
        return 0;
    }
    const int b = bad_bit[syn];
    if (b < 0) {
        return 2;
    }
    if (b >= 4) {
        v ^= 1u << b;
    }
    for (int i = 0; i < 6; i++) {
        data6[i] = (uint8_t)((v >> (9 - i)) & 1u);
    }
    return 1;
}

/* BCH_63_16_11::decode_with_result (include/dsd-neo/fec/BCH_63_16.hpp:288-329): input bit i is coefficient 62 - i,
 * 22 syndromes over GF(64), Berlekamp iteration in Rockliff's index form, Chien search; the 16 data bits are coefficients
 * 62..47.  Returns 1 success / 0 failure; *error_count = corrected bits (0 on failure, output untouched). */
int
oracle_bch_63_16_decode(const uint8_t* in63, uint8_t* out16, int* error_count) {
    enum { NN = 63, TT = 11, N2T = 22 };
    gf_init();
    int recd[NN], s[N2T + 1];
    *error_count = 0;
    for (int i = 0; i < NN; i++) {
        recd[i] = in63[NN - 1 - i] ? 1 : 0;
    }
    int has_err = 0;
    for (int i = 1; i <= N2T; i++) {
        int syn = 0;
        for (int j = 0; j < NN; j++) {
            if (recd[j]) {
                syn ^= gf_exp[(i * j) % NN];
            }
        }
        has_err |= syn != 0;
        s[i] = gf_log[syn];
    }
    if (has_err) {
        int elp[N2T + 2][N2T], d[N2T + 2], l[N2T + 2], u_lu[N2T + 2];
        d[0] = 0;
        d[1] = s[1];
        elp[0][0] = 0;
        elp[1][0] = 1;
        for (int i = 1; i < N2T; i++) {
            elp[0][i] = -1;
            elp[1][i] = 0;
        }
        l[0] = l[1] = 0;
        u_lu[0] = -1;
        u_lu[1] = 0;
        int u = 0;
        do {
            u++;
            if (d[u] == -1) {
                l[u + 1] = l[u];
                for (int i = 0; i <= l[u]; i++) {
                    elp[u + 1][i] = elp[u][i];
                }
                for (int i = 0; i <= l[u]; i++) {
                    if (elp[u][i] >= 0) {
                        elp[u][i] = gf_log[elp[u][i]];
                    }
                }
            } else {
                int q = u - 1;
                while (q > 0 && d[q] == -1) {
                    q--;
                }
                if (q > 0) {
                    for (int j = q - 1; j > 0; j--) {
                        if (d[j] != -1 && u_lu[q] < u_lu[j]) {
                            q = j;
                        }
                    }
                }
                const int cand = l[q] + u - q;
                l[u + 1] = l[u] > cand ? l[u] : cand;
                for (int i = 0; i < N2T; i++) {
                    elp[u + 1][i] = 0;
                }
                for (int i = 0; i <= l[q]; i++) {
                    if (elp[q][i] != -1) {
                        elp[u + 1][i + u - q] = gf_exp[(d[u] + NN - d[q] + elp[q][i]) % NN];
                    }
                }
                for (int i = 0; i <= l[u]; i++) {
                    elp[u + 1][i] ^= elp[u][i];
                }
                for (int i = 0; i <= l[u]; i++) {
                    if (elp[u][i] >= 0) {
                        elp[u][i] = gf_log[elp[u][i]];
                    }
                }
            }
            u_lu[u + 1] = u - l[u + 1];
            if (u < N2T) {
                int disc = (s[u + 1] != -1) ? gf_exp[s[u + 1]] : 0;
                for (int i = 1; i <= l[u + 1]; i++) {
                    if (s[u + 1 - i] != -1 && elp[u + 1][i] != 0) {
                        disc ^= gf_exp[(s[u + 1 - i] + gf_log[elp[u + 1][i]]) % NN];
                    }
                }
                d[u + 1] = gf_log[disc];
            }
        } while (u < N2T && l[u + 1] <= TT);
        u++;
        if (l[u] > TT) {
            return 0;
        }
        for (int i = 0; i <= l[u]; i++) {
            if (elp[u][i] >= 0) {
                elp[u][i] = gf_log[elp[u][i]];
            }
        }
        int reg[TT + 1] = {0}, loc[TT], count = 0;
        for (int i = 1; i <= l[u]; i++) {
            reg[i] = elp[u][i];
        }
        for (int i = 1; i <= NN; i++) {
            int q = 1;
            for (int j = 1; j <= l[u]; j++) {
                if (reg[j] != -1) {
                    reg[j] = (reg[j] + j) % NN;
                    q ^= gf_exp[reg[j]];
                }
            }
            if (q == 0) {
                if (count >= TT) {
                    break;
                }
                loc[count++] = NN - i;
            }
        }
        if (count != l[u]) {
            return 0;
        }
        for (int i = 0; i < count; i++) {
            if (loc[i] >= 0 && loc[i] < NN) {
                recd[loc[i]] ^= 1;
            }
        }
        *error_count = count;
    }
    for (int i = 0; i < 16; i++) {
        out16[i] = (uint8_t)recd[NN - 1 - i];
    }
    return 1;
}

/* ------------------------------------------------------------------ K = 5 soft Viterbi (M17 / YSF) */

/* viterbi_decode (src/core/util/dsd_misc.c:118-143) with viterbi_decode_bit (:191-236), viterbi_chainback (:246-275):
 * costs are uint16 "probability of a 1" (0 = strong 0, 0xFFFF = strong 1); ties (m0 >= m1) take the "1" predecessor;
 * chain-back starts in state 0 and writes bit `len/2 + 3 - step`; only the first (len/2-1)/8+1 output bytes are cleared,
 * later bytes are OR-ed into.  Returns the minimum final path metric. */
uint32_t
oracle_viterbi_k5_decode(uint8_t* out, const uint16_t* in, int len) {
    static const uint16_t C0[8] = {0, 0, 0, 0, 0xFFFF, 0xFFFF, 0xFFFF, 0xFFFF};
    static const uint16_t C1[8] = {0, 0xFFFF, 0xFFFF, 0, 0, 0xFFFF, 0xFFFF, 0};
    uint32_t pm[16] = {0}, cm[16];
    static uint16_t hist[244];
    memset(hist, 0, sizeof(hist));
    int pos = 0;
    for (int i = 0; i + 1 < len; i += 2, pos++) {
        uint16_t s0 = in[i], s1 = in[i + 1];
        uint16_t h = 0;
        for (int k = 0; k < 8; k++) {
            uint32_t d0 = C0[k] > s0 ? (uint32_t)(C0[k] - s0) : (uint32_t)(s0 - C0[k]);
            uint32_t d1 = C1[k] > s1 ? (uint32_t)(C1[k] - s1) : (uint32_t)(s1 - C1[k]);
            uint32_t metric = (uint16_t)d0 + (uint32_t)(uint16_t)d1;
            uint32_t m0 = pm[k] + metric, m1 = pm[k + 8] + (0x1FFFE - metric);
            uint32_t m2 = pm[k] + (0x1FFFE - metric), m3 = pm[k + 8] + metric;
            if (m0 >= m1) {
                h |= (uint16_t)(1u << (2 * k));
                cm[2 * k] = m1;
            } else {
                cm[2 * k] = m0;
            }
            if (m2 >= m3) {
                h |= (uint16_t)(1u << (2 * k + 1));
                cm[2 * k + 1] = m3;
            } else {
                cm[2 * k + 1] = m2;
            }
        }
        hist[pos] = h;
        memcpy(pm, cm, sizeof(pm));
    }
    int nbits = len / 2;
    uint8_t state = 0;
    int bit_pos = nbits + 4;
    memset(out, 0, (size_t)((nbits - 1) / 8 + 1));
    while (pos > 0) {
        bit_pos--;
        pos--;
        uint16_t bit = hist[pos] & (uint16_t)(1u << (state >> 4));
        state >>= 1;
        if (bit) {
            state |= 0x80;
            out[bit_pos / 8] |= (uint8_t)(1u << (7 - (bit_pos % 8)));
        }
    }
    uint32_t best = pm[0];
    for (int i = 1; i < 16; i++) {
        if (pm[i] < best) {
            best = pm[i];
        }
    }
    return best;
}

/* viterbi_decode_punctured (dsd_misc.c:156-182) */
uint32_t
oracle_viterbi_k5_decode_punctured(uint8_t* out, const uint16_t* in, const uint8_t* punct, int in_len, int p_len) {
    uint16_t umsg[488];
    memset(umsg, 0, sizeof(umsg));
    int p = 0, u = 0, i = 0;
    while (i < in_len) {
        if (punct[p]) {
            umsg[u] = in[i++];
        } else {
            umsg[u] = 0x7FFF;
        }
        u++;
        p = (p + 1) % p_len;
    }
    return oracle_viterbi_k5_decode(out, umsg, u) - (uint32_t)(u - in_len) * 0x7FFFu;
}

/* ------------------------------------------------------------------ NXDN K = 5 convolution */

/* CNXDNConvolution_start + decode / decode_soft per symbol pair + chainback (src/protocol/nxdn/nxdn_convolution.c:58-164).
 * metrics_io (2 x 16 uint16) are the decoder's two ping-pong metric arrays m_metrics1 / m_metrics2: the reference zeroes
 * them only once at start-up (CNXDNConvolution_init, engine.c:2808); CNXDNConvolution_start() merely points "old" back at
 * m_metrics1, so a frame starts from whatever the last ODD step of the previous frame left there (stale by one step when
 * the previous frame had an odd number of steps).  rel == NULL selects the hard decoder.  Returns nothing; out receives n_bits_out bits (MSB first), taken from the LAST n_bits_out steps. */
void
oracle_nxdn_conv_decode(const uint8_t* sym /*[2*n_steps]*/, const uint8_t* rel /*[2*n_steps] or NULL*/, int n_steps,
                        int n_bits_out, uint16_t* metrics_io, uint8_t* out) {
    static const uint8_t T1[8] = {0, 0, 0, 0, 2, 2, 2, 2}, T2[8] = {0, 2, 2, 0, 0, 2, 2, 0};
    static uint64_t dec[300];
    uint16_t* om = metrics_io;      /* m_metrics1 */
    uint16_t* nm = metrics_io + 16; /* m_metrics2 */
    for (int t = 0; t < n_steps; t++) {
        uint8_t s0 = sym[2 * t], s1 = sym[2 * t + 1];
        uint64_t d = 0;
        for (int i = 0; i < 8; i++) {
            int d0 = (int)T1[i] - (int)s0, d1 = (int)T2[i] - (int)s1;
            d0 = d0 < 0 ? -d0 : d0;
            d1 = d1 < 0 ? -d1 : d1;
            uint32_t m0, m1, m2, m3;
            if (!rel) {
                uint16_t metric = (uint16_t)(d0 + d1);
                m0 = (uint16_t)(om[i] + metric);
                m1 = (uint16_t)(om[i + 8] + (4u - metric));
                m2 = (uint16_t)(om[i] + (4u - metric));
                m3 = (uint16_t)(om[i + 8] + metric);
            } else {
                uint32_t metric = ((uint32_t)d0 * rel[2 * t] + (uint32_t)d1 * rel[2 * t + 1]) / 128u;
                if (metric > 8u) {
                    metric = 8u;
                }
                m0 = om[i] + metric;
                m1 = om[i + 8] + (8u - metric);
                m2 = om[i] + (8u - metric);
                m3 = om[i + 8] + metric;
            }
            unsigned dec0 = m0 >= m1, dec1 = m2 >= m3;
            nm[2 * i] = (uint16_t)(dec0 ? m1 : m0);
            nm[2 * i + 1] = (uint16_t)(dec1 ? m3 : m2);
            d |= ((uint64_t)dec1 << (2 * i + 1)) | ((uint64_t)dec0 << (2 * i));
        }
        dec[t] = d;
        uint16_t* tmp = om;
        om = nm;
        nm = tmp;
    }
    uint32_t state = 0;
    int t = n_steps;
    int nb = n_bits_out;
    while (nb-- > 0) {
        --t;
        uint32_t i = state >> 4;
        uint8_t bit = (uint8_t)((dec[t] >> i) & 1u);
        state = ((uint32_t)bit << 7) | (state >> 1);
        if (bit) {
            out[nb >> 3] |= (uint8_t)(0x80u >> (nb & 7));
        } else {
            out[nb >> 3] &= (uint8_t)~(0x80u >> (nb & 7));
        }
    }
}

/* ------------------------------------------------------------------ P25 Phase 1 NID decode (hard + NAC retry + Chase search) */

/* One 63-bit NID candidate through decode_nid_codeword (src/protocol/p25/phase1/p25p1_check_nid.cpp:250-303): BCH(63,16,11)
 * correction, DUID membership in TIA-102.BAAA-A Table 8-4, final parity bit (1 for LDU1/LDU2 only).  status: 0 = fail, 1 = ok,
 * 2 = parity override.  *bch_failed tells a BCH failure from an invalid DUID. */
static int
nid_codeword(const uint8_t* code63, int parity, int* nac, int* duid, int* errs, int* bch_failed) {
    static const uint8_t duid_valid[16] = {1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 1, 0, 1, 0, 0, 1};
    uint8_t dec[16];
    int count = 0;
    *nac = 0;
    *duid = 0;
    *errs = 0;
    *bch_failed = 0;
    if (!oracle_bch_63_16_decode(code63, dec, &count)) {
        *bch_failed = 1;
        return 0;
    }
    *errs = count;
    for (int i = 0; i < 12; i++) {
        *nac = (*nac << 1) | dec[i];
    }
    *duid = (dec[12] << 3) | (dec[13] << 2) | (dec[14] << 1) | dec[15];
    if (!duid_valid[*duid]) {
        *errs = 0;
        return 0;
    }
    const int want_parity = (*duid == 0x5 || *duid == 0xA) ? 1 : 0;
    return (want_parity == parity) ? 1 : 2;
}

static int
nid_nac_of(const uint8_t* code63) {
    int nac = 0;
    for (int i = 0; i < 12; i++) {
        nac = (nac << 1) | (code63[i] ? 1 : 0);
    }
    return nac;
}

/* p25p1_nid_decode (p25p1_check_nid.cpp:322-354).  reliab63 may be NULL (hard decode only).  `threshold` is
 * p25p1_get_erasure_threshold() (64 unless configured, p25p1_soft.cpp:20-39).  Returns the status; nac / duid / errs as
 * in struct p25p1_nid_result. */
int
oracle_p25p1_nid_decode(const uint8_t* code63, const uint8_t* reliab63, int observed_nac, int parity, int parity_reliab,
                        int threshold, int* nac, int* duid, int* errs) {
    int failed = 0;
    const int nac_ok = observed_nac > 0 && observed_nac <= 0xFFF && observed_nac != 0xFFF;
    const int rx_nac = nid_nac_of(code63);
    uint8_t retry[63];
    memcpy(retry, code63, 63);
    for (int i = 0; i < 12; i++) {
        retry[i] = (uint8_t)((observed_nac >> (11 - i)) & 1);
    }
    /* decode_nid_hard (:305-320): one retry with the known NAC written over the received one, after a BCH failure only */
    int status = nid_codeword(code63, parity, nac, duid, errs, &failed);
    if (status == 0 && failed && nac_ok && rx_nac != observed_nac) {
        status = nid_codeword(retry, parity, nac, duid, errs, &failed);
    }
    if (status > 0 || !reliab63) {
        return status;
    }
    /* build_soft_nid_pool (:123-153): positions by (reliability, index); those under the threshold first (at most 8), then
     * filled up to 6 with the next weakest */
    int order[63], pool[8], n_pool = 0;
    for (int i = 0; i < 63; i++) {
        order[i] = i;
    }
    for (int i = 0; i < 63; i++) {
        for (int j = i + 1; j < 63; j++) {
            const int a = order[j], b = order[i];
            const int cmp = (reliab63[a] != reliab63[b]) ? (int)reliab63[a] - (int)reliab63[b] : a - b;
            if (cmp < 0) {
                order[i] = a;
                order[j] = b;
            }
        }
    }
    uint8_t picked[63] = {0};
    for (int i = 0; i < 63 && n_pool < 8; i++) {
        if ((int)reliab63[order[i]] < threshold) {
            pool[n_pool++] = order[i];
            picked[i] = 1;
        }
    }
    for (int i = 0; i < 63 && n_pool < 6; i++) {
        if (!picked[i]) {
            pool[n_pool++] = order[i];
        }
    }
    if (n_pool <= 0) {
        return status;
    }
    /* soft_nid_search_from_base (:200-228) from the received word, then from the NAC-rewritten word; candidates of at
     * most 3 flips whose summed reliability stays within threshold x flips; best = lowest score (+ parity reliability on
     * a parity override), then status ok, then fewer BCH corrections, then fewer flips, then first found */
    int found = 0, b_status = 0, b_nac = 0, b_duid = 0, b_errs = 0, b_score = 0, b_changes = 0;
    for (int base = 0; base < 2; base++) {
        if (base == 1 && !(nac_ok && rx_nac != observed_nac)) {
            break;
        }
        const uint8_t* from = base ? retry : code63;
        for (int mask = 0; mask < (1 << n_pool); mask++) {
            int weight = 0, score = 0;
            uint8_t cand[63];
            memcpy(cand, from, 63);
            for (int b = 0; b < n_pool; b++) {
                if (mask & (1 << b)) {
                    weight++;
                    cand[pool[b]] ^= 1;
                    score += reliab63[pool[b]];
                }
            }
            if (weight > 3 || (weight > 0 && score > threshold * weight)) {
                continue;
            }
            int c_nac, c_duid, c_errs, c_failed;
            const int c_status = nid_codeword(cand, parity, &c_nac, &c_duid, &c_errs, &c_failed);
            if (c_status <= 0) {
                continue;
            }
            if (c_status == 2) {
                score += parity_reliab;
            }
            if (!found || score < b_score || (score == b_score && c_status == 1 && b_status != 1)
                || (score == b_score && c_status == b_status && c_errs < b_errs)
                || (score == b_score && c_status == b_status && c_errs == b_errs && weight < b_changes)) {
                found = 1;
                b_status = c_status;
                b_nac = c_nac;
                b_duid = c_duid;
                b_errs = c_errs;
                b_score = score;
                b_changes = weight;
            }
        }
    }
    if (!found) {
        return status;
    }
    *nac = b_nac;
    *duid = b_duid;
    *errs = b_errs;
    return b_status;
}

/* ------------------------------------------------------------------ P25 Phase 1 frame cutter (sequential restatement) */

/* Reads a frame the way the reference does, one dibit at a time after the sync:
 *   NID   p25p1_read_nid_fields (src/engine/dispatch/dispatch_p25p1.c:121-143): 6 NAC dibits, 2 DUID dibits, 3 dibits, one
 *         status symbol, 20 dibits, one last dibit whose low bit is the parity bit; bit reliabilities min(|llr|, 255) (:59-66)
 *   data  tsbk_read_repetition_samples (src/protocol/p25/phase1/p25p1_tsbk.c:135-152) with skipdibit = 36 - 14 (:1054): a
 *         dibit is payload while skipdibit / 36 == 0, else it is a status symbol and the counter restarts.
 * PARITY UNPINNED for this function alone: both reference routines are static functions of translation units that need
 * the whole decoder state (getDibitSoft), so they are restated here and checked by round trip instead -- frames built
 * with the reference's own NID generator rule (tests/test_support/p25_nid_generator.hpp) and trellis table come back with
 * their NAC / DUID / TSBK bytes through this cutter + the pinned decoders (tests/test_oracle_fec.py).
 * `pos_last_sync` = index of the last sync dibit.  Returns bit 0: NID complete, bit 1: payload complete. */
int
oracle_p25p1_frame_cut(const uint8_t* dibits, const int16_t* llr /* [n][2] */, int count, int pos_last_sync, int n_payload,
                       uint8_t* code63, uint8_t* reliab63, uint8_t* parity, uint8_t* parity_reliab, uint8_t* payload_dibits,
                       int16_t* payload_llr) {
    int p = pos_last_sync + 1; /* next dibit to read */
    int flags = 0;
    memset(code63, 0, 63);
    memset(reliab63, 0, 63);
    *parity = 0;
    *parity_reliab = 0;
    if (n_payload > 0) {
        memset(payload_dibits, 0, (size_t)n_payload);
        memset(payload_llr, 0, (size_t)n_payload * 2 * sizeof(int16_t));
    }
    if (pos_last_sync - 23 < 0 || p + 33 > count) {
        return 0;
    }
    int idx = 0;
    for (int k = 0; k < 33; k++) {
        const int dib = dibits[p], l0 = llr[2 * p], l1 = llr[2 * p + 1];
        p++;
        if (k == 11) {
            continue; /* status symbol between the 11th and 12th NID dibit */
        }
        const int r0 = (l0 < 0 ? -l0 : l0) > 255 ? 255 : (l0 < 0 ? -l0 : l0);
        const int r1 = (l1 < 0 ? -l1 : l1) > 255 ? 255 : (l1 < 0 ? -l1 : l1);
        code63[idx] = (uint8_t)((dib >> 1) & 1);
        reliab63[idx] = (uint8_t)r0;
        idx++;
        if (idx < 63) {
            code63[idx] = (uint8_t)(dib & 1);
            reliab63[idx] = (uint8_t)r1;
            idx++;
        } else {
            *parity = (uint8_t)(dib & 1);
            *parity_reliab = (uint8_t)r1;
        }
    }
    flags |= 1;
    /* payload: does it fit?  walk with the reference's counter */
    int skip = 36 - 14, k = 0, q = p;
    while (k < n_payload && q < count) {
        if (skip / 36 == 0) {
            k++;
        } else {
            skip = 0;
        }
        skip++;
        q++;
    }
    if (n_payload <= 0 || k < n_payload) {
        return flags;
    }
    skip = 36 - 14;
    k = 0;
    while (k < n_payload) {
        if (skip / 36 == 0) {
            payload_dibits[k] = dibits[p];
            payload_llr[2 * k] = llr[2 * p];
            payload_llr[2 * k + 1] = llr[2 * p + 1];
            k++;
        } else {
            skip = 0;
        }
        skip++;
        p++;
    }
    return flags | 2;
}

/* ------------------------------------------------------------------ Hamming(10,6,3) soft decode (P25p1 LDU hex words) */

/* hamming_10_6_3_soft (src/protocol/p25/phase1/p25p1_soft.cpp:444-475): the hard decode seeds the best candidate; then the 32
 * flip masks over the 5 least reliable bits (find_k_least_reliable, :174-206: (reliability, index) order, those under the
 * erasure threshold first) with at most 2 flips are tried, a candidate counts only if it is a clean codeword; lowest summed
 * reliability of changed bits wins, then fewer flips; a hard single-bit correction is kept unless overriding is enabled and
 * the soft winner is more than 8 cheaper.  Returns 0 unchanged / 1 corrected / 2 failed (out = in). */
int
oracle_hamming_10_6_3_soft(const uint8_t* bits10, const int* reliab10, int hard_override_enabled, int threshold, uint8_t* out10) {
    int rel[10];
    for (int i = 0; i < 10; i++) {
        rel[i] = reliab10[i] < 0 ? 0 : (reliab10[i] > 255 ? 255 : reliab10[i]);
    }
    int best_pen = 999999, best_flips = 99, found = 0, hard_valid = 0, hard_corrected = 0, hard_pen = 999999;
    uint8_t best[10] = {0}, hard[10] = {0};
    {
        uint8_t d[6], p[4];
        memcpy(d, bits10, 6);
        memcpy(p, bits10 + 6, 4);
        const int rc = oracle_hamming_10_6_3_decode(d, p);
        if (rc == 0 || rc == 1) {
            memcpy(hard, d, 6);
            hard[6] = d[0] ^ d[1] ^ d[2] ^ d[5];
            hard[7] = d[0] ^ d[1] ^ d[3] ^ d[5];
            hard[8] = d[0] ^ d[2] ^ d[3] ^ d[4];
            hard[9] = d[1] ^ d[2] ^ d[3] ^ d[4];
            hard_valid = 1;
            hard_corrected = rc == 1;
            hard_pen = 0;
            int flips = 0;
            for (int i = 0; i < 10; i++) {
                if (hard[i] != bits10[i]) {
                    hard_pen += rel[i];
                    flips++;
                }
            }
            best_pen = hard_pen;
            best_flips = flips;
            memcpy(best, hard, 10);
            found = 1;
        }
    }
    int order[10], least[5], n_least = 0;
    for (int i = 0; i < 10; i++) {
        order[i] = i;
    }
    for (int i = 0; i < 10; i++) {
        for (int j = i + 1; j < 10; j++) {
            const int a = order[j], b = order[i];
            if (rel[a] < rel[b] || (rel[a] == rel[b] && a < b)) {
                order[i] = a;
                order[j] = b;
            }
        }
    }
    for (int i = 0; i < 10 && n_least < 5; i++) {
        if (rel[order[i]] < threshold) {
            least[n_least++] = order[i];
        }
    }
    for (int i = 0; i < 10 && n_least < 5; i++) {
        if (rel[order[i]] >= threshold) {
            least[n_least++] = order[i];
        }
    }
    for (int mask = 0; mask < 32; mask++) {
        uint8_t cand[10];
        memcpy(cand, bits10, 10);
        int flips = 0;
        for (int b = 0; b < 5; b++) {
            if (mask & (1 << b)) {
                cand[least[b]] ^= 1;
                flips++;
            }
        }
        if (flips > 2) {
            continue;
        }
        uint8_t d[6];
        memcpy(d, cand, 6);
        if (oracle_hamming_10_6_3_decode(d, cand + 6) != 0) {
            continue;
        }
        int pen = 0;
        for (int i = 0; i < 10; i++) {
            if (cand[i] != bits10[i]) {
                pen += rel[i];
            }
        }
        if (pen < best_pen || (pen == best_pen && flips < best_flips)) {
            best_pen = pen;
            best_flips = flips;
            memcpy(best, cand, 10);
            found = 1;
        }
    }
    if (!found) {
        memcpy(out10, bits10, 10);
        return 2;
    }
    if (hard_valid && hard_corrected && memcmp(best, hard, 10) != 0) {
        if (!hard_override_enabled || best_pen + 8 >= hard_pen) {
            memcpy(out10, hard, 10);
            return 1;
        }
    }
    memcpy(out10, best, 10);
    return memcmp(best, bits10, 10) == 0 ? 0 : 1;
}

/* ------------------------------------------------------------------ Golay(24,6) / (24,12) soft decode (P25p1 HDU / LDU words) */

/* DSDGolay24::encode_6 / encode_12 (include/dsd-neo/fec/Golay24.hpp:407-434): 12 parity bits for `length` data bits */
static void
golay24_encode_parity(int length, const uint8_t* word, uint8_t* parity12) {
    unsigned data = 0;
    for (int i = 0; i < length; i++) {
        data = (data << 1) | word[length - 1 - i];
    }
    data <<= (12 - length);
    unsigned cw = g23_syndrome(data) | data; /* Golay24::golay (:22-37) */
    if (parity32(cw)) {
        cw ^= 0x800000u;
    }
    unsigned mask = 1u << 12;
    for (int i = 0; i < 12; i++, mask <<= 1) {
        parity12[i] = (cw & mask) ? 1 : 0;
    }
}

/* check_and_fix_golay_24_6_soft / _24_12_soft (src/protocol/p25/phase1/p25p1_soft.cpp:477-593): hard decode seeds the best
 * candidate; the 8 least reliable of the 6 + 12 / 12 + 12 bits (find_k_least_reliable, :174-206) are flipped in every
 * combination of at most 4, each candidate is Golay-decoded and re-encoded, lowest summed reliability of changed bits
 * wins, then fewer changed bits; a hard correction is kept unless overriding is enabled and the soft winner is more than
 * 8 cheaper.  Returns 0 ok (data corrected in place, *fixed = changed bits or the hard count) / 1 no valid candidate. */
int
oracle_p25_golay24_soft(int length, uint8_t* data, const uint8_t* parity, const int* reliab, int hard_override_enabled,
                        int threshold, int* fixed) {
    *fixed = 0;
    if (length != 6 && length != 12) {
        return 1;
    }
    const int n = length + 12;
    uint8_t orig[24], best[12] = {0}, hard[12] = {0};
    int rel[24];
    memcpy(orig, data, (size_t)length);
    memcpy(orig + length, parity, 12);
    for (int i = 0; i < n; i++) {
        rel[i] = reliab[i] < 0 ? 0 : (reliab[i] > 255 ? 255 : reliab[i]);
    }
    int best_pen = 999999, best_fixed = 0, found = 0, hard_valid = 0, hard_corrected = 0, hard_pen = 999999, hard_fixed = 0;
    {
        uint8_t w[12];
        memcpy(w, data, (size_t)length);
        if (oracle_p25_golay24_decode(length, w, parity, &hard_fixed) == 0) {
            uint8_t dec[24];
            memcpy(dec, w, (size_t)length);
            golay24_encode_parity(length, w, dec + length);
            hard_valid = 1;
            hard_corrected = hard_fixed > 0;
            hard_pen = 0;
            int diff = 0;
            for (int i = 0; i < n; i++) {
                if (dec[i] != orig[i]) {
                    hard_pen += rel[i];
                    diff++;
                }
            }
            best_pen = hard_pen;
            best_fixed = diff;
            memcpy(hard, w, (size_t)length);
            memcpy(best, w, (size_t)length);
            found = 1;
        }
    }
    int order[24], least[8], n_least = 0;
    for (int i = 0; i < n; i++) {
        order[i] = i;
    }
    for (int i = 0; i < n; i++) {
        for (int j = i + 1; j < n; j++) {
            const int a = order[j], b = order[i];
            if (rel[a] < rel[b] || (rel[a] == rel[b] && a < b)) {
                order[i] = a;
                order[j] = b;
            }
        }
    }
    for (int i = 0; i < n && n_least < 8; i++) {
        if (rel[order[i]] < threshold) {
            least[n_least++] = order[i];
        }
    }
    for (int i = 0; i < n && n_least < 8; i++) {
        if (rel[order[i]] >= threshold) {
            least[n_least++] = order[i];
        }
    }
    for (int mask = 0; mask < 256; mask++) {
        if (__builtin_popcount((unsigned)mask) > 4) {
            continue;
        }
        uint8_t cand[24];
        memcpy(cand, orig, (size_t)n);
        for (int b = 0; b < 8; b++) {
            if (mask & (1 << b)) {
                cand[least[b]] ^= 1;
            }
        }
        uint8_t w[12];
        int cf = 0;
        memcpy(w, cand, (size_t)length);
        if (oracle_p25_golay24_decode(length, w, cand + length, &cf) != 0) {
            continue;
        }
        uint8_t dec[24];
        memcpy(dec, w, (size_t)length);
        golay24_encode_parity(length, w, dec + length);
        int pen = 0, diff = 0;
        for (int i = 0; i < n; i++) {
            if (dec[i] != orig[i]) {
                pen += rel[i];
                diff++;
            }
        }
        if (pen < best_pen || (pen == best_pen && diff < best_fixed)) {
            best_pen = pen;
            best_fixed = diff;
            memcpy(best, w, (size_t)length);
            found = 1;
        }
    }
    if (!found) {
        return 1;
    }
    if (hard_valid && hard_corrected && memcmp(best, hard, (size_t)length) != 0
        && (!hard_override_enabled || best_pen + 8 >= hard_pen)) {
        memcpy(data, hard, (size_t)length);
        *fixed = hard_fixed;
        return 0;
    }
    memcpy(data, best, (size_t)length);
    *fixed = best_fixed;
    return 0;
}

/* ------------------------------------------------------------------ DMR BS data burst cutter (sequential restatement) */

/* Collects one DMR base-station data burst around a BS DATA sync the way dmr_data_sync does (src/protocol/dmr/dmr_data.c:
 * 54-65,118-157,159-179,218-226,261-268,321-340): 90 dibits back from the dibit after the sync -- 12 CACH dibits
 * (de-interleaved with dmr_cach_interleave, src/protocol/dmr/dmr_cach.c:9-11), 49 info dibits, 5 slot-type dibits, the
 * 24 sync dibits -- then, after the sync, 5 slot-type dibits and 49 info dibits.  `inverted` = opts->inverted_dmr (XOR 2
 * on the cached part only, :87-89).  rel98 = per-dibit reliabilities of the 98 info dibits.
 * PARITY UNPINNED for this function alone (static functions that need the whole decoder state); checked by round trip:
 * bursts built from BPTC(196,96) / Golay(20,8) / Hamming(7,4) codewords come back through the pinned decoders.
 * Returns 1 if the stream holds the whole burst. */
int
oracle_dmr_burst_cut(const uint8_t* dibits, const uint8_t* reliab, int count, int pos_last_sync, int inverted, uint8_t* cach24,
                     uint8_t* info196, uint8_t* rel98, uint8_t* slot20) {
    static const uint8_t cach_interleave[24] = {0, 7, 8, 9, 1, 10, 11, 12, 2, 13, 14, 15, 3, 16, 4, 17, 18, 19, 5, 20, 21, 22, 6, 23};
    memset(cach24, 0, 24);
    memset(info196, 0, 196);
    memset(rel98, 0, 98);
    memset(slot20, 0, 20);
    const int live = pos_last_sync + 1; /* state->dmr_payload_p when dmr_data_sync starts */
    int p = live - 90;
    if (p < 0 || live + 54 > count) {
        return 0;
    }
    for (int i = 0; i < 12; i++, p++) {
        const int d = dibits[p] ^ (inverted ? 2 : 0);
        cach24[cach_interleave[2 * i]] = (uint8_t)((d >> 1) & 1);
        cach24[cach_interleave[2 * i + 1]] = (uint8_t)(d & 1);
    }
    for (int i = 0; i < 49; i++, p++) {
        const int d = dibits[p] ^ (inverted ? 2 : 0);
        info196[2 * i] = (uint8_t)((d >> 1) & 1);
        info196[2 * i + 1] = (uint8_t)(d & 1);
        rel98[i] = reliab[p];
    }
    for (int i = 0; i < 5; i++, p++) {
        const int d = dibits[p] ^ (inverted ? 2 : 0);
        slot20[2 * i] = (uint8_t)((d >> 1) & 1);
        slot20[2 * i + 1] = (uint8_t)(d & 1);
    }
    p += 24; /* the sync itself */
    for (int i = 0; i < 5; i++, p++) {
        const int d = dibits[p];
        slot20[2 * i + 10] = (uint8_t)((d >> 1) & 1);
        slot20[2 * i + 11] = (uint8_t)(d & 1);
    }
    for (int i = 0; i < 49; i++, p++) {
        const int d = dibits[p];
        info196[2 * i + 98] = (uint8_t)((d >> 1) & 1);
        info196[2 * i + 99] = (uint8_t)(d & 1);
        rel98[i + 49] = reliab[p];
    }
    return 1;
}

/* ---- DMR rate 3/4 trellis: dmr_r34_viterbi_decode / _decode_soft (src/protocol/dmr/dmr_34_viterbi.c:402-474) -----------------
 * 98 dibits -> de-interleave (dsd_trellis_interleave_98: received i lands at table[i], the table p25_interleave_98 above
 * generates) -> 49 four-bit symbols (first dibit high) -> 8-state Viterbi over the ETSI TS 102 361-1 B.2.5 encoder (state =
 * previous tribit, output point = fsm[prev * 8 + next]), start state 0, traceback from end state 0, the first 48 states packed
 * three bits each into 18 bytes.  Hard: Hamming distance between constellation POINT indices (:213-236); soft: the expected
 * point mapped back to its dibit pair and each differing bit charged the reliability of its dibit (:180-198, 238-263).
 * Survivor choice: strict '<' with the previous state ascending (the lowest previous state wins ties).
 * Pinned by tests/test_oracle_fec.py against the reference vectors (tests/protocol/dmr/dmr_r34_reference_vectors.h) and the
 * compiled reference on random noisy inputs. */
static const uint8_t r34_point_of_nibble[16] = {11, 12, 0, 7, 14, 9, 5, 2, 10, 13, 1, 6, 15, 8, 4, 3}; /* B.2.5 constellation */
static const uint8_t r34_fsm[64] = {0, 8,  4, 12, 2, 10, 6, 14, 4, 12, 2, 10, 6, 14, 0, 8, 1, 9,  5, 13, 3, 11,
                                    7, 15, 5, 13, 3, 11, 7, 15, 1, 9,  3, 11, 7, 15, 1, 9, 5, 13, 7, 15, 1, 9,
                                    5, 13, 3, 11, 2, 10, 6, 14, 0, 8,  4, 12, 6, 14, 0, 8, 4, 12, 2, 10}; /* B.2.5 state table */

int
oracle_dmr_r34_decode(const uint8_t* dibits98, const uint8_t* reliab98, uint8_t* out18) {
    uint8_t t98[98], dei[98], rdei[98], nib_of_point[16];
    p25_interleave_98(t98);
    for (int i = 0; i < 16; i++) {
        nib_of_point[r34_point_of_nibble[i]] = (uint8_t)i;
    }
    for (int i = 0; i < 98; i++) {
        dei[t98[i]] = dibits98[i] & 3;
        rdei[t98[i]] = reliab98 ? reliab98[i] : 0;
    }
    const int INF = 1000000000;
    int prev[8], cur[8];
    uint8_t back[49][8];
    memset(back, 0, sizeof(back));
    for (int s = 0; s < 8; s++) {
        prev[s] = INF;
    }
    prev[0] = 0;
    for (int t = 0; t < 49; t++) {
        const int nib = (dei[2 * t] << 2) | dei[2 * t + 1];
        const int point = r34_point_of_nibble[nib];
        for (int s = 0; s < 8; s++) {
            cur[s] = INF;
        }
        for (int ps = 0; ps < 8; ps++) {
            if (prev[ps] >= INF) {
                continue;
            }
            for (int ns = 0; ns < 8; ns++) {
                const int expect = r34_fsm[ps * 8 + ns];
                int cost;
                if (!reliab98) {
                    cost = __builtin_popcount((unsigned)((expect ^ point) & 15));
                } else {
                    const int x = nib_of_point[expect] ^ nib;
                    cost = ((x >> 3) & 1) * rdei[2 * t] + ((x >> 2) & 1) * rdei[2 * t] + ((x >> 1) & 1) * rdei[2 * t + 1]
                           + (x & 1) * rdei[2 * t + 1];
                }
                const int m = prev[ps] + cost;
                if (m < cur[ns]) {
                    cur[ns] = m;
                    back[t][ns] = (uint8_t)ps;
                }
            }
        }
        memcpy(prev, cur, sizeof(prev));
    }
    uint8_t states[49];
    int s = 0;
    for (int t = 48; t >= 0; t--) {
        states[t] = (uint8_t)s;
        s = back[t][s];
    }
    for (int g = 0; g < 6; g++) {
        uint32_t v = 0;
        for (int k = 0; k < 8; k++) {
            v = (v << 3) | (uint32_t)(states[g * 8 + k] & 7);
        }
        out18[3 * g] = (uint8_t)(v >> 16), out18[3 * g + 1] = (uint8_t)(v >> 8), out18[3 * g + 2] = (uint8_t)v;
    }
    return 0;
}

/* ---- RS(12,9) over GF(2^8): rs_12_9_calc_syndrome / _check_syndrome / _correct_errors (src/fec/rs-12-9.c:237-323) -----------
 * Field: x^8 + x^4 + x^3 + x^2 + 1, alpha = 2; exp[255] = 1, log[0] = 0, so the reference's inverse of 0 evaluates to 1.
 * Syndromes S_j = Horner of the 12 bytes at alpha^(j+1), j = 0..2 (first byte = highest power).  The corrector runs
 * Berlekamp-Massey over the three syndromes, finds the locator's roots by trying r = 1..255 (location = 255 - r), and applies
 * Forney with the evaluator = (locator x syndrome) mod z^3; any location >= 12 => cannot correct; no roots => "no errors
 * found" (result 0) even though the syndrome was not zero.
 * Returns: 0 syndrome zero (nothing done), else 1 + the reference's result code (1 no roots, 2 corrected, 3 cannot correct);
 * *errors_found as the reference reports it.  Pinned by tests/test_oracle_fec.py (reference codeword test_fec_bptc_rs.c:194 and
 * the compiled reference on random corruptions). */
static uint8_t rs129_exp[256], rs129_log[256];
static int rs129_ready;

static void
rs129_init(void) {
    if (rs129_ready) {
        return;
    }
    int v = 1;
    memset(rs129_log, 0, sizeof(rs129_log));
    for (int i = 0; i < 255; i++) {
        rs129_exp[i] = (uint8_t)v;
        rs129_log[v] = (uint8_t)i;
        v <<= 1;
        if (v & 0x100) {
            v ^= 0x11D;
        }
    }
    rs129_exp[255] = 1;
    rs129_ready = 1;
}

static uint8_t
rs129_mul(uint8_t a, uint8_t b) {
    return (a == 0 || b == 0) ? 0 : rs129_exp[(rs129_log[a] + rs129_log[b]) % 255];
}

static uint8_t
rs129_inv(uint8_t a) {
    return rs129_exp[255 - rs129_log[a]];
}

int
oracle_rs_12_9_decode(uint8_t* cw12, uint8_t* syndrome3, uint8_t* errors_found) {
    rs129_init();
    uint8_t S[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < 12; i++) {
            S[j] = cw12[i] ^ rs129_mul(rs129_exp[j + 1], S[j]);
        }
    }
    if (syndrome3) {
        syndrome3[0] = S[0], syndrome3[1] = S[1], syndrome3[2] = S[2];
    }
    *errors_found = 0;
    if (!(S[0] | S[1] | S[2])) {
        return 0;
    }
    /* Berlekamp-Massey, 3 iterations */
    uint8_t loc[6] = {1, 0, 0, 0, 0, 0}, D[6] = {0, 1, 0, 0, 0, 0}, psi2[6];
    int L = 0, k = -1;
    for (int n = 0; n < 3; n++) {
        uint8_t d = 0;
        for (int i = 0; i <= L; i++) {
            d ^= rs129_mul(loc[i], S[n - i]);
        }
        if (d != 0) {
            for (int i = 0; i < 6; i++) {
                psi2[i] = loc[i] ^ rs129_mul(d, D[i]);
            }
            if (L < n - k) {
                const int L2 = n - k;
                k = n - L;
                const uint8_t di = rs129_inv(d);
                for (int i = 0; i < 6; i++) {
                    D[i] = rs129_mul(loc[i], di);
                }
                L = L2;
            }
            memcpy(loc, psi2, 6);
        }
        for (int i = 5; i > 0; i--) {
            D[i] = D[i - 1];
        }
        D[0] = 0;
    }
    /* evaluator = (locator * syndrome) mod z^3 */
    uint8_t ev[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; i + j < 3; j++) {
            ev[i + j] ^= rs129_mul(S[j], loc[i]);
        }
    }
    /* roots */
    uint8_t locs[256];
    int nroots = 0;
    for (int r = 1; r < 256; r++) {
        uint8_t sum = 0;
        for (int kk = 0; kk < 4; kk++) {
            sum ^= rs129_mul(rs129_exp[(kk * r) % 255], loc[kk]);
        }
        if (sum == 0) {
            locs[nroots++] = (uint8_t)(255 - r);
        }
    }
    *errors_found = (uint8_t)nroots;
    if (nroots == 0) {
        return 1;
    }
    if (nroots > 3) {
        return 3;
    }
    for (int r = 0; r < nroots; r++) {
        if (locs[r] >= 12) {
            return 3;
        }
    }
    for (int r = 0; r < nroots; r++) {
        const int i = locs[r];
        uint8_t num = 0, den = 0;
        for (int j = 0; j < 6; j++) {
            num ^= rs129_mul(ev[j], rs129_exp[((255 - i) * j) % 255]);
        }
        for (int j = 1; j < 6; j += 2) {
            den ^= rs129_mul(loc[j], rs129_exp[((255 - i) * (j - 1)) % 255]);
        }
        cw12[12 - i - 1] ^= rs129_mul(num, rs129_inv(den));
    }
    return 2;
}
