/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's P25 Phase 1 frame handlers over a sliced dibit stream,
 * composed from the (pinned) leaf restatements of oracle_fec.c:
 *
 *   NID read + decode + DUID switch      src/engine/dispatch/dispatch_p25p1.c:121-143,203-223,401-426
 *   processTSBK (1..3 half-rate blocks)  src/protocol/p25/phase1/p25p1_tsbk.c:108-161,1051-1081
 *   processHDU                           src/protocol/p25/phase1/p25p1_hdu.c:54-96,108-210,270-303
 *   processLDU1 / processLDU2            src/protocol/p25/phase1/p25p1_ldu.c:89-222, p25p1_ldu1.c:54-240, p25p1_ldu2.c:54-280
 *   processTDULC                         src/protocol/p25/phase1/p25p1_tdulc.c:75-238,284-300
 *   LSD (16,8) cyclic code               src/protocol/p25/p25_lsd.c:27-161
 *   crc16_lb_bridge                      src/protocol/p25/p25_crc.c:11-75
 *
 * One sequential cursor walks the stream the way the reference's getDibitSoft() calls do (status symbol dropped whenever its
 * counter reaches 35), so the status-symbol stripping is restated independently of the index arithmetic the device uses.
 * Pinned against the UNMODIFIED handlers replayed by oracle/_ref/libdsdneo_ref_p25.so (tests/test_oracle_p25p1_frame.py).
 */
#include <string.h>

#include "oracle.h"

typedef struct {
    const uint8_t* dibits;
    const int16_t* llr;
    int count, pos, status_count, overrun;
} cursor;

/* read_dibit_soft (p25p1_hdu.c:54-79): drop a status symbol when the counter says so, then take one dibit */
static int
next_dibit(cursor* c, int* l0, int* l1) {
    if (c->status_count == 35) {
        c->pos++;
        c->status_count = 1;
    } else {
        c->status_count++;
    }
    if (c->pos >= c->count) {
        c->overrun = 1;
        c->pos++;
        *l0 = *l1 = 0;
        return 0;
    }
    const int d = c->dibits[c->pos] & 3;
    *l0 = c->llr[2 * c->pos];
    *l1 = c->llr[2 * c->pos + 1];
    c->pos++;
    return d;
}

static int
iabs(int v) {
    return v < 0 ? -v : v;
}

static int
clamp255(int v) {
    return v > 255 ? 255 : v;
}

/* ComputeCrcCCITT16b + crc16_ok over the first 80 bits against the next 16 (p25_crc.c:11-60): 0 ok, 65535 mismatch */
static int
crc16_80(const uint8_t* bytes12) {
    unsigned crc = 0;
    for (int i = 0; i < 80; i++) {
        const unsigned bit = (bytes12[i >> 3] >> (7 - (i & 7))) & 1u;
        if (((crc >> 15) & 1u) ^ bit) {
            crc = ((crc << 1) ^ 0x1021u) & 0xFFFFu;
        } else {
            crc = (crc << 1) & 0xFFFFu;
        }
    }
    crc ^= 0xFFFFu;
    const unsigned rx = ((unsigned)bytes12[10] << 8) | bytes12[11];
    return rx == crc ? 0 : 65535;
}

/* p25_lsd_fec_16x8 (p25_lsd.c:27-78); the table lsd_parity[d] is (d * x^8) mod (x^8 + x^5 + x^4 + x^3 + 1) */
static unsigned
lsd_parity_of(unsigned d) {
    unsigned r = d << 8;
    for (int i = 15; i >= 8; i--) {
        if ((r >> i) & 1u) {
            r ^= 0x139u << (i - 8);
        }
    }
    return r & 0xFFu;
}

static int
lsd_hard(uint8_t* bits16) {
    unsigned data = 0, parity = 0;
    for (int i = 0; i < 8; i++) {
        data = (data << 1) | (bits16[i] & 1u);
        parity = (parity << 1) | (bits16[8 + i] & 1u);
    }
    const unsigned synd = parity ^ lsd_parity_of(data);
    if (synd == 0) {
        return 1;
    }
    if ((synd & (synd - 1)) == 0) {
        int b = 7;
        while (!((synd >> b) & 1u)) {
            b--;
        }
        bits16[8 + (7 - b)] ^= 1;
        return 1;
    }
    for (int pos = 0; pos < 8; pos++) {
        if (lsd_parity_of(1u << (7 - pos)) == synd) {
            bits16[pos] ^= 1;
            return 1;
        }
    }
    return 0;
}

/* p25_lsd_fec_16x8_soft (p25_lsd.c:80-161) */
static int
lsd_soft(uint8_t* bits16, const int16_t* llr16, int threshold) {
    if (lsd_hard(bits16)) {
        return 1;
    }
    int cand[16], n = 0;
    for (int i = 0; i < 16; i++) {
        if (iabs(llr16[i]) < threshold) {
            cand[n++] = i;
        }
    }
    for (int i = 0; i < n; i++) {
        for (int j = i + 1; j < n; j++) {
            const int ri = iabs(llr16[cand[i]]), rj = iabs(llr16[cand[j]]);
            if (rj < ri || (rj == ri && cand[j] < cand[i])) {
                const int t = cand[i];
                cand[i] = cand[j];
                cand[j] = t;
            }
        }
    }
    if (n > 6) {
        n = 6;
    }
    if (n <= 0) {
        return 0;
    }
    uint8_t best[16];
    int best_pen = 999999, found = 0;
    for (int mask = 1; mask < (1 << n); mask++) {
        uint8_t tmp[16];
        memcpy(tmp, bits16, 16);
        int pen = 0;
        for (int b = 0; b < n; b++) {
            if (mask & (1 << b)) {
                tmp[cand[b]] ^= 1;
                pen += iabs(llr16[cand[b]]);
            }
        }
        if (pen >= best_pen) {
            continue;
        }
        if (lsd_hard(tmp)) {
            memcpy(best, tmp, 16);
            best_pen = pen;
            found = 1;
        }
    }
    if (!found) {
        return 0;
    }
    memcpy(bits16, best, 16);
    return 1;
}

/* P25 Phase 1 IMBE interleave schedule (TIA-102.BAAA; include/dsd-neo/protocol/p25/p25p1_const.h:30-53), stored as the flat
 * bit index row * 23 + column of imbe_fr[8][23] for the first and second bit of each of the 72 dibits */
static const uint8_t k_imbe_hi[72] = {
    22, 66, 102, 43, 87, 115, 20, 64, 100, 41, 85, 151, 18, 62, 98, 39, 83, 149, 16, 60, 96, 37, 81, 147,
    14, 58, 94, 35, 79, 145, 12, 56, 92, 33, 77, 143, 10, 54, 128, 31, 75, 141, 8, 52, 126, 29, 73, 139,
    6, 50, 124, 27, 71, 167, 4, 48, 122, 25, 69, 165, 2, 46, 120, 23, 105, 163, 0, 90, 118, 67, 103, 161};
static const uint8_t k_imbe_lo[72] = {
    44, 88, 116, 21, 65, 101, 42, 86, 152, 19, 63, 99, 40, 84, 150, 17, 61, 97, 38, 82, 148, 15, 59, 95,
    36, 80, 146, 13, 57, 93, 34, 78, 144, 11, 55, 129, 32, 76, 142, 9, 53, 127, 30, 74, 140, 7, 51, 125,
    28, 72, 138, 5, 49, 123, 26, 70, 166, 3, 47, 121, 24, 106, 164, 1, 91, 119, 68, 104, 162, 45, 89, 117};

static void
read_imbe(cursor* c, oracle_p25p1_voice* v, int k) {
    uint8_t bit[184], rel[184];
    memset(bit, 0, sizeof(bit));
    memset(rel, 0, sizeof(rel));
    for (int j = 0; j < 72; j++) {
        int l0, l1;
        const int d = next_dibit(c, &l0, &l1);
        bit[k_imbe_hi[j]] = (uint8_t)((d >> 1) & 1);
        rel[k_imbe_hi[j]] = (uint8_t)clamp255(iabs(l0));
        bit[k_imbe_lo[j]] = (uint8_t)(d & 1);
        rel[k_imbe_lo[j]] = (uint8_t)clamp255(iabs(l1));
    }
    if (!v) {
        return;
    }
    for (int r = 0; r < 8; r++) {
        uint32_t w = 0;
        for (int col = 0; col < 23; col++) {
            w |= (uint32_t)bit[r * 23 + col] << col;
            v->reliab[k][r][col] = rel[r * 23 + col];
        }
        v->bits[k][r] = w;
    }
}

/* read_and_correct_hex_word (p25p1_ldu.c:190-222): 3 data + 2 parity dibits, Hamming(10,6,3) hard then soft.  Returns the
 * corrected 6-bit word; *sym_rel = min |llr| of the six RAW data-bit LLRs (p25p1_hamming_rs_symbol_reliability). */
static int
read_hamming_word(cursor* c, int threshold, int* sym_rel, int* soft_changed) {
    uint8_t bits[10], out[10], data[6], parity[4];
    int rel[10], minrel = 255;
    for (int d = 0; d < 5; d++) {
        int l0, l1;
        const int dib = next_dibit(c, &l0, &l1);
        bits[2 * d] = (uint8_t)((dib >> 1) & 1);
        bits[2 * d + 1] = (uint8_t)(dib & 1);
        rel[2 * d] = iabs(l0);
        rel[2 * d + 1] = iabs(l1);
        if (d < 3) {
            if (clamp255(iabs(l0)) < minrel) {
                minrel = clamp255(iabs(l0));
            }
            if (clamp255(iabs(l1)) < minrel) {
                minrel = clamp255(iabs(l1));
            }
        }
    }
    *sym_rel = minrel;
    memcpy(data, bits, 6);
    memcpy(parity, bits + 6, 4);
    const int hard = oracle_hamming_10_6_3_decode(data, parity);
    if (hard == 1 || hard == 2) {
        const int soft = oracle_hamming_10_6_3_soft(bits, rel, 1, threshold, out);
        if (soft != 2) {
            uint8_t hard_bits[10];
            memcpy(hard_bits, data, 6);
            memcpy(hard_bits + 6, parity, 4);
            if (hard == 2 || memcmp(out, hard_bits, 10) != 0) {
                (*soft_changed)++;
            }
            memcpy(data, out, 6);
        }
    }
    int v = 0;
    for (int i = 0; i < 6; i++) {
        v = (v << 1) | (data[i] & 1);
    }
    return v;
}

/* read_and_correct_hex_word of the HDU (p25p1_hdu.c:108-210): 3 data + 6 parity dibits, Golay(24,6) hard then soft */
static int
read_golay6_word(cursor* c, int threshold, int* sym_rel, int* soft_changed) {
    uint8_t data[6], parity[12], raw_data[6];
    int rel[18], minrel = 255, idx = 0;
    for (int d = 0; d < 9; d++) {
        int l0, l1;
        const int dib = next_dibit(c, &l0, &l1);
        if (d < 3) {
            data[2 * d] = (uint8_t)((dib >> 1) & 1);
            data[2 * d + 1] = (uint8_t)(dib & 1);
            if (clamp255(iabs(l0)) < minrel) {
                minrel = clamp255(iabs(l0));
            }
            if (clamp255(iabs(l1)) < minrel) {
                minrel = clamp255(iabs(l1));
            }
        } else {
            parity[2 * (d - 3)] = (uint8_t)((dib >> 1) & 1);
            parity[2 * (d - 3) + 1] = (uint8_t)(dib & 1);
        }
        rel[idx++] = iabs(l0);
        rel[idx++] = iabs(l1);
    }
    *sym_rel = minrel;
    memcpy(raw_data, data, 6);
    int fixed = 0;
    const int hard = oracle_p25_golay24_decode(6, data, parity, &fixed);
    if (hard != 0 || fixed > 0) {
        uint8_t sd[6];
        int sfixed = 0;
        memcpy(sd, raw_data, 6);
        if (oracle_p25_golay24_soft(6, sd, parity, rel, 1, threshold, &sfixed) == 0) {
            if (hard != 0) {
                (*soft_changed)++;
            }
            memcpy(data, sd, 6);
        }
    }
    int v = 0;
    for (int i = 0; i < 6; i++) {
        v = (v << 1) | (data[i] & 1);
    }
    return v;
}

/* hex words (value, MSB first) <-> the byte-per-bit arrays the RS wrappers take */
static void
words_to_bits(const uint8_t* words, int n, uint8_t* bits) {
    for (int i = 0; i < n; i++) {
        for (int b = 0; b < 6; b++) {
            bits[6 * i + b] = (uint8_t)((words[i] >> (5 - b)) & 1);
        }
    }
}

static void
bits_to_words(const uint8_t* bits, int n, uint8_t* words) {
    for (int i = 0; i < n; i++) {
        int v = 0;
        for (int b = 0; b < 6; b++) {
            v = (v << 1) | (bits[6 * i + b] & 1);
        }
        words[i] = (uint8_t)v;
    }
}

/* check_and_fix_* then p25p1_rs_*_soft_reliability (p25p1_hdu.c:270-285, p25p1_ldu1.c:228-245, p25p1_ldu2.c:258-272) */
static void
run_rs(oracle_p25p1_frame* f, int kind, int n_total, int n_data, const uint8_t* data_rel, const uint8_t* par_rel, int threshold) {
    uint8_t dbits[120], pbits[96];
    const int n_par = n_total - n_data;
    words_to_bits(f->rs_in_data, n_data, dbits);
    words_to_bits(f->rs_in_parity, n_par, pbits);
    f->rs_kind = (uint8_t)kind;
    int rc = oracle_p25_rs_decode(n_total, n_data, dbits, pbits);
    f->rs_status = rc == 0 ? 0 : 2;
    if (rc != 0 && oracle_p25_rs_soft_reliability(n_total, n_data, dbits, pbits, data_rel, par_rel, threshold) == 0) {
        f->rs_status = 1;
    }
    bits_to_words(dbits, n_data, f->rs_data);
}

static void
decode_ldu(cursor* c, oracle_p25p1_frame* f, oracle_p25p1_voice* v, int ldu2, int threshold) {
    const int n_data = ldu2 ? 16 : 12, n_par = 24 - n_data;
    uint8_t data_rel[16], par_rel[12];
    int soft_changed = 0, w = 0;
    uint8_t lsd_bits[32];
    int16_t lsd_llr[32];
    for (int imbe = 0; imbe < 9; imbe++) {
        read_imbe(c, v, imbe);
        if (imbe >= 1 && imbe <= 6) { /* four hex words follow voice frames 2..7 */
            for (int k = 0; k < 4; k++, w++) {
                int rel;
                const int word = read_hamming_word(c, threshold, &rel, &soft_changed);
                if (w < n_data) {
                    f->rs_in_data[n_data - 1 - w] = (uint8_t)word;
                    data_rel[n_data - 1 - w] = (uint8_t)rel;
                } else {
                    f->rs_in_parity[23 - w] = (uint8_t)word;
                    par_rel[23 - w] = (uint8_t)rel;
                }
            }
        } else if (imbe == 7) { /* low speed data: 2 x (8 data bits + 8 parity bits) */
            for (int k = 0; k < 16; k++) {
                int l0, l1;
                const int d = next_dibit(c, &l0, &l1);
                lsd_bits[2 * k] = (uint8_t)((d >> 1) & 1);
                lsd_bits[2 * k + 1] = (uint8_t)(d & 1);
                lsd_llr[2 * k] = (int16_t)l0;
                lsd_llr[2 * k + 1] = (int16_t)l1;
            }
        }
    }
    f->n_word_soft = (uint8_t)soft_changed;
    run_rs(f, ldu2 ? 3 : 2, 24, n_data, data_rel, par_rel, threshold);
    f->lsd_ok = 0;
    for (int k = 0; k < 2; k++) {
        if (lsd_soft(lsd_bits + 16 * k, lsd_llr + 16 * k, threshold)) {
            f->lsd_ok |= (uint8_t)(1 << k);
        }
        int val = 0;
        for (int i = 0; i < 8; i++) {
            val = (val << 1) | (lsd_bits[16 * k + i] & 1);
        }
        f->lsd[k] = (uint8_t)val;
    }
    c->pos++; /* trailing status symbol (p25p1_ldu1.c:218-226) */
}

/* read_and_correct_dodeca_word (p25p1_tdulc.c:75-157): 6 data + 6 parity dibits, Golay(24,12) hard, then the soft decoder on the
 * raw word when the hard decode failed or corrected anything.  half_rel[h] = dodeca_half_reliability (:159-169): the weakest
 * clamped |LLR| of data dibits 3h .. 3h+2. */
static void
read_golay12_word(cursor* c, int threshold, uint8_t* data12, int* half_rel, int* soft_changed) {
    uint8_t parity[12], raw[12];
    int rel[24];
    half_rel[0] = half_rel[1] = 255;
    for (int d = 0; d < 12; d++) {
        int l0, l1;
        const int dib = next_dibit(c, &l0, &l1);
        uint8_t* dst = d < 6 ? data12 + 2 * d : parity + 2 * (d - 6);
        dst[0] = (uint8_t)((dib >> 1) & 1);
        dst[1] = (uint8_t)(dib & 1);
        rel[2 * d] = iabs(l0);
        rel[2 * d + 1] = iabs(l1);
        if (d < 6) {
            int* h = &half_rel[d / 3];
            if (clamp255(iabs(l0)) < *h) {
                *h = clamp255(iabs(l0));
            }
            if (clamp255(iabs(l1)) < *h) {
                *h = clamp255(iabs(l1));
            }
        }
    }
    memcpy(raw, data12, 12);
    int fixed = 0;
    const int hard = oracle_p25_golay24_decode(12, data12, parity, &fixed);
    if (hard != 0 || fixed > 0) {
        uint8_t sd[12];
        int sfixed = 0;
        memcpy(sd, raw, 12);
        if (oracle_p25_golay24_soft(12, sd, parity, rel, 1, threshold, &sfixed) == 0) {
            if (hard != 0) {
                (*soft_changed)++;
            }
            memcpy(data12, sd, 12);
        }
    }
}

/* processTDULC (p25p1_tdulc.c:198-238,284-300): six data dodeca words 5..0, six parity dodeca words 5..0; swap_hex_words turns
 * every dodeca word into two hex symbols (bits 6..11 first), RS(24,12,13) hard then ranked erasures, ten null dibits and the
 * trailing status symbol.  Record: rs_in_* / rs_data hold the 12 + 12 hex symbols in that (swapped) order; the link control
 * word is dodeca 5..0, i.e. rs_data[11], rs_data[10], rs_data[9] ... read as (2i+1, 2i) pairs. */
static void
decode_tdulc(cursor* c, oracle_p25p1_frame* f, int threshold) {
    uint8_t word[12][12], data_rel[12], par_rel[12];
    int soft_changed = 0;
    for (int k = 0; k < 12; k++) { /* air order: data[5] .. data[0], parity[5] .. parity[0] */
        const int i = 5 - (k % 6);
        int hr[2];
        read_golay12_word(c, threshold, word[k], hr, &soft_changed);
        uint8_t* in = k < 6 ? f->rs_in_data : f->rs_in_parity;
        uint8_t* rl = k < 6 ? data_rel : par_rel;
        int hi = 0, lo = 0;
        for (int b = 0; b < 6; b++) {
            lo = (lo << 1) | (word[k][b] & 1);      /* bits 0..5 */
            hi = (hi << 1) | (word[k][6 + b] & 1);  /* bits 6..11 */
        }
        in[2 * i] = (uint8_t)hi;
        in[2 * i + 1] = (uint8_t)lo;
        rl[2 * i] = (uint8_t)hr[1];
        rl[2 * i + 1] = (uint8_t)hr[0];
    }
    f->n_word_soft = (uint8_t)soft_changed;
    run_rs(f, 2, 24, 12, data_rel, par_rel, threshold);
    for (int i = 0; i < 10; i++) { /* read_zeros(20): ten dibits */
        int l0, l1;
        (void)next_dibit(c, &l0, &l1);
    }
    c->pos++; /* trailing status symbol (p25p1_tdulc.c:249-251) */
}

static void
decode_hdu(cursor* c, oracle_p25p1_frame* f, int threshold) {
    uint8_t data_rel[20], par_rel[16];
    int soft_changed = 0;
    for (int w = 0; w < 36; w++) {
        int rel;
        const int word = read_golay6_word(c, threshold, &rel, &soft_changed);
        if (w < 20) {
            f->rs_in_data[19 - w] = (uint8_t)word;
            data_rel[19 - w] = (uint8_t)rel;
        } else {
            f->rs_in_parity[35 - w] = (uint8_t)word;
            par_rel[35 - w] = (uint8_t)rel;
        }
    }
    f->n_word_soft = (uint8_t)soft_changed;
    run_rs(f, 1, 36, 20, data_rel, par_rel, threshold);
    c->pos += 6; /* five filler dibits and the trailing status symbol (p25p1_hdu.c:213-224) */
}

static void
decode_tsbk(cursor* c, oracle_p25p1_frame* f) {
    int skip = 36 - 14;
    f->n_tsbk = 0;
    f->tsbk_crc_ok = 0;
    for (int block = 0; block < 3; block++) {
        int16_t llr[196];
        int k = 0;
        for (int i = 0; i < 101; i++) { /* tsbk_read_repetition_samples (p25p1_tsbk.c:135-152) */
            int d = 0, l0 = 0, l1 = 0;
            if (c->pos < c->count) {
                d = c->dibits[c->pos] & 3;
                l0 = c->llr[2 * c->pos];
                l1 = c->llr[2 * c->pos + 1];
            } else {
                c->overrun = 1;
            }
            (void)d;
            c->pos++;
            if ((skip / 36) == 0) {
                if (k < 98) {
                    llr[2 * k] = (int16_t)l0;
                    llr[2 * k + 1] = (int16_t)l1;
                }
                k++;
            } else {
                skip = 0;
            }
            skip++;
        }
        uint8_t cand[8 * 12];
        uint32_t metric[8];
        const int n = oracle_p25_12_soft_llr_list(llr, cand, metric, 8);
        uint8_t* out = f->tsbk[block];
        if (n > 0) {
            int sel = 0;
            for (int i = 0; i < n; i++) {
                if (crc16_80(cand + 12 * i) == 0) {
                    sel = i;
                    break;
                }
            }
            memcpy(out, cand + 12 * sel, 12);
        } else {
            oracle_p25_12_soft_llr(llr, out);
        }
        if (crc16_80(out) == 0) {
            f->tsbk_crc_ok |= (uint8_t)(1 << block);
        }
        f->n_tsbk = (uint8_t)(block + 1);
        if ((out[0] >> 7) & 1) {
            break;
        }
    }
}

/*
 * One frame whose LAST sync dibit sits at stream index pos_last_sync.  `voice` (may be NULL) receives the nine IMBE frames of
 * an LDU.  Returns the dibits consumed after the sync, or -1 when the stream ended inside the frame (record unusable).
 */
int
oracle_p25p1_decode_frame(const uint8_t* dibits, const int16_t* llr, int count, int pos_last_sync, int observed_nac, int threshold,
                          oracle_p25p1_frame* f, oracle_p25p1_voice* voice) {
    memset(f, 0, sizeof(*f));
    f->voice_index = -1;
    f->duid = 0xFF;
    uint8_t code63[63], rel63[63], parity, prel;
    uint8_t pd[4];
    int16_t pl[8];
    const int flags = oracle_p25p1_frame_cut(dibits, llr, count, pos_last_sync, 0, code63, rel63, &parity, &prel, pd, pl);
    if (!(flags & 1)) {
        return -1;
    }
    int nac = 0, duid = 0, errs = 0;
    const int st = oracle_p25p1_nid_decode(code63, rel63, observed_nac, parity, prel, threshold, &nac, &duid, &errs);
    f->nid_status = (int8_t)st;
    f->nac = (int16_t)nac;
    f->nid_errs = (int16_t)errs;
    cursor c = {dibits, llr, count, pos_last_sync + 1 + 33, 21, 0};
    if (st <= 0) {
        return 33;
    }
    f->duid = (uint8_t)duid;
    switch (duid) {
        case 0x0: decode_hdu(&c, f, threshold); break;
        case 0x5: decode_ldu(&c, f, voice, 0, threshold); break;
        case 0xA: decode_ldu(&c, f, voice, 1, threshold); break;
        case 0x7: decode_tsbk(&c, f); break;
        case 0xF: decode_tdulc(&c, f, threshold); break;
        default: break; /* TDU has no payload; MPDU payloads are not decoded here */
    }
    if (c.overrun || c.pos > count) {
        return -1;
    }
    return c.pos - (pos_last_sync + 1);
}
