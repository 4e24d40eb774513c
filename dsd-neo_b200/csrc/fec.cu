// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * Batched FEC leaves of the hot path (K13, K16, K17, K19): one thread per codeword / burst / trellis block.
 * Pure integer work, bit-exact with the reference (statuses, tie-breaks and in-place correction semantics
 * included).  Reference entry points replaced, each by a `*_batch` twin over n independent items:
 *   Hamming_{7_4,12_8,13_9,15_11,16_11_4}_decode, Golay_{20_8,24_12}_decode, QR_16_7_6_decode,
 *   Golay_24_12_encode                                  src/fec/fec.c:145-824 (API include/dsd-neo/fec/block_codes.h:29-56)
 *   BPTCDeInterleaveDMRData + BPTC_196x96_Extract_Data   src/fec/bptc.c:51-59,136-149
 *   p25_12_soft_llr, p25_12_soft_llr_list                src/protocol/p25/p25_12.c:204-283,144-202
 *   check_and_fix_redsolomon_36_20_17, check_and_fix_reedsolomon_24_12_13 / _24_16_9
 *                                                        src/protocol/p25/phase1/p25p1_check_hdu.cpp:38-45, p25p1_check_ldu.cpp:37-62
 *                                                        (engine include/dsd-neo/fec/ReedSolomon.hpp:61-816)
 * Data layouts at the C-ABI are the reference's own (one bit per byte, hex words as 6 bytes MSB first, int16 LLRs);
 * words are packed to registers on load.  These kernels move tens of bytes per item and are latency-bound; they
 * are sized one thread per item so that thousands of channels' frames fill the machine.
 */
#include <stdlib.h>
#include <string.h>

#include "../host/fec_tables.h"
#include "common.cuh"

using namespace dsdneo;

namespace {

dsdneo_fec_tables* g_d_tables = nullptr; /* device copy */

int
ensure_tables() {
    if (g_d_tables) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    dsdneo_fec_tables* h = (dsdneo_fec_tables*)malloc(sizeof(dsdneo_fec_tables));
    if (!h) {
        set_error("fec: out of host memory");
        return DSDNEO_B200_ENOMEM;
    }
    dsdneo_fec_build_tables(h);
    dsdneo_fec_tables* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, sizeof(dsdneo_fec_tables));
    if (e == cudaSuccess) {
        e = cudaMemcpy(d, h, sizeof(dsdneo_fec_tables), cudaMemcpyHostToDevice);
    }
    free(h);
    if (e != cudaSuccess) {
        cudaFree(d);
        return cuda_fail(e, "fec tables upload", __FILE__, __LINE__);
    }
    g_d_tables = d;
    return 0;
}

/* ------------------------------------------------------------------ block codes */

__device__ __forceinline__ unsigned
pack_bits(const uint8_t* b, int n) {
    unsigned w = 0;
    for (int j = 0; j < n; j++) {
        w |= (unsigned)(b[j] & 1u) << j; /* the reference's mod-2 sums see only the LSB of each byte */
    }
    return w;
}

__global__ void
hamming_decode_kernel(const dsdneo_fec_tables* __restrict__ T, int code, uint8_t* bits, uint8_t* decoded, uint8_t* ok_out,
                      int n_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) {
        return;
    }
    const dsdneo_hamming_table& t = T->ham[code];
    uint8_t* b = bits + (size_t)i * t.n;
    const unsigned w = pack_bits(b, t.n);
    unsigned s = 0;
    for (int row = 0; row < t.r; row++) {
        s |= (unsigned)(__popc(w & t.row_mask[row]) & 1) << (t.r - 1 - row);
    }
    int ok = 1;
    if (s) {
        const unsigned pos = t.pos_of[s];
        if (pos == 0xFFu) {
            ok = 0;
        } else {
            b[pos] ^= 1; /* in place, upper bits of the byte untouched like the reference */
        }
    }
    ok_out[i] = (uint8_t)ok;
    if (!ok && t.stop_on_fail) {
        return; /* the reference breaks out before copying the information bits */
    }
    if (decoded && code != DSDNEO_FEC_HAMMING_7_4) {
        uint8_t* d = decoded + (size_t)i * t.k;
        for (int j = 0; j < t.k; j++) {
            d[j] = b[j];
        }
    }
}

__global__ void
gq_decode_kernel(const dsdneo_fec_tables* __restrict__ T, int which, uint8_t* bits, uint8_t* ok_out, int n_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) {
        return;
    }
    const dsdneo_gq_table& t = T->gq[which];
    uint8_t* b = bits + (size_t)i * t.n;
    const unsigned w = pack_bits(b, t.n);
    unsigned s = 0;
    for (int row = 0; row < t.r; row++) {
        s |= (unsigned)(__popc(w & t.row_mask[row]) & 1) << (t.r - 1 - row);
    }
    int ok = 1;
    if (s) {
        int flips = 0;
        for (; flips < t.maxw; flips++) {
            const unsigned pos = t.corr[s][flips];
            if (pos == 0xFFu) {
                break;
            }
            b[pos] ^= 1;
        }
        ok = flips != 0;
        if (which == 0 && flips > 2) {
            ok = 0; /* Golay_20_8_decode: three flips are applied but reported as failure (fec.c:578-584) */
        }
    }
    ok_out[i] = (uint8_t)ok;
}

__global__ void
golay_24_12_encode_kernel(const dsdneo_fec_tables* __restrict__ T, const uint8_t* data, uint8_t* out, int n_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) {
        return;
    }
    const dsdneo_gq_table& t = T->gq[1];
    const unsigned w = pack_bits(data + (size_t)i * 12, 12);
    uint8_t* o = out + (size_t)i * 24;
    for (int j = 0; j < 12; j++) {
        o[j] = (uint8_t)((w >> j) & 1u);
    }
    for (int row = 0; row < 12; row++) {
        o[12 + row] = (uint8_t)(__popc(w & t.row_mask[row] & 0xFFFu) & 1);
    }
}

/* ------------------------------------------------------------------ BPTC(196,96) */

__device__ __forceinline__ int
ham_fix_word(const dsdneo_hamming_table& t, unsigned& w) {
    unsigned s = 0;
    for (int row = 0; row < t.r; row++) {
        s |= (unsigned)(__popc(w & t.row_mask[row]) & 1) << (t.r - 1 - row);
    }
    if (!s) {
        return 1;
    }
    const unsigned pos = t.pos_of[s];
    if (pos == 0xFFu) {
        return 0;
    }
    w ^= 1u << pos;
    return 1;
}

__global__ void
bptc_196x96_kernel(const dsdneo_fec_tables* __restrict__ T, const uint8_t* in, int interleaved, uint8_t* out96, uint8_t* r3,
                   uint32_t* errs_out, int n_bursts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bursts) {
        return;
    }
    const uint8_t* src = in + (size_t)i * 196;
    unsigned rows[13]; /* bit j of rows[r] = matrix[r][j] */
    for (int r = 0; r < 13; r++) {
        rows[r] = 0;
    }
    /* matrix[r][j] = deinterleaved[1 + 15 r + j]; deinterleaved[(13 t) mod 196] = received[t] (bptc.c:51-59) */
    for (int t = 0; t < 196; t++) {
        const int pos = interleaved ? (13 * t) % 196 : t;
        if (pos == 0) {
            continue; /* reserved bit R(3) */
        }
        const int r = (pos - 1) / 15, j = (pos - 1) - 15 * r;
        rows[r] |= (unsigned)(src[t] & 1u) << j;
    }
    const dsdneo_hamming_table& h15 = T->ham[DSDNEO_FEC_HAMMING_15_11];
    const dsdneo_hamming_table& h13 = T->ham[DSDNEO_FEC_HAMMING_13_9];
    unsigned errs = 0;
    for (int pass = 0; pass < 2; pass++) {
        unsigned e = 0;
        for (int r = 0; r < 9; r++) {
            unsigned w = rows[r];
            (void)ham_fix_word(h15, w); /* perfect code: always "correctable"; only bits 0..10 are written back */
            rows[r] = (rows[r] & ~0x7FFu) | (w & 0x7FFu);
        }
        unsigned last = 0;
        int have_last = 0;
        for (int j = 0; j < 15; j++) {
            unsigned c = 0;
            for (int r = 0; r < 13; r++) {
                c |= ((rows[r] >> j) & 1u) << r;
            }
            unsigned put;
            if (ham_fix_word(h13, c)) {
                last = c & 0x1FFu;
                have_last = 1;
                put = last;
            } else {
                e++;
                if (!have_last) {
                    continue; /* reference reads an uninitialised buffer here; column left unchanged (documented) */
                }
                put = last; /* stale output buffer of the callee: previous good column (bptc.c:100-113) */
            }
            for (int r = 0; r < 9; r++) {
                rows[r] = (rows[r] & ~(1u << j)) | (((put >> r) & 1u) << j);
            }
        }
        if (pass == 1) {
            errs = e;
        }
    }
    uint8_t* o = out96 + (size_t)i * 96;
    int k = 0;
    for (int j = 3; j < 11; j++) {
        o[k++] = (uint8_t)((rows[0] >> j) & 1u);
    }
    for (int r = 1; r < 9; r++) {
        for (int j = 0; j < 11; j++) {
            o[k++] = (uint8_t)((rows[r] >> j) & 1u);
        }
    }
    if (r3) {
        r3[3 * i + 0] = (uint8_t)((rows[0] >> 2) & 1u);
        r3[3 * i + 1] = (uint8_t)((rows[0] >> 1) & 1u);
        r3[3 * i + 2] = (uint8_t)(rows[0] & 1u);
    }
    errs_out[i] = errs;
}

/* BPTC_128x77_Extract_Data (src/fec/bptc.c:167-252): rows 0-6 Hamming(16,11,4), row 7 column parity; an
 * uncorrectable row receives the information bits of the most recent correctable row (the callee's stale output
 * buffer); with no correctable row before it the reference's result is undefined and the row is left unchanged. */
__global__ void
bptc_128x77_kernel(const dsdneo_fec_tables* __restrict__ T, const uint8_t* in, uint8_t* out77, uint32_t* errs_out, int n_items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) {
        return;
    }
    const uint8_t* src = in + (size_t)i * 128;
    const dsdneo_hamming_table& h16 = T->ham[DSDNEO_FEC_HAMMING_16_11_4];
    unsigned rows[8];
    for (int r = 0; r < 8; r++) {
        rows[r] = pack_bits(src + 16 * r, 16);
    }
    unsigned errs = 0, last = 0;
    int have_last = 0;
    for (int r = 0; r < 7; r++) {
        unsigned w = rows[r];
        if (ham_fix_word(h16, w)) {
            last = w & 0x7FFu;
            have_last = 1;
            rows[r] = (rows[r] & ~0x7FFu) | last;
        } else {
            errs++;
            if (have_last) {
                rows[r] = (rows[r] & ~0x7FFu) | last;
            }
        }
    }
    uint8_t* o = out77 + (size_t)i * 77;
    int k = 0;
    for (int r = 0; r < 2; r++) {
        for (int j = 0; j < 11; j++) {
            o[k++] = (uint8_t)((rows[r] >> j) & 1u);
        }
    }
    for (int r = 2; r < 7; r++) {
        for (int j = 0; j < 10; j++) {
            o[k++] = (uint8_t)((rows[r] >> j) & 1u);
        }
    }
    for (int r = 2; r < 7; r++) {
        o[k++] = (uint8_t)((rows[r] >> 10) & 1u);
    }
    unsigned par = 0;
    for (int r = 0; r < 7; r++) {
        par ^= rows[r];
    }
    errs += (unsigned)__popc((par ^ rows[7]) & 0xFFFFu);
    errs_out[i] = errs;
}

/* BPTC_16x2_Extract_Data (src/fec/bptc.c:272-333) with the reverse-channel de-interleave tables (:33-38). */
__global__ void
bptc_16x2_kernel(const dsdneo_fec_tables* __restrict__ T, const uint8_t* in, uint8_t* out32, uint32_t* errs_out, int parity_odd,
                 int n_items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) {
        return;
    }
    const uint8_t* src = in + (size_t)i * 32;
    unsigned m = 0;
    for (int t = 0; t < 32; t++) {
        /* DeInterleaveReverseChannelBptc[t] = t + 16 * (t odd) mod 32 pattern; Placement[v] = v/2 + 16 * (v odd) */
        const int v = (t & 1) ? ((t + 16) & 31) : t;
        const int pos = (v >> 1) + ((v & 1) << 4);
        m |= (unsigned)(src[t] & 1u) << pos;
    }
    const dsdneo_hamming_table& h16 = T->ham[DSDNEO_FEC_HAMMING_16_11_4];
    unsigned w = m & 0xFFFFu, errs = 0;
    if (ham_fix_word(h16, w)) {
        m = (m & ~0x7FFu) | (w & 0x7FFu);
    } else {
        errs = 1; /* reference copies an uninitialised buffer here: left as de-interleaved (documented) */
    }
    uint8_t* o = out32 + (size_t)i * 32;
    for (int j = 0; j < 32; j++) {
        o[j] = (uint8_t)((m >> j) & 1u);
    }
    const unsigned same = (unsigned)__popc(~((m & 0xFFFFu) ^ (m >> 16)) & 0xFFFFu);
    errs_out[i] = errs + (parity_odd ? same : 16u - same);
}

/* ------------------------------------------------------------------ P25 half-rate trellis */

/* transition nibble table p25_12.c:19, entry i in nibble i */
__device__ __forceinline__ unsigned
p25_dtm(int prev, int next) {
    return (unsigned)((0x86B54A793D0EF1C2ull >> (4 * ((prev << 2) | next))) & 0xFull);
}

/* de-interleaved position of received dibit i (trellis34.c:8-13): groups of dibit pairs with stride 8 */
__device__ __forceinline__ int
p25_deinterleave_pos(int i) {
    /* group sizes in dibits: 26, 24, 24, 24 */
    int g, o;
    if (i < 26) {
        g = 0, o = i;
    } else {
        g = 1 + (i - 26) / 24;
        o = (i - 26) % 24;
    }
    return 2 * g + 8 * (o >> 1) + (o & 1);
}

__device__ __forceinline__ uint32_t
llr_cost(int llr, unsigned bit) {
    return bit ? (llr < 0 ? (uint32_t)(-llr) : 0u) : (llr > 0 ? (uint32_t)llr : 0u);
}

__device__ __forceinline__ void
p25_load_deinterleaved(const int16_t* llr196, int16_t* dei) {
    for (int i = 0; i < 98; i++) {
        const int p = p25_deinterleave_pos(i);
        dei[2 * p] = llr196[2 * i];
        dei[2 * p + 1] = llr196[2 * i + 1];
    }
}

__device__ __forceinline__ uint32_t
p25_branch(const int16_t* dei, int sym, int pv, int nx) {
    const unsigned e = p25_dtm(pv, nx);
    const int16_t* l = dei + 4 * sym;
    return llr_cost(l[0], (e >> 3) & 1u) + llr_cost(l[1], (e >> 2) & 1u) + llr_cost(l[2], (e >> 1) & 1u) + llr_cost(l[3], e & 1u);
}

/* p25_12_soft_llr (src/protocol/p25/p25_12.c:204-283) for one block: 12 bytes out, returns best_final >> 8 */
__device__ __noinline__ int32_t
p25_12_decode(const int16_t* llr196, uint8_t* o, int* final_state = nullptr) {
    int16_t dei[196];
    p25_load_deinterleaved(llr196, dei);
    uint32_t pm0 = 0, pm1 = 256, pm2 = 256, pm3 = 256; /* path metrics stay in registers */
    uint8_t bp[49];                                    /* 4 x 2-bit predecessors per step */
    for (int s = 0; s < 49; s++) {
        uint32_t cm[4];
        unsigned packed = 0;
#pragma unroll
        for (int nx = 0; nx < 4; nx++) {
            uint32_t best = pm0 + p25_branch(dei, s, 0, nx);
            unsigned bprev = 0;
            uint32_t m = pm1 + p25_branch(dei, s, 1, nx);
            if (m < best) { /* strict <: the lowest predecessor wins ties (p25_12.c:244-247) */
                best = m, bprev = 1;
            }
            m = pm2 + p25_branch(dei, s, 2, nx);
            if (m < best) {
                best = m, bprev = 2;
            }
            m = pm3 + p25_branch(dei, s, 3, nx);
            if (m < best) {
                best = m, bprev = 3;
            }
            cm[nx] = best;
            packed |= bprev << (2 * nx);
        }
        bp[s] = (uint8_t)packed;
        pm0 = cm[0], pm1 = cm[1], pm2 = cm[2], pm3 = cm[3];
    }
    uint32_t bf = pm0;
    int st = 0;
    if (pm1 < bf) {
        bf = pm1, st = 1;
    }
    if (pm2 < bf) {
        bf = pm2, st = 2;
    }
    if (pm3 < bf) {
        bf = pm3, st = 3;
    }
    if (final_state) {
        *final_state = st;
    }
    uint8_t td[49];
    for (int s = 49; s-- > 0;) {
        td[s] = (uint8_t)st;
        st = (bp[s] >> (2 * st)) & 3;
    }
    for (int b = 0; b < 12; b++) {
        o[b] = (uint8_t)((td[4 * b] << 6) | (td[4 * b + 1] << 4) | (td[4 * b + 2] << 2) | td[4 * b + 3]);
    }
    return (int32_t)(bf >> 8);
}

__global__ void
p25_12_soft_llr_kernel(const int16_t* llr, uint8_t* out12, int32_t* metric_out, int n_blocks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks) {
        return;
    }
    metric_out[i] = p25_12_decode(llr + (size_t)i * 196, out12 + (size_t)i * 12);
}

constexpr int kListK = 8;

/* p25_12_soft_llr_list (src/protocol/p25/p25_12.c:31-202) for one block: `out` receives up to min(max_candidates, 8)
 * candidates sorted by metric (stable), returns their count.  llr196 = the 98 received dibits' LLR pairs in air order. */
__device__ __noinline__ int
p25_12_list_decode(const int16_t* llr196, dsdneo_b200_p25_12_candidate* out, int max_candidates) {
    int16_t dei[196];
    p25_load_deinterleaved(llr196, dei);
    uint32_t ma[4][kListK], mb[4][kListK];
    uint8_t bp[49][4][kListK];
    for (int s = 0; s < 4; s++) {
        for (int r = 0; r < kListK; r++) {
            ma[s][r] = 0xFFFFFFFFu;
        }
        ma[s][0] = s == 0 ? 0u : 256u;
    }
    uint32_t (*pm)[kListK] = ma;
    uint32_t (*cm)[kListK] = mb;
    for (int sym = 0; sym < 49; sym++) {
        for (int s = 0; s < 4; s++) {
            for (int r = 0; r < kListK; r++) {
                cm[s][r] = 0xFFFFFFFFu;
                bp[sym][s][r] = 0;
            }
        }
        for (int pv = 0; pv < 4; pv++) {
            for (int nx = 0; nx < 4; nx++) {
                const uint32_t cost = p25_branch(dei, sym, pv, nx);
                for (int rk = 0; rk < kListK; rk++) {
                    if (pm[pv][rk] == 0xFFFFFFFFu) {
                        continue;
                    }
                    const uint32_t m = pm[pv][rk] + cost;
                    int at = -1;
                    for (int q = 0; q < kListK; q++) {
                        if (m < cm[nx][q]) { /* stable insertion, strict < (p25_12.c:31-52) */
                            at = q;
                            break;
                        }
                    }
                    if (at < 0) {
                        continue;
                    }
                    for (int q = kListK - 1; q > at; q--) {
                        cm[nx][q] = cm[nx][q - 1];
                        bp[sym][nx][q] = bp[sym][nx][q - 1];
                    }
                    cm[nx][at] = m;
                    bp[sym][nx][at] = (uint8_t)((pv << 3) | rk);
                }
            }
        }
        uint32_t (*t)[kListK] = pm;
        pm = cm;
        cm = t;
    }
    if (max_candidates > kListK) {
        max_candidates = kListK;
    }
    int count = 0;
    for (int s = 0; s < 4; s++) {
        for (int rk = 0; rk < kListK; rk++) {
            if (pm[s][rk] == 0xFFFFFFFFu) {
                continue;
            }
            uint8_t bytes[12];
            for (int b = 0; b < 12; b++) {
                bytes[b] = 0;
            }
            int st = s, r = rk;
            for (int sym = 49; sym-- > 0;) {
                if (sym < 48) {
                    bytes[sym >> 2] |= (uint8_t)(st << (6 - 2 * (sym & 3)));
                }
                const uint8_t p = bp[sym][st][r];
                st = (p >> 3) & 3;
                r = p & 7;
            }
            bool dup = false;
            for (int c = 0; c < count && !dup; c++) {
                bool same = true;
                for (int b = 0; b < 12; b++) {
                    same = same && (out[c].bytes[b] == bytes[b]);
                }
                dup = same;
            }
            if (dup) {
                continue;
            }
            const uint32_t metric = pm[s][rk];
            int at = count;
            for (int c = 0; c < count; c++) {
                if (metric < out[c].metric) {
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
                out[c] = out[c - 1];
            }
            for (int b = 0; b < 12; b++) {
                out[at].bytes[b] = bytes[b];
            }
            out[at].metric = metric;
        }
    }
    return count;
}

__global__ void __launch_bounds__(64)
p25_12_soft_llr_list_kernel(const int16_t* llr, dsdneo_b200_p25_12_candidate* cands, int32_t* count_out, int max_candidates,
                            int n_blocks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks) {
        return;
    }
    count_out[i] = p25_12_list_decode(llr + (size_t)i * 196, cands + (size_t)i * kListK, max_candidates);
}

/* ------------------------------------------------------------------ RS(63,k) over GF(64) */

struct RsShape {
    int n_total, n_data, tt;
};

/* Berlekamp iteration in Rockliff's index-form bookkeeping, as ReedSolomon_63<TT>::decode
 * (ReedSolomon.hpp:353-582,738-771): same failure conditions (degree > t, root count != degree), corrections may land
 * in the zero padding of the shortened code exactly like the reference. */
/* hex words (6 byte-per-bit chars, any non-zero byte = 1, MSB first: ReedSolomon.hpp:799-815) -> 6-bit symbols, parity
 * first (ReedSolomon.hpp:838-851) */
__device__ __forceinline__ void
rs_pack_symbols(const RsShape& sh, const uint8_t* dbits, const uint8_t* pbits, uint8_t* sym) {
    const int n_par = sh.n_total - sh.n_data;
    for (int i = 0; i < sh.n_total; i++) {
        const uint8_t* src = (i < n_par) ? pbits + 6 * i : dbits + 6 * (i - n_par);
        int v = 0;
        for (int b = 0; b < 6; b++) {
            v = (v << 1) | (src[b] != 0);
        }
        sym[i] = (uint8_t)v;
    }
}

/* Hard decision decode of one shortened word: sym = n_total symbols in polynomial form (parity first); out_sym receives
 * the n_data data symbols (corrected, or as received when the word is irrecoverable).  Returns 0 / 1. */
__device__ __noinline__ int
rs63_hard_decode(const signed char* __restrict__ EXP, const signed char* __restrict__ LOG, const RsShape& sh, const uint8_t* sym,
                 int* out_sym) {
    constexpr int NN = 63;
    const int n_par = sh.n_total - sh.n_data, tt = sh.tt, n2t = 2 * sh.tt;
    signed char recd[NN]; /* index form, -1 = zero */
    const int used = sh.n_total;
    for (int i = 0; i < NN; i++) {
        recd[i] = LOG[i < used ? sym[i] : 0];
    }
    int s[18];
    int syn_err = 0;
    s[0] = 0;
    for (int i = 1; i <= n2t; i++) {
        int acc = 0;
        for (int j = 0; j < used; j++) {
            if (recd[j] != -1) {
                acc ^= EXP[(recd[j] + i * j) % NN];
            }
        }
        syn_err |= acc;
        s[i] = LOG[acc];
    }
    /* hex_to_bin of the (possibly corrected) data symbols is a no-op for 0/1 inputs; inputs with other non-zero
     * byte values are normalised to 1 like the reference's bin_to_hex/hex_to_bin round trip */
    for (int i = 0; i < sh.n_data; i++) {
        out_sym[i] = recd[n_par + i] == -1 ? 0 : EXP[recd[n_par + i]];
    }
    int rc = 0;
    if (syn_err) {
        int elp[18][16], d[18], l[18], u_lu[18];
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
                    elp[u][i] = LOG[elp[u][i]];
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
                        elp[u + 1][i + u - q] = EXP[(d[u] + NN - d[q] + elp[q][i]) % NN];
                    }
                }
                for (int i = 0; i <= l[u]; i++) {
                    elp[u + 1][i] ^= elp[u][i];
                    elp[u][i] = LOG[elp[u][i]];
                }
            }
            u_lu[u + 1] = u - l[u + 1];
            if (u < n2t) {
                int dd = (s[u + 1] != -1) ? EXP[s[u + 1]] : 0;
                for (int i = 1; i <= l[u + 1]; i++) {
                    if (s[u + 1 - i] != -1 && elp[u + 1][i] != 0) {
                        dd ^= EXP[(s[u + 1 - i] + LOG[elp[u + 1][i]]) % NN];
                    }
                }
                d[u + 1] = LOG[dd];
            }
        } while (u < n2t && l[u + 1] <= tt);
        u++;
        if (l[u] > tt) {
            rc = 1;
        } else {
            const int deg = l[u];
            int lp[9], reg[9], root[8], loc[8], count = 0;
            for (int i = 0; i <= deg; i++) {
                lp[i] = LOG[elp[u][i]];
                reg[i] = lp[i];
            }
            for (int i = 1; i <= NN; i++) {
                int q = 1;
                for (int j = 1; j <= deg; j++) {
                    if (reg[j] != -1) {
                        reg[j] = (reg[j] + j) % NN;
                        q ^= EXP[reg[j]];
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
                rc = 1;
            } else {
                int z[9];
                for (int i = 1; i <= deg; i++) {
                    int zi = 0;
                    if (s[i] != -1) {
                        zi ^= EXP[s[i]];
                    }
                    if (lp[i] != -1) {
                        zi ^= EXP[lp[i]];
                    }
                    for (int j = 1; j < i; j++) {
                        if (s[j] != -1 && lp[i - j] != -1) {
                            zi ^= EXP[(lp[i - j] + s[j]) % NN];
                        }
                    }
                    z[i] = LOG[zi];
                }
                for (int i = 0; i < deg; i++) {
                    int num = 1;
                    for (int j = 1; j <= deg; j++) {
                        if (z[j] != -1) {
                            num ^= EXP[(z[j] + j * root[i]) % NN];
                        }
                    }
                    if (num != 0) {
                        int den = 0;
                        for (int j = 0; j < deg; j++) {
                            if (j != i) {
                                den += LOG[1 ^ EXP[(loc[j] + root[i]) % NN]];
                            }
                        }
                        den %= NN;
                        const int e = EXP[(LOG[num] - den + NN) % NN];
                        const int pos = loc[i] - n_par;
                        if (pos >= 0 && pos < sh.n_data) {
                            out_sym[pos] ^= e;
                        }
                    }
                }
            }
        }
    }
    return rc;
}

__global__ void __launch_bounds__(64)
p25_rs_decode_kernel(const dsdneo_fec_tables* __restrict__ T, RsShape sh, uint8_t* data_bits, const uint8_t* parity_bits,
                     uint8_t* status, int n_words) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) {
        return;
    }
    const int n_par = sh.n_total - sh.n_data;
    uint8_t* dbits = data_bits + (size_t)w * sh.n_data * 6;
    uint8_t sym[36];
    int out_sym[36];
    rs_pack_symbols(sh, dbits, parity_bits + (size_t)w * n_par * 6, sym);
    const int rc = rs63_hard_decode(T->gf_exp, T->gf_log, sh, sym, out_sym);
    for (int i = 0; i < sh.n_data; i++) {
        for (int b = 0; b < 6; b++) {
            dbits[6 * i + b] = (uint8_t)((out_sym[i] >> (5 - b)) & 1);
        }
    }
    status[w] = (uint8_t)rc;
}

/* ---- bounded errors-and-erasures decoding (ReedSolomon.hpp:621-683,773-795) and the ranked-erasure soft wrappers ---- */

struct Gf64 {
    const signed char* EXP;
    const signed char* LOG;
    __device__ __forceinline__ int mul(int a, int b) const { return (a == 0 || b == 0) ? 0 : EXP[(LOG[a] + LOG[b]) % 63]; }
    __device__ __forceinline__ int div(int a, int b) const { return (a == 0 || b == 0) ? 0 : EXP[(LOG[a] - LOG[b] + 63) % 63]; }
    __device__ __forceinline__ int apow(int e) const { return EXP[e % 63]; } /* e >= 0 everywhere below */
};

__device__ int
rs63_syndromes(const Gf64& gf, const uint8_t* w, int used, int n2t, uint8_t* syn) {
    int err = 0;
    syn[0] = 0;
    for (int i = 1; i <= n2t; i++) {
        int acc = 0;
        for (int j = 0; j < used; j++) {
            if (w[j]) {
                acc ^= gf.mul(w[j], gf.apow(i * j));
            }
        }
        syn[i] = (uint8_t)acc;
        err |= acc;
    }
    return err != 0;
}

/* word: n_total symbols (parity first), corrected in place on success.  Returns 0 / 1 (word restored on failure).
 * Corrections that the reference would apply to the zero padding of the shortened code make its final syndrome
 * re-check see a non-codeword only if they are non-zero there; that case is reproduced by tracking the padding. */
__device__ __noinline__ int
rs63_erasure_decode(const Gf64& gf, const RsShape& sh, uint8_t* word63 /* [63], positions >= n_total are zero on entry */,
                    const uint8_t* erasures, int n_er) {
    const int n2t = 2 * sh.tt;
    uint8_t saved[63];
    for (int i = 0; i < 63; i++) {
        saved[i] = word63[i];
    }
    uint8_t syn[17];
    int status = 1;
    do {
        if (!rs63_syndromes(gf, word63, 63, n2t, syn)) {
            status = 0;
            break;
        }
        uint8_t el[17];
        for (int i = 0; i <= n2t; i++) {
            el[i] = 0;
        }
        el[0] = 1;
        for (int e = 0, deg = 0; e < n_er; e++, deg++) {
            const int f = gf.apow(erasures[e]);
            for (int i = deg; i >= 0; i--) {
                el[i + 1] ^= (uint8_t)gf.mul(el[i], f);
            }
        }
        uint8_t ms[16];
        for (int i = 0; i < n2t; i++) {
            int v = 0;
            for (int j = 0; j <= n_er && j <= i; j++) {
                v ^= gf.mul(el[j], syn[(i - j) + 1]);
            }
            ms[i] = (uint8_t)v;
        }
        uint8_t c[17], b[17], t[17];
        for (int i = 0; i <= n2t; i++) {
            c[i] = b[i] = 0;
        }
        c[0] = b[0] = 1;
        int l = 0, m = 1, bb = 1, fail = 0;
        const uint8_t* sy = ms + n_er;
        const int ns = n2t - n_er;
        for (int n = 0; n < ns; n++) {
            int disc = sy[n];
            for (int i = 1; i <= l; i++) {
                disc ^= gf.mul(c[i], sy[n - i]);
            }
            if (disc == 0) {
                m++;
                continue;
            }
            for (int i = 0; i <= n2t; i++) {
                t[i] = c[i];
            }
            if (bb == 0) {
                fail = 1;
                break;
            }
            const int coef = gf.div(disc, bb);
            for (int i = 0; i + m <= n2t; i++) {
                if (b[i]) {
                    c[i + m] ^= (uint8_t)gf.mul(coef, b[i]);
                }
            }
            if (2 * l <= n) {
                l = n + 1 - l;
                for (int i = 0; i <= n2t; i++) {
                    b[i] = t[i];
                }
                bb = disc;
                m = 1;
            } else {
                m++;
            }
        }
        if (fail) {
            break;
        }
        int udeg = 0, edeg = 0;
        for (int i = n2t; i >= 0; i--) {
            if (c[i]) {
                udeg = i;
                break;
            }
        }
        if (2 * udeg + n_er > n2t) {
            break;
        }
        for (int i = n2t; i >= 0; i--) {
            if (el[i]) {
                edeg = i;
                break;
            }
        }
        if (edeg + udeg > n2t) {
            break;
        }
        uint8_t comb[17];
        for (int i = 0; i <= n2t; i++) {
            comb[i] = 0;
        }
        for (int i = 0; i <= edeg; i++) {
            for (int j = 0; j <= udeg; j++) {
                comb[i + j] ^= (uint8_t)gf.mul(el[i], c[j]);
            }
        }
        int cdeg = 0;
        for (int i = n2t; i >= 0; i--) {
            if (comb[i]) {
                cdeg = i;
                break;
            }
        }
        uint8_t locs[16];
        int n_loc = 0;
        if (cdeg != 0) {
            for (int pos = 0; pos < 63; pos++) {
                const int x = gf.apow(63 - pos);
                int v = 0, xp = 1;
                for (int i = 0; i <= cdeg; i++) {
                    v ^= gf.mul(comb[i], xp);
                    xp = gf.mul(xp, x);
                }
                if (v == 0) {
                    if (n_loc >= n2t) {
                        n_loc++;
                        break;
                    }
                    locs[n_loc++] = (uint8_t)pos;
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
        uint8_t mat[16][17];
        for (int r = 0; r < n_loc; r++) {
            for (int k = 0; k < n_loc; k++) {
                mat[r][k] = (uint8_t)gf.apow((r + 1) * locs[k]);
            }
            mat[r][n_loc] = syn[r + 1];
        }
        int singular = 0;
        for (int col = 0; col < n_loc && !singular; col++) {
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
                    const uint8_t tmp = mat[col][k];
                    mat[col][k] = mat[piv][k];
                    mat[piv][k] = tmp;
                }
            }
            const int pv = mat[col][col];
            for (int k = col; k <= n_loc; k++) {
                mat[col][k] = (uint8_t)gf.div(mat[col][k], pv);
            }
            for (int r = 0; r < n_loc; r++) {
                if (r == col || mat[r][col] == 0) {
                    continue;
                }
                const int f = mat[r][col];
                for (int k = col; k <= n_loc; k++) {
                    mat[r][k] ^= (uint8_t)gf.mul(f, mat[col][k]);
                }
            }
        }
        if (singular) {
            break;
        }
        for (int i = 0; i < n_loc; i++) {
            word63[locs[i]] ^= mat[i][n_loc];
        }
        if (rs63_syndromes(gf, word63, 63, n2t, syn)) {
            break;
        }
        status = 0;
    } while (0);
    if (status) {
        for (int i = 0; i < 63; i++) {
            word63[i] = saved[i];
        }
    }
    return status;
}

/* The ranked-erasure retry of p25p1_rs_*_soft_reliability (p25p1_check_hdu.cpp:56-77, p25p1_check_ldu.cpp:73-94,
 * p25p1_soft.cpp:83-170) for a word the hard decoder rejected: rank all symbols by (reliability, position) with parity
 * positions first, erase the n = 1..ranked weakest, first success wins.  word63: parity first, zero padded; on success
 * out_sym receives the n_data corrected data symbols.  Returns 0 / 1. */
__device__ __noinline__ int
rs63_ranked_erasure_decode(const Gf64& gf, const RsShape& sh, uint8_t* word63, const uint8_t* dr, const uint8_t* pr, int threshold,
                           int* out_sym) {
    const int n_par = sh.n_total - sh.n_data, n2t = 2 * sh.tt;
    uint8_t er[16];
    /* selection-sort the n2t weakest of all symbols by (reliability, position): same order as the reference's
     * full stable sort truncated to its first entries */
    int hits = 0;
    for (int i = 0; i < sh.n_total; i++) {
        hits += ((i < n_par) ? pr[i] : dr[i - n_par]) < threshold;
    }
    int ranked = hits > sh.tt ? hits : sh.tt;
    if (ranked > n2t) {
        ranked = n2t;
    }
    unsigned long long taken = 0;
    for (int k = 0; k < ranked; k++) {
        int best = -1, best_rel = 256;
        for (int i = 0; i < sh.n_total; i++) {
            if ((taken >> i) & 1ull) {
                continue;
            }
            const int r = (i < n_par) ? pr[i] : dr[i - n_par];
            if (r < best_rel) {
                best_rel = r;
                best = i;
            }
        }
        taken |= 1ull << best;
        er[k] = (uint8_t)best;
    }
    for (int n = 1; n <= ranked; n++) {
        if (rs63_erasure_decode(gf, sh, word63, er, n) == 0) {
            for (int i = 0; i < sh.n_data; i++) {
                out_sym[i] = word63[n_par + i];
            }
            return 0;
        }
    }
    return 1;
}

/* mode 0: DSDReedSolomon_*::decode_soft with a caller-supplied erasure list (check_and_fix_*_soft,
 *         phase1/p25p1_check_hdu.cpp:47-54, p25p1_check_ldu.cpp:46-71): hard decode, then one erasure decode.
 * mode 1: p25p1_rs_*_soft_reliability (p25p1_check_hdu.cpp:56-77, p25p1_check_ldu.cpp:73-94): rank all symbols by
 *         (reliability, position) with parity positions first (p25p1_soft.cpp:83-170), try n = 1..ranked erasures.
 * Data bits are rewritten (normalised 0/1) exactly where the reference rewrites them: always in mode 0 (its hard decoder
 * writes back even on failure), only on success in mode 1 (it works on a copy). */
__global__ void __launch_bounds__(64)
p25_rs_soft_kernel(const dsdneo_fec_tables* __restrict__ T, RsShape sh, int mode, uint8_t* data_bits, const uint8_t* parity_bits,
                   const int32_t* erasures_in, const int32_t* n_erasures_in, int erasure_pitch, const uint8_t* data_rel,
                   const uint8_t* par_rel, int threshold, uint8_t* status, int n_words) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) {
        return;
    }
    const Gf64 gf = {T->gf_exp, T->gf_log};
    const int n_par = sh.n_total - sh.n_data, n2t = 2 * sh.tt;
    uint8_t* dbits = data_bits + (size_t)w * sh.n_data * 6;
    uint8_t word[63];
    int out_sym[36];
    for (int i = 0; i < 63; i++) {
        word[i] = 0;
    }
    rs_pack_symbols(sh, dbits, parity_bits + (size_t)w * n_par * 6, word);
    int rc = rs63_hard_decode(T->gf_exp, T->gf_log, sh, word, out_sym);
    bool write_back = (rc == 0) || (mode == 0);
    if (rc != 0) {
        uint8_t er[16];
        int n_list = 0;
        if (mode == 0) {
            const int n = n_erasures_in[w];
            bool valid = (n > 0 && n <= n2t);
            unsigned long long seen = 0;
            for (int i = 0; valid && i < n; i++) {
                const int pos = erasures_in[(size_t)w * erasure_pitch + i];
                if (pos < 0 || pos >= 63 || ((seen >> pos) & 1ull)) {
                    valid = false; /* validate_erasures, ReedSolomon.hpp:584-600 */
                } else {
                    seen |= 1ull << pos;
                    er[i] = (uint8_t)pos;
                }
            }
            n_list = valid ? n : 0;
            if (n_list > 0 && rs63_erasure_decode(gf, sh, word, er, n_list) == 0) {
                rc = 0;
                for (int i = 0; i < sh.n_data; i++) {
                    out_sym[i] = word[n_par + i];
                }
            }
        } else {
            const uint8_t* dr = data_rel + (size_t)w * sh.n_data;
            const uint8_t* pr = par_rel + (size_t)w * n_par;
            if (rs63_ranked_erasure_decode(gf, sh, word, dr, pr, threshold, out_sym) == 0) {
                rc = 0;
                write_back = true;
            }
        }
    }
    if (write_back) {
        for (int i = 0; i < sh.n_data; i++) {
            for (int b = 0; b < 6; b++) {
                dbits[6 * i + b] = (uint8_t)((out_sym[i] >> (5 - b)) & 1);
            }
        }
    }
    status[w] = (uint8_t)rc;
}

/* ------------------------------------------------------------------ P25 word codes: Golay(24,6/12), Hamming(10,6,3), BCH(63,16,11) */

__device__ __forceinline__ unsigned
g23_syndrome(unsigned cw) { /* Golay24::syndrome, include/dsd-neo/fec/Golay24.hpp:58-72 (POLY 0xAE3) */
    cw &= 0x7fffffu;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        cw = (cw & 1u) ? ((cw ^ 0xAE3u) >> 1) : (cw >> 1);
    }
    return cw << 12;
}

/* Golay24::correct (Golay24.hpp:108-169): *errs is the weight of the last syndrome examined, also on failure. */
__device__ __noinline__ unsigned
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
            *errs = __popc(s & 0x7fffffu);
            if (*errs <= w) {
                cw ^= s;
                /* rotate right by i within 23 bits */
                cw &= 0x7fffffu;
                return i ? (((cw >> i) | (cw << (23 - i))) & 0x7fffffu) : cw;
            }
            cw = ((cw & 0x400000u) ? ((cw << 1) | 1u) : (cw << 1)) & 0x7fffffu;
            s = g23_syndrome(cw);
        }
        j++;
    }
    return saver;
}

/* code 0/1: DSDGolay24::decode_6 / decode_12 (Golay24.hpp:336-405) = check_and_fix_golay_24_6 / _24_12
 * (phase1/p25p1_check_hdu.cpp:26-36): status 0 ok / 1 uncorrectable or non-binary input (data untouched), fixed = errs.
 * code 2: hamming_10_6_3_decode (src/fec/hamming_10_6_3.cpp:14-105): status 0 clean / 1 corrected / 2 uncorrectable. */
__global__ void
p25_word_decode_kernel(int code, uint8_t* data_bits, const uint8_t* parity_bits, uint8_t* status, int32_t* fixed, int n_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_words) {
        return;
    }
    if (code == DSDNEO_P25_WORD_HAMMING_10_6_3) {
        uint8_t* d = data_bits + (size_t)i * 6;
        const uint8_t* p = parity_bits + (size_t)i * 4;
        unsigned v = 0;
        bool binary = true;
        for (int k = 0; k < 6; k++) {
            binary = binary && d[k] <= 1;
            v = (v << 1) | (d[k] & 1u);
        }
        for (int k = 0; k < 4; k++) {
            binary = binary && p[k] <= 1;
            v = (v << 1) | (p[k] & 1u);
        }
        int st = 2;
        if (binary) {
            const int syn = ((__popc(v & 0x398u) & 1) << 3) | ((__popc(v & 0x354u) & 1) << 2) | ((__popc(v & 0x2E2u) & 1) << 1)
                            | (__popc(v & 0x1E1u) & 1);
            /* bad_bit_table {-2,0,1,5,2,-1,-1,6,3,-1,-1,7,4,8,9,-1} as nibbles, 0xF = uncorrectable */
            const int b = (int)((0xF9847FF36FF2510Full >> (4 * syn)) & 0xFull);
            if (syn == 0) {
                st = 0;
            } else if (b != 0xF) {
                st = 1;
                if (b >= 4) {
                    v ^= 1u << b;
                }
                for (int k = 0; k < 6; k++) {
                    d[k] = (uint8_t)((v >> (9 - k)) & 1u);
                }
            }
        }
        status[i] = (uint8_t)st;
        if (fixed) {
            fixed[i] = st == 1 ? 1 : 0;
        }
        return;
    }
    const int length = (code == DSDNEO_P25_WORD_GOLAY_24_6) ? 6 : 12;
    uint8_t* w = data_bits + (size_t)i * length;
    const uint8_t* p = parity_bits + (size_t)i * 12;
    bool binary = true;
    unsigned cw = 0;
    for (int k = 0; k < 12; k++) {
        binary = binary && p[11 - k] <= 1;
        cw = (cw << 1) | (p[11 - k] & 1u);
    }
    for (int k = 0; k < length; k++) {
        binary = binary && w[length - 1 - k] <= 1;
        cw = (cw << 1) | (w[length - 1 - k] & 1u);
    }
    cw <<= (12 - length);
    int errs = 0, st = 1;
    if (binary) {
        const unsigned pbit = cw & 0x800000u;
        cw = g23_correct(cw & ~0x800000u, &errs) | pbit;
        const int odd = __popc(cw & 0xffffffu) & 1;
        if (!(odd && (cw & 0x3fu) != 0)) {
            st = 0;
            unsigned mask = 1u << (12 - length);
            for (int k = 0; k < length; k++, mask <<= 1) {
                w[k] = (cw & mask) ? 1 : 0;
            }
        }
    }
    status[i] = (uint8_t)st;
    if (fixed) {
        fixed[i] = errs;
    }
}

/* BCH_63_16_11::decode_with_result (include/dsd-neo/fec/BCH_63_16.hpp:288-329), the P25 NID code: same Berlekamp
 * bookkeeping as the RS(63,k) decoder above with t = 11 and binary error values.  `recd` bit j = coefficient j (input bit i of
 * the reference's byte-per-bit array is coefficient 62 - i); corrected in place on success.  Returns success; *count_out =
 * corrected bits (0 on failure). */
__device__ __forceinline__ int
bch_63_16_decode_word(const dsdneo_fec_tables* __restrict__ T, unsigned long long& recd, int* count_out) {
    constexpr int NN = 63, TT = 11, N2T = 22;
    const signed char* EXP = T->gf_exp;
    const signed char* LOG = T->gf_log;
    signed char s[N2T + 1];
    int has_err = 0;
    for (int i = 1; i <= N2T; i++) {
        int syn = 0;
        for (int j = 0; j < NN; j++) {
            if ((recd >> j) & 1ull) {
                syn ^= EXP[(i * j) % NN];
            }
        }
        has_err |= syn;
        s[i] = LOG[syn];
    }
    int ok = 1, count = 0;
    if (has_err) {
        signed char elp[N2T + 2][N2T];
        signed char d[N2T + 2], l[N2T + 2], u_lu[N2T + 2];
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
                l[u + 1] = (signed char)(l[u] > cand ? l[u] : cand);
                for (int i = 0; i < N2T; i++) {
                    elp[u + 1][i] = 0;
                }
                for (int i = 0; i <= l[q]; i++) {
                    if (elp[q][i] != -1) {
                        elp[u + 1][i + u - q] = EXP[(d[u] + NN - d[q] + elp[q][i]) % NN];
                    }
                }
                for (int i = 0; i <= l[u]; i++) {
                    elp[u + 1][i] ^= elp[u][i];
                }
            }
            for (int i = 0; i <= l[u]; i++) { /* index_elp_row(u): polynomial -> index form, in place */
                if (elp[u][i] >= 0) {
                    elp[u][i] = LOG[elp[u][i]];
                }
            }
            u_lu[u + 1] = (signed char)(u - l[u + 1]);
            if (u < N2T) {
                int disc = (s[u + 1] != -1) ? EXP[s[u + 1]] : 0;
                for (int i = 1; i <= l[u + 1]; i++) {
                    if (s[u + 1 - i] != -1 && elp[u + 1][i] != 0) {
                        disc ^= EXP[(s[u + 1 - i] + LOG[elp[u + 1][i]]) % NN];
                    }
                }
                d[u + 1] = LOG[disc];
            }
        } while (u < N2T && l[u + 1] <= TT);
        u++;
        if (l[u] > TT) {
            ok = 0;
        } else {
            const int deg = l[u];
            int reg[TT + 1];
            for (int i = 1; i <= deg; i++) {
                reg[i] = elp[u][i] >= 0 ? LOG[elp[u][i]] : elp[u][i];
            }
            unsigned long long flips = 0;
            for (int i = 1; i <= NN; i++) {
                int q = 1;
                for (int j = 1; j <= deg; j++) {
                    if (reg[j] != -1) {
                        reg[j] = (reg[j] + j) % NN;
                        q ^= EXP[reg[j]];
                    }
                }
                if (q == 0) {
                    if (count >= TT) {
                        break;
                    }
                    flips |= 1ull << (NN - i);
                    count++;
                }
            }
            if (count != deg) {
                ok = 0;
                count = 0;
            } else {
                recd ^= flips;
            }
        }
    }
    *count_out = ok ? count : 0;
    return ok;
}

__global__ void __launch_bounds__(64)
bch_63_16_kernel(const dsdneo_fec_tables* __restrict__ T, const uint8_t* in63, uint8_t* out16, uint8_t* ok_out, int32_t* err_count,
                 int n_words) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) {
        return;
    }
    constexpr int NN = 63;
    const uint8_t* in = in63 + (size_t)w * 63;
    unsigned long long recd = 0; /* bit j = coefficient j */
    for (int i = 0; i < NN; i++) {
        recd |= (unsigned long long)(in[NN - 1 - i] ? 1 : 0) << i;
    }
    int count = 0;
    const int ok = bch_63_16_decode_word(T, recd, &count);
    ok_out[w] = (uint8_t)ok;
    if (err_count) {
        err_count[w] = ok ? count : 0;
    }
    if (ok) {
        uint8_t* o = out16 + (size_t)w * 16;
        for (int i = 0; i < 16; i++) {
            o[i] = (uint8_t)((recd >> (NN - 1 - i)) & 1ull);
        }
    }
}

/* ------------------------------------------------------------------ P25 Phase 1 NID decode (p25p1_nid_decode) */

struct NidDecoded {
    int status, nac, duid, errs;
};

/* decode_nid_codeword (src/protocol/p25/phase1/p25p1_check_nid.cpp:250-303); word bit (62 - i) = reference bit i */
__device__ __forceinline__ NidDecoded
nid_codeword(const dsdneo_fec_tables* __restrict__ T, unsigned long long word, int parity, int* bch_failed) {
    NidDecoded r = {0, 0, 0, 0};
    int count = 0;
    *bch_failed = 0;
    if (!bch_63_16_decode_word(T, word, &count)) {
        *bch_failed = 1;
        return r;
    }
    r.errs = count;
    r.nac = (int)((word >> 51) & 0xFFFull);
    r.duid = (int)((word >> 47) & 0xFull);
    /* TIA-102.BAAA-A Table 8-4: HDU 0, TDU 3, LDU1 5, TSDU 7, LDU2 A, PDU C, TDULC F */
    if (!((0x94A9u >> r.duid) & 1u)) {
        r.errs = 0;
        return r;
    }
    const int want_parity = (r.duid == 0x5 || r.duid == 0xA) ? 1 : 0;
    r.status = (want_parity == parity) ? 1 : 2;
    return r;
}

/* p25p1_nid_decode (p25p1_check_nid.cpp:322-354): hard decode, one retry with the known NAC after a BCH failure, then the
 * bounded Chase search over the (at most 8) least reliable positions.  One warp per 32 NIDs: every lane hard-decodes its own
 * NID (the common case ends there); the NIDs that still fail are then searched one after the other by the whole warp -- the
 * lanes rank the 63 reliabilities, stride over the <= 2 x 256 flip masks, each running its own BCH decodes, and the winner is
 * the minimum of the packed key {score, not-ok, corrections, flips, order of enumeration}: the reference's replacement rule
 * is exactly that lexicographic order with "first found" breaking ties. */
__global__ void __launch_bounds__(128)
p25p1_nid_decode_kernel(const dsdneo_fec_tables* __restrict__ T, const uint8_t* code63, const uint8_t* reliab63,
                        const int32_t* observed_nac, const uint8_t* parity_in, const uint8_t* parity_reliab, int threshold,
                        int8_t* status_out, int32_t* nac_out, uint8_t* duid_out, int32_t* errs_out, int n) {
    const int lane = threadIdx.x & 31;
    const int warp_first = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    if (warp_first >= n) {
        return;
    }
    const int w = warp_first + lane;
    const bool valid = w < n;
    unsigned long long word = 0;
    int parity = 0, prel = 0, obs = 0;
    if (valid) {
        const uint8_t* in = code63 + (size_t)w * 63;
        for (int i = 0; i < 63; i++) {
            word |= (unsigned long long)(in[i] ? 1 : 0) << (62 - i);
        }
        parity = parity_in[w] ? 1 : 0;
        prel = parity_reliab ? parity_reliab[w] : 0;
        obs = observed_nac ? observed_nac[w] : 0;
    }
    const bool nac_ok = obs > 0 && obs <= 0xFFF && obs != 0xFFF;
    const int rx_nac = (int)((word >> 51) & 0xFFFull);
    const unsigned long long retry = (word & ~(0xFFFull << 51)) | ((unsigned long long)(obs & 0xFFF) << 51);
    const bool have_retry = nac_ok && rx_nac != obs;

    NidDecoded res = {0, 0, 0, 0};
    if (valid) { /* decode_nid_hard (:305-320) */
        int failed = 0;
        res = nid_codeword(T, word, parity, &failed);
        if (res.status == 0 && failed && have_retry) {
            res = nid_codeword(T, retry, parity, &failed);
        }
    }
    unsigned todo = __ballot_sync(0xffffffffu, valid && res.status <= 0 && reliab63 != nullptr);
    __shared__ uint8_t s_order[4][64];
    uint8_t* order = s_order[threadIdx.x >> 5];
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned long long c_word = __shfl_sync(0xffffffffu, word, src);
        const unsigned long long c_retry = __shfl_sync(0xffffffffu, retry, src);
        const int c_parity = __shfl_sync(0xffffffffu, parity, src);
        const int c_prel = __shfl_sync(0xffffffffu, prel, src);
        const bool c_have_retry = __shfl_sync(0xffffffffu, have_retry ? 1 : 0, src) != 0;
        const uint8_t* rel = reliab63 + (size_t)(warp_first + src) * 63;
        /* rank of position i in the (reliability, index) order = number of positions that precede it */
        __syncwarp();
        for (int i = lane; i < 63; i += 32) {
            const int ri = rel[i];
            int rank = 0;
            for (int j = 0; j < 63; j++) {
                const int rj = rel[j];
                rank += (rj < ri || (rj == ri && j < i)) ? 1 : 0;
            }
            order[rank] = (uint8_t)i;
        }
        __syncwarp();
        /* build_soft_nid_pool (:123-153), every lane redundantly (at most 63 + 63 steps) */
        int pool[8], n_pool = 0;
        unsigned long long picked = 0;
        for (int i = 0; i < 63 && n_pool < 8; i++) {
            if ((int)rel[order[i]] < threshold) {
                pool[n_pool++] = order[i];
                picked |= 1ull << i;
            }
        }
        for (int i = 0; i < 63 && n_pool < 6; i++) {
            if (!((picked >> i) & 1ull)) {
                pool[n_pool++] = order[i];
            }
        }
        unsigned best_key = 0xffffffffu;
        NidDecoded best = {0, 0, 0, 0};
        const int n_masks = 1 << n_pool;
        const int n_seq = n_pool > 0 ? (c_have_retry ? 2 : 1) * n_masks : 0;
        for (int seq = lane; seq < n_seq; seq += 32) {
            const int mask = seq & (n_masks - 1);
            if (__popc(mask) > 3) {
                continue;
            }
            unsigned long long cand = (seq >= n_masks) ? c_retry : c_word;
            int score = 0;
            for (int bq = 0; bq < n_pool; bq++) {
                if (mask & (1 << bq)) {
                    cand ^= 1ull << (62 - pool[bq]);
                    score += rel[pool[bq]];
                }
            }
            const int weight = __popc(mask);
            if (weight > 0 && score > threshold * weight) { /* candidate_flip_allowed (:108-117) */
                continue;
            }
            int failed = 0;
            const NidDecoded d = nid_codeword(T, cand, c_parity, &failed);
            if (d.status <= 0) {
                continue;
            }
            if (d.status == 2) {
                score += c_prel;
            }
            /* score <= 3 * 255 + 255 (11 bits), not-ok 1 bit, corrections <= 11 (4 bits), flips <= 3 (2 bits), seq < 512 */
            const unsigned key = ((unsigned)score << 17) | ((d.status == 1 ? 0u : 1u) << 16) | ((unsigned)d.errs << 12)
                                 | ((unsigned)weight << 10) | (unsigned)seq;
            if (key < best_key) {
                best_key = key;
                best = d;
            }
        }
        unsigned win = best_key;
        for (int o = 16; o > 0; o >>= 1) {
            win = min(win, __shfl_xor_sync(0xffffffffu, win, o));
        }
        if (win != 0xffffffffu) {
            const int from = __ffs(__ballot_sync(0xffffffffu, best_key == win)) - 1;
            const int b_status = __shfl_sync(0xffffffffu, best.status, from);
            const int b_nac = __shfl_sync(0xffffffffu, best.nac, from);
            const int b_duid = __shfl_sync(0xffffffffu, best.duid, from);
            const int b_errs = __shfl_sync(0xffffffffu, best.errs, from);
            if (lane == src) {
                res.status = b_status;
                res.nac = b_nac;
                res.duid = b_duid;
                res.errs = b_errs;
            }
        }
    }
    if (valid) {
        status_out[w] = (int8_t)res.status;
        nac_out[w] = res.nac;
        duid_out[w] = (uint8_t)res.duid;
        errs_out[w] = res.errs;
    }
}

/* ------------------------------------------------------------------ Golay(24,6) / (24,12) soft decode */

/* check_and_fix_golay_24_6_soft / _24_12_soft (src/protocol/p25/phase1/p25p1_soft.cpp:477-593), one thread per word, on the
 * packed codeword of DSDGolay24 (parity[i] = bit 12 + i, data[i] = bit 12 - length + i): hard decode as seed, then every
 * combination of at most 4 flips among the 8 least reliable bits, each Golay-decoded and re-encoded; lowest summed
 * reliability of changed bits wins, then fewer changed bits; hard-correction precedence as in the reference. */
__device__ __noinline__ int
golay24_soft_word(unsigned orig, const int (&rel)[24], int length, int hard_override, int threshold, unsigned& result_out, int& fixed_out) {
    const int n = length + 12;
    auto bit_of = [&](int idx) { return idx < length ? (12 - length + idx) : (12 + idx - length); };
    const unsigned data_mask = 0xfffu & ~((1u << (12 - length)) - 1u);
    auto penalty = [&](unsigned diff, int& count) {
        int pen = 0;
        count = __popc(diff);
        for (int i = 0; i < n; i++) {
            pen += ((diff >> bit_of(i)) & 1u) ? rel[i] : 0;
        }
        return pen;
    };
    /* decode one packed candidate: returns false if uncorrectable; else the decoded data field and the diff vs orig of
     * {decoded data, re-encoded parity} */
    auto try_word = [&](unsigned cw, unsigned& dec_data, unsigned& diff, int& errs) {
        const unsigned pbit = cw & 0x800000u;
        cw = g23_correct(cw & ~0x800000u, &errs) | pbit;
        const int odd = __popc(cw & 0xffffffu) & 1;
        if (odd && (cw & 0x3fu) != 0) {
            return false;
        }
        dec_data = cw & data_mask;
        unsigned enc = g23_syndrome(dec_data) | dec_data; /* Golay24::golay + overall parity = Golay24::encode */
        if (__popc(enc & 0xffffffu) & 1) {
            enc ^= 0x800000u;
        }
        diff = ((dec_data ^ orig) & data_mask) | ((enc ^ orig) & 0xfff000u);
        return true;
    };
    int best_pen = 999999, best_fixed = 0, found = 0, hard_valid = 0, hard_corrected = 0, hard_pen = 999999, hard_fixed = 0;
    unsigned best = 0, hard = 0;
    {
        unsigned dd, diff;
        int errs = 0;
        if (try_word(orig, dd, diff, errs)) {
            int cnt;
            hard_fixed = errs;
            hard_valid = 1;
            hard_corrected = errs > 0;
            hard_pen = penalty(diff, cnt);
            best_pen = hard_pen;
            best_fixed = cnt;
            hard = dd;
            best = dd;
            found = 1;
        }
    }
    int order[24], least[8], n_least = 0;
    for (int i = 0; i < n; i++) {
        int rank = 0;
        for (int j = 0; j < n; j++) {
            rank += (rel[j] < rel[i] || (rel[j] == rel[i] && j < i)) ? 1 : 0;
        }
        order[rank] = i;
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
    unsigned flip_bit[8];
    for (int b = 0; b < 8; b++) {
        flip_bit[b] = 1u << bit_of(least[b]);
    }
    for (int mask = 0; mask < 256; mask++) {
        if (__popc(mask) > 4) {
            continue;
        }
        unsigned cand = orig;
        for (int b = 0; b < 8; b++) {
            if (mask & (1 << b)) {
                cand ^= flip_bit[b];
            }
        }
        unsigned dd, diff;
        int errs = 0;
        if (!try_word(cand, dd, diff, errs)) {
            continue;
        }
        int cnt;
        const int pen = penalty(diff, cnt);
        if (pen < best_pen || (pen == best_pen && cnt < best_fixed)) {
            best_pen = pen;
            best_fixed = cnt;
            best = dd;
            found = 1;
        }
    }
    int st = 1, fx = 0;
    unsigned result = 0;
    if (found) {
        st = 0;
        if (hard_valid && hard_corrected && best != hard && (!hard_override || best_pen + 8 >= hard_pen)) {
            result = hard;
            fx = hard_fixed;
        } else {
            result = best;
            fx = best_fixed;
        }
    }
    result_out = result;
    fixed_out = fx;
    return st;
}

/* DSDGolay24::decode_6 / decode_12 (Golay24.hpp:336-405) on the packed codeword (parity[i] = bit 12 + i, data[i] = bit
 * 12 - length + i): returns 0 ok / 1 uncorrectable; cw_out = corrected word (valid when 0), errs = the decoder's count. */
__device__ __forceinline__ int
golay24_hard_word(unsigned cw, unsigned& cw_out, int& errs) {
    const unsigned pbit = cw & 0x800000u;
    cw = g23_correct(cw & ~0x800000u, &errs) | pbit;
    const int odd = __popc(cw & 0xffffffu) & 1;
    cw_out = cw;
    return (odd && (cw & 0x3fu) != 0) ? 1 : 0;
}

__global__ void
p25_golay_soft_kernel(int length, uint8_t* data_bits, const uint8_t* parity_bits, const int32_t* reliab, int hard_override,
                      int threshold, uint8_t* status, int32_t* fixed, int n_words) {
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n_words) {
        return;
    }
    const int n = length + 12;
    uint8_t* d = data_bits + (size_t)wi * length;
    const uint8_t* pb = parity_bits + (size_t)wi * 12;
    const int32_t* rin = reliab + (size_t)wi * n;
    int rel[24];
    auto bit_of = [&](int idx) { return idx < length ? (12 - length + idx) : (12 + idx - length); };
    unsigned orig = 0;
    bool binary = true;
    for (int i = 0; i < 24; i++) {
        rel[i] = 0;
    }
    for (int i = 0; i < n; i++) {
        const int r = rin[i];
        rel[i] = r < 0 ? 0 : (r > 255 ? 255 : r);
        const unsigned b = i < length ? d[i] : pb[i - length];
        binary = binary && b <= 1;
        orig |= (b & 1u) << bit_of(i);
    }
    if (!binary) { /* word_bits_are_valid fails: every decode in the reference returns 1 and nothing is found */
        status[wi] = 1;
        fixed[wi] = 0;
        return;
    }
    unsigned result = 0;
    int fx = 0;
    const int st = golay24_soft_word(orig, rel, length, hard_override, threshold, result, fx);
    if (st == 0) {
        for (int i = 0; i < length; i++) {
            d[i] = (uint8_t)((result >> (12 - length + i)) & 1u);
        }
    }
    status[wi] = (uint8_t)st;
    fixed[wi] = fx;
}

/* ------------------------------------------------------------------ Hamming(10,6,3) soft decode */

/* hamming_10_6_3_decode (src/fec/hamming_10_6_3.cpp:14-105) on a packed word v (bit 9 - i = reference bit i): returns the
 * status (0 clean / 1 corrected / 2 uncorrectable) and leaves the corrected DATA bits in v (parity bits untouched). */
__device__ __forceinline__ int
ham1063_hard(unsigned& v) {
    const int syn = ((__popc(v & 0x398u) & 1) << 3) | ((__popc(v & 0x354u) & 1) << 2) | ((__popc(v & 0x2E2u) & 1) << 1)
                    | (__popc(v & 0x1E1u) & 1);
    if (syn == 0) {
        return 0;
    }
    const int b = (int)((0xF9847FF36FF2510Full >> (4 * syn)) & 0xFull);
    if (b == 0xF) {
        return 2;
    }
    if (b >= 4) {
        v ^= 1u << b;
    }
    return 1;
}

/* hamming_10_6_3_soft (src/protocol/p25/phase1/p25p1_soft.cpp:444-475) on a packed word (bit 9 - i = reference bit i) with
 * per-bit reliabilities already clamped to 0..255: returns 0 unchanged / 1 corrected / 2 failed, `result` = output word. */
__device__ __noinline__ int
ham1063_soft_word(unsigned orig, const int (&rel)[10], int hard_override, int threshold, unsigned& result_out) {
    auto penalty = [&](unsigned diff) {
        int p = 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            p += ((diff >> (9 - i)) & 1u) ? rel[i] : 0;
        }
        return p;
    };
    int best_pen = 999999, best_flips = 99, found = 0, hard_valid = 0, hard_corrected = 0, hard_pen = 999999;
    unsigned best = 0, hard = 0;
    {
        unsigned v = orig;
        const int rc = ham1063_hard(v);
        if (rc != 2) {
            const unsigned d = v >> 4; /* data bits d0..d5 = bits 5..0 of d */
            const unsigned d0 = (d >> 5) & 1u, d1 = (d >> 4) & 1u, d2 = (d >> 3) & 1u, d3 = (d >> 2) & 1u, d4 = (d >> 1) & 1u, d5 = d & 1u;
            hard = (d << 4) | ((d0 ^ d1 ^ d2 ^ d5) << 3) | ((d0 ^ d1 ^ d3 ^ d5) << 2) | ((d0 ^ d2 ^ d3 ^ d4) << 1) | (d1 ^ d2 ^ d3 ^ d4);
            hard_valid = 1;
            hard_corrected = rc == 1;
            hard_pen = penalty(hard ^ orig);
            best_pen = hard_pen;
            best_flips = __popc(hard ^ orig);
            best = hard;
            found = 1;
        }
    }
    /* find_k_least_reliable (:174-206): rank by (reliability, index); positions under the threshold first */
    int least[5], n_least = 0;
    int order[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 10; j++) {
            rank += (rel[j] < rel[i] || (rel[j] == rel[i] && j < i)) ? 1 : 0;
        }
        order[rank] = i;
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
        if (__popc(mask) > 2) {
            continue;
        }
        unsigned cand = orig;
        for (int b = 0; b < 5; b++) {
            if (mask & (1 << b)) {
                cand ^= 1u << (9 - least[b]);
            }
        }
        unsigned v = cand;
        if (ham1063_hard(v) != 0) {
            continue;
        }
        const int pen = penalty(cand ^ orig);
        const int flips = __popc(mask);
        if (pen < best_pen || (pen == best_pen && flips < best_flips)) {
            best_pen = pen;
            best_flips = flips;
            best = cand;
            found = 1;
        }
    }
    unsigned result = orig;
    int st = 2;
    if (found) {
        if (hard_valid && hard_corrected && best != hard && (!hard_override || best_pen + 8 >= hard_pen)) {
            result = hard;
            st = 1;
        } else {
            result = best;
            st = (best == orig) ? 0 : 1;
        }
    }
    result_out = result;
    return st;
}

/* one thread per word */
__global__ void
hamming_10_6_3_soft_kernel(const uint8_t* bits10, const int32_t* reliab10, int hard_override, int threshold, uint8_t* out10,
                           uint8_t* status, int n_words) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) {
        return;
    }
    const uint8_t* in = bits10 + (size_t)w * 10;
    const int32_t* rin = reliab10 + (size_t)w * 10;
    int rel[10];
    unsigned orig = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int r = rin[i];
        rel[i] = r < 0 ? 0 : (r > 255 ? 255 : r);
        orig = (orig << 1) | (in[i] & 1u);
    }
    unsigned result = orig;
    const int st = ham1063_soft_word(orig, rel, hard_override, threshold, result);
    uint8_t* o = out10 + (size_t)w * 10;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        o[i] = (uint8_t)((result >> (9 - i)) & 1u);
    }
    status[w] = (uint8_t)st;
}

/* ------------------------------------------------------------------ P25 Phase 1 frame cutter (status-symbol stripping) */

/* Index (from the first sync dibit) of the k-th non-status dibit at or after frame offset `first`: the air interface inserts
 * one status symbol after every 35 dibits, i.e. at frame offsets 35, 71, 107, ...  (the reference counts the same thing with
 * `skipdibit`, src/protocol/p25/phase1/p25p1_tsbk.c:135-152 with skipdibit = 36 - 14 at offset 57, :1054). */
__device__ __forceinline__ int
p25p1_payload_offset(int first, int k) {
    /* status symbols below offset `first`: first / 36 */
    const int ordinal = first - first / 36 + k; /* 0-based ordinal among the non-status dibits of the frame */
    return ordinal + ordinal / 35;
}

/* One thread block per (channel, hit) slot.  NID: the 32 dibits after the sync with the status symbol at frame offset 35
 * dropped (p25p1_read_nid_fields, src/engine/dispatch/dispatch_p25p1.c:121-143): 63 BCH bits + reliabilities min(|llr|, 255)
 * (:59-66) + the parity bit.  Payload: n_payload non-status dibits from frame offset 57 on, with their LLR pairs. */
__global__ void __launch_bounds__(128)
p25p1_frame_cut_kernel(const uint8_t* dibits, size_t dibit_pitch, const int16_t* llr, size_t llr_pitch, const int32_t* counts,
                       const int32_t* hits /* [ch][max_hits][2] = {position of the last sync dibit, sync type} */,
                       const int32_t* n_hits, int n_channels, int max_hits, int n_payload, uint8_t* nid_code63,
                       uint8_t* nid_reliab63, uint8_t* nid_parity, uint8_t* nid_parity_reliab, uint8_t* nid_valid,
                       uint8_t* payload_dibits, int16_t* payload_llr, uint8_t* payload_valid, int region_off) {
    const int slot = blockIdx.x;
    const int ch = slot / max_hits, h = slot - ch * max_hits;
    if (ch >= n_channels) {
        return;
    }
    const int tid = threadIdx.x;
    const bool present = h < min(n_hits[ch], max_hits);
    const int count = counts[ch];
    const long start = present ? (long)hits[((size_t)ch * max_hits + h) * 2] + region_off - 23 : 0; /* first sync dibit */
    const uint8_t* d = dibits + (size_t)ch * dibit_pitch;
    const int16_t* l = llr + (size_t)ch * llr_pitch * 2;
    const bool nid_ok = present && start >= 0 && start + 57 <= count;
    const int last_payload = p25p1_payload_offset(57, n_payload - 1);
    const bool pay_ok = nid_ok && n_payload > 0 && start + last_payload + 1 <= count;
    if (tid == 0) {
        nid_valid[slot] = nid_ok ? 1 : 0;
        payload_valid[slot] = pay_ok ? 1 : 0;
    }
    if (tid < 32) {
        /* NID dibit k: frame offset 24 + k, skipping the status symbol at 35 */
        const int off = 24 + tid + (tid >= 11 ? 1 : 0);
        int dib = 0, l0 = 0, l1 = 0;
        if (nid_ok) {
            dib = d[start + off];
            l0 = l[(start + off) * 2];
            l1 = l[(start + off) * 2 + 1];
        }
        const int r0 = min(abs(l0), 255), r1 = min(abs(l1), 255);
        uint8_t* code = nid_code63 + (size_t)slot * 63;
        uint8_t* rel = nid_reliab63 + (size_t)slot * 63;
        code[2 * tid] = (uint8_t)((dib >> 1) & 1);
        rel[2 * tid] = (uint8_t)r0;
        if (tid < 31) {
            code[2 * tid + 1] = (uint8_t)(dib & 1);
            rel[2 * tid + 1] = (uint8_t)r1;
        } else {
            nid_parity[slot] = (uint8_t)(dib & 1);
            nid_parity_reliab[slot] = (uint8_t)r1;
        }
    }
    for (int k = tid; k < n_payload; k += blockDim.x) {
        int dib = 0, l0 = 0, l1 = 0;
        if (pay_ok) {
            const long pos = start + p25p1_payload_offset(57, k);
            dib = d[pos];
            l0 = l[pos * 2];
            l1 = l[pos * 2 + 1];
        }
        payload_dibits[(size_t)slot * n_payload + k] = (uint8_t)dib;
        payload_llr[((size_t)slot * n_payload + k) * 2] = (int16_t)l0;
        payload_llr[((size_t)slot * n_payload + k) * 2 + 1] = (int16_t)l1;
    }
}

/* ------------------------------------------------------------------ DMR BS data burst cutter */

/* dmr_data_sync's collection phase (src/protocol/dmr/dmr_data.c:54-65,118-157,159-179,218-226,261-268) for every BS DATA
 * sync hit of every channel: 90 dibits back from the dibit after the sync -- 12 CACH dibits de-interleaved with
 * dmr_cach_interleave (src/protocol/dmr/dmr_cach.c:9-11), 49 info dibits, 5 slot-type dibits, the sync -- then 5 slot-type
 * and 49 info dibits after it.  One thread block per (channel, hit) slot, one thread per dibit. */
__global__ void __launch_bounds__(128)
dmr_burst_cut_kernel(const uint8_t* dibits, size_t dibit_pitch, const uint8_t* reliab, size_t rel_pitch, const int32_t* counts,
                     const int32_t* hits, const int32_t* n_hits, int n_channels, int max_hits, int inverted, uint8_t* cach24,
                     uint8_t* info196, uint8_t* rel98, uint8_t* slot20, uint8_t* valid_out) {
    const int slot = blockIdx.x;
    const int ch = slot / max_hits, h = slot - ch * max_hits;
    if (ch >= n_channels) {
        return;
    }
    const int t = threadIdx.x;
    const bool present = h < min(n_hits[ch], max_hits);
    const long live = present ? (long)hits[((size_t)ch * max_hits + h) * 2] + 1 : 0;
    const long start = live - 90;
    const bool ok = present && start >= 0 && live + 54 <= counts[ch];
    if (t == 0) {
        valid_out[slot] = ok ? 1 : 0;
    }
    /* burst-relative dibit index: 0..11 CACH, 12..60 info, 61..65 slot type, 66..89 sync, 90..94 slot type, 95..143 info */
    for (int k = t; k < 144; k += blockDim.x) {
        if (k >= 66 && k < 90) {
            continue;
        }
        int d = 0, r = 0;
        if (ok) {
            d = dibits[(size_t)ch * dibit_pitch + start + k];
            r = reliab[(size_t)ch * rel_pitch + start + k];
            if (inverted && k < 90) {
                d ^= 2;
            }
        }
        const uint8_t b1 = (uint8_t)((d >> 1) & 1), b0 = (uint8_t)(d & 1);
        if (k < 12) {
            /* dmr_cach_interleave {0,7,8,9,1,10,11,12,2,13,14,15,3,16,4,17,18,19,5,20,21,22,6,23}, 5 bits per entry */
            const unsigned long long lo = 0x07B9A262D414A0E0ull; /* entries 0..11 */
            const unsigned long long hi = 0x0B9AD5A167289203ull; /* entries 12..23 */
            const int e0 = 2 * k, e1 = 2 * k + 1;
            const int i0 = (int)(((e0 < 12 ? lo : hi) >> (5 * (e0 % 12))) & 31ull);
            const int i1 = (int)(((e1 < 12 ? lo : hi) >> (5 * (e1 % 12))) & 31ull);
            cach24[(size_t)slot * 24 + i0] = b1;
            cach24[(size_t)slot * 24 + i1] = b0;
        } else if (k < 61) {
            const int i = k - 12;
            info196[(size_t)slot * 196 + 2 * i] = b1;
            info196[(size_t)slot * 196 + 2 * i + 1] = b0;
            rel98[(size_t)slot * 98 + i] = (uint8_t)r;
        } else if (k < 66) {
            const int i = k - 61;
            slot20[(size_t)slot * 20 + 2 * i] = b1;
            slot20[(size_t)slot * 20 + 2 * i + 1] = b0;
        } else if (k < 95) {
            const int i = k - 90;
            slot20[(size_t)slot * 20 + 10 + 2 * i] = b1;
            slot20[(size_t)slot * 20 + 11 + 2 * i] = b0;
        } else {
            const int i = k - 95;
            info196[(size_t)slot * 196 + 98 + 2 * i] = b1;
            info196[(size_t)slot * 196 + 99 + 2 * i] = b0;
            rel98[(size_t)slot * 98 + 49 + i] = (uint8_t)r;
        }
    }
}

/* ------------------------------------------------------------------ K = 5 convolutional decoders */

constexpr int kVitMaxSteps = 244; /* viterbi_history[244], src/core/util/dsd_misc.c:108 */
constexpr int kNxdnMaxSteps = 300; /* m_decisions[8 * 300], src/protocol/nxdn/nxdn_convolution.c:52 */

/* viterbi_decode / viterbi_decode_punctured (src/core/util/dsd_misc.c:118-182) with viterbi_decode_bit (:191-236) and
 * viterbi_chainback (:246-275).  One thread per frame: 16 path metrics in registers, one 16-bit decision word per step. */
__global__ void __launch_bounds__(64)
viterbi_k5_kernel(const uint16_t* in, size_t in_pitch, int in_len, const uint8_t* punct, int p_len, uint8_t* out, size_t out_pitch,
                  uint32_t* metric_out, int n_frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) {
        return;
    }
    const uint16_t* src = in + (size_t)f * in_pitch;
    uint32_t pm[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        pm[i] = 0;
    }
    uint16_t hist[kVitMaxSteps];
    /* de-puncturing: erased positions carry the neutral cost 0x7FFF (dsd_misc.c:163-177) */
    int p = 0, i_in = 0, u = 0, pos = 0;
    auto next_cost = [&](bool& ok) -> uint16_t {
        if (i_in >= in_len) {
            ok = false;
            return 0;
        }
        ok = true;
        uint16_t v;
        if (!punct || punct[p]) {
            v = src[i_in++];
        } else {
            v = 0x7FFF;
        }
        u++;
        if (punct) {
            p = (p + 1) % p_len;
        }
        return v;
    };
    for (;;) {
        bool ok0, ok1;
        const uint16_t s0 = next_cost(ok0);
        if (!ok0) {
            break;
        }
        const uint16_t s1 = next_cost(ok1);
        if (!ok1) {
            break; /* odd tail element is ignored, as `i + 1 < len` does */
        }
        if (pos >= kVitMaxSteps) {
            break;
        }
        uint32_t cm[16];
        unsigned h = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t c0 = (k >= 4) ? 0xFFFFu : 0u;
            const uint32_t c1 = ((0x66u >> k) & 1u) ? 0xFFFFu : 0u; /* {0,F,F,0,0,F,F,0} */
            const uint32_t d0 = c0 > s0 ? c0 - s0 : s0 - c0;
            const uint32_t d1 = c1 > s1 ? c1 - s1 : s1 - c1;
            const uint32_t metric = d0 + d1;
            const uint32_t m0 = pm[k] + metric, m1 = pm[k + 8] + (0x1FFFEu - metric);
            const uint32_t m2 = pm[k] + (0x1FFFEu - metric), m3 = pm[k + 8] + metric;
            if (m0 >= m1) { /* ties take the "1" predecessor */
                h |= 1u << (2 * k);
                cm[2 * k] = m1;
            } else {
                cm[2 * k] = m0;
            }
            if (m2 >= m3) {
                h |= 1u << (2 * k + 1);
                cm[2 * k + 1] = m3;
            } else {
                cm[2 * k + 1] = m2;
            }
        }
        hist[pos++] = (uint16_t)h;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            pm[k] = cm[k];
        }
    }
    const int total = u;             /* length of the de-punctured message */
    const int nbits = total / 2;
    uint8_t* o = out + (size_t)f * out_pitch;
    const int clear = (nbits - 1) / 8 + 1; /* only these bytes are cleared; later ones are OR-ed into (dsd_misc.c:252) */
    for (int k = 0; k < clear && k < (int)out_pitch; k++) {
        o[k] = 0;
    }
    unsigned state = 0;
    int bit_pos = nbits + 4;
    while (pos > 0) {
        bit_pos--;
        pos--;
        const unsigned bit = hist[pos] & (1u << (state >> 4));
        state >>= 1;
        if (bit) {
            state |= 0x80u;
            if (bit_pos / 8 < (int)out_pitch) {
                o[bit_pos / 8] |= (uint8_t)(1u << (7 - (bit_pos % 8)));
            }
        }
    }
    uint32_t best = pm[0];
#pragma unroll
    for (int k = 1; k < 16; k++) {
        best = pm[k] < best ? pm[k] : best;
    }
    metric_out[f] = best - (uint32_t)(total - in_len) * 0x7FFFu;
}

/* CNXDNConvolution_start / _decode / _decode_soft / _chainback (src/protocol/nxdn/nxdn_convolution.c:58-164).
 * metrics: [n][32] = the reference's two ping-pong arrays m_metrics1 | m_metrics2, carried from frame to frame. */
__global__ void __launch_bounds__(64)
nxdn_conv_kernel(const uint8_t* sym, const uint8_t* rel, size_t pitch, int n_steps, int n_bits_out, uint16_t* metrics, uint8_t* out,
                 size_t out_pitch, int n_frames) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) {
        return;
    }
    const uint8_t* s = sym + (size_t)f * pitch;
    const uint8_t* r = rel ? rel + (size_t)f * pitch : nullptr;
    uint16_t a[16], b[16];
    uint16_t* mio = metrics + (size_t)f * 32;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        a[i] = mio[i];
        b[i] = mio[16 + i];
    }
    unsigned short dec[kNxdnMaxSteps]; /* 16 decision bits per step (the reference keeps them in a uint64) */
    for (int t = 0; t < n_steps; t++) {
        const int s0 = s[2 * t], s1 = s[2 * t + 1];
        const bool even = (t & 1) == 0; /* even steps read m_metrics1 (a) and write m_metrics2 (b) */
        unsigned d = 0;
        uint16_t nm[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int t1 = (i >= 4) ? 2 : 0, t2 = ((0x66 >> i) & 1) ? 2 : 0;
            int d0 = t1 - s0, d1 = t2 - s1;
            d0 = d0 < 0 ? -d0 : d0;
            d1 = d1 < 0 ? -d1 : d1;
            const uint32_t oi = even ? a[i] : b[i], oj = even ? a[i + 8] : b[i + 8];
            uint32_t m0, m1, m2, m3;
            if (!r) {
                const uint16_t metric = (uint16_t)(d0 + d1);
                m0 = (uint16_t)(oi + metric);
                m1 = (uint16_t)(oj + (4u - metric));
                m2 = (uint16_t)(oi + (4u - metric));
                m3 = (uint16_t)(oj + metric);
            } else {
                uint32_t metric = ((uint32_t)d0 * r[2 * t] + (uint32_t)d1 * r[2 * t + 1]) / 128u;
                metric = metric > 8u ? 8u : metric;
                m0 = oi + metric;
                m1 = oj + (8u - metric);
                m2 = oi + (8u - metric);
                m3 = oj + metric;
            }
            const unsigned dec0 = m0 >= m1, dec1 = m2 >= m3;
            nm[2 * i] = (uint16_t)(dec0 ? m1 : m0);
            nm[2 * i + 1] = (uint16_t)(dec1 ? m3 : m2);
            d |= (dec1 << (2 * i + 1)) | (dec0 << (2 * i));
        }
        dec[t] = (unsigned short)d;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (even) {
                b[i] = nm[i];
            } else {
                a[i] = nm[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        mio[i] = a[i];
        mio[16 + i] = b[i];
    }
    uint8_t* o = out + (size_t)f * out_pitch;
    unsigned state = 0;
    int t = n_steps, nb = n_bits_out;
    while (nb-- > 0) {
        --t;
        const unsigned bit = (dec[t] >> (state >> 4)) & 1u;
        state = (bit << 7) | (state >> 1);
        const uint8_t mask = (uint8_t)(0x80u >> (nb & 7));
        o[nb >> 3] = bit ? (uint8_t)(o[nb >> 3] | mask) : (uint8_t)(o[nb >> 3] & ~mask);
    }
}

/* RAII device scratch for the *_host entry points */
struct DevBuf {
    void* p = nullptr;
    cudaError_t err = cudaSuccess;
    explicit DevBuf(size_t bytes) { err = cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { cudaFree(p); }
    template <typename U> U* as() { return static_cast<U*>(p); }
};

inline int
grid_for(int n, int block) {
    return (n + block - 1) / block;
}

int
block_code_shape(int code, int* n, int* k) {
    static const int N[8] = {7, 12, 13, 15, 16, 20, 24, 16};
    static const int K[8] = {4, 8, 9, 11, 11, 8, 12, 7};
    if (code < 0 || code >= DSDNEO_FEC_BLOCK_CODE_COUNT) {
        set_error("fec: unknown block code %d", code);
        return DSDNEO_B200_EINVAL;
    }
    *n = N[code];
    *k = K[code];
    return 0;
}


/* ------------------------------------------------------------------ P25 Phase 1 frame decoder (one warp per frame) */

/*
 * Everything the reference's frame handlers do between the frame sync and the vocoder / message parsers, for every sync hit
 * of every channel at once, straight from the slicer's dibit + LLR streams (no host step, no intermediate frame copies):
 *   processTSBK   src/protocol/p25/phase1/p25p1_tsbk.c:108-161,1051-1081   1..3 half-rate blocks, list-8 + CRC-16 pick
 *   processHDU    src/protocol/p25/phase1/p25p1_hdu.c:108-303              36 Golay(24,6) words hard + soft, RS(36,20,17)
 *   processLDU1/2 src/protocol/p25/phase1/p25p1_ldu.c:89-222, p25p1_ldu1.c:54-245, p25p1_ldu2.c:54-280
 *                                                                           9 IMBE de-interleaves, 24 Hamming(10,6,3) words hard
 *                                                                           + soft, RS(24,12,13) / (24,16,9) hard + ranked
 *                                                                           erasures, 2 LSD (16,8) words hard + soft
 * The status symbol after every 35 dibits is skipped by index arithmetic (p25p1_payload_offset).  One warp per frame: lanes
 * take one code word (or one trellis block, or one LSD word) each; lane 0 runs the Reed-Solomon decoder.
 */
struct P25FrameParams {
    const uint8_t* dibits;
    size_t dibit_pitch;
    const int16_t* llr;
    size_t llr_pitch; /* in dibits */
    const int32_t* counts;
    const int32_t* hits;   /* [ch][max_hits][2] = {position of the last sync dibit relative to region_off, sync type} */
    const int32_t* n_hits;
    int region_off;
    const long long* stream_base; /* [ch] absolute stream index of buffer index 0, or NULL */
    int n_channels, max_hits;
    const int8_t* nid_status;
    const uint8_t* nid_valid; /* optional: 0 = the NID did not fit in the stream (status forced to 0) */
    const int32_t* nid_nac;
    const uint8_t* nid_duid;
    const int32_t* nid_errs;
    int32_t* frame_off; /* [ch] first frame record of the channel (exclusive scan of min(n_hits, max_hits)) */
    int32_t* voice_off; /* [ch] first voice record of the channel */
    int32_t* totals;    /* {frames, voice records} */
    dsdneo_b200_p25p1_frame* frames;
    dsdneo_b200_p25p1_voice* voices;
    int frame_capacity, voice_capacity;
    int threshold, hard_override;
};

__device__ __forceinline__ bool
p25_is_ldu(int status, int duid) {
    return status > 0 && (duid == 0x5 || duid == 0xA);
}

__device__ __forceinline__ int
p25_slot_status(const P25FrameParams& p, int slot) {
    return (p.nid_valid && !p.nid_valid[slot]) ? 0 : p.nid_status[slot];
}

/* exclusive scans over the channels: frame records and voice records per channel (single CTA, n_channels in chunks) */
__global__ void __launch_bounds__(1024)
p25p1_frame_index_kernel(const P25FrameParams p) {
    __shared__ int s_f[1024], s_v[1024];
    __shared__ int carry_f, carry_v;
    const int t = threadIdx.x;
    if (t == 0) {
        carry_f = 0, carry_v = 0;
    }
    __syncthreads();
    for (int base = 0; base < p.n_channels; base += 1024) {
        const int c = base + t;
        int nf = 0, nv = 0;
        if (c < p.n_channels) {
            nf = min(p.n_hits[c], p.max_hits);
            for (int h = 0; h < nf; h++) {
                const int slot = c * p.max_hits + h;
                nv += p25_is_ldu(p25_slot_status(p, slot), p.nid_duid[slot]) ? 1 : 0;
            }
        }
        s_f[t] = nf, s_v[t] = nv;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) { /* Hillis-Steele inclusive scan */
            const int af = t >= o ? s_f[t - o] : 0, av = t >= o ? s_v[t - o] : 0;
            __syncthreads();
            s_f[t] += af, s_v[t] += av;
            __syncthreads();
        }
        if (c < p.n_channels) {
            p.frame_off[c] = carry_f + s_f[t] - nf;
            p.voice_off[c] = carry_v + s_v[t] - nv;
        }
        __syncthreads();
        if (t == 1023) {
            carry_f += s_f[1023], carry_v += s_v[1023];
        }
        __syncthreads();
    }
    if (t == 0) {
        p.totals[0] = carry_f, p.totals[1] = carry_v;
    }
}

/* ComputeCrcCCITT16b + crc16_ok over the first 80 bits of a TSBK against its last 16 (src/protocol/p25/p25_crc.c:11-75) */
__device__ __forceinline__ bool
p25_crc16_ok(const uint8_t* bytes12) {
    unsigned crc = 0;
    for (int i = 0; i < 80; i++) {
        const unsigned bit = (bytes12[i >> 3] >> (7 - (i & 7))) & 1u;
        crc = (((crc >> 15) & 1u) ^ bit) ? (((crc << 1) ^ 0x1021u) & 0xFFFFu) : ((crc << 1) & 0xFFFFu);
    }
    return (crc ^ 0xFFFFu) == (((unsigned)bytes12[10] << 8) | bytes12[11]);
}

/* p25_lsd_fec_16x8 (src/protocol/p25/p25_lsd.c:27-78) on a packed word (bit 15 - i = reference bit i); the reference's
 * lsd_parity[d] table is (d * x^8) mod (x^8 + x^5 + x^4 + x^3 + 1) */
__device__ __forceinline__ unsigned
lsd_parity_of(unsigned d) {
    unsigned r = d << 8;
#pragma unroll
    for (int i = 15; i >= 8; i--) {
        r ^= ((r >> i) & 1u) ? (0x139u << (i - 8)) : 0u;
    }
    return r & 0xFFu;
}

__device__ __noinline__ bool
lsd_hard(unsigned& w) {
    const unsigned synd = (w & 0xFFu) ^ lsd_parity_of(w >> 8);
    if (synd == 0) {
        return true;
    }
    if ((synd & (synd - 1)) == 0) {
        w ^= synd; /* single parity bit */
        return true;
    }
    for (int pos = 0; pos < 8; pos++) {
        if (lsd_parity_of(1u << (7 - pos)) == synd) {
            w ^= 1u << (15 - pos);
            return true;
        }
    }
    return false;
}

/* p25_lsd_fec_16x8_soft (p25_lsd.c:80-161): up to 6 bits under the erasure threshold, every flip subset, least penalty */
__device__ __noinline__ bool
lsd_soft(unsigned& w, const int (&rel)[16], int threshold) {
    if (lsd_hard(w)) {
        return true;
    }
    int cand[16], n = 0;
    for (int i = 0; i < 16; i++) {
        if (rel[i] < threshold) {
            cand[n++] = i;
        }
    }
    for (int i = 0; i < n; i++) {
        for (int j = i + 1; j < n; j++) {
            const int ri = rel[cand[i]], rj = rel[cand[j]];
            if (rj < ri || (rj == ri && cand[j] < cand[i])) {
                const int t = cand[i];
                cand[i] = cand[j];
                cand[j] = t;
            }
        }
    }
    n = n > 6 ? 6 : n;
    if (n <= 0) {
        return false;
    }
    unsigned best = 0;
    int best_pen = 999999;
    bool found = false;
    for (int mask = 1; mask < (1 << n); mask++) {
        unsigned tmp = w;
        int pen = 0;
        for (int b = 0; b < n; b++) {
            if (mask & (1 << b)) {
                tmp ^= 1u << (15 - cand[b]);
                pen += rel[cand[b]];
            }
        }
        if (pen >= best_pen) {
            continue;
        }
        if (lsd_hard(tmp)) {
            best = tmp, best_pen = pen, found = true;
        }
    }
    if (found) {
        w = best;
    }
    return found;
}

/* P25 Phase 1 IMBE interleave schedule (TIA-102.BAAA; include/dsd-neo/protocol/p25/p25p1_const.h:30-53) as the flat bit
 * index row * 23 + column of imbe_fr[8][23] for the first / second bit of each of the 72 dibits of a voice frame */
__constant__ uint8_t c_imbe_hi[72] = {22, 66, 102, 43, 87,  115, 20, 64, 100, 41, 85, 151, 18, 62, 98,  39, 83,  149,
                                      16, 60, 96,  37, 81,  147, 14, 58, 94,  35, 79, 145, 12, 56, 92,  33, 77,  143,
                                      10, 54, 128, 31, 75,  141, 8,  52, 126, 29, 73, 139, 6,  50, 124, 27, 71,  167,
                                      4,  48, 122, 25, 69,  165, 2,  46, 120, 23, 105, 163, 0,  90, 118, 67, 103, 161};
__constant__ uint8_t c_imbe_lo[72] = {44, 88, 116, 21, 65,  101, 42, 86, 152, 19, 63, 99,  40, 84, 150, 17, 61,  97,
                                      38, 82, 148, 15, 59,  95,  36, 80, 146, 13, 57, 93,  34, 78, 144, 11, 55,  129,
                                      32, 76, 142, 9,  53,  127, 30, 74, 140, 7,  51, 125, 28, 72, 138, 5,  49,  123,
                                      26, 70, 166, 3,  47,  121, 24, 106, 164, 1,  91, 119, 68, 104, 162, 45, 89, 117};

/* air layout of an LDU after the NID, in status-stripped dibits (p25p1_ldu1.c:176-213, p25p1_ldu2.c:206-234) */
__constant__ short c_ldu_imbe_off[9] = {0, 72, 164, 256, 348, 440, 532, 624, 712};
__constant__ short c_ldu_word_off[6] = {144, 236, 328, 420, 512, 604};
constexpr int kLduLsdOff = 696;
constexpr int kLduPayload = 784, kHduPayload = 329, kTsbkBlock = 98, kTdulcPayload = 154; /* TDULC: 12 x 12 dibits + 10 nulls */

struct P25Stream {
    const uint8_t* d;
    const int16_t* l;
    int start; /* buffer index of the first sync dibit */
    __device__ __forceinline__ int get(int k, int& l0, int& l1) const {
        const int pos = start + p25p1_payload_offset(57, k);
        l0 = l[2 * pos];
        l1 = l[2 * pos + 1];
        return d[pos] & 3;
    }
};

__device__ __forceinline__ int
clamp_rel(int l) {
    const int a = l < 0 ? -l : l;
    return a > 255 ? 255 : a;
}

/* check_and_fix_* then p25p1_rs_*_soft_reliability on 6-bit symbols (parity first), lane-serial.  Returns the record status:
 * 0 hard decode ok, 1 recovered by ranked erasures, 2 irrecoverable; data_out = the n_data data symbols afterwards. */
__device__ __noinline__ int
p25_rs_frame_decode(const dsdneo_fec_tables* __restrict__ T, int n_total, int n_data, const uint8_t* sym, const uint8_t* data_rel,
                    const uint8_t* par_rel, int threshold, uint8_t* data_out) {
    const RsShape sh = {n_total, n_data, (n_total - n_data) / 2};
    int out_sym[36];
    int rc = rs63_hard_decode(T->gf_exp, T->gf_log, sh, sym, out_sym);
    int status = rc == 0 ? 0 : 2;
    if (rc != 0) {
        const Gf64 gf = {T->gf_exp, T->gf_log};
        uint8_t word[63];
        for (int i = 0; i < 63; i++) {
            word[i] = i < n_total ? sym[i] : 0;
        }
        if (rs63_ranked_erasure_decode(gf, sh, word, data_rel, par_rel, threshold, out_sym) == 0) {
            status = 1;
        }
    }
    for (int i = 0; i < n_data; i++) {
        data_out[i] = (uint8_t)out_sym[i];
    }
    return status;
}

constexpr int kFrameWarps = 4;

__global__ void __launch_bounds__(kFrameWarps * 32)
p25p1_frame_decode_kernel(const dsdneo_fec_tables* __restrict__ T, const P25FrameParams p) {
    __shared__ uint8_t s_bit[kFrameWarps][184], s_rel[kFrameWarps][184];
    __shared__ uint8_t s_word[kFrameWarps][36], s_wrel[kFrameWarps][36];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * kFrameWarps + warp;
    const int ch = slot / p.max_hits, h = slot - ch * p.max_hits;
    if (ch >= p.n_channels || h >= min(p.n_hits[ch], p.max_hits)) {
        return;
    }
    const int r = p.frame_off[ch] + h;
    if (r >= p.frame_capacity) {
        return;
    }
    const int status = p25_slot_status(p, slot);
    const int duid = status > 0 ? p.nid_duid[slot] : 0xFF;
    const int count = p.counts[ch];
    const int pos_last = p.hits[2 * slot] + p.region_off;
    P25Stream st;
    st.d = p.dibits + (size_t)ch * p.dibit_pitch;
    st.l = p.llr + (size_t)ch * p.llr_pitch * 2;
    st.start = pos_last - 23;
    /* voice record index: LDUs of this channel before this hit */
    int vi = -1;
    {
        const bool mine = lane < h && p25_is_ldu(p25_slot_status(p, ch * p.max_hits + lane), p.nid_duid[ch * p.max_hits + lane]);
        const unsigned before = __ballot_sync(0xffffffffu, mine);
        if (duid == 0x5 || duid == 0xA) {
            vi = p.voice_off[ch] + __popc(before);
            if (vi >= p.voice_capacity) {
                vi = -1;
            }
        }
    }
    dsdneo_b200_p25p1_frame* f = p.frames + r;
    const int payload = (duid == 0x5 || duid == 0xA) ? kLduPayload
                        : (duid == 0x0 ? kHduPayload : (duid == 0x7 ? 3 * kTsbkBlock : (duid == 0xF ? kTdulcPayload : 0)));
    /* a TSDU is read block by block and ends at the block flagged last: only its first block must fit up front */
    const int must_fit = duid == 0x7 ? kTsbkBlock : payload;
    const bool fits = st.start >= 0 && (must_fit == 0 || st.start + p25p1_payload_offset(57, must_fit - 1) + 1 <= count);
    /* header + cleared payload fields */
    for (int i = lane; i < (int)sizeof(*f); i += 32) {
        reinterpret_cast<uint8_t*>(f)[i] = 0;
    }
    __syncwarp();
    if (lane == 0) {
        f->position = (p.stream_base ? p.stream_base[ch] : 0ll) + pos_last;
        f->channel = ch;
        f->voice_index = fits ? vi : -1;
        f->nac = (int16_t)p.nid_nac[slot];
        f->nid_errs = (int16_t)p.nid_errs[slot];
        f->nid_status = (int8_t)status;
        f->duid = (uint8_t)duid;
        f->reserved[0] = fits ? 0 : 1; /* the stream ended inside the frame: payload fields not decoded */
    }
    if (!fits || payload == 0) {
        return;
    }
    if (duid == 0x7) { /* ---- TSDU: one trellis block per lane ---- */
        uint8_t out12[12];
        bool crc = false, last = true;
        const bool blk_fits = lane < 3 && st.start + p25p1_payload_offset(57, (lane + 1) * kTsbkBlock - 1) + 1 <= count;
        if (blk_fits) {
            int16_t llr196[196];
            for (int i = 0; i < kTsbkBlock; i++) {
                int l0, l1;
                (void)st.get(lane * kTsbkBlock + i, l0, l1);
                llr196[2 * i] = (int16_t)l0;
                llr196[2 * i + 1] = (int16_t)l1;
            }
            /* Shortcut for the common case.  Within one predecessor the rank-0 survivor of the list decoder has the smallest
             * metric and equal metrics keep insertion order (predecessor, then rank, strict <), so rank 0 of every state is the
             * plain Viterbi survivor (p25_12_soft_llr's tie rule).  The final candidates are inserted state by state and
             * de-duplicated on the 48 packed symbols BEFORE comparing metrics, so a best path ending in state > 0 can be
             * shadowed by a worse twin that ends in a lower state; a best path that ends in state 0 -- where every valid block
             * ends, its 49th dibit is the flush -- is inserted first and nothing can displace it.  tsbk_select_crc_candidate
             * takes the first candidate that passes the CRC: if the plain path ends in state 0 and passes, it is the answer.
             * Everything else runs the full list decoder. */
            int final_state = 0;
            (void)p25_12_decode(llr196, out12, &final_state);
            if (final_state != 0 || !p25_crc16_ok(out12)) {
                dsdneo_b200_p25_12_candidate cands[kListK];
                const int n = p25_12_list_decode(llr196, cands, kListK);
                if (n > 0) {
                    int sel = 0;
                    for (int c = 0; c < n; c++) {
                        if (p25_crc16_ok(cands[c].bytes)) {
                            sel = c;
                            break;
                        }
                    }
                    for (int b = 0; b < 12; b++) {
                        out12[b] = cands[sel].bytes[b];
                    }
                }
            }
            crc = p25_crc16_ok(out12);
            last = (out12[0] >> 7) & 1;
        }
        const unsigned fitmask = __ballot_sync(0xffffffffu, blk_fits) & 7u;
        const unsigned lastmask = __ballot_sync(0xffffffffu, blk_fits && last) & 7u;
        const unsigned crcmask = __ballot_sync(0xffffffffu, blk_fits && crc) & 7u;
        int n_blocks = lastmask ? __ffs(lastmask) : 3; /* processTSBK stops after the block flagged last */
        const int n_avail = __ffs(~fitmask) - 1;       /* leading blocks that lie inside the stream */
        if (n_blocks > n_avail) {                      /* the stream ended before the last block of this TSDU */
            n_blocks = n_avail;
            if (lane == 0) {
                f->reserved[0] = 1;
            }
        }
        if (lane < n_blocks) {
            for (int b = 0; b < 12; b++) {
                f->tsbk[lane][b] = out12[b];
            }
        }
        if (lane == 0) {
            f->n_tsbk = (uint8_t)n_blocks;
            f->tsbk_crc_ok = (uint8_t)(crcmask & ((1u << n_blocks) - 1u));
        }
        return;
    }
    int soft_changed = 0;
    if (duid == 0x0) { /* ---- HDU: 36 Golay(24,6) words, 9 dibits each ---- */
        for (int w = lane; w < 36; w += 32) {
            unsigned cw = 0;
            int rel[24];
            int minrel = 255;
#pragma unroll
            for (int i = 0; i < 24; i++) {
                rel[i] = 0;
            }
            for (int d = 0; d < 9; d++) {
                int l0, l1;
                const int dib = st.get(w * 9 + d, l0, l1);
                const int b0 = (dib >> 1) & 1, b1 = dib & 1;
                /* bit index i of the word (data 0..5, parity 6..17) sits at packed bit 6 + i */
                cw |= (unsigned)b0 << (6 + 2 * d);
                cw |= (unsigned)b1 << (6 + 2 * d + 1);
                rel[2 * d] = clamp_rel(l0);
                rel[2 * d + 1] = clamp_rel(l1);
                if (d < 3) {
                    minrel = min(minrel, min(rel[2 * d], rel[2 * d + 1]));
                }
            }
            unsigned fixed_cw = cw;
            int errs = 0;
            const int hard = golay24_hard_word(cw, fixed_cw, errs);
            unsigned data = hard == 0 ? (fixed_cw & 0xFC0u) : (cw & 0xFC0u);
            if (hard != 0 || errs > 0) {
                unsigned sres = 0;
                int sfx = 0;
                if (golay24_soft_word(cw, rel, 6, p.hard_override, p.threshold, sres, sfx) == 0) {
                    soft_changed += hard != 0 ? 1 : 0;
                    data = sres & 0xFC0u;
                }
            }
            int v = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                v = (v << 1) | (int)((data >> (6 + i)) & 1u);
            }
            s_word[warp][w] = (uint8_t)v;
            s_wrel[warp][w] = (uint8_t)minrel;
        }
        for (int o = 16; o > 0; o >>= 1) {
            soft_changed += __shfl_xor_sync(0xffffffffu, soft_changed, o);
        }
        __syncwarp();
        if (lane == 0) {
            uint8_t sym[36], dr[20], pr[16];
            for (int i = 0; i < 20; i++) { /* hex_data[i] = air word 19 - i, hex_parity[i] = air word 35 - i */
                f->rs_in_data[i] = s_word[warp][19 - i];
                dr[i] = s_wrel[warp][19 - i];
                sym[16 + i] = s_word[warp][19 - i];
            }
            for (int i = 0; i < 16; i++) {
                f->rs_in_parity[i] = s_word[warp][35 - i];
                pr[i] = s_wrel[warp][35 - i];
                sym[i] = s_word[warp][35 - i];
            }
            f->rs_kind = 1;
            f->rs_status = (uint8_t)p25_rs_frame_decode(T, 36, 20, sym, dr, pr, p.threshold, f->rs_data);
            f->n_word_soft = (uint8_t)soft_changed;
        }
        return;
    }
    if (duid == 0xF) { /* ---- TDULC: 12 Golay(24,12) dodeca words, 12 dibits each (p25p1_tdulc.c:75-157,198-238) ---- */
        if (lane < 12) {
            const int k = lane; /* air order: data word 5 .. 0, parity word 5 .. 0 */
            unsigned cw = 0;
            int rel[24], half_rel[2] = {255, 255};
            for (int d = 0; d < 12; d++) {
                int l0, l1;
                const int dib = st.get(k * 12 + d, l0, l1);
                /* bit index i of the word (data 0..11, parity 12..23) sits at packed bit i */
                cw |= (unsigned)((dib >> 1) & 1) << (2 * d);
                cw |= (unsigned)(dib & 1) << (2 * d + 1);
                rel[2 * d] = clamp_rel(l0);
                rel[2 * d + 1] = clamp_rel(l1);
                if (d < 6) {
                    half_rel[d / 3] = min(half_rel[d / 3], min(rel[2 * d], rel[2 * d + 1]));
                }
            }
            unsigned fixed_cw = cw;
            int errs = 0;
            const int hard = golay24_hard_word(cw, fixed_cw, errs);
            unsigned data = hard == 0 ? (fixed_cw & 0xFFFu) : (cw & 0xFFFu);
            if (hard != 0 || errs > 0) {
                unsigned sres = 0;
                int sfx = 0;
                if (golay24_soft_word(cw, rel, 12, p.hard_override, p.threshold, sres, sfx) == 0) {
                    soft_changed += hard != 0 ? 1 : 0;
                    data = sres & 0xFFFu;
                }
            }
            /* swap_hex_words: the dodeca word becomes two hex symbols, bits 6..11 first (reference bit 0 = MSB of a symbol) */
            int lo = 0, hi = 0;
#pragma unroll
            for (int b = 0; b < 6; b++) {
                lo = (lo << 1) | (int)((data >> b) & 1u);
                hi = (hi << 1) | (int)((data >> (6 + b)) & 1u);
            }
            const int i = 5 - (k % 6), base = k < 6 ? 0 : 12;
            s_word[warp][base + 2 * i] = (uint8_t)hi;
            s_word[warp][base + 2 * i + 1] = (uint8_t)lo;
            s_wrel[warp][base + 2 * i] = (uint8_t)half_rel[1];
            s_wrel[warp][base + 2 * i + 1] = (uint8_t)half_rel[0];
        }
        for (int o = 16; o > 0; o >>= 1) {
            soft_changed += __shfl_xor_sync(0xffffffffu, soft_changed, o);
        }
        __syncwarp();
        if (lane == 0) {
            uint8_t sym[24], dr[12], pr[12];
            for (int j = 0; j < 12; j++) {
                f->rs_in_data[j] = s_word[warp][j];
                dr[j] = s_wrel[warp][j];
                sym[12 + j] = s_word[warp][j];
                f->rs_in_parity[j] = s_word[warp][12 + j];
                pr[j] = s_wrel[warp][12 + j];
                sym[j] = s_word[warp][12 + j];
            }
            f->rs_kind = 2;
            f->rs_status = (uint8_t)p25_rs_frame_decode(T, 24, 12, sym, dr, pr, p.threshold, f->rs_data);
            f->n_word_soft = (uint8_t)soft_changed;
        }
        return;
    }
    /* ---- LDU1 / LDU2 ---- */
    const bool ldu2 = duid == 0xA;
    const int n_data = ldu2 ? 16 : 12;
    dsdneo_b200_p25p1_voice* voice = vi >= 0 ? p.voices + vi : nullptr;
    for (int v = 0; v < 9; v++) {
        for (int i = lane; i < 184; i += 32) {
            s_bit[warp][i] = 0, s_rel[warp][i] = 0;
        }
        __syncwarp();
        for (int j = lane; j < 72; j += 32) {
            int l0, l1;
            const int dib = st.get(c_ldu_imbe_off[v] + j, l0, l1);
            s_bit[warp][c_imbe_hi[j]] = (uint8_t)((dib >> 1) & 1);
            s_rel[warp][c_imbe_hi[j]] = (uint8_t)clamp_rel(l0);
            s_bit[warp][c_imbe_lo[j]] = (uint8_t)(dib & 1);
            s_rel[warp][c_imbe_lo[j]] = (uint8_t)clamp_rel(l1);
        }
        __syncwarp();
        if (voice) {
            if (lane < 8) {
                unsigned w = 0;
                for (int col = 0; col < 23; col++) {
                    w |= (unsigned)s_bit[warp][lane * 23 + col] << col;
                }
                voice->bits[v][lane] = w;
            }
            uint8_t* rel_out = &voice->reliab[v][0][0];
            for (int i = lane; i < 184; i += 32) {
                rel_out[i] = s_rel[warp][i];
            }
        }
        __syncwarp();
    }
    if (lane < 24) { /* one Hamming(10,6,3) hex word per lane: read_and_correct_hex_word, p25p1_ldu.c:190-222 */
        const int w = lane;
        const int base = c_ldu_word_off[w >> 2] + (w & 3) * 5;
        unsigned v10 = 0;
        int rel[10], minrel = 255;
        for (int d = 0; d < 5; d++) {
            int l0, l1;
            const int dib = st.get(base + d, l0, l1);
            v10 = (v10 << 2) | (unsigned)dib;
            rel[2 * d] = clamp_rel(l0);
            rel[2 * d + 1] = clamp_rel(l1);
            if (d < 3) {
                minrel = min(minrel, min(rel[2 * d], rel[2 * d + 1]));
            }
        }
        unsigned hv = v10;
        const int hard = ham1063_hard(hv);
        unsigned data = hv >> 4;
        if (hard == 1 || hard == 2) {
            unsigned sres = v10;
            const int soft = ham1063_soft_word(v10, rel, p.hard_override, p.threshold, sres);
            if (soft != 2) {
                soft_changed += (hard == 2 || sres != hv) ? 1 : 0;
                data = sres >> 4;
            }
        }
        s_word[warp][w] = (uint8_t)(data & 0x3Fu);
        s_wrel[warp][w] = (uint8_t)minrel;
    }
    for (int o = 16; o > 0; o >>= 1) {
        soft_changed += __shfl_xor_sync(0xffffffffu, soft_changed, o);
    }
    __syncwarp();
    if (lane == 0) {
        uint8_t sym[24], dr[16], pr[12];
        const int n_par = 24 - n_data;
        for (int i = 0; i < n_data; i++) { /* hex_data[i] = air word n_data - 1 - i, hex_parity[i] = air word 23 - i */
            f->rs_in_data[i] = s_word[warp][n_data - 1 - i];
            dr[i] = s_wrel[warp][n_data - 1 - i];
            sym[n_par + i] = s_word[warp][n_data - 1 - i];
        }
        for (int i = 0; i < n_par; i++) {
            f->rs_in_parity[i] = s_word[warp][23 - i];
            pr[i] = s_wrel[warp][23 - i];
            sym[i] = s_word[warp][23 - i];
        }
        f->rs_kind = ldu2 ? 3 : 2;
        f->rs_status = (uint8_t)p25_rs_frame_decode(T, 24, n_data, sym, dr, pr, p.threshold, f->rs_data);
        f->n_word_soft = (uint8_t)soft_changed;
    }
    bool lsd_good = false;
    if (lane >= 30) { /* low speed data: two (16,8) words, 8 dibits each (p25p1_ldu1.c:128-175, 306-320) */
        const int k = lane - 30;
        unsigned w = 0;
        int rel[16];
        for (int d = 0; d < 8; d++) {
            int l0, l1;
            const int dib = st.get(kLduLsdOff + 8 * k + d, l0, l1);
            w = (w << 2) | (unsigned)dib;
            rel[2 * d] = l0 < 0 ? -l0 : l0; /* the LSD search compares unclamped |llr| with the threshold */
            rel[2 * d + 1] = l1 < 0 ? -l1 : l1;
        }
        lsd_good = lsd_soft(w, rel, p.threshold);
        f->lsd[k] = (uint8_t)(w >> 8);
    }
    const unsigned okmask = __ballot_sync(0xffffffffu, lsd_good);
    if (lane == 1) {
        f->lsd_ok = (uint8_t)((okmask >> 30) & 3u);
    }
}

}  // namespace

/* ---- DMR rate 3/4 trellis (dmr_r34_viterbi_decode / _soft, src/protocol/dmr/dmr_34_viterbi.c:402-474) ---------------------
 * 8 lanes per codeword, lane = next state: 49 add-compare-select steps with the previous metrics exchanged by shuffles, the
 * survivor's previous state kept three bits per step in registers, traceback by shuffles from end state 0.  Constellation and
 * state tables: ETSI TS 102 361-1 B.2.5 (the reference's dsd_trellis34_constellation / _fsm, src/fec/trellis34.c:15-21). */
__constant__ uint8_t c_r34_point_of_nibble[16] = {11, 12, 0, 7, 14, 9, 5, 2, 10, 13, 1, 6, 15, 8, 4, 3};
__constant__ uint8_t c_r34_nibble_of_point[16] = {2, 10, 7, 15, 14, 6, 11, 3, 13, 5, 8, 0, 1, 9, 4, 12};
__constant__ uint8_t c_r34_fsm[64] = {0, 8,  4, 12, 2, 10, 6, 14, 4, 12, 2, 10, 6, 14, 0, 8, 1, 9,  5, 13, 3, 11,
                                      7, 15, 5, 13, 3, 11, 7, 15, 1, 9,  3, 11, 7, 15, 1, 9, 5, 13, 7, 15, 1, 9,
                                      5, 13, 3, 11, 2, 10, 6, 14, 0, 8,  4, 12, 6, 14, 0, 8, 4, 12, 2, 10};

__device__ __forceinline__ int
r34_deinterleaved_source(int k) { /* inverse of dsd_trellis_interleave_98: which received dibit lands at position k */
    /* table[i] = 8 * ((i % 26) / 2) + 2 * (i / 26) + (i & 1) for the 26/26/24/22 column walk (src/fec/trellis34.c:8-13) */
    const int col = k >> 3, within = k & 7; /* k = 8 * col + 2 * row + parity */
    const int row = within >> 1, par = within & 1;
    const int base = row == 0 ? 0 : (row == 1 ? 26 : (row == 2 ? 50 : 74));
    return base + 2 * col + par;
}

__global__ void __launch_bounds__(128)
dmr_r34_kernel(const uint8_t* __restrict__ dibits, const uint8_t* __restrict__ reliab, uint8_t* __restrict__ out18, int n) {
    __shared__ uint8_t s_dei[16][98], s_rel[16][98], s_states[16][49];
    const int grp = threadIdx.x >> 3, ns = threadIdx.x & 7;
    const int w = blockIdx.x * 16 + grp;
    const bool live = w < n;
    const unsigned lane = threadIdx.x & 31;
    const unsigned gbase = lane & ~7u;
    if (live) {
        for (int k = ns; k < 98; k += 8) {
            const int src = r34_deinterleaved_source(k);
            s_dei[grp][k] = dibits[(size_t)w * 98 + src] & 3;
            s_rel[grp][k] = reliab ? reliab[(size_t)w * 98 + src] : 0;
        }
    }
    __syncthreads();
    const int INF = 1000000000;
    int metric = ns == 0 ? 0 : INF;
    unsigned long long bp0 = 0, bp1 = 0, bp2 = 0; /* 3 bits per step: steps 0..20, 21..41, 42..48 */
    const bool soft = reliab != nullptr;
    for (int t = 0; t < 49; t++) {
        const int d0 = live ? s_dei[grp][2 * t] : 0, d1 = live ? s_dei[grp][2 * t + 1] : 0;
        const int nib = (d0 << 2) | d1;
        const int point = c_r34_point_of_nibble[nib];
        const int rhi = live ? s_rel[grp][2 * t] : 0, rlo = live ? s_rel[grp][2 * t + 1] : 0;
        int best = INF, best_ps = 0;
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int mp = __shfl_sync(0xffffffffu, metric, gbase + ps);
            const int expect = c_r34_fsm[ps * 8 + ns];
            int cost;
            if (!soft) {
                cost = __popc((unsigned)((expect ^ point) & 15));
            } else {
                const int x = c_r34_nibble_of_point[expect] ^ nib;
                cost = (((x >> 3) & 1) + ((x >> 2) & 1)) * rhi + (((x >> 1) & 1) + (x & 1)) * rlo;
            }
            const int m = mp + cost;
            if (mp < INF && m < best) {
                best = m, best_ps = ps;
            }
        }
        metric = best;
        const unsigned long long v = (unsigned long long)best_ps;
        if (t < 21) {
            bp0 |= v << (3 * t);
        } else if (t < 42) {
            bp1 |= v << (3 * (t - 21));
        } else {
            bp2 |= v << (3 * (t - 42));
        }
    }
    /* traceback: every lane follows the same path; lane `state` owns the back pointer */
    int state = 0;
    for (int t = 48; t >= 0; t--) {
        if (ns == 0 && live) {
            s_states[grp][t] = (uint8_t)state;
        }
        const unsigned long long word = t < 21 ? bp0 : (t < 42 ? bp1 : bp2);
        const int sh = 3 * (t < 21 ? t : (t < 42 ? t - 21 : t - 42));
        const int mine = (int)((word >> sh) & 7ull);
        state = __shfl_sync(0xffffffffu, mine, gbase + state);
    }
    __syncthreads();
    if (live && ns < 6) {
        unsigned v = 0;
        for (int k = 0; k < 8; k++) {
            v = (v << 3) | (unsigned)(s_states[grp][ns * 8 + k] & 7);
        }
        out18[(size_t)w * 18 + 3 * ns] = (uint8_t)(v >> 16);
        out18[(size_t)w * 18 + 3 * ns + 1] = (uint8_t)(v >> 8);
        out18[(size_t)w * 18 + 3 * ns + 2] = (uint8_t)v;
    }
}

/* ---- RS(12,9) over GF(2^8) (src/fec/rs-12-9.c:237-323): one thread per codeword; exp / log tables built per CTA ------------ */
__device__ __forceinline__ uint8_t
rs129_mul(const uint8_t* ex, const uint8_t* lg, uint8_t a, uint8_t b) {
    return (a == 0 || b == 0) ? 0 : ex[(lg[a] + lg[b]) % 255];
}

__global__ void __launch_bounds__(128)
rs_12_9_kernel(uint8_t* __restrict__ cw, uint8_t* __restrict__ syndrome3, uint8_t* __restrict__ result, uint8_t* __restrict__ errors_found,
               int n) {
    __shared__ uint8_t ex[256], lg[256];
    if (threadIdx.x == 0) { /* x^8 + x^4 + x^3 + x^2 + 1; exp[255] = 1 and log[0] = 0 as in the reference's tables */
        int v = 1;
        lg[0] = 0;
        for (int i = 0; i < 255; i++) {
            ex[i] = (uint8_t)v;
            lg[v] = (uint8_t)i;
            v <<= 1;
            if (v & 0x100) {
                v ^= 0x11D;
            }
        }
        ex[255] = 1;
    }
    __syncthreads();
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) {
        return;
    }
    uint8_t c[12];
    for (int i = 0; i < 12; i++) {
        c[i] = cw[(size_t)w * 12 + i];
    }
    uint8_t S[3] = {0, 0, 0};
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < 12; i++) {
            S[j] = c[i] ^ rs129_mul(ex, lg, ex[j + 1], S[j]);
        }
    }
    if (syndrome3) {
        syndrome3[(size_t)w * 3] = S[0], syndrome3[(size_t)w * 3 + 1] = S[1], syndrome3[(size_t)w * 3 + 2] = S[2];
    }
    if (!(S[0] | S[1] | S[2])) {
        result[w] = 0;
        errors_found[w] = 0;
        return;
    }
    uint8_t loc[6] = {1, 0, 0, 0, 0, 0}, D[6] = {0, 1, 0, 0, 0, 0}, psi2[6];
    int L = 0, k = -1;
    for (int nn = 0; nn < 3; nn++) {
        uint8_t d = 0;
        for (int i = 0; i <= L; i++) {
            d ^= rs129_mul(ex, lg, loc[i], S[nn - i]);
        }
        if (d != 0) {
            for (int i = 0; i < 6; i++) {
                psi2[i] = loc[i] ^ rs129_mul(ex, lg, d, D[i]);
            }
            if (L < nn - k) {
                const int L2 = nn - k;
                k = nn - L;
                const uint8_t di = ex[255 - lg[d]];
                for (int i = 0; i < 6; i++) {
                    D[i] = rs129_mul(ex, lg, loc[i], di);
                }
                L = L2;
            }
            for (int i = 0; i < 6; i++) {
                loc[i] = psi2[i];
            }
        }
        for (int i = 5; i > 0; i--) {
            D[i] = D[i - 1];
        }
        D[0] = 0;
    }
    uint8_t ev[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; i + j < 3; j++) {
            ev[i + j] ^= rs129_mul(ex, lg, S[j], loc[i]);
        }
    }
    uint8_t locs[4];
    int nroots = 0;
    for (int r = 1; r < 256; r++) {
        uint8_t sum = 0;
        for (int kk = 0; kk < 4; kk++) {
            sum ^= rs129_mul(ex, lg, ex[(kk * r) % 255], loc[kk]);
        }
        if (sum == 0) {
            if (nroots < 4) {
                locs[nroots] = (uint8_t)(255 - r);
            }
            nroots++;
        }
    }
    errors_found[w] = (uint8_t)nroots;
    if (nroots == 0) {
        result[w] = 1;
        return;
    }
    bool bad = nroots > 3;
    for (int r = 0; r < nroots && r < 4 && !bad; r++) {
        bad = locs[r] >= 12;
    }
    if (bad) {
        result[w] = 3;
        return;
    }
    for (int r = 0; r < nroots; r++) {
        const int i = locs[r];
        uint8_t num = 0, den = 0;
        for (int j = 0; j < 3; j++) {
            num ^= rs129_mul(ex, lg, ev[j], ex[((255 - i) * j) % 255]);
        }
        for (int j = 1; j < 6; j += 2) {
            den ^= rs129_mul(ex, lg, loc[j], ex[((255 - i) * (j - 1)) % 255]);
        }
        c[12 - i - 1] ^= rs129_mul(ex, lg, num, ex[255 - lg[den]]);
    }
    for (int i = 0; i < 12; i++) {
        cw[(size_t)w * 12 + i] = c[i];
    }
    result[w] = 2;
}

extern "C" {

int
dsdneo_b200_fec_block_code_len(int code) {
    int n, k;
    int rc = block_code_shape(code, &n, &k);
    return rc ? rc : n;
}

int
dsdneo_b200_fec_block_code_k(int code) {
    int n, k;
    int rc = block_code_shape(code, &n, &k);
    return rc ? rc : k;
}

int
dsdneo_b200_fec_block_decode_batch(int code, uint8_t* d_bits, uint8_t* d_decoded, uint8_t* d_ok, int n_words, void* stream) {
    int n, k;
    int rc = block_code_shape(code, &n, &k);
    if (rc) {
        return rc;
    }
    if (!d_bits || !d_ok || n_words < 0) {
        set_error("fec_block_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    if (code <= DSDNEO_FEC_HAMMING_16_11_4) {
        KernelTimer kt("hamming_decode_kernel", s);
        hamming_decode_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(g_d_tables, code, d_bits, d_decoded, d_ok, n_words);
    } else {
        KernelTimer kt("gq_decode_kernel", s);
        gq_decode_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(g_d_tables, code - DSDNEO_FEC_GOLAY_20_8, d_bits, d_ok, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_fec_block_decode_batch_host(int code, uint8_t* h_bits, uint8_t* h_decoded, uint8_t* h_ok, int n_words) {
    int n, k;
    int rc = block_code_shape(code, &n, &k);
    if (rc) {
        return rc;
    }
    if (!h_bits || !h_ok || n_words < 0) {
        set_error("fec_block_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    rc = ensure_device();
    if (rc) {
        return rc;
    }
    DevBuf bits((size_t)n_words * n), dec((size_t)n_words * k), ok((size_t)n_words);
    DSDNEO_CUDA(bits.err);
    DSDNEO_CUDA(dec.err);
    DSDNEO_CUDA(ok.err);
    DSDNEO_CUDA(cudaMemcpy(bits.p, h_bits, (size_t)n_words * n, cudaMemcpyHostToDevice));
    if (h_decoded) {
        DSDNEO_CUDA(cudaMemcpy(dec.p, h_decoded, (size_t)n_words * k, cudaMemcpyHostToDevice)); /* untouched-on-failure semantics */
    }
    rc = dsdneo_b200_fec_block_decode_batch(code, bits.as<uint8_t>(), h_decoded ? dec.as<uint8_t>() : NULL, ok.as<uint8_t>(),
                                            n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_bits, bits.p, (size_t)n_words * n, cudaMemcpyDeviceToHost));
    if (h_decoded) {
        DSDNEO_CUDA(cudaMemcpy(h_decoded, dec.p, (size_t)n_words * k, cudaMemcpyDeviceToHost));
    }
    DSDNEO_CUDA(cudaMemcpy(h_ok, ok.p, (size_t)n_words, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_fec_golay_24_12_encode_batch(const uint8_t* d_data, uint8_t* d_out, int n_words, void* stream) {
    if (!d_data || !d_out || n_words < 0) {
        set_error("golay_24_12_encode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("golay_24_12_encode_kernel", s);
        golay_24_12_encode_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(g_d_tables, d_data, d_out, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_dmr_r34_decode_batch(const uint8_t* d_dibits98, const uint8_t* d_reliab98, uint8_t* d_out18, int n_blocks, void* stream) {
    if (!d_dibits98 || !d_out18 || n_blocks < 0) {
        set_error("dmr_r34_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_blocks == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("dmr_r34_kernel", s);
        dmr_r34_kernel<<<grid_for(n_blocks, 16), 128, 0, s>>>(d_dibits98, d_reliab98, d_out18, n_blocks);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_dmr_r34_decode_batch_host(const uint8_t* h_dibits98, const uint8_t* h_reliab98, uint8_t* h_out18, int n_blocks) {
    if (!h_dibits98 || !h_out18 || n_blocks < 0) {
        set_error("dmr_r34_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_blocks == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_blocks;
    DevBuf in(n * 98), rel(n * 98), out(n * 18);
    DSDNEO_CUDA(in.err);
    DSDNEO_CUDA(rel.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(cudaMemcpy(in.p, h_dibits98, n * 98, cudaMemcpyHostToDevice));
    if (h_reliab98) {
        DSDNEO_CUDA(cudaMemcpy(rel.p, h_reliab98, n * 98, cudaMemcpyHostToDevice));
    }
    rc = dsdneo_b200_dmr_r34_decode_batch(in.as<uint8_t>(), h_reliab98 ? rel.as<uint8_t>() : NULL, out.as<uint8_t>(), n_blocks, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out18, out.p, n * 18, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_rs_12_9_decode_batch(uint8_t* d_codewords, uint8_t* d_syndrome3, uint8_t* d_result, uint8_t* d_errors_found, int n_words,
                                 void* stream) {
    if (!d_codewords || !d_result || !d_errors_found || n_words < 0) {
        set_error("rs_12_9_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("rs_12_9_kernel", s);
        rs_12_9_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(d_codewords, d_syndrome3, d_result, d_errors_found, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_rs_12_9_decode_batch_host(uint8_t* h_codewords, uint8_t* h_syndrome3, uint8_t* h_result, uint8_t* h_errors_found,
                                      int n_words) {
    if (!h_codewords || !h_result || !h_errors_found || n_words < 0) {
        set_error("rs_12_9_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words;
    DevBuf cw(n * 12), syn(n * 3), res(n), ef(n);
    DSDNEO_CUDA(cw.err);
    DSDNEO_CUDA(syn.err);
    DSDNEO_CUDA(res.err);
    DSDNEO_CUDA(ef.err);
    DSDNEO_CUDA(cudaMemcpy(cw.p, h_codewords, n * 12, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_rs_12_9_decode_batch(cw.as<uint8_t>(), syn.as<uint8_t>(), res.as<uint8_t>(), ef.as<uint8_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_codewords, cw.p, n * 12, cudaMemcpyDeviceToHost));
    if (h_syndrome3) {
        DSDNEO_CUDA(cudaMemcpy(h_syndrome3, syn.p, n * 3, cudaMemcpyDeviceToHost));
    }
    DSDNEO_CUDA(cudaMemcpy(h_result, res.p, n, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_errors_found, ef.p, n, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_bptc_196x96_batch(const uint8_t* d_in, int interleaved, uint8_t* d_out96, uint8_t* d_r3, uint32_t* d_errs,
                              int n_bursts, void* stream) {
    if (!d_in || !d_out96 || !d_errs || n_bursts < 0) {
        set_error("bptc_196x96_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_bursts == 0) {
        return 0;
    }
    int rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("bptc_196x96_kernel", s);
        bptc_196x96_kernel<<<grid_for(n_bursts, 64), 64, 0, s>>>(g_d_tables, d_in, interleaved, d_out96, d_r3, d_errs, n_bursts);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_bptc_196x96_batch_host(const uint8_t* h_in, int interleaved, uint8_t* h_out96, uint8_t* h_r3, uint32_t* h_errs,
                                   int n_bursts) {
    if (!h_in || !h_out96 || !h_errs || n_bursts < 0) {
        set_error("bptc_196x96_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_bursts == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_bursts;
    DevBuf in(n * 196), out(n * 96), r3(n * 3), er(n * 4);
    DSDNEO_CUDA(in.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(r3.err);
    DSDNEO_CUDA(er.err);
    DSDNEO_CUDA(cudaMemcpy(in.p, h_in, n * 196, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_bptc_196x96_batch(in.as<uint8_t>(), interleaved, out.as<uint8_t>(), r3.as<uint8_t>(), er.as<uint32_t>(),
                                       n_bursts, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out96, out.p, n * 96, cudaMemcpyDeviceToHost));
    if (h_r3) {
        DSDNEO_CUDA(cudaMemcpy(h_r3, r3.p, n * 3, cudaMemcpyDeviceToHost));
    }
    DSDNEO_CUDA(cudaMemcpy(h_errs, er.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int
bptc_small_batch(int which, const uint8_t* d_in, uint8_t* d_out, uint32_t* d_errs, int parity_odd, int n_items, void* stream) {
    if (!d_in || !d_out || !d_errs || n_items < 0) {
        set_error("bptc_%s_batch: bad argument", which ? "16x2" : "128x77");
        return DSDNEO_B200_EINVAL;
    }
    if (n_items == 0) {
        return 0;
    }
    int rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    if (which == 0) {
        KernelTimer kt("bptc_128x77_kernel", s);
        bptc_128x77_kernel<<<grid_for(n_items, 128), 128, 0, s>>>(g_d_tables, d_in, d_out, d_errs, n_items);
    } else {
        KernelTimer kt("bptc_16x2_kernel", s);
        bptc_16x2_kernel<<<grid_for(n_items, 128), 128, 0, s>>>(g_d_tables, d_in, d_out, d_errs, parity_odd ? 1 : 0, n_items);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

static int
bptc_small_batch_host(int which, const uint8_t* h_in, uint8_t* h_out, uint32_t* h_errs, int parity_odd, int n_items) {
    if (!h_in || !h_out || !h_errs || n_items < 0) {
        set_error("bptc_%s_batch_host: bad argument", which ? "16x2" : "128x77");
        return DSDNEO_B200_EINVAL;
    }
    if (n_items == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_items, in_b = which ? 32 : 128, out_b = which ? 32 : 77;
    DevBuf in(n * in_b), out(n * out_b), er(n * 4);
    DSDNEO_CUDA(in.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(er.err);
    DSDNEO_CUDA(cudaMemcpy(in.p, h_in, n * in_b, cudaMemcpyHostToDevice));
    rc = bptc_small_batch(which, in.as<uint8_t>(), out.as<uint8_t>(), er.as<uint32_t>(), parity_odd, n_items, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out, out.p, n * out_b, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_errs, er.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_bptc_128x77_batch(const uint8_t* d_in128, uint8_t* d_out77, uint32_t* d_errs, int n_items, void* stream) {
    return bptc_small_batch(0, d_in128, d_out77, d_errs, 0, n_items, stream);
}

int
dsdneo_b200_bptc_128x77_batch_host(const uint8_t* h_in128, uint8_t* h_out77, uint32_t* h_errs, int n_items) {
    return bptc_small_batch_host(0, h_in128, h_out77, h_errs, 0, n_items);
}

int
dsdneo_b200_bptc_16x2_batch(const uint8_t* d_in32, uint8_t* d_out32, uint32_t* d_errs, int parity_odd, int n_items, void* stream) {
    return bptc_small_batch(1, d_in32, d_out32, d_errs, parity_odd, n_items, stream);
}

int
dsdneo_b200_bptc_16x2_batch_host(const uint8_t* h_in32, uint8_t* h_out32, uint32_t* h_errs, int parity_odd, int n_items) {
    return bptc_small_batch_host(1, h_in32, h_out32, h_errs, parity_odd, n_items);
}

int
dsdneo_b200_p25_12_soft_llr_batch(const int16_t* d_llr196, uint8_t* d_out12, int32_t* d_metric, int n_blocks, void* stream) {
    if (!d_llr196 || !d_out12 || !d_metric || n_blocks < 0) {
        set_error("p25_12_soft_llr_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_blocks == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25_12_soft_llr_kernel", s);
        p25_12_soft_llr_kernel<<<grid_for(n_blocks, 64), 64, 0, s>>>(d_llr196, d_out12, d_metric, n_blocks);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25_12_soft_llr_batch_host(const int16_t* h_llr196, uint8_t* h_out12, int32_t* h_metric, int n_blocks) {
    if (!h_llr196 || !h_out12 || !h_metric || n_blocks < 0) {
        set_error("p25_12_soft_llr_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_blocks == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_blocks;
    DevBuf llr(n * 392), out(n * 12), met(n * 4);
    DSDNEO_CUDA(llr.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(met.err);
    DSDNEO_CUDA(cudaMemcpy(llr.p, h_llr196, n * 392, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_p25_12_soft_llr_batch(llr.as<int16_t>(), out.as<uint8_t>(), met.as<int32_t>(), n_blocks, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out12, out.p, n * 12, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_metric, met.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_p25_12_soft_llr_list_batch(const int16_t* d_llr196, dsdneo_b200_p25_12_candidate* d_cands, int32_t* d_count,
                                       int max_candidates, int n_blocks, void* stream) {
    if (!d_llr196 || !d_cands || !d_count || max_candidates <= 0 || n_blocks < 0) {
        set_error("p25_12_soft_llr_list_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_blocks == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25_12_soft_llr_list_kernel", s);
        p25_12_soft_llr_list_kernel<<<grid_for(n_blocks, 64), 64, 0, s>>>(d_llr196, d_cands, d_count, max_candidates, n_blocks);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25_12_soft_llr_list_batch_host(const int16_t* h_llr196, dsdneo_b200_p25_12_candidate* h_cands, int32_t* h_count,
                                            int max_candidates, int n_blocks) {
    if (!h_llr196 || !h_cands || !h_count || n_blocks < 0) {
        set_error("p25_12_soft_llr_list_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_blocks == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_blocks;
    DevBuf llr(n * 392), cands(n * kListK * sizeof(dsdneo_b200_p25_12_candidate)), cnt(n * 4);
    DSDNEO_CUDA(llr.err);
    DSDNEO_CUDA(cands.err);
    DSDNEO_CUDA(cnt.err);
    DSDNEO_CUDA(cudaMemcpy(llr.p, h_llr196, n * 392, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemset(cands.p, 0, n * kListK * sizeof(dsdneo_b200_p25_12_candidate)));
    rc = dsdneo_b200_p25_12_soft_llr_list_batch(llr.as<int16_t>(), cands.as<dsdneo_b200_p25_12_candidate>(), cnt.as<int32_t>(),
                                                max_candidates, n_blocks, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_cands, cands.p, n * kListK * sizeof(dsdneo_b200_p25_12_candidate), cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_count, cnt.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int
rs_shape(int variant, RsShape* sh) {
    switch (variant) {
        case DSDNEO_P25_RS_36_20_17: *sh = RsShape{36, 20, 8}; return 0;
        case DSDNEO_P25_RS_24_12_13: *sh = RsShape{24, 12, 6}; return 0;
        case DSDNEO_P25_RS_24_16_9: *sh = RsShape{24, 16, 4}; return 0;
        default: set_error("p25_rs_decode: unknown variant %d", variant); return DSDNEO_B200_EINVAL;
    }
}

int
dsdneo_b200_p25_rs_decode_batch(int variant, uint8_t* d_data_bits, const uint8_t* d_parity_bits, uint8_t* d_status, int n_words,
                                void* stream) {
    RsShape sh;
    int rc = rs_shape(variant, &sh);
    if (rc) {
        return rc;
    }
    if (!d_data_bits || !d_parity_bits || !d_status || n_words < 0) {
        set_error("p25_rs_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25_rs_decode_kernel", s);
        p25_rs_decode_kernel<<<grid_for(n_words, 64), 64, 0, s>>>(g_d_tables, sh, d_data_bits, d_parity_bits, d_status, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25_rs_decode_batch_host(int variant, uint8_t* h_data_bits, const uint8_t* h_parity_bits, uint8_t* h_status,
                                     int n_words) {
    RsShape sh;
    int rc = rs_shape(variant, &sh);
    if (rc) {
        return rc;
    }
    if (!h_data_bits || !h_parity_bits || !h_status || n_words < 0) {
        set_error("p25_rs_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words, db = (size_t)sh.n_data * 6, pb = (size_t)(sh.n_total - sh.n_data) * 6;
    DevBuf data(n * db), par(n * pb), st(n);
    DSDNEO_CUDA(data.err);
    DSDNEO_CUDA(par.err);
    DSDNEO_CUDA(st.err);
    DSDNEO_CUDA(cudaMemcpy(data.p, h_data_bits, n * db, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(par.p, h_parity_bits, n * pb, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_p25_rs_decode_batch(variant, data.as<uint8_t>(), par.as<uint8_t>(), st.as<uint8_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_data_bits, data.p, n * db, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_status, st.p, n, cudaMemcpyDeviceToHost));
    return 0;
}

static int
rs_soft_launch(int variant, int mode, uint8_t* d_data_bits, const uint8_t* d_parity_bits, const int32_t* d_erasures,
               int erasure_pitch, const int32_t* d_n_erasures, const uint8_t* d_data_rel, const uint8_t* d_par_rel, int threshold,
               uint8_t* d_status, int n_words, void* stream) {
    RsShape sh;
    int rc = rs_shape(variant, &sh);
    if (rc) {
        return rc;
    }
    const bool args_ok = d_data_bits && d_parity_bits && d_status && n_words >= 0 &&
                         (mode == 0 ? (d_erasures && d_n_erasures && erasure_pitch >= 1) : (d_data_rel && d_par_rel));
    if (!args_ok) {
        set_error("p25_rs soft decode: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25_rs_soft_kernel", s);
        p25_rs_soft_kernel<<<grid_for(n_words, 64), 64, 0, s>>>(g_d_tables, sh, mode, d_data_bits, d_parity_bits, d_erasures,
                                                                d_n_erasures, erasure_pitch, d_data_rel, d_par_rel, threshold,
                                                                d_status, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25_rs_decode_erasures_batch(int variant, uint8_t* d_data_bits, const uint8_t* d_parity_bits, const int32_t* d_erasures,
                                         int erasure_pitch, const int32_t* d_n_erasures, uint8_t* d_status, int n_words,
                                         void* stream) {
    return rs_soft_launch(variant, 0, d_data_bits, d_parity_bits, d_erasures, erasure_pitch, d_n_erasures, NULL, NULL, 0, d_status,
                          n_words, stream);
}

int
dsdneo_b200_p25_rs_soft_reliability_batch(int variant, uint8_t* d_data_bits, const uint8_t* d_parity_bits,
                                          const uint8_t* d_data_reliab, const uint8_t* d_parity_reliab, int erasure_threshold,
                                          uint8_t* d_status, int n_words, void* stream) {
    return rs_soft_launch(variant, 1, d_data_bits, d_parity_bits, NULL, 0, NULL, d_data_reliab, d_parity_reliab, erasure_threshold,
                          d_status, n_words, stream);
}

static int
rs_soft_host(int variant, int mode, uint8_t* h_data_bits, const uint8_t* h_parity_bits, const int32_t* h_erasures, int erasure_pitch,
             const int32_t* h_n_erasures, const uint8_t* h_data_rel, const uint8_t* h_par_rel, int threshold, uint8_t* h_status,
             int n_words) {
    RsShape sh;
    int rc = rs_shape(variant, &sh);
    if (rc) {
        return rc;
    }
    const bool args_ok = h_data_bits && h_parity_bits && h_status && n_words >= 0 &&
                         (mode == 0 ? (h_erasures && h_n_erasures && erasure_pitch >= 1) : (h_data_rel && h_par_rel));
    if (!args_ok) {
        set_error("p25_rs soft decode (host): bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words, n_par = (size_t)(sh.n_total - sh.n_data), db = (size_t)sh.n_data * 6, pb = n_par * 6;
    DevBuf data(n * db), par(n * pb), st(n);
    DevBuf a(mode == 0 ? n * (size_t)erasure_pitch * 4 : n * (size_t)sh.n_data), b(mode == 0 ? n * 4 : n * n_par);
    DSDNEO_CUDA(data.err);
    DSDNEO_CUDA(par.err);
    DSDNEO_CUDA(st.err);
    DSDNEO_CUDA(a.err);
    DSDNEO_CUDA(b.err);
    DSDNEO_CUDA(cudaMemcpy(data.p, h_data_bits, n * db, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(par.p, h_parity_bits, n * pb, cudaMemcpyHostToDevice));
    if (mode == 0) {
        DSDNEO_CUDA(cudaMemcpy(a.p, h_erasures, n * (size_t)erasure_pitch * 4, cudaMemcpyHostToDevice));
        DSDNEO_CUDA(cudaMemcpy(b.p, h_n_erasures, n * 4, cudaMemcpyHostToDevice));
        rc = rs_soft_launch(variant, 0, data.as<uint8_t>(), par.as<uint8_t>(), a.as<int32_t>(), erasure_pitch, b.as<int32_t>(), NULL,
                            NULL, 0, st.as<uint8_t>(), n_words, NULL);
    } else {
        DSDNEO_CUDA(cudaMemcpy(a.p, h_data_rel, n * (size_t)sh.n_data, cudaMemcpyHostToDevice));
        DSDNEO_CUDA(cudaMemcpy(b.p, h_par_rel, n * n_par, cudaMemcpyHostToDevice));
        rc = rs_soft_launch(variant, 1, data.as<uint8_t>(), par.as<uint8_t>(), NULL, 0, NULL, a.as<uint8_t>(), b.as<uint8_t>(),
                            threshold, st.as<uint8_t>(), n_words, NULL);
    }
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_data_bits, data.p, n * db, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_status, st.p, n, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_p25_rs_decode_erasures_batch_host(int variant, uint8_t* h_data_bits, const uint8_t* h_parity_bits,
                                              const int32_t* h_erasures, int erasure_pitch, const int32_t* h_n_erasures,
                                              uint8_t* h_status, int n_words) {
    return rs_soft_host(variant, 0, h_data_bits, h_parity_bits, h_erasures, erasure_pitch, h_n_erasures, NULL, NULL, 0, h_status,
                        n_words);
}

int
dsdneo_b200_p25_rs_soft_reliability_batch_host(int variant, uint8_t* h_data_bits, const uint8_t* h_parity_bits,
                                               const uint8_t* h_data_reliab, const uint8_t* h_parity_reliab, int erasure_threshold,
                                               uint8_t* h_status, int n_words) {
    return rs_soft_host(variant, 1, h_data_bits, h_parity_bits, NULL, 0, NULL, h_data_reliab, h_parity_reliab, erasure_threshold,
                        h_status, n_words);
}

int
dsdneo_b200_p25_word_decode_batch(int code, uint8_t* d_data_bits, const uint8_t* d_parity_bits, uint8_t* d_status, int32_t* d_fixed,
                                  int n_words, void* stream) {
    if (code < DSDNEO_P25_WORD_GOLAY_24_6 || code > DSDNEO_P25_WORD_HAMMING_10_6_3 || !d_data_bits || !d_parity_bits || !d_status
        || n_words < 0) {
        set_error("p25_word_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25_word_decode_kernel", s);
        p25_word_decode_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(code, d_data_bits, d_parity_bits, d_status, d_fixed, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25_word_decode_batch_host(int code, uint8_t* h_data_bits, const uint8_t* h_parity_bits, uint8_t* h_status,
                                       int32_t* h_fixed, int n_words) {
    if (code < DSDNEO_P25_WORD_GOLAY_24_6 || code > DSDNEO_P25_WORD_HAMMING_10_6_3 || !h_data_bits || !h_parity_bits || !h_status
        || n_words < 0) {
        set_error("p25_word_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words, db = (code == DSDNEO_P25_WORD_GOLAY_24_12) ? 12 : 6,
                 pb = (code == DSDNEO_P25_WORD_HAMMING_10_6_3) ? 4 : 12;
    DevBuf data(n * db), par(n * pb), st(n), fx(n * 4);
    DSDNEO_CUDA(data.err);
    DSDNEO_CUDA(par.err);
    DSDNEO_CUDA(st.err);
    DSDNEO_CUDA(fx.err);
    DSDNEO_CUDA(cudaMemcpy(data.p, h_data_bits, n * db, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(par.p, h_parity_bits, n * pb, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_p25_word_decode_batch(code, data.as<uint8_t>(), par.as<uint8_t>(), st.as<uint8_t>(), fx.as<int32_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_data_bits, data.p, n * db, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_status, st.p, n, cudaMemcpyDeviceToHost));
    if (h_fixed) {
        DSDNEO_CUDA(cudaMemcpy(h_fixed, fx.p, n * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int
dsdneo_b200_bch_63_16_decode_batch(const uint8_t* d_in63, uint8_t* d_out16, uint8_t* d_ok, int32_t* d_err_count, int n_words,
                                   void* stream) {
    if (!d_in63 || !d_out16 || !d_ok || n_words < 0) {
        set_error("bch_63_16_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("bch_63_16_kernel", s);
        bch_63_16_kernel<<<grid_for(n_words, 64), 64, 0, s>>>(g_d_tables, d_in63, d_out16, d_ok, d_err_count, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_bch_63_16_decode_batch_host(const uint8_t* h_in63, uint8_t* h_out16, uint8_t* h_ok, int32_t* h_err_count, int n_words) {
    if (!h_in63 || !h_out16 || !h_ok || n_words < 0) {
        set_error("bch_63_16_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words;
    DevBuf in(n * 63), out(n * 16), ok(n), ec(n * 4);
    DSDNEO_CUDA(in.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(ok.err);
    DSDNEO_CUDA(ec.err);
    DSDNEO_CUDA(cudaMemcpy(in.p, h_in63, n * 63, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(out.p, h_out16, n * 16, cudaMemcpyHostToDevice)); /* failed words leave the caller's bits untouched */
    rc = dsdneo_b200_bch_63_16_decode_batch(in.as<uint8_t>(), out.as<uint8_t>(), ok.as<uint8_t>(), ec.as<int32_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out16, out.p, n * 16, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_ok, ok.p, n, cudaMemcpyDeviceToHost));
    if (h_err_count) {
        DSDNEO_CUDA(cudaMemcpy(h_err_count, ec.p, n * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int
dsdneo_b200_p25_golay_soft_batch(int code, uint8_t* d_data_bits, const uint8_t* d_parity_bits, const int32_t* d_reliab,
                                 int hard_override_enabled, int erasure_threshold, uint8_t* d_status, int32_t* d_fixed,
                                 int n_words, void* stream) {
    if ((code != DSDNEO_P25_WORD_GOLAY_24_6 && code != DSDNEO_P25_WORD_GOLAY_24_12) || !d_data_bits || !d_parity_bits || !d_reliab
        || !d_status || !d_fixed || n_words < 0) {
        set_error("p25_golay_soft_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    const int length = (code == DSDNEO_P25_WORD_GOLAY_24_6) ? 6 : 12;
    {
        KernelTimer kt("p25_golay_soft_kernel", s);
        p25_golay_soft_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(length, d_data_bits, d_parity_bits, d_reliab,
                                                                    hard_override_enabled, erasure_threshold, d_status, d_fixed,
                                                                    n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25_golay_soft_batch_host(int code, uint8_t* h_data_bits, const uint8_t* h_parity_bits, const int32_t* h_reliab,
                                      int hard_override_enabled, int erasure_threshold, uint8_t* h_status, int32_t* h_fixed,
                                      int n_words) {
    if ((code != DSDNEO_P25_WORD_GOLAY_24_6 && code != DSDNEO_P25_WORD_GOLAY_24_12) || !h_data_bits || !h_parity_bits || !h_reliab
        || !h_status || !h_fixed || n_words < 0) {
        set_error("p25_golay_soft_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words, length = (code == DSDNEO_P25_WORD_GOLAY_24_6) ? 6 : 12;
    DevBuf dat(n * length), par(n * 12), rel(n * (length + 12) * 4), st(n), fx(n * 4);
    DSDNEO_CUDA(dat.err);
    DSDNEO_CUDA(par.err);
    DSDNEO_CUDA(rel.err);
    DSDNEO_CUDA(st.err);
    DSDNEO_CUDA(fx.err);
    DSDNEO_CUDA(cudaMemcpy(dat.p, h_data_bits, n * length, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(par.p, h_parity_bits, n * 12, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(rel.p, h_reliab, n * (length + 12) * 4, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_p25_golay_soft_batch(code, dat.as<uint8_t>(), par.as<uint8_t>(), rel.as<int32_t>(), hard_override_enabled,
                                          erasure_threshold, st.as<uint8_t>(), fx.as<int32_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_data_bits, dat.p, n * length, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_status, st.p, n, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_fixed, fx.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_hamming_10_6_3_soft_batch(const uint8_t* d_bits10, const int32_t* d_reliab10, int hard_override_enabled,
                                      int erasure_threshold, uint8_t* d_out10, uint8_t* d_status, int n_words, void* stream) {
    if (!d_bits10 || !d_reliab10 || !d_out10 || !d_status || n_words < 0) {
        set_error("hamming_10_6_3_soft_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("hamming_10_6_3_soft_kernel", s);
        hamming_10_6_3_soft_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(d_bits10, d_reliab10, hard_override_enabled,
                                                                         erasure_threshold, d_out10, d_status, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_hamming_10_6_3_soft_batch_host(const uint8_t* h_bits10, const int32_t* h_reliab10, int hard_override_enabled,
                                           int erasure_threshold, uint8_t* h_out10, uint8_t* h_status, int n_words) {
    if (!h_bits10 || !h_reliab10 || !h_out10 || !h_status || n_words < 0) {
        set_error("hamming_10_6_3_soft_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words;
    DevBuf in(n * 10), rel(n * 40), out(n * 10), st(n);
    DSDNEO_CUDA(in.err);
    DSDNEO_CUDA(rel.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(st.err);
    DSDNEO_CUDA(cudaMemcpy(in.p, h_bits10, n * 10, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(rel.p, h_reliab10, n * 40, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_hamming_10_6_3_soft_batch(in.as<uint8_t>(), rel.as<int32_t>(), hard_override_enabled, erasure_threshold,
                                               out.as<uint8_t>(), st.as<uint8_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out10, out.p, n * 10, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_status, st.p, n, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_p25p1_nid_decode_batch(const uint8_t* d_code63, const uint8_t* d_reliab63, const int32_t* d_observed_nac,
                                   const uint8_t* d_parity, const uint8_t* d_parity_reliab, int erasure_threshold,
                                   int8_t* d_status, int32_t* d_nac, uint8_t* d_duid, int32_t* d_error_count, int n_words,
                                   void* stream) {
    if (!d_code63 || !d_parity || !d_status || !d_nac || !d_duid || !d_error_count || n_words < 0) {
        set_error("p25p1_nid_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_tables();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25p1_nid_decode_kernel", s);
        p25p1_nid_decode_kernel<<<grid_for(n_words, 128), 128, 0, s>>>(g_d_tables, d_code63, d_reliab63, d_observed_nac, d_parity,
                                                                    d_parity_reliab, erasure_threshold, d_status, d_nac, d_duid,
                                                                    d_error_count, n_words);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25p1_nid_decode_batch_host(const uint8_t* h_code63, const uint8_t* h_reliab63, const int32_t* h_observed_nac,
                                        const uint8_t* h_parity, const uint8_t* h_parity_reliab, int erasure_threshold,
                                        int8_t* h_status, int32_t* h_nac, uint8_t* h_duid, int32_t* h_error_count, int n_words) {
    if (!h_code63 || !h_parity || !h_status || !h_nac || !h_duid || !h_error_count || n_words < 0) {
        set_error("p25p1_nid_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_words == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_words;
    DevBuf code(n * 63), rel(n * 63), obs(n * 4), par(n), prel(n), st(n), nac(n * 4), duid(n), ec(n * 4);
    DSDNEO_CUDA(code.err);
    DSDNEO_CUDA(rel.err);
    DSDNEO_CUDA(obs.err);
    DSDNEO_CUDA(par.err);
    DSDNEO_CUDA(prel.err);
    DSDNEO_CUDA(st.err);
    DSDNEO_CUDA(nac.err);
    DSDNEO_CUDA(duid.err);
    DSDNEO_CUDA(ec.err);
    DSDNEO_CUDA(cudaMemcpy(code.p, h_code63, n * 63, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(par.p, h_parity, n, cudaMemcpyHostToDevice));
    if (h_reliab63) {
        DSDNEO_CUDA(cudaMemcpy(rel.p, h_reliab63, n * 63, cudaMemcpyHostToDevice));
    }
    if (h_observed_nac) {
        DSDNEO_CUDA(cudaMemcpy(obs.p, h_observed_nac, n * 4, cudaMemcpyHostToDevice));
    }
    if (h_parity_reliab) {
        DSDNEO_CUDA(cudaMemcpy(prel.p, h_parity_reliab, n, cudaMemcpyHostToDevice));
    }
    rc = dsdneo_b200_p25p1_nid_decode_batch(code.as<uint8_t>(), h_reliab63 ? rel.as<uint8_t>() : NULL,
                                            h_observed_nac ? obs.as<int32_t>() : NULL, par.as<uint8_t>(),
                                            h_parity_reliab ? prel.as<uint8_t>() : NULL, erasure_threshold, st.as<int8_t>(),
                                            nac.as<int32_t>(), duid.as<uint8_t>(), ec.as<int32_t>(), n_words, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_status, st.p, n, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_nac, nac.p, n * 4, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_duid, duid.p, n, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_error_count, ec.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_p25p1_frames_decode_batch(const uint8_t* d_dibits, size_t dibit_pitch, const int16_t* d_llr, size_t llr_pitch,
                                      const int32_t* d_counts, const dsdneo_b200_sync_hit* d_hits, const int32_t* d_n_hits,
                                      int n_channels, int max_hits, int region_offset, const long long* d_stream_base,
                                      const int8_t* d_nid_status, const uint8_t* d_nid_valid, const int32_t* d_nid_nac,
                                      const uint8_t* d_nid_duid, const int32_t* d_nid_errs, int erasure_threshold, int hard_override_enabled,
                                      int32_t* d_frame_off, int32_t* d_voice_off, int32_t* d_totals,
                                      dsdneo_b200_p25p1_frame* d_frames, int frame_capacity, dsdneo_b200_p25p1_voice* d_voices,
                                      int voice_capacity, void* stream) {
    if (!d_dibits || !d_llr || !d_counts || !d_hits || !d_n_hits || !d_nid_status || !d_nid_nac || !d_nid_duid || !d_nid_errs
        || !d_frame_off || !d_voice_off || !d_totals || !d_frames || n_channels <= 0 || max_hits <= 0 || max_hits > 32
        || frame_capacity < 0 || voice_capacity < 0 || (voice_capacity > 0 && !d_voices)) {
        set_error("p25p1_frames_decode_batch: bad argument (max_hits must be 1..32)");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_tables();
    if (rc) {
        return rc;
    }
    P25FrameParams p;
    p.dibits = d_dibits, p.dibit_pitch = dibit_pitch, p.llr = d_llr, p.llr_pitch = llr_pitch, p.counts = d_counts;
    p.hits = reinterpret_cast<const int32_t*>(d_hits), p.n_hits = d_n_hits, p.region_off = region_offset;
    p.stream_base = d_stream_base, p.n_channels = n_channels, p.max_hits = max_hits;
    p.nid_status = d_nid_status, p.nid_valid = d_nid_valid, p.nid_nac = d_nid_nac, p.nid_duid = d_nid_duid, p.nid_errs = d_nid_errs;
    p.frame_off = d_frame_off, p.voice_off = d_voice_off, p.totals = d_totals;
    p.frames = d_frames, p.voices = d_voices, p.frame_capacity = frame_capacity, p.voice_capacity = voice_capacity;
    p.threshold = erasure_threshold, p.hard_override = hard_override_enabled ? 1 : 0;
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25p1_frame_index_kernel", s);
        p25p1_frame_index_kernel<<<1, 1024, 0, s>>>(p);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    {
        KernelTimer kt("p25p1_frame_decode_kernel", s);
        const int slots = n_channels * max_hits;
        p25p1_frame_decode_kernel<<<(slots + kFrameWarps - 1) / kFrameWarps, kFrameWarps * 32, 0, s>>>(g_d_tables, p);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_p25p1_frame_cut_region(const uint8_t* d_dibits, size_t dibit_pitch, const int16_t* d_llr, size_t llr_pitch,
                                  const int32_t* d_counts, const void* d_hits, const int32_t* d_n_hits, int n_channels,
                                  int max_hits, int n_payload, uint8_t* d_nid_code63, uint8_t* d_nid_reliab63,
                                  uint8_t* d_nid_parity, uint8_t* d_nid_parity_reliab, uint8_t* d_nid_valid,
                                  uint8_t* d_payload_dibits, int16_t* d_payload_llr, uint8_t* d_payload_valid, int region_off,
                                  void* stream) {
    if (!d_dibits || !d_llr || !d_counts || !d_hits || !d_n_hits || n_channels <= 0 || max_hits <= 0 || n_payload < 0
        || !d_nid_code63 || !d_nid_reliab63 || !d_nid_parity || !d_nid_parity_reliab || !d_nid_valid || !d_payload_valid
        || (n_payload > 0 && (!d_payload_dibits || !d_payload_llr))) {
        set_error("p25p1_frame_cut_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("p25p1_frame_cut_kernel", s);
        p25p1_frame_cut_kernel<<<(unsigned)(n_channels * max_hits), 128, 0, s>>>(
            d_dibits, dibit_pitch, d_llr, llr_pitch, d_counts, (const int32_t*)d_hits, d_n_hits, n_channels, max_hits, n_payload,
            d_nid_code63, d_nid_reliab63, d_nid_parity, d_nid_parity_reliab, d_nid_valid, d_payload_dibits, d_payload_llr,
            d_payload_valid, region_off);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25p1_frame_cut_batch(const uint8_t* d_dibits, size_t dibit_pitch, const int16_t* d_llr, size_t llr_pitch,
                                  const int32_t* d_counts, const void* d_hits, const int32_t* d_n_hits, int n_channels,
                                  int max_hits, int n_payload, uint8_t* d_nid_code63, uint8_t* d_nid_reliab63,
                                  uint8_t* d_nid_parity, uint8_t* d_nid_parity_reliab, uint8_t* d_nid_valid,
                                  uint8_t* d_payload_dibits, int16_t* d_payload_llr, uint8_t* d_payload_valid, void* stream) {
    return dsdneo_p25p1_frame_cut_region(d_dibits, dibit_pitch, d_llr, llr_pitch, d_counts, d_hits, d_n_hits, n_channels, max_hits,
                                         n_payload, d_nid_code63, d_nid_reliab63, d_nid_parity, d_nid_parity_reliab, d_nid_valid,
                                         d_payload_dibits, d_payload_llr, d_payload_valid, 0, stream);
}

int
dsdneo_b200_dmr_burst_cut_batch(const uint8_t* d_dibits, size_t dibit_pitch, const uint8_t* d_reliability, size_t reliability_pitch,
                                const int32_t* d_counts, const void* d_hits, const int32_t* d_n_hits, int n_channels, int max_hits,
                                int inverted_dmr, uint8_t* d_cach24, uint8_t* d_info196, uint8_t* d_rel98, uint8_t* d_slot_type20,
                                uint8_t* d_valid, void* stream) {
    if (!d_dibits || !d_reliability || !d_counts || !d_hits || !d_n_hits || n_channels <= 0 || max_hits <= 0 || !d_cach24
        || !d_info196 || !d_rel98 || !d_slot_type20 || !d_valid) {
        set_error("dmr_burst_cut_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("dmr_burst_cut_kernel", s);
        dmr_burst_cut_kernel<<<(unsigned)(n_channels * max_hits), 128, 0, s>>>(
            d_dibits, dibit_pitch, d_reliability, reliability_pitch, d_counts, (const int32_t*)d_hits, d_n_hits, n_channels, max_hits,
            inverted_dmr ? 1 : 0, d_cach24, d_info196, d_rel98, d_slot_type20, d_valid);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_viterbi_k5_decode_batch(const uint16_t* d_cost, size_t cost_pitch, int in_len, const uint8_t* d_punct, int p_len,
                                    uint8_t* d_out, size_t out_pitch, uint32_t* d_metric, int n_frames, void* stream) {
    if (!d_cost || !d_out || !d_metric || in_len < 2 || n_frames < 0 || cost_pitch < (size_t)in_len || (d_punct && p_len <= 0)) {
        set_error("viterbi_k5_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("viterbi_k5_kernel", s);
        viterbi_k5_kernel<<<grid_for(n_frames, 64), 64, 0, s>>>(d_cost, cost_pitch, in_len, d_punct, p_len, d_out, out_pitch, d_metric,
                                                                n_frames);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_viterbi_k5_decode_batch_host(const uint16_t* h_cost, size_t cost_pitch, int in_len, const uint8_t* h_punct, int p_len,
                                         uint8_t* h_out, size_t out_pitch, uint32_t* h_metric, int n_frames) {
    if (!h_cost || !h_out || !h_metric || n_frames < 0) {
        set_error("viterbi_k5_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_frames;
    DevBuf cost(n * cost_pitch * 2), out(n * out_pitch), met(n * 4), pun(h_punct ? (size_t)p_len : 1);
    DSDNEO_CUDA(cost.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(met.err);
    DSDNEO_CUDA(pun.err);
    DSDNEO_CUDA(cudaMemcpy(cost.p, h_cost, n * cost_pitch * 2, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(out.p, h_out, n * out_pitch, cudaMemcpyHostToDevice)); /* bytes past the cleared prefix are OR-ed into */
    if (h_punct) {
        DSDNEO_CUDA(cudaMemcpy(pun.p, h_punct, (size_t)p_len, cudaMemcpyHostToDevice));
    }
    rc = dsdneo_b200_viterbi_k5_decode_batch(cost.as<uint16_t>(), cost_pitch, in_len, h_punct ? pun.as<uint8_t>() : NULL, p_len,
                                             out.as<uint8_t>(), out_pitch, met.as<uint32_t>(), n_frames, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_out, out.p, n * out_pitch, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_metric, met.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_b200_nxdn_conv_decode_batch(const uint8_t* d_sym, const uint8_t* d_rel, size_t pitch, int n_steps, int n_bits_out,
                                   uint16_t* d_metrics, uint8_t* d_out, size_t out_pitch, int n_frames, void* stream) {
    if (!d_sym || !d_metrics || !d_out || n_steps < 1 || n_steps > kNxdnMaxSteps || n_bits_out < 0 || n_bits_out > n_steps
        || pitch < (size_t)(2 * n_steps) || out_pitch * 8 < (size_t)n_bits_out || n_frames < 0) {
        set_error("nxdn_conv_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("nxdn_conv_kernel", s);
        nxdn_conv_kernel<<<grid_for(n_frames, 64), 64, 0, s>>>(d_sym, d_rel, pitch, n_steps, n_bits_out, d_metrics, d_out, out_pitch,
                                                               n_frames);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_nxdn_conv_decode_batch_host(const uint8_t* h_sym, const uint8_t* h_rel, size_t pitch, int n_steps, int n_bits_out,
                                        uint16_t* h_metrics, uint8_t* h_out, size_t out_pitch, int n_frames) {
    if (!h_sym || !h_metrics || !h_out || n_frames < 0) {
        set_error("nxdn_conv_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_frames;
    DevBuf sym(n * pitch), rel(h_rel ? n * pitch : 1), met(n * 64), out(n * out_pitch);
    DSDNEO_CUDA(sym.err);
    DSDNEO_CUDA(rel.err);
    DSDNEO_CUDA(met.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(cudaMemcpy(sym.p, h_sym, n * pitch, cudaMemcpyHostToDevice));
    if (h_rel) {
        DSDNEO_CUDA(cudaMemcpy(rel.p, h_rel, n * pitch, cudaMemcpyHostToDevice));
    }
    DSDNEO_CUDA(cudaMemcpy(met.p, h_metrics, n * 64, cudaMemcpyHostToDevice));
    DSDNEO_CUDA(cudaMemcpy(out.p, h_out, n * out_pitch, cudaMemcpyHostToDevice));
    rc = dsdneo_b200_nxdn_conv_decode_batch(sym.as<uint8_t>(), h_rel ? rel.as<uint8_t>() : NULL, pitch, n_steps, n_bits_out,
                                            met.as<uint16_t>(), out.as<uint8_t>(), out_pitch, n_frames, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_metrics, met.p, n * 64, cudaMemcpyDeviceToHost));
    DSDNEO_CUDA(cudaMemcpy(h_out, out.p, n * out_pitch, cudaMemcpyDeviceToHost));
    return 0;
}

} /* extern "C" */
