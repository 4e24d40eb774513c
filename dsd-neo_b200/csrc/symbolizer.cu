// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * Sample side of the hot path (K9 + K10 + K11), batched over channels: discriminator samples -> matched filter ->
 * symbol (window mean + jitter timing) -> threshold tracking -> 4-level slice + soft metrics.
 *
 * Reference being replaced, per channel (4-level C4FM family, RTL FSK-discriminator input, rf_mod == 0):
 *   p25_filter/dmr_filter/... -> apply_sps_fir   src/dsp/dsd_filters.c:172-201, selection src/dsp/dsd_symbol.c:301-337
 *   getSymbol                                    src/dsp/dsd_symbol.c:1853-1880 (+ :197-225,:347-516,:1306-1387,:1769-1796)
 *   use_symbol / dsd_state_push_minmax_window    src/core/frames/dsd_dibit.c:195-299, include/dsd-neo/core/state.h:1388-1454
 *   digitize / compute_dibit_soft_metric         src/core/frames/dsd_dibit.c:455-721,963-1041
 *   getDibitSoft                                 src/core/frames/dsd_dibit.c:1043-1089
 *
 * Two kernels:
 *   sps_fir_kernel      time-parallel.  The matched filter is a plain causal FIR over the sample stream (the symbol
 *                       timing never feeds back into it), so it is evaluated for every sample up front, in the
 *                       reference's accumulation order (taps oldest -> newest, mul then add, no FMA).
 *   symbolize_kernel    time-serial, one WARP per channel: everything with a loop-carried dependence -- the +-1-sample
 *                       jitter nudge, clip, window mean, the threshold tracker -- runs warp-uniform; the lanes share out the
 *                       memory work (coalesced sample staging, the 128-symbol window and the ring chunks in registers)
 *                       and digitize + store the outputs 32 symbols at a time.
 * Bit-exact with the reference (floats included).  The SNR-dependent reliability weight (dsd_dibit.c:504-546) is a per-channel
 * parameter (dsdneo_b200_symbolizer_set_snr); its default is the reference's value for "no SNR hook installed" (w256 = 0).
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

constexpr int kMaxTaps = DSDNEO_B200_SYM_MAX_TAPS; /* 256 */
constexpr int kCarry = 256;                        /* leftover samples carried between launches: the longest symbol, or the
                                                      samples an unfinished acquisition still needs (kAcqMargin) */
constexpr int kSbuf = 128;
constexpr int kMinMax = 1024;

/* two smallest / two largest elements of a multiset, order independent (the reference's scan, dsd_dibit.c:264-289,
 * finds exactly these: duplicates count as separate elements) */
__device__ __forceinline__ void
two_min_push(float& m1, float& m2, float v) {
    m2 = fminf(m2, fmaxf(m1, v));
    m1 = fminf(m1, v);
}

__device__ __forceinline__ void
two_max_push(float& m1, float& m2, float v) {
    m2 = fmaxf(m2, fminf(m1, v));
    m1 = fmaxf(m1, v);
}

/* per-channel scalars, struct-of-arrays on the device */
struct SymScalars {
    int* filter;      /* index into the filter table, -1 = none */
    int* window_l;
    int* track;
    int* negative;
    int* rf_mod;
    int* sps_num;
    int* sps_den;
    int* sps_accum;
    int* sps;
    int* center_idx;
    int* jitter;
    float* lastsample;
    float* vmin;
    float* vmax;
    float* center;
    float* umid;
    float* lmid;
    float* minref;
    float* maxref;
    int* sidx;
    int* midx;
    int* sum_window;
    double* minbuf_sum;
    double* maxbuf_sum;
    int* carry_n;
    int* snr_num;
    long long* symbolcnt;
};

struct FirParams {
    const float* in;     /* [n_ch][in_pitch] discriminator samples */
    size_t in_pitch;
    float* out;          /* [n_ch][out_pitch] filtered */
    size_t out_pitch;
    const float* taps;   /* [n_filters][kMaxTaps] */
    const int* taps_len; /* [n_filters] */
    const int* filter;   /* [n_ch] */
    const float* hist;   /* [n_ch][kMaxTaps] last taps_len-1 inputs */
    int n;
};

constexpr int kFirTile = 1024;
constexpr int kFirThreads = 256;

__global__ void __launch_bounds__(kFirThreads)
sps_fir_kernel(const FirParams p) {
    __shared__ float W[kFirTile + kMaxTaps];
    __shared__ float T[kMaxTaps];
    const int ch = blockIdx.y;
    const int t0 = blockIdx.x * kFirTile;
    const int tid = threadIdx.x;
    const int f = p.filter[ch];
    const float* x = p.in + (size_t)ch * p.in_pitch;
    float* y = p.out + (size_t)ch * p.out_pitch;
    if (f < 0) { /* no matched filter selected (dsd_symbol.c:301-337 falls through): identity */
        for (int j = tid; j < kFirTile; j += kFirThreads) {
            const int n = t0 + j;
            if (n < p.n) {
                y[n] = x[n];
            }
        }
        return;
    }
    const int L = p.taps_len[f];
    for (int i = tid; i < L; i += kFirThreads) {
        T[i] = p.taps[f * kMaxTaps + i];
    }
    const float* hist = p.hist + (size_t)ch * kMaxTaps;
    for (int j = tid; j < kFirTile + L - 1; j += kFirThreads) {
        const int g = t0 - (L - 1) + j; /* W[j] = x[g] */
        float v = 0.0f;
        if (g >= 0) {
            v = g < p.n ? x[g] : 0.0f;
        } else {
            v = hist[(L - 1) + g]; /* hist[L-2] == x[-1] */
        }
        W[j] = v;
    }
    __syncthreads();
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int i = 0; i < L; i++) { /* taps oldest -> newest, separate multiply and add (dsd_filters.c:191-199) */
        const float t = T[i];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc[j] = __fadd_rn(acc[j], __fmul_rn(t, W[tid + kFirThreads * j + i]));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int n = t0 + tid + kFirThreads * j;
        if (n < p.n) {
            y[n] = acc[j];
        }
    }
}

/*
 * Fast form of the matched filter: one CTA per (channel, 2048 outputs); every thread produces 8 consecutive outputs from a
 * sliding register window (two LDS.128 of samples and two broadcast LDS.128 of taps per 8 taps and 128 FP32 operations), so the
 * kernel runs at the FP32 pipe instead of the shared-memory port (the 4-outputs-per-thread form above does one LDS per 2 FP32
 * operations).  The sample tile is staged with ONE bulk asynchronous copy (cp.async.bulk global -> shared, completion on an
 * mbarrier, TMA unit) issued by thread 0; the few samples that come from the carried history or lie beyond the stream are
 * patched in by the threads.  Same per-output operation order as the reference: taps oldest -> newest, multiply then add.
 */
constexpr int kFir8Threads = 256;
constexpr int kFir8Tile = kFir8Threads * 8;

__device__ __forceinline__ void
mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void
mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void
bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(kFir8Threads)
sps_fir8_kernel(const FirParams p) {
    /* S[k] = x[a0 + k], a0 = the 16-byte aligned sample index at or below t0 - (L - 1); W[j] = S[j + extra] */
    __shared__ __align__(128) float S[kFir8Tile + kMaxTaps + 16];
    __shared__ float P[(kFir8Tile + kMaxTaps + 16) / 8 * 9 + 16];
    __shared__ __align__(16) float T[kMaxTaps + 8];
    __shared__ __align__(8) unsigned long long bar;
    const int ch = blockIdx.y;
    const int t0 = blockIdx.x * kFir8Tile;
    const int tid = threadIdx.x;
    const int f = p.filter[ch];
    const float* x = p.in + (size_t)ch * p.in_pitch;
    float* y = p.out + (size_t)ch * p.out_pitch;
    if (f < 0) { /* no matched filter selected (dsd_symbol.c:301-337 falls through): identity */
        for (int j = tid; j < kFir8Tile; j += kFir8Threads) {
            const int n = t0 + j;
            if (n < p.n) {
                y[n] = x[n];
            }
        }
        return;
    }
    const int L = p.taps_len[f];
    const int g0 = t0 - (L - 1);
    const int a0 = g0 & ~3;                       /* floor to a multiple of 4 (also for negative g0) */
    const int extra = g0 - a0;                    /* 0..3 */
    const int need = kFir8Tile + L - 1 + extra;   /* S entries the tile reads */
    /* the part of [a0, a0 + need) that lies inside the row: [lo, hi), both multiples of 4 (the row pitch is one too) */
    const int lo = a0 < 0 ? 0 : a0;
    int hi = a0 + ((need + 3) & ~3);
    const int row_end = (int)((p.n + 3) & ~3);
    hi = hi > row_end ? row_end : hi;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
    if (tid == 0) {
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned bytes = hi > lo ? (unsigned)(hi - lo) * 4u : 0u;
        if (bytes) {
            mbar_expect_tx(bar_s, bytes);
            bulk_g2s((unsigned)__cvta_generic_to_shared(&S[lo - a0]), x + lo, bytes, bar_s);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
        }
    }
    for (int i = tid; i < kMaxTaps + 8; i += kFir8Threads) {
        T[i] = i < L ? p.taps[f * kMaxTaps + i] : 0.0f;
    }
    mbar_wait(bar_s, 0);
    /* Re-layout into P with one pad word per 8 samples (index k -> k + k / 8): thread t then walks its window from 9 t, so the
     * 32 lanes of every scalar LDS hit 32 different banks (the raw tile at stride 8 would be an 8-way conflict).  The same pass
     * patches what the bulk copy could not supply: carried history left of the stream, zeros right of it. */
    const float* hist = p.hist + (size_t)ch * kMaxTaps;
    for (int k = tid; k < kFir8Tile + L - 1 + 8; k += kFir8Threads) {
        const int g = g0 + k;
        float v;
        if (g < 0) {
            const int h = (L - 1) + g; /* hist[L-2] == x[-1] */
            v = h >= 0 ? hist[h] : 0.0f;
        } else if (g >= p.n) {
            v = 0.0f;
        } else {
            v = S[k + extra];
        }
        P[k + (k >> 3)] = v;
    }
    __syncthreads();
    const float* W = P + 9 * tid; /* sample 8 t + i of the window sits at W[i + i / 8] */
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        acc[j] = 0.0f;
    }
    float w[16];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        w[j] = W[j];
    }
    const int full = L >> 3;
    for (int q = 0; q < full; q++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            w[8 + j] = W[9 * q + 9 + j];
        }
        const float4 ta = *reinterpret_cast<const float4*>(&T[8 * q]);
        const float4 tb = *reinterpret_cast<const float4*>(&T[8 * q + 4]);
        const float t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc[j] = __fadd_rn(acc[j], __fmul_rn(t[i], w[i + j]));
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            w[j] = w[8 + j];
        }
    }
    const int rem = L & 7;
    if (rem) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            w[8 + j] = W[9 * full + 9 + j];
        }
#pragma unroll
        for (int i = 0; i < 7; i++) {
            if (i < rem) {
                const float t = T[8 * full + i];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[j] = __fadd_rn(acc[j], __fmul_rn(t, w[i + j]));
                }
            }
        }
    }
    const int n0 = t0 + 8 * tid;
    if (n0 + 8 <= p.n && (p.out_pitch & 3) == 0) {
        reinterpret_cast<float4*>(y + n0)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(y + n0)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (n0 + j < p.n) {
                y[n0 + j] = acc[j];
            }
        }
    }
}

__global__ void
sps_fir_hist_kernel(const float* in, size_t in_pitch, float* hist_all, const int* filter, const int* taps_len, int n_ch, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * (blockDim.x >> 5) + warp;
    if (ch >= n_ch) {
        return;
    }
    const int f = filter[ch];
    if (f < 0) {
        return;
    }
    const int hl = taps_len[f] - 1;
    float* hist = hist_all + (size_t)ch * kMaxTaps;
    const float* x = in + (size_t)ch * in_pitch;
    if (n >= hl) {
        for (int k = lane; k < hl; k += 32) {
            hist[k] = x[n - hl + k];
        }
    } else {
        float keep[kMaxTaps / 32];
        int c = 0;
        for (int k = lane; k < hl - n; k += 32) {
            keep[c++] = hist[k + n];
        }
        __syncwarp();
        c = 0;
        for (int k = lane; k < hl - n; k += 32) {
            hist[k] = keep[c++];
        }
        for (int k = lane; k < n; k += 32) {
            hist[hl - n + k] = x[k];
        }
    }
}

struct SymParams {
    SymScalars s;
    const float* filt;   /* [n_ch][filt_pitch] matched-filter output for this launch */
    size_t filt_pitch;
    float* carry;        /* [n_ch][kCarry], right-aligned: the last carry_n entries are the unconsumed tail */
    float* sbuf;         /* [n_ch][kSbuf] */
    float* minbuf;       /* [n_ch][kMinMax] */
    float* maxbuf;       /* [n_ch][kMinMax] */
    float* symbols;      /* outputs, [n_ch][out_pitch] */
    uint8_t* dibits;
    uint8_t* reliab;
    int16_t* llr;        /* [n_ch][out_pitch][2] */
    int* count;          /* [n_ch] */
    const int* snr_num;  /* [n_ch] reliability weight numerator 204 + (w256 >> 2), dsd_dibit.c:504-546 */
    const int* start_off; /* optional [n_ch]: samples of this launch already consumed by the acquisition kernel */
    const int* out_off;   /* optional [n_ch]: outputs already written by it */
    const int* acquired;  /* optional [n_ch]: channels still hunting are skipped by symbolize_kernel */
    size_t out_pitch;
    int n_ch, n, mode, have_sync, rate, symrate, ssize, msize;
};

__device__ __forceinline__ int
clamp255(int v) {
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

/* compute_dibit_soft_metric's per-bit magnitudes (dsd_dibit.c:609-642) for both bits at once: the spacing scan and the
 * scale 255 / min_spacing^2 do not depend on the bit index, so they are evaluated once. */
__device__ __forceinline__ void
bit_metrics(float sym, const float (&ideal)[4], int& mag0, int& mag1) {
    const float big = 3.4028234663852886e38f;
    float d[4], min_spacing = big;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float e = __fsub_rn(sym, ideal[i]);
        d[i] = __fmul_rn(e, e);
#pragma unroll
        for (int j = i + 1; j < 4; j++) {
            const float sp = fabsf(__fsub_rn(ideal[i], ideal[j]));
            if (sp > 1e-6f && sp < min_spacing) {
                min_spacing = sp;
            }
        }
    }
    if (min_spacing == big) {
        min_spacing = 2.0f;
    }
    const float scale = __fdiv_rn(255.0f, __fmul_rn(min_spacing, min_spacing));
    /* bit 0 = MSB of the dibit index: {0,1} vs {2,3}; bit 1 = LSB: {0,2} vs {1,3}; strict < keeps the first minimum */
    const float b0_0 = d[1] < d[0] ? d[1] : d[0], b0_1 = d[3] < d[2] ? d[3] : d[2];
    const float b1_0 = d[2] < d[0] ? d[2] : d[0], b1_1 = d[3] < d[1] ? d[3] : d[1];
    mag0 = clamp255(__float2int_rn(__fmul_rn(fabsf(__fsub_rn(b0_0, b0_1)), scale))); /* lrintf: round to nearest even */
    mag1 = clamp255(__float2int_rn(__fmul_rn(fabsf(__fsub_rn(b1_0, b1_1)), scale)));
}

/* thresholds from the tracked extremes (dsd_dibit.c:268-272); x / 2 and x / 8 are exact scalings, the same correctly
 * rounded values as the reference's divisions */
__device__ __forceinline__ void
cq_thresholds(float vmin, float vmax, float& center, float& umid, float& lmid) {
    center = __fmul_rn(__fadd_rn(vmax, vmin), 0.5f);
    umid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmax, center), 5.0f), 0.125f), center);
    lmid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmin, center), 5.0f), 0.125f), center);
}

/* c4fm_reliability_from_thresholds (dsd_dibit.c:455-502) + apply_c4fm_snr_weight (:504-546) = dmr_compute_reliability for
 * rf_mod 0 / 2 (:548-568) */
__device__ __forceinline__ int
c4fm_reliability(float sym, float vmin, float vmax, float center, float umid, float lmid, int snr_num) {
    const float eps = 1e-6f;
    int rel;
    if (sym > umid) {
        float span = __fsub_rn(vmax, umid);
        span = span < eps ? eps : span;
        rel = __float2int_rn(__fdiv_rn(__fmul_rn(__fsub_rn(sym, umid), 255.0f), span));
    } else if (sym > center) {
        const float d1 = __fsub_rn(sym, center), d2 = __fsub_rn(umid, sym);
        float span = __fsub_rn(umid, center);
        span = span < eps ? eps : span;
        rel = __float2int_rn(__fdiv_rn(__fmul_rn(d1 < d2 ? d1 : d2, 510.0f), span));
    } else if (sym >= lmid) {
        const float d1 = __fsub_rn(center, sym), d2 = __fsub_rn(sym, lmid);
        float span = __fsub_rn(center, lmid);
        span = span < eps ? eps : span;
        rel = __float2int_rn(__fdiv_rn(__fmul_rn(d1 < d2 ? d1 : d2, 510.0f), span));
    } else {
        float span = __fsub_rn(lmid, vmin);
        span = span < eps ? eps : span;
        rel = __float2int_rn(__fdiv_rn(__fmul_rn(__fsub_rn(lmid, sym), 255.0f), span));
    }
    rel = clamp255(rel);
    return clamp255((rel * snr_num) >> 8);
}

/* digitize (dsd_dibit.c:963-976,1018-1041) + compute_dibit_soft_metric (:644-721) + c4fm_reliability_from_thresholds
 * (:455-502) + apply_c4fm_snr_weight (:504-546, snr_num = 204 + (w256 >> 2)) for one symbol given the thresholds in force
 * after use_symbol.  Reads nothing else, so the serial kernel hands 32 symbols at a time to its 32 lanes. */
__device__ __forceinline__ void
digitize_one(float sym, float vmin, float vmax, float center, float umid, float lmid, int negative, int snr_num, int& dibit_out,
             int& rel_out, int& l0_out, int& l1_out) {
    int dibit;
    if (sym > center) {
        dibit = sym > umid ? (negative ? 3 : 1) : (negative ? 2 : 0);
    } else {
        dibit = sym < lmid ? (negative ? 1 : 3) : (negative ? 0 : 2);
    }
    const float plus_one = __fmul_rn(0.5f, __fadd_rn(center, umid)), minus_one = __fmul_rn(0.5f, __fadd_rn(lmid, center));
    float ideal[4];
    if (negative) {
        ideal[0] = minus_one, ideal[1] = vmin, ideal[2] = plus_one, ideal[3] = vmax;
    } else {
        ideal[0] = plus_one, ideal[1] = vmax, ideal[2] = minus_one, ideal[3] = vmin;
    }
    int mag0, mag1;
    bit_metrics(sym, ideal, mag0, mag1);
    int rel = c4fm_reliability(sym, vmin, vmax, center, umid, lmid, snr_num);
    const int min_mag = mag0 < mag1 ? mag0 : mag1;
    if (min_mag > 0 && rel < min_mag) {
        mag0 = (mag0 * rel) / min_mag;
        mag1 = (mag1 * rel) / min_mag;
    }
    mag0 = clamp255(mag0);
    mag1 = clamp255(mag1);
    dibit_out = dibit;
    l0_out = ((dibit >> 1) & 1) ? mag0 : -mag0;
    l1_out = (dibit & 1) ? mag1 : -mag1;
    rel_out = clamp255(mag1 < mag0 ? mag1 : mag0);
}

/* order-preserving map float -> int (an involution), so the warp-wide extrema can use redux.sync.min/max.s32 */
__device__ __forceinline__ int
fkey(float f) {
    const int k = __float_as_int(f);
    return k ^ ((k >> 31) & 0x7fffffff);
}

__device__ __forceinline__ float
funkey(int k) {
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}

constexpr int kSymWarps = 4;    /* channels (warps) per CTA */
constexpr int kRing = 1024;     /* staged samples per channel (shared-memory ring) */
constexpr int kRingPad = 8;     /* slots 0..7 are mirrored behind the ring so a 5-sample window never wraps */
constexpr int kLook = 704;      /* samples requested ahead of the read position */
constexpr int kPending = 6;     /* cp.async groups (32 samples each) that may still be in flight after a refill */

/* per-warp shared memory: the sample ring, the 128-symbol window, the cached chunk of the min / max rings and one group
 * of 32 finished symbols waiting for digitize + store */
struct WarpShared {
    float ring[kRing + kRingPad];
    float sbuf[kSbuf];
    float mnr[32], mxr[32];
    float o_sym[32], o_min[32], o_max[32], o_center[32], o_umid[32], o_lmid[32];
};

/*
 * use_symbol's threshold tracker (dsd_dibit.c:243-299, core/state.h:1388-1454) for ONE channel held by ONE warp.
 *
 * The reference rescans sbuf[0..cap) for its two smallest and two largest entries on every symbol (126 compare-and-branch
 * steps).  Here the two smallest / largest are carried: replacing entry e by v leaves them untouched unless e was one of
 * them (e <= mn2 or e >= mx2), so the common case is one two_min_push / two_max_push; otherwise the warp rescans (4 entries
 * per lane, then redux.sync on order-preserving keys).  The two smallest / largest of a multiset do not depend on
 * evaluation order, so both paths give the reference's values.  The 1024-entry min / max rings stay in global memory
 * ([channel][entry]); the 32-entry chunk being replaced is cached in shared memory and written back coalesced.
 * Every lane carries the same scalar state and performs the same (same-value) shared-memory writes: control flow is
 * warp-uniform, nothing diverges, and no lane ever reads a location it has not written itself.
 */
struct WarpTracker {
    float mn1, mn2, mx1, mx2;  /* two smallest / largest of sbuf[0..cap) */
    int cur_chunk, dirty;
    int cap, window, lane;
    double inv_window;
    bool pow2;
    float* minbuf;
    float* maxbuf;
    WarpShared* sh;

    __device__ __forceinline__ void rescan() {
        const float big = 3.4028234663852886e38f;
        float a1 = big, a2 = big, z1 = -big, z2 = -big;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j * 32 + lane < cap) {
                const float v = sh->sbuf[j * 32 + lane];
                two_min_push(a1, a2, v);
                two_max_push(z1, z2, v);
            }
        }
        const int k1 = fkey(a1), k2 = fkey(a2), K1 = fkey(z1), K2 = fkey(z2);
        const int m1 = __reduce_min_sync(0xffffffffu, k1);
        const int M1 = __reduce_max_sync(0xffffffffu, K1);
        const unsigned own_min = __ballot_sync(0xffffffffu, k1 == m1);
        const unsigned own_max = __ballot_sync(0xffffffffu, K1 == M1);
        const int m2 = __reduce_min_sync(0xffffffffu, lane == (__ffs(own_min) - 1) ? k2 : k1);
        const int M2 = __reduce_max_sync(0xffffffffu, lane == (__ffs(own_max) - 1) ? K2 : K1);
        mn1 = funkey(m1), mn2 = funkey(m2), mx1 = funkey(M1), mx2 = funkey(M2);
    }

    __device__ __forceinline__ void flush_chunk() {
        if (dirty && cur_chunk >= 0 && cur_chunk * 32 + lane < kMinMax) {
            minbuf[cur_chunk * 32 + lane] = sh->mnr[lane];
            maxbuf[cur_chunk * 32 + lane] = sh->mxr[lane];
        }
        dirty = 0;
    }

    __device__ __forceinline__ void invalidate_cache() {
        cur_chunk = -1, dirty = 0;
    }

    /* makes the chunk holding ring entry idx current */
    __device__ __forceinline__ void seek(int idx) {
        const int c = idx >> 5;
        if (c == cur_chunk) {
            return;
        }
        flush_chunk();
        __syncwarp();
        const float a = minbuf[c * 32 + lane], b = maxbuf[c * 32 + lane];
        sh->mnr[lane] = a;
        sh->mxr[lane] = b;
        __syncwarp();
        cur_chunk = c;
    }

    __device__ __forceinline__ void init(int lane_, int ssize, int msize, float* minbuf_row, float* maxbuf_row, const float* sbuf_row,
                                         WarpShared* sh_, bool track) {
        lane = lane_;
        sh = sh_;
        cap = ssize < 0 ? 0 : (ssize > kSbuf ? kSbuf : ssize);
        window = msize < 1 ? 1 : (msize > kMinMax ? kMinMax : msize);
        pow2 = (window & (window - 1)) == 0;
        inv_window = 1.0 / (double)window; /* exact for a power of two: x / 2^k == x * 2^-k */
        minbuf = minbuf_row, maxbuf = maxbuf_row;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sh->sbuf[j * 32 + lane] = sbuf_row[j * 32 + lane];
        }
        __syncwarp();
        invalidate_cache();
        mn1 = mn2 = mx1 = mx2 = 0.0f;
        if (track && cap >= 2) {
            rescan();
        }
    }

    __device__ __forceinline__ void store_sbuf(float* sbuf_row) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sbuf_row[j * 32 + lane] = sh->sbuf[j * 32 + lane];
        }
    }

    /* dsd_state_recompute_minmax_sums: sequential f64 sums over ring[0..window), the reference's order */
    __device__ __forceinline__ void recompute_sums(double& min_sum, double& max_sum) {
        flush_chunk();
        __syncwarp();
        double a = 0.0, b = 0.0;
        for (int base = 0; base < window; base += 32) {
            const float vm = base + lane < window ? minbuf[base + lane] : 0.0f;
            const float vx = base + lane < window ? maxbuf[base + lane] : 0.0f;
            const int m = min(32, window - base);
            for (int k = 0; k < m; k++) {
                a += (double)__shfl_sync(0xffffffffu, vm, k);
                b += (double)__shfl_sync(0xffffffffu, vx, k);
            }
        }
        min_sum = a, max_sum = b;
    }

    /* fills ring[0..count) with constants (slicer reset, warm start) */
    __device__ __forceinline__ void fill_rings(float vmin, float vmax, int count) {
        invalidate_cache();
        for (int k = lane; k < count; k += 32) {
            minbuf[k] = vmin;
            maxbuf[k] = vmax;
        }
        __syncwarp();
    }

    /* the extrema part of use_symbol after sbuf[si] (old value e) was replaced by sym */
    __device__ __forceinline__ void extrema(int si, float e, float sym, float& lmin, float& lmax) {
        lmin = 0.0f, lmax = 0.0f;
        if (cap >= 2) {
            if (si < cap) {
                if (e > mn2 && e < mx2) {
                    two_min_push(mn1, mn2, sym);
                    two_max_push(mx1, mx2, sym);
                } else {
                    rescan();
                }
            }
            lmin = __fmul_rn(__fadd_rn(mn1, mn2), 0.5f);
            lmax = __fmul_rn(__fadd_rn(mx1, mx2), 0.5f);
        }
    }

    /* sbuf[sidx] = sym, then the tracker part of use_symbol (general form: any cap / window, lazily recomputed sums) */
    __device__ __forceinline__ void push(float sym, int sidx, int& midx, int& sum_window, double& min_sum, double& max_sum,
                                         float& vmin, float& vmax) {
        /* get_dibit_and_analog_signal stores the symbol first (sidx may sit outside [0, cap) when ssize < 128) */
        const int si = sidx & (kSbuf - 1);
        const float e = sh->sbuf[si];
        sh->sbuf[si] = sym;
        float lmin, lmax;
        extrema(si, e, sym, lmin, lmax);
        if (sum_window != window) {
            recompute_sums(min_sum, max_sum);
            sum_window = window;
            if (midx < 0 || midx >= window) {
                midx = 0;
            }
            invalidate_cache();
        }
        const int idx = (midx < 0 || midx >= window) ? 0 : midx;
        seek(idx);
        const float old_min = sh->mnr[idx & 31], old_max = sh->mxr[idx & 31];
        min_sum += (double)lmin - (double)old_min;
        max_sum += (double)lmax - (double)old_max;
        sh->mnr[idx & 31] = lmin;
        sh->mxr[idx & 31] = lmax;
        dirty = 1;
        midx = idx + 1 >= window ? 0 : idx + 1;
        if (pow2) {
            vmin = (float)(min_sum * inv_window);
            vmax = (float)(max_sum * inv_window);
        } else {
            vmin = (float)(min_sum / (double)window);
            vmax = (float)(max_sum / (double)window);
        }
    }
};

/* x / 5 correctly rounded for the symbol mean: q = x * RN(1/5), one exact-remainder correction step.  Checked against the
 * IEEE division on all 2^32 operands (dsdneo_b200_selftest_div5); operands outside the verified magnitude range take the
 * IEEE operator. */
__device__ __forceinline__ float
div5_rn(float x) {
    const float ax = fabsf(x);
    if (ax > 1e-30f && ax < 1e30f) {
        const float rc = 0.2f;
        const float q = __fmul_rn(x, rc);
        const float r = __fmaf_rn(-5.0f, q, x);
        return __fmaf_rn(r, rc, q);
    }
    return __fdiv_rn(x, 5.0f);
}

/*
 * One warp per channel.  Everything with a loop-carried dependence -- the +-1-sample jitter nudge, clip, window mean,
 * the threshold tracker -- is evaluated by all 32 lanes redundantly (same registers, warp-uniform branches); what the
 * lanes share out is memory: coalesced cp.async staging of the sample stream into a shared-memory ring, the rescan of
 * the 128-entry symbol window, the ring chunks, and the per-symbol outputs, which are collected per group of 32 symbols
 * and digitized + stored one symbol per lane (digitize and the soft metrics only read the symbol and the thresholds in
 * force after use_symbol).  A single warp issues one dependent instruction every ~5 cycles, so the design minimises
 * INSTRUCTIONS per symbol: the synchronised steady state (getDibitSoft on a locked channel: no timing nudge, zero-crossing
 * detector idle) runs in a tight batch loop of up to 32 symbols with every invariant hoisted -- 5 shared-memory loads,
 * clip, 5 adds, the constant division, the tracker update -- and everything else (hunting, fractional samples per symbol,
 * the first symbols after a reset, unusual window sizes) takes the general per-sample path below it.
 */
template <int NW, bool TRACK>
__device__ __forceinline__ int
sym_fast_batch(WarpShared* sh, WarpTracker& tr, int nb, int sps, int lo, int& pos, int& sidx, int midx0, int g0, float& vmin, float& vmax,
               float& last_vmin, float& last_vmax, double& min_sum, double& max_sum) {
    /* returns the number of symbols done (stops early only when vmin > vmax, which the general path handles).
     * Everything that does not depend on the previous symbol's thresholds -- the five window samples, the ring entries
     * about to be replaced, the symbol about to leave the 128-entry window -- is loaded one symbol ahead, so the loop-carried
     * chain is clip -> five adds -> divide -> extrema -> two f64 updates -> thresholds and nothing else. */
    int k = 0;
    int si = sidx;
    float e = sh->sbuf[si];
    int off = (pos + lo) & (kRing - 1);
    const float* b = sh->ring + off; /* the ring is mirrored kRingPad entries past its end: b[0..4] never wraps */
    float r0 = b[0], r1 = b[1], r2 = b[2], r3 = b[3], r4 = NW > 4 ? b[4] : 0.0f;
    int mi = TRACK ? (midx0 & 31) : 0;
    float old_min = TRACK ? sh->mnr[mi] : 0.0f, old_max = TRACK ? sh->mxr[mi] : 0.0f;
#pragma unroll 1
    for (; k < nb; k++) {
        if (!(vmin <= vmax)) {
            break;
        }
        float sum = __fadd_rn(0.0f, fminf(fmaxf(r0, vmin), vmax));
        sum = __fadd_rn(sum, fminf(fmaxf(r1, vmin), vmax));
        sum = __fadd_rn(sum, fminf(fmaxf(r2, vmin), vmax));
        sum = __fadd_rn(sum, fminf(fmaxf(r3, vmin), vmax));
        float sym;
        if (NW > 4) {
            sum = __fadd_rn(sum, fminf(fmaxf(r4, vmin), vmax));
            /* div5_rn's multiply + correction step, unconditionally; its operand-range test runs beside it (off the carried
             * chain) and only decides whether this symbol leaves the batch for the general path, which divides with the IEEE
             * operator.  A sum of exactly +0 (squelched channel; the sum starts from +0 and can never be -0) stays here:
             * q = +0, remainder +0, result +0 == +0 / 5. */
            const float q = __fmul_rn(sum, 0.2f);
            sym = __fmaf_rn(__fmaf_rn(-5.0f, q, sum), 0.2f, q);
            const float ax = fabsf(sum);
            if (__builtin_expect(!(ax < 1e30f && (ax > 1e-30f || ax == 0.0f)), 0)) {
                break;
            }
        } else {
            sym = __fmul_rn(sum, 0.25f); /* exact scaling == the correctly rounded quotient */
        }
        last_vmin = vmin, last_vmax = vmax;
        /* next symbol's operands (independent of this symbol's result) */
        off = (off + sps) & (kRing - 1);
        b = sh->ring + off;
        r0 = b[0], r1 = b[1], r2 = b[2], r3 = b[3];
        if (NW > 4) {
            r4 = b[4];
        }
        const int si_next = (si + 1) & (kSbuf - 1);
        const float e_next = sh->sbuf[si_next]; /* si_next != si: read before this symbol's store, same value either way */
        sh->sbuf[si] = sym;
        if (TRACK) {
            if (e > tr.mn2 && e < tr.mx2) {
                two_min_push(tr.mn1, tr.mn2, sym);
                two_max_push(tr.mx1, tr.mx2, sym);
            } else {
                tr.rescan();
            }
            const float lmin = __fmul_rn(__fadd_rn(tr.mn1, tr.mn2), 0.5f);
            const float lmax = __fmul_rn(__fadd_rn(tr.mx1, tr.mx2), 0.5f);
            min_sum += (double)lmin - (double)old_min;
            max_sum += (double)lmax - (double)old_max;
            const int mi_next = (mi + 1) & 31;
            old_min = sh->mnr[mi_next], old_max = sh->mxr[mi_next]; /* a different entry than the one stored below */
            sh->mnr[mi] = lmin;
            sh->mxr[mi] = lmax;
            vmin = (float)(min_sum * tr.inv_window);
            vmax = (float)(max_sum * tr.inv_window);
            sh->o_min[g0 + k] = vmin;
            sh->o_max[g0 + k] = vmax;
            mi = mi_next;
        }
        sh->o_sym[g0 + k] = sym;
        si = si_next;
        e = e_next;
        pos += sps;
    }
    sidx = si;
    return k;
}

__global__ void __launch_bounds__(kSymWarps * 32)
symbolize_kernel(const SymParams p) {
    __shared__ WarpShared s_warp[kSymWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kSymWarps + warp;
    if (c >= p.n_ch) {
        return;
    }
    WarpShared* sh = &s_warp[warp];
    float* ring = sh->ring;
    if (p.acquired && !p.acquired[c]) {
        return; /* still hunting: sym_acquire_kernel consumed this launch's samples and wrote the outputs and the count */
    }

    /* state -> registers (every lane holds the same copy) */
    const int window_l = p.s.window_l[c], track = p.s.track[c], negative = p.s.negative[c], rf_mod = p.s.rf_mod[c];
    const int snr_num = p.snr_num[c];
    int sps_num = p.s.sps_num[c], sps_den = p.s.sps_den[c], sps_accum = p.s.sps_accum[c];
    int sps = p.s.sps[c], center_idx = p.s.center_idx[c], jitter = p.s.jitter[c];
    float lastsample = p.s.lastsample[c];
    float vmin = p.s.vmin[c], vmax = p.s.vmax[c], center = p.s.center[c], umid = p.s.umid[c], lmid = p.s.lmid[c];
    float minref = p.s.minref[c], maxref = p.s.maxref[c];
    int sidx = p.s.sidx[c], midx = p.s.midx[c], sum_window = p.s.sum_window[c];
    double minbuf_sum = p.s.minbuf_sum[c], maxbuf_sum = p.s.maxbuf_sum[c];
    int carry_n = p.s.carry_n[c];
    carry_n = carry_n < 0 ? 0 : (carry_n > kCarry ? kCarry : carry_n);
    long long symbolcnt = p.s.symbolcnt[c];
    const bool soft = p.mode == DSDNEO_SYM_MODE_GET_DIBIT_SOFT;
    const int have_sync = soft ? 1 : p.have_sync;

    WarpTracker tr;
    tr.init(lane, p.ssize, p.msize, p.minbuf + (size_t)c * kMinMax, p.maxbuf + (size_t)c * kMinMax, p.sbuf + (size_t)c * kSbuf, sh,
            soft && track);

    /* sample stream in q-space: q in [kCarry - carry_n, kCarry) = carried tail, q >= kCarry = this launch's samples */
    const float* filt = p.filt + (size_t)c * p.filt_pitch;
    float* carry = p.carry + (size_t)c * kCarry;
    const int q0 = kCarry - carry_n;
    const int q_end = kCarry + p.n;
    int pos = q0 + (p.start_off ? p.start_off[c] : 0);
    /* staging: 16 bytes per lane and step (128 samples), at most kFillPending steps in flight.  Rows of the filter buffer
     * and of the carry start on 128-byte lines and kCarry is a multiple of 4, so every 4-sample chunk is 16-byte aligned
     * on both sides and lies wholly in the carry or wholly in this launch's samples */
    constexpr int kFill = 128, kFillPending = 2;
    const int q_fill_end = (q_end + kFill - 1) & ~(kFill - 1);
    int whi = pos & ~(kFill - 1);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    auto refill = [&]() {
        while (whi < pos + kLook && whi < q_fill_end) {
            const int q = whi + 4 * lane;
            const float* src = filt;
            int bytes = 0;
            if (q < kCarry) {
                if (q + 4 > q0) {
                    src = carry + q, bytes = 16; /* entries before q0 are stale carry slots nobody reads */
                }
            } else if (q < q_end) {
                src = filt + (q - kCarry);
                bytes = min(16, 4 * (q_end - q));
            }
            const int slot = q & (kRing - 1);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(ring_s + 4u * (unsigned)slot), "l"(src), "r"(bytes) : "memory");
            if (slot < kRingPad) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(ring_s + 4u * (unsigned)(slot + kRing)), "l"(src), "r"(bytes)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            whi += kFill;
        }
        if (whi >= q_fill_end) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group %0;" ::"n"(kFillPending) : "memory");
        }
        __syncwarp();
    };
    auto at = [&](int q) -> float { return ring[q & (kRing - 1)]; };

    /* samples-per-symbol of the reference's fractional accumulator (dsd_symbol.c:1343-1387); invariant part hoisted */
    int whole0 = p.rate / p.symrate, rem0 = p.rate % p.symrate;
    if (whole0 < 2) {
        whole0 = 2, rem0 = 0;
    }
    if (whole0 > 64) {
        whole0 = 64, rem0 = 0;
    }
    const int reserve = whole0 + 2; /* longest possible symbol (one more sample when the accumulator wraps, one for a nudge) */
    float* o_sym = p.symbols + (size_t)c * p.out_pitch;
    uint8_t* o_dib = p.dibits ? p.dibits + (size_t)c * p.out_pitch : nullptr;
    uint8_t* o_rel = p.reliab ? p.reliab + (size_t)c * p.out_pitch : nullptr;
    short2* o_llr = p.llr ? reinterpret_cast<short2*>(p.llr) + (size_t)c * p.out_pitch : nullptr;
    const int nsym0 = p.out_off ? p.out_off[c] : 0; /* outputs the acquisition kernel already wrote for this launch */
    int nsym = nsym0;
    const int out_cap = (int)(p.out_pitch > 0x7fffffffu ? 0x7fffffffu : p.out_pitch);

    /* finished symbols wait in a 32-slot group buffer; when a group is complete lane k digitizes and stores the symbol in
     * slot k.  With the threshold tracker on, slot = position in the tracker's 32-entry ring chunk, so a batch of the fast
     * loop ends at a group boundary and at a chunk boundary at the same time */
    const int gphase = (soft && track && (tr.window & 31) == 0) ? (((midx < 0 || midx >= tr.window) ? 0 : midx) & 31) : 0;
    int flushed = nsym0;
    auto slot_of = [&](int ns) -> int { return (ns - nsym0 + gphase) & 31; };
    auto flush_group = [&]() { /* symbols [flushed, nsym) sit in slots [end - n, end) */
        __syncwarp();
        const int n_in_group = nsym - flushed;
        const int end = slot_of(nsym) == 0 ? 32 : slot_of(nsym);
        const int start = end - n_in_group;
        if (n_in_group > 0 && lane >= start && lane < end) {
            const int o = flushed + (lane - start);
            const float g_sym = sh->o_sym[lane];
            o_sym[o] = g_sym;
            if (soft) {
                float g_min, g_max, g_center, g_umid, g_lmid;
                if (track) { /* centre / mid thresholds follow from {min, max} by the reference's own formulas */
                    g_min = sh->o_min[lane], g_max = sh->o_max[lane];
                    cq_thresholds(g_min, g_max, g_center, g_umid, g_lmid);
                } else {
                    g_min = sh->o_min[lane], g_max = sh->o_max[lane], g_center = sh->o_center[lane];
                    g_umid = sh->o_umid[lane], g_lmid = sh->o_lmid[lane];
                }
                int dibit, rel, l0, l1;
                digitize_one(g_sym, g_min, g_max, g_center, g_umid, g_lmid, negative, snr_num, dibit, rel, l0, l1);
                o_dib[o] = (uint8_t)dibit;
                o_rel[o] = (uint8_t)rel;
                o_llr[o] = make_short2((short)l0, (short)l1);
            }
        }
        flushed = nsym;
        __syncwarp();
    };

    const int lo0 = max((whole0 - 1) / 2 - window_l, 0), hi0 = min((whole0 - 1) / 2 + 2, whole0 - 1);
    const int nw0 = hi0 - lo0 + 1;
    const bool fast_cfg = soft && rf_mod == 0 && rem0 == 0 && whole0 != 20 && whole0 != 5 && (nw0 == 4 || nw0 == 5) && tr.cap == kSbuf
                          && (!track || tr.pow2);

    refill();
    while (nsym < out_cap && (q_end - pos) >= reserve) {
        refill();
        /* ---- steady state of a locked channel: a batch of symbols through the tight loop ---- */
        if (fast_cfg && jitter >= 0 && sps_num == p.rate && sps_den == p.symrate && (!track || sum_window == tr.window)) {
            int nb = min(32 - slot_of(nsym), out_cap - nsym);
            nb = min(nb, (q_end - pos - reserve) / whole0 + 1);
            /* every sample of the batch must have landed in the ring */
            const int landed = (whi >= q_fill_end) ? whi : whi - kFill * kFillPending;
            nb = min(nb, (landed - pos) / whole0);
            int mi0 = 0;
            if (track) {
                const int idx = (midx < 0 || midx >= tr.window) ? 0 : midx;
                nb = min(nb, min(32 - (idx & 31), tr.window - idx));
                tr.seek(idx);
                mi0 = idx;
            }
            if (nb > 0) {
                const int g0 = slot_of(nsym);
                if (!track) { /* thresholds do not move: hand them to digitize as they are */
                    for (int k = lane; k < nb; k += 32) {
                        sh->o_min[g0 + k] = vmin, sh->o_max[g0 + k] = vmax, sh->o_center[g0 + k] = center;
                        sh->o_umid[g0 + k] = umid, sh->o_lmid[g0 + k] = lmid;
                    }
                    __syncwarp();
                }
                float lvmin = vmin, lvmax = vmax;
                const int pos0 = pos;
                int done;
                if (track) {
                    done = nw0 == 5 ? sym_fast_batch<5, true>(sh, tr, nb, whole0, lo0, pos, sidx, mi0, g0, vmin, vmax, lvmin, lvmax,
                                                              minbuf_sum, maxbuf_sum)
                                    : sym_fast_batch<4, true>(sh, tr, nb, whole0, lo0, pos, sidx, mi0, g0, vmin, vmax, lvmin, lvmax,
                                                              minbuf_sum, maxbuf_sum);
                } else {
                    done = nw0 == 5 ? sym_fast_batch<5, false>(sh, tr, nb, whole0, lo0, pos, sidx, mi0, g0, vmin, vmax, lvmin, lvmax,
                                                               minbuf_sum, maxbuf_sum)
                                    : sym_fast_batch<4, false>(sh, tr, nb, whole0, lo0, pos, sidx, mi0, g0, vmin, vmax, lvmin, lvmax,
                                                               minbuf_sum, maxbuf_sum);
                }
                if (done > 0) {
                    (void)pos0;
                    /* the symbol's last sample becomes lastsample, clipped with the thresholds that symbol saw */
                    const float s = at(pos - 1);
                    lastsample = s > lvmax ? lvmax : (s < lvmin ? lvmin : s);
                    sps = whole0;
                    center_idx = (sps - 1) / 2;
                    symbolcnt += done;
                    if (track) {
                        tr.dirty = 1;
                        midx = mi0 + done >= tr.window ? 0 : mi0 + done;
                        center = __fmul_rn(__fadd_rn(vmax, vmin), 0.5f); /* x / 2.0f, exact */
                        umid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmax, center), 5.0f), 0.125f), center); /* .. / 8.0f, exact */
                        lmid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmin, center), 5.0f), 0.125f), center);
                        maxref = __fmul_rn(vmax, 0.80f);
                        minref = __fmul_rn(vmin, 0.80f);
                    } else {
                        maxref = vmax;
                        minref = vmin;
                    }
                    nsym += done;
                    if (slot_of(nsym) == 0) {
                        flush_group();
                    }
                    continue;
                }
            }
        }
        /* ---- general path, one symbol: symbol_apply_rtl_fsk_discriminator_timing (dsd_symbol.c:1328-1387) ---- */
        if (sps_num != p.rate || sps_den != p.symrate) {
            sps_num = p.rate;
            sps_den = p.symrate;
            sps_accum = 0;
            jitter = -1;
            center = 0.0f, vmin = -30000.0f, vmax = 30000.0f, lmid = -20000.0f, umid = 20000.0f;
            minref = -24000.0f, maxref = 24000.0f;
            tr.fill_rings(vmin, vmax, kMinMax);
            midx = 0;
            sum_window = 0;
        }
        {
            int whole = whole0;
            if (rem0 > 0 && sps_den > 0) {
                int acc = sps_accum + rem0;
                if (acc >= sps_den) {
                    whole++;
                    acc -= sps_den;
                }
                sps_accum = acc;
                if (whole > 64) {
                    whole = 64;
                }
            }
            sps = whole;
            center_idx = (sps - 1) / 2;
        }
        /* ---- symbol_process_live_samples (dsd_symbol.c:1769-1792) ---- */
        float sum = 0.0f;
        int cnt = 0;
        for (int i = 0; i < sps; i++) {
            if (i == 0 && have_sync == 0 && jitter >= 0) { /* dsd_symbol.c:462-516 */
                if (sps == 20) {
                    if (jitter >= 7 && jitter <= 10) {
                        i--;
                    } else if (jitter >= 11 && jitter <= 14) {
                        i++;
                    }
                } else if (rf_mod == 2) { /* symbol_adjust_timing_gfsk */
                    if (jitter >= center_idx - 1 && jitter <= center_idx) {
                        i--;
                    } else if (jitter >= center_idx + 1 && jitter <= center_idx + 2) {
                        i++;
                    }
                } else {
                    if (jitter > 0 && jitter <= center_idx) {
                        i--;
                    } else if (jitter > center_idx && jitter < sps) {
                        i++;
                    }
                }
                jitter = -1;
            }
            float s = at(pos);
            pos++;
            if (have_sync == 1 && rf_mod == 0) { /* symbol_apply_sync_clip: C4FM only */
                s = s > vmax ? vmax : (s < vmin ? vmin : s);
            }
            if (s > center) { /* symbol_update_jitter, rf_mod == 0 branches */
                if (!(s > __fmul_rn(maxref, 1.25f)) && jitter < 0 && lastsample < center) {
                    jitter = i;
                }
            } else {
                if (!(s < __fmul_rn(minref, 1.25f)) && jitter < 0 && lastsample > center) {
                    jitter = i;
                }
            }
            if (sps == 20 && i >= 7 && i <= 13) { /* symbol_accumulate_sample */
                sum = __fadd_rn(sum, s);
                cnt++;
            }
            if (sps == 5 && i == 2) {
                sum = __fadd_rn(sum, s);
                cnt++;
            } else if (rf_mod == 0) { /* symbol_accumulate_c4fm_window */
                if (i >= center_idx - window_l && i <= center_idx + 2) {
                    sum = __fadd_rn(sum, s);
                    cnt++;
                }
            } else if (sps <= 4 ? (i == center_idx) : (i == center_idx - 1 || i == center_idx + 1)) {
                sum = __fadd_rn(sum, s); /* symbol_accumulate_other_window with select_window_gfsk: the two edge samples */
                cnt++;
            }
            lastsample = s;
        }
        const float sym = cnt > 0 ? __fdiv_rn(sum, (float)cnt) : 0.0f;
        symbolcnt++;
        if (soft) {
            /* ---- get_dibit_and_analog_signal: sbuf, use_symbol (dsd_dibit.c:243-299) ---- */
            if (track) {
                tr.push(sym, sidx, midx, sum_window, minbuf_sum, maxbuf_sum, vmin, vmax);
                center = __fmul_rn(__fadd_rn(vmax, vmin), 0.5f); /* x / 2.0f, exact */
                umid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmax, center), 5.0f), 0.125f), center); /* .. / 8.0f, exact */
                lmid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmin, center), 5.0f), 0.125f), center);
                maxref = __fmul_rn(vmax, 0.80f);
                minref = __fmul_rn(vmin, 0.80f);
            } else {
                sh->sbuf[sidx & (kSbuf - 1)] = sym;
                maxref = vmax;
                minref = vmin;
            }
            if (tr.cap > 0) {
                sidx = (sidx >= tr.cap - 1) ? 0 : sidx + 1;
            }
        }
        {
            const int g = slot_of(nsym);
            sh->o_sym[g] = sym, sh->o_min[g] = vmin, sh->o_max[g] = vmax;
            sh->o_center[g] = center, sh->o_umid[g] = umid, sh->o_lmid[g] = lmid;
        }
        nsym++;
        if (slot_of(nsym) == 0) {
            flush_group();
        }
    }
    flush_group();

    /* leftover samples -> carry (right-aligned), through registers: the ring may be refilled first */
    refill();
    int left = q_end - pos;
    left = left < 0 ? 0 : (left > kCarry ? kCarry : left);
    float keep[kCarry / 32];
#pragma unroll
    for (int j = 0; j < kCarry / 32; j++) {
        const int k = j * 32 + lane;
        keep[j] = k < left ? at(q_end - left + k) : 0.0f;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kCarry / 32; j++) {
        const int k = j * 32 + lane;
        if (k < left) {
            carry[kCarry - left + k] = keep[j];
        }
    }
    tr.flush_chunk();
    tr.store_sbuf(p.sbuf + (size_t)c * kSbuf);
    if (lane == 0) {
        p.s.sps_num[c] = sps_num, p.s.sps_den[c] = sps_den, p.s.sps_accum[c] = sps_accum;
        p.s.sps[c] = sps, p.s.center_idx[c] = center_idx, p.s.jitter[c] = jitter;
        p.s.lastsample[c] = lastsample;
        p.s.vmin[c] = vmin, p.s.vmax[c] = vmax, p.s.center[c] = center, p.s.umid[c] = umid, p.s.lmid[c] = lmid;
        p.s.minref[c] = minref, p.s.maxref[c] = maxref;
        p.s.sidx[c] = sidx, p.s.midx[c] = midx, p.s.sum_window[c] = sum_window;
        p.s.minbuf_sum[c] = minbuf_sum, p.s.maxbuf_sum[c] = maxbuf_sum;
        p.s.carry_n[c] = left;
        p.s.symbolcnt[c] = symbolcnt;
        p.count[c] = nsym;
    }
}

/* ------------------------------------------------------------------ acquisition: getFrameSync() on the device */

constexpr int kAcqMaxPatterns = 8;
constexpr int kAcqMargin = 208; /* hunting stops this many samples before the end of a launch: an acquisition always finishes
                                   its matched-filter start-up (taps - 1 samples + one symbol) inside the launch; <= kCarry */

struct AcqPattern {
    unsigned bits;  /* last 24 hunt-sliced symbols, '1' -> 1, '3' -> 0, oldest in bit 23 */
    int sync_type, kind;
    int filter, window_l, track, negative;
};

struct AcqParams {
    SymParams sp;          /* sp.filt = the RAW discriminator samples of this launch (no matched filter before the first sync) */
    const float* taps;     /* [n_filters][kMaxTaps] */
    const int* taps_len;
    AcqPattern pat[kAcqMaxPatterns];
    int n_pat;
    int* acquired;         /* [n_ch] carried: 0 hunting, 1 synchronised */
    int* hunt_since;       /* [n_ch] carried: symbols since the hunt context (re)started (getFrameSync gives up after 1800) */
    unsigned* hunt_bits;   /* [n_ch] carried: hunt-sliced symbols, newest in bit 0 */
    int* hunt_count;       /* [n_ch] carried: symbols in the window (<= 24) */
    float* lbuf;           /* [n_ch][24] carried level ring */
    int* lidx;             /* [n_ch] */
    int* level_count;      /* [n_ch] */
    float* hist;           /* [n_ch][128] carried symbol history ring (warm start, resample-on-sync) */
    int* hist_head;        /* [n_ch] total symbols pushed */
    int* start_off;        /* [n_ch] out: samples of this launch consumed here */
    int* out_off;          /* [n_ch] out: outputs written here */
    dsdneo_b200_acq_info* info; /* [n_ch] out */
    float* fir_hist;       /* [n_ch][kMaxTaps] the matched filter's carried raw history (sps_fir_kernel) */
    int filtered_input;    /* the samples are already matched-filtered (a hunt after the first sync: lastsynctype known): no
                            * filter start-up at the sync, the filter history is not touched */
};

/*
 * getFrameSync() from the never-synchronised state (src/dsp/dsd_frame_sync.c:3098-3148), one warp per hunting channel,
 * warp-uniform like symbolize_kernel: getSymbol(have_sync = 0) with its timing nudges on the RAW samples; the level ring,
 * sbuf and rolling payload dibit / reliability bookkeeping; hunt-time slice and 24-symbol pattern compare; on a match the
 * basic lock, the sync warm start of the slicer thresholds (src/dsp/sync_calibration.c:155-226), for DMR the re-slicing of
 * the 66 dibits in front of the sync (src/dsp/dmr_sync.c:60-126), and the switch to the decoder class of that sync type.
 * The reference turns the matched filter on at that moment with an all-zero delay line; the first taps - 1 samples after the
 * sync are therefore filtered here, directly and in the reference's accumulation order, and the first few synchronised
 * symbols (getDibitSoft) are produced here too; symbolize_kernel takes over where the steady-state filter output is valid.
 */
__global__ void __launch_bounds__(kSymWarps * 32)
sym_acquire_kernel(const AcqParams a) {
    __shared__ WarpShared s_warp[kSymWarps];
    __shared__ float s_hist[kSymWarps][128], s_lbuf[kSymWarps][24], s_sorted[kSymWarps][24];
    const SymParams& p = a.sp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * kSymWarps + warp;
    if (c >= p.n_ch) {
        return;
    }
    if (a.acquired[c]) {
        if (lane == 0) {
            a.start_off[c] = 0, a.out_off[c] = 0;
            a.info[c].acquired = 1, a.info[c].sync_type = -1, a.info[c].hit_index = -1, a.info[c].hunt_symbols = 0;
            a.info[c].warm_start = 0, a.info[c].resample_ok = 0;
        }
        return;
    }
    WarpShared* sh = &s_warp[warp];
    float* ring = sh->ring;
    float* hist = s_hist[warp];
    float* lbuf = s_lbuf[warp];
    float* sorted = s_sorted[warp];
    const unsigned full = 0xffffffffu;

    int window_l = p.s.window_l[c], track = p.s.track[c], negative = p.s.negative[c];
    const int rf_mod = p.s.rf_mod[c], snr_num = p.snr_num[c];
    int sps_num = p.s.sps_num[c], sps_den = p.s.sps_den[c], sps_accum = p.s.sps_accum[c];
    int sps = p.s.sps[c], center_idx = p.s.center_idx[c], jitter = p.s.jitter[c];
    float lastsample = p.s.lastsample[c];
    float vmin = p.s.vmin[c], vmax = p.s.vmax[c], center = p.s.center[c], umid = p.s.umid[c], lmid = p.s.lmid[c];
    float minref = p.s.minref[c], maxref = p.s.maxref[c];
    int sidx = p.s.sidx[c], midx = p.s.midx[c], sum_window = p.s.sum_window[c];
    double minbuf_sum = p.s.minbuf_sum[c], maxbuf_sum = p.s.maxbuf_sum[c];
    int carry_n = p.s.carry_n[c];
    carry_n = carry_n < 0 ? 0 : (carry_n > kCarry ? kCarry : carry_n);
    long long symbolcnt = p.s.symbolcnt[c];
    int hunt_since = a.hunt_since[c], hunt_count = a.hunt_count[c], lidx = a.lidx[c], level_count = a.level_count[c];
    unsigned hunt_bits = a.hunt_bits[c];
    int hist_head = a.hist_head[c];
    for (int k = lane; k < 128; k += 32) {
        hist[k] = a.hist[(size_t)c * 128 + k];
    }
    if (lane < 24) {
        lbuf[lane] = a.lbuf[(size_t)c * 24 + lane];
    }
    WarpTracker tr;
    tr.init(lane, p.ssize, p.msize, p.minbuf + (size_t)c * kMinMax, p.maxbuf + (size_t)c * kMinMax, p.sbuf + (size_t)c * kSbuf, sh, false);

    const float* raw = p.filt + (size_t)c * p.filt_pitch;
    float* carry = p.carry + (size_t)c * kCarry;
    const int q0 = kCarry - carry_n;
    const int q_end = kCarry + p.n;
    const int q_fill_end = (q_end + 31) & ~31;
    int pos = q0;
    int whi = q0 & ~31;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    auto refill = [&]() {
        while (whi < pos + kLook && whi < q_fill_end) {
            const int q = whi + lane;
            const float* src = raw;
            unsigned bytes = 0;
            if (q < kCarry) {
                if (q >= q0) {
                    src = carry + q, bytes = 4;
                }
            } else if (q < q_end) {
                src = raw + (q - kCarry), bytes = 4;
            }
            const int slot = q & (kRing - 1);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ring_s + 4u * (unsigned)slot), "l"(src), "r"(bytes) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            whi += 32;
        }
        if (whi >= q_fill_end) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
        }
        __syncwarp();
    };
    auto at = [&](int q) -> float { return ring[q & (kRing - 1)]; };

    int whole0 = p.rate / p.symrate, rem0 = p.rate % p.symrate;
    if (whole0 < 2) {
        whole0 = 2, rem0 = 0;
    }
    if (whole0 > 64) {
        whole0 = 64, rem0 = 0;
    }
    const int reserve = whole0 + 2;
    float* o_sym = p.symbols + (size_t)c * p.out_pitch;
    uint8_t* o_dib = p.dibits ? p.dibits + (size_t)c * p.out_pitch : nullptr;
    uint8_t* o_rel = p.reliab ? p.reliab + (size_t)c * p.out_pitch : nullptr;
    short2* o_llr = p.llr ? reinterpret_cast<short2*>(p.llr) + (size_t)c * p.out_pitch : nullptr;
    const int out_cap = (int)(p.out_pitch > 0x7fffffffu ? 0x7fffffffu : p.out_pitch);
    int nsym = 0;

    /* matched filter of the acquired class, evaluated directly from the raw samples with a delay line that was all zero when
     * the sync was accepted (apply_sps_fir, dsd_filters.c:172-201: taps oldest -> newest, multiply then add) */
    int acq_pos = -1, filt_id = -1, filt_len = 0;
    auto filtered = [&](int q) -> float {
        if (filt_id < 0) {
            return at(q);
        }
        const float* t = a.taps + (size_t)filt_id * kMaxTaps;
        const int d = q - acq_pos; /* samples fed to the filter before this one */
        float acc = 0.0f;
        for (int i = max(0, filt_len - 1 - d); i < filt_len; i++) {
            acc = __fadd_rn(acc, __fmul_rn(t[i], at(q - (filt_len - 1) + i)));
        }
        return acc;
    };

    /* one getSymbol(): dsd_symbol.c:1328-1387 (timing), :1769-1796 (samples) */
    auto take_symbol = [&](int have_sync) -> float {
        if (sps_num != p.rate || sps_den != p.symrate) {
            sps_num = p.rate, sps_den = p.symrate, sps_accum = 0, jitter = -1;
            center = 0.0f, vmin = -30000.0f, vmax = 30000.0f, lmid = -20000.0f, umid = 20000.0f;
            minref = -24000.0f, maxref = 24000.0f;
            tr.fill_rings(vmin, vmax, kMinMax);
            midx = 0, sum_window = 0;
        }
        int whole = whole0;
        if (rem0 > 0 && sps_den > 0) {
            int acc = sps_accum + rem0;
            if (acc >= sps_den) {
                whole++;
                acc -= sps_den;
            }
            sps_accum = acc;
            whole = whole > 64 ? 64 : whole;
        }
        sps = whole;
        center_idx = (sps - 1) / 2;
        const int l_edge = rf_mod == 2 ? 1 : window_l;
        float sum = 0.0f;
        int cnt = 0;
        for (int i = 0; i < sps; i++) {
            if (i == 0 && have_sync == 0 && jitter >= 0) {
                if (sps == 20) {
                    if (jitter >= 7 && jitter <= 10) {
                        i--;
                    } else if (jitter >= 11 && jitter <= 14) {
                        i++;
                    }
                } else if (rf_mod == 2) {
                    if (jitter >= center_idx - 1 && jitter <= center_idx) {
                        i--;
                    } else if (jitter >= center_idx + 1 && jitter <= center_idx + 2) {
                        i++;
                    }
                } else {
                    if (jitter > 0 && jitter <= center_idx) {
                        i--;
                    } else if (jitter > center_idx && jitter < sps) {
                        i++;
                    }
                }
                jitter = -1;
            }
            float sv = filtered(pos);
            pos++;
            if (have_sync == 1 && rf_mod == 0) {
                sv = sv > vmax ? vmax : (sv < vmin ? vmin : sv);
            }
            if (sv > center) {
                if (!(sv > __fmul_rn(maxref, 1.25f)) && jitter < 0 && lastsample < center) {
                    jitter = i;
                }
            } else {
                if (!(sv < __fmul_rn(minref, 1.25f)) && jitter < 0 && lastsample > center) {
                    jitter = i;
                }
            }
            if (sps == 20 && i >= 7 && i <= 13) {
                sum = __fadd_rn(sum, sv);
                cnt++;
            }
            if (sps == 5 && i == 2) {
                sum = __fadd_rn(sum, sv);
                cnt++;
            } else if (rf_mod == 0) {
                if (i >= center_idx - l_edge && i <= center_idx + 2) {
                    sum = __fadd_rn(sum, sv);
                    cnt++;
                }
            } else if (sps <= 4 ? (i == center_idx) : (i == center_idx - 1 || i == center_idx + 1)) {
                sum = __fadd_rn(sum, sv);
                cnt++;
            }
            lastsample = sv;
        }
        symbolcnt++;
        return cnt > 0 ? __fdiv_rn(sum, (float)cnt) : 0.0f;
    };
    auto payload_dibit = [&](float v) -> int { return v > center ? (v > umid ? 1 : 0) : (v < lmid ? 3 : 2); };

    int acquired = 0, sync_type = -1, hit_index = -1, warm = 0, resample_ok = 0;
    refill();
    /* ---------------- hunt ---------------- */
    while (!acquired && nsym < out_cap && (q_end - pos) >= kAcqMargin) {
        refill();
        const float symbol = take_symbol(0);
        hist[hist_head & 127] = symbol; /* dsd_symbol_history_push */
        hist_head++;
        /* frame_sync_update_symbol_ring (dsd_frame_sync.c:1747-1764) */
        lbuf[lidx] = symbol;
        level_count = level_count < 24 ? level_count + 1 : 24;
        sh->sbuf[sidx & (kSbuf - 1)] = symbol;
        lidx = lidx == 23 ? 0 : lidx + 1;
        sidx = (sidx == p.ssize - 1) ? 0 : sidx + 1;
        /* rolling payload dibit + reliability (frame_sync_store_dmr_payload_symbol, :2161-2190) */
        const int d = payload_dibit(symbol);
        const int rel = c4fm_reliability(symbol, vmin, vmax, center, umid, lmid, snr_num);
        if (lane == 0) {
            o_sym[nsym] = symbol;
            if (o_dib) {
                o_dib[nsym] = (uint8_t)d;
                o_rel[nsym] = (uint8_t)rel;
                o_llr[nsym] = make_short2((short)(((d >> 1) & 1) ? rel : -rel), (short)((d & 1) ? rel : -rel));
            }
        }
        nsym++;
        hunt_bits = (hunt_bits << 1) | (symbol > 0.0f ? 1u : 0u);
        hunt_count = hunt_count < 24 ? hunt_count + 1 : 24;
        hunt_since++;
        __syncwarp();
        if (hunt_since >= 8) { /* frame_sync_eval_window (:2638-2676) */
            /* sorted copy of the level ring: every lane ranks one entry (equal values are interchangeable) */
            if (lane < level_count) {
                const float v = lbuf[lane];
                int rank = 0;
                for (int j = 0; j < level_count; j++) {
                    const float w = lbuf[j];
                    rank += (w < v || (w == v && j < lane)) ? 1 : 0;
                }
                sorted[rank] = v;
            }
            __syncwarp();
            float lmin, lmax;
            { /* dsd_frame_sync_estimate_sorted_window_levels (src/dsp/frame_sync_level.c) */
                const int n = level_count;
                if (n < 3) {
                    float sum = 0.0f;
                    for (int i = 0; i < n; i++) {
                        sum = __fadd_rn(sum, sorted[i]);
                    }
                    lmin = lmax = __fdiv_rn(sum, (float)n);
                } else {
                    int min_idx = 0, max_idx = n - 3;
                    if (n >= 13) {
                        min_idx = 2, max_idx = n - 5;
                    }
                    if (max_idx + 2 >= n) {
                        max_idx = n - 3;
                    }
                    lmin = __fdiv_rn(__fadd_rn(__fadd_rn(sorted[min_idx], sorted[min_idx + 1]), sorted[min_idx + 2]), 3.0f);
                    lmax = __fdiv_rn(__fadd_rn(__fadd_rn(sorted[max_idx], sorted[max_idx + 1]), sorted[max_idx + 2]), 3.0f);
                }
            }
            maxref = vmax, minref = vmin; /* frame_sync_window_levels, FSK profiles (:2332-2335) */
            int hit = -1;
            if (hunt_count >= 24) {
                for (int k = 0; k < a.n_pat && hit < 0; k++) {
                    hit = ((hunt_bits & 0xFFFFFFu) == a.pat[k].bits) ? k : -1;
                }
            }
            if (hit >= 0) {
                const AcqPattern& pt = a.pat[hit];
                vmax = __fmul_rn(__fadd_rn(vmax, lmax), 0.5f); /* frame_sync_set_basic_lock: (x + y) / 2 */
                vmin = __fmul_rn(__fadd_rn(vmin, lmin), 0.5f);
                if (pt.kind == 1 || rf_mod == 0) { /* dsd_sync_warm_start_thresholds_outer_only(24) */
                    float sum_pos = 0.0f, sum_neg = 0.0f;
                    int n_pos = 0, n_neg = 0;
                    for (int i = 0; i < 24; i++) { /* newest first, the reference's order */
                        const float v = hist[(hist_head - 1 - i) & 127];
                        if (v > 0.0f) {
                            sum_pos = __fadd_rn(sum_pos, v);
                            n_pos++;
                        } else {
                            sum_neg = __fadd_rn(sum_neg, v);
                            n_neg++;
                        }
                    }
                    if (n_pos > 0 && n_neg > 0) {
                        const float mp = __fdiv_rn(sum_pos, (float)n_pos), mn = __fdiv_rn(sum_neg, (float)n_neg);
                        if (!(fabsf(__fsub_rn(mp, mn)) < 1.0f)) {
                            vmax = mp, vmin = mn;
                            center = __fmul_rn(__fadd_rn(vmax, vmin), 0.5f);
                            umid = __fadd_rn(center, __fmul_rn(__fsub_rn(vmax, center), 0.625f));
                            lmid = __fadd_rn(center, __fmul_rn(__fsub_rn(vmin, center), 0.625f));
                            maxref = __fmul_rn(vmax, 0.80f);
                            minref = __fmul_rn(vmin, 0.80f);
                            tr.fill_rings(vmin, vmax, p.msize > kMinMax ? kMinMax : p.msize);
                            sum_window = 0;
                            warm = 1;
                        }
                    }
                }
                if (pt.kind == 1 && hist_head >= 90) { /* dmr_resample_cach: symbols [-90 .. -25] of the stream */
                    for (int i = lane; i < 66; i += 32) {
                        const int back = 89 - i; /* 0 = newest */
                        const uint8_t d2 = (uint8_t)payload_dibit(hist[(hist_head - 1 - back) & 127]);
                        a.info[c].resampled[i] = d2;
                        const int at_out = nsym - 1 - back; /* the part that lies in this launch is rewritten in place */
                        if (at_out >= 0 && o_dib) {
                            o_dib[at_out] = d2;
                        }
                    }
                    resample_ok = 1;
                }
                window_l = pt.window_l, track = pt.track, negative = pt.negative;
                filt_id = pt.filter;
                filt_len = filt_id >= 0 ? a.taps_len[filt_id] : 0;
                if (filt_len <= 0) {
                    filt_id = -1;
                }
                acq_pos = pos;
                acquired = 1, sync_type = pt.sync_type, hit_index = nsym - 1;
            }
        }
        if (!acquired && hunt_since >= 1800) { /* frame_sync_handle_no_sync_timeout: the next call starts with an empty window */
            hunt_since = 0, hunt_count = 0, lidx = 0, level_count = 0, hunt_bits = 0;
        }
    }
    /* ---------------- the first synchronised symbols, through the starting-up matched filter ---------------- */
    if (acquired) {
        if (track && tr.cap >= 2) {
            tr.rescan();
        }
        while (nsym < out_cap && ((pos - acq_pos) < filt_len - 1 || pos < kCarry) && (q_end - pos) >= reserve) {
            refill();
            const float sym = take_symbol(1);
            if (track) {
                tr.push(sym, sidx, midx, sum_window, minbuf_sum, maxbuf_sum, vmin, vmax);
                cq_thresholds(vmin, vmax, center, umid, lmid);
                maxref = __fmul_rn(vmax, 0.80f);
                minref = __fmul_rn(vmin, 0.80f);
            } else {
                sh->sbuf[sidx & (kSbuf - 1)] = sym;
                maxref = vmax;
                minref = vmin;
            }
            if (tr.cap > 0) {
                sidx = (sidx >= tr.cap - 1) ? 0 : sidx + 1;
            }
            int dibit, rel, l0, l1;
            digitize_one(sym, vmin, vmax, center, umid, lmid, negative, snr_num, dibit, rel, l0, l1);
            if (lane == 0) {
                o_sym[nsym] = sym;
                if (o_dib) {
                    o_dib[nsym] = (uint8_t)dibit;
                    o_rel[nsym] = (uint8_t)rel;
                    o_llr[nsym] = make_short2((short)l0, (short)l1);
                }
            }
            nsym++;
        }
    }
    /* ---------------- hand over ---------------- */
    refill();
    int left = 0;
    if (!acquired) { /* still hunting: the unconsumed raw tail is carried */
        left = q_end - pos;
        left = left < 0 ? 0 : (left > kCarry ? kCarry : left);
        float keep[kCarry / 32];
#pragma unroll
        for (int j = 0; j < kCarry / 32; j++) {
            const int k = j * 32 + lane;
            keep[j] = k < left ? at(q_end - left + k) : 0.0f;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kCarry / 32; j++) {
            const int k = j * 32 + lane;
            if (k < left) {
                carry[kCarry - left + k] = keep[j];
            }
        }
    }
    if (acquired && filt_id >= 0 && !a.filtered_input) { /* sps_fir_kernel runs next over this launch: x[-1], x[-2] .. are the carried raw samples */
        const int hl = filt_len - 1;
        for (int k = lane; k < hl; k += 32) {
            const int q = kCarry - hl + k;
            a.fir_hist[(size_t)c * kMaxTaps + k] = q >= q0 ? carry[q] : 0.0f;
        }
    }
    tr.flush_chunk();
    tr.store_sbuf(p.sbuf + (size_t)c * kSbuf);
    __syncwarp();
    for (int k = lane; k < 128; k += 32) {
        a.hist[(size_t)c * 128 + k] = hist[k];
    }
    if (lane < 24) {
        a.lbuf[(size_t)c * 24 + lane] = lbuf[lane];
    }
    if (lane == 0) {
        p.s.sps_num[c] = sps_num, p.s.sps_den[c] = sps_den, p.s.sps_accum[c] = sps_accum;
        p.s.sps[c] = sps, p.s.center_idx[c] = center_idx, p.s.jitter[c] = jitter;
        p.s.lastsample[c] = lastsample;
        p.s.vmin[c] = vmin, p.s.vmax[c] = vmax, p.s.center[c] = center, p.s.umid[c] = umid, p.s.lmid[c] = lmid;
        p.s.minref[c] = minref, p.s.maxref[c] = maxref;
        p.s.sidx[c] = sidx, p.s.midx[c] = midx, p.s.sum_window[c] = sum_window;
        p.s.minbuf_sum[c] = minbuf_sum, p.s.maxbuf_sum[c] = maxbuf_sum;
        p.s.symbolcnt[c] = symbolcnt;
        a.hunt_since[c] = hunt_since, a.hunt_count[c] = hunt_count, a.hunt_bits[c] = hunt_bits;
        a.lidx[c] = lidx, a.level_count[c] = level_count, a.hist_head[c] = hist_head;
        a.acquired[c] = acquired;
        a.info[c].acquired = acquired, a.info[c].sync_type = sync_type, a.info[c].hit_index = hit_index;
        a.info[c].hunt_symbols = acquired ? hit_index + 1 : nsym;
        a.info[c].warm_start = (uint8_t)warm, a.info[c].resample_ok = (uint8_t)resample_ok;
        a.out_off[c] = nsym;
        if (acquired) {
            p.s.window_l[c] = window_l, p.s.track[c] = track, p.s.negative[c] = negative;
            p.s.filter[c] = filt_id;
            p.s.carry_n[c] = 0;              /* symbolize_kernel continues inside this launch's filtered samples */
            a.start_off[c] = pos - kCarry;   /* >= 0: the loop above ran until the carried raw samples were used up */
        } else {
            p.s.carry_n[c] = left;
            a.start_off[c] = p.n;
            p.count[c] = nsym;
        }
    }
}

/* exhaustive check of div5_rn against the IEEE operator: every one of the 2^32 operands */
__global__ void
div5_selftest_kernel(unsigned long long* n_mismatch, unsigned* first_bad) {
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long it = 0; it < (1ull << 32) / stride; it++, i += stride) {
        const float x = __uint_as_float(i);
        const float a = div5_rn(x), b = __fdiv_rn(x, 5.0f);
        const bool same = (__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b);
        if (!same) {
            bad++;
            atomicMin(first_bad, i);
        }
    }
    if (bad) {
        atomicAdd(n_mismatch, bad);
    }
}

__global__ void
sym_reset_kernel(SymScalars s, float* minbuf, float* maxbuf, float* sbuf, float* carry, float* hist, int n_ch) {
    const int c = blockIdx.x;
    const int t = threadIdx.x;
    if (c >= n_ch) {
        return;
    }
    if (t == 0) {
        /* initState() values (src/core/util/dsd_init.c:519-592) */
        s.sps_num[c] = 0, s.sps_den[c] = 0, s.sps_accum[c] = 0;
        s.sps[c] = 10, s.center_idx[c] = 4, s.jitter[c] = -1;
        s.lastsample[c] = 0.0f;
        s.vmin[c] = -15000.0f, s.vmax[c] = 15000.0f, s.center[c] = 0.0f, s.umid[c] = 0.0f, s.lmid[c] = 0.0f;
        s.minref[c] = -12000.0f, s.maxref[c] = 12000.0f;
        s.sidx[c] = 0, s.midx[c] = 0, s.sum_window[c] = 0;
        s.minbuf_sum[c] = 0.0, s.maxbuf_sum[c] = 0.0;
        s.carry_n[c] = 0;
        s.symbolcnt[c] = 0;
    }
    for (int k = t; k < kMinMax; k += blockDim.x) {
        minbuf[(size_t)c * kMinMax + k] = -15000.0f;
        maxbuf[(size_t)c * kMinMax + k] = 15000.0f;
    }
    for (int k = t; k < kSbuf; k += blockDim.x) {
        sbuf[(size_t)c * kSbuf + k] = 0.0f;
    }
    for (int k = t; k < kCarry; k += blockDim.x) {
        carry[(size_t)c * kCarry + k] = 0.0f;
    }
    for (int k = t; k < kMaxTaps; k += blockDim.x) {
        hist[(size_t)c * kMaxTaps + k] = 0.0f;
    }
}

/* ---- symbol-rate CQPSK input (output kind 2): the sample side behind cqpsk_chain_kernel ---------------------------------
 *
 * Reference, per symbol (one stream float = one symbol):
 *   symbol_try_rtl_symbol_rate_fast_path   src/dsp/dsd_symbol.c:1583-1625 (fixed thresholds :744-765, then take)
 *   use_symbol, rf_mod == 1                src/core/frames/dsd_dibit.c:243-299: the min / max tracker always runs and
 *                                           overwrites the fixed thresholds before digitize sees them
 *   digitize                               :1018-1041: cqpsk_slice(symbol - center) (:329-349) through the OP25 dibit map
 *                                           (core/p25_cqpsk_dibit.h) when the CQPSK chain is active on a P25 sync, else regions
 *   compute_dibit_soft_metric              :685-721 with build_cqpsk_dibit_ideals (:660-683) / standard ideals,
 *                                           reliability = cqpsk_reliability_raw (:376-401) x CQPSK SNR weight (:404-427)
 * The symbol value does not feed back into anything but the tracker, so the per-channel chain is the extrema of the
 * 128-entry symbol window and the two f64 running sums (one warp per channel, see cqpsk_slice_kernel). */
struct CqSliceParams {
    const float* symbols; /* [n_ch][sym_pitch] */
    size_t sym_pitch;
    const int* n_symbols; /* [n_ch] */
    uint8_t* dibits;      /* [n_ch][out_pitch] */
    uint8_t* reliab;
    int16_t* llr;         /* [n_ch][out_pitch][2] */
    size_t out_pitch;
    /* per-channel class */
    const uint8_t* negative;
    const uint8_t* p25_slice;
    const uint8_t* map_idx;
    /* carried state */
    float* sbuf;          /* [n_ch][128] */
    float* minbuf;        /* [n_ch][1024] */
    float* maxbuf;
    int* sidx;
    int* midx;
    int* sum_window;
    double* minbuf_sum;
    double* maxbuf_sum;
    float* thr;           /* [8][n_ch]: min, max, center, umid, lmid, minref, maxref, lastsample */
    int n_ch, ssize, msize;
    int snr_scale_num;    /* 204 + (w256 >> 2), or 0 when the SNR hook reports <= -50 dB (no weighting) */
    float2* minmax;       /* [n_ch][out_pitch] scratch: {min, max} after use_symbol, per symbol (tracker -> digitize kernel) */
};

/*
 * One WARP per channel (round 1: one lane per channel, 2.9 ms per 4800 symbols): the tracker of use_symbol is the only
 * loop-carried state, and WarpTracker runs it the way symbolize_kernel does -- every lane holds the same scalars, the lanes
 * share out the 128-entry window rescan (needed only when the symbol that leaves was one of the four extremes) and the
 * 32-entry chunks of the 1024-entry f64-averaged rings.  Symbols are read 32 at a time (one per lane, coalesced) and handed
 * round by shuffle; {min, max} after every symbol are collected one per lane and stored coalesced for the digitize kernel.
 */
constexpr int kCqWarps = 4;

__global__ void __launch_bounds__(kCqWarps * 32)
cqpsk_slice_kernel(const CqSliceParams p) {
    __shared__ WarpShared s_warp[kCqWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * kCqWarps + warp;
    if (ch >= p.n_ch) {
        return;
    }
    WarpShared* sh = &s_warp[warp];
    const int n_ch = p.n_ch;
    const int n = p.n_symbols[ch];
    int sidx = p.sidx[ch], midx = p.midx[ch], sum_window = p.sum_window[ch];
    double min_sum = p.minbuf_sum[ch], max_sum = p.maxbuf_sum[ch];
    float vmin = p.thr[0 * (size_t)n_ch + ch], vmax = p.thr[1 * (size_t)n_ch + ch];
    float center = p.thr[2 * (size_t)n_ch + ch], umid = p.thr[3 * (size_t)n_ch + ch], lmid = p.thr[4 * (size_t)n_ch + ch];
    float minref = p.thr[5 * (size_t)n_ch + ch], maxref = p.thr[6 * (size_t)n_ch + ch], last = p.thr[7 * (size_t)n_ch + ch];
    WarpTracker tr;
    tr.init(lane, p.ssize, p.msize, p.minbuf + (size_t)ch * kMinMax, p.maxbuf + (size_t)ch * kMinMax, p.sbuf + (size_t)ch * kSbuf, sh, true);
    const float* in = p.symbols + (size_t)ch * p.sym_pitch;
    float2* mm = p.minmax + (size_t)ch * p.out_pitch;
    float nxt = (lane < n) ? in[lane] : 0.0f;
    for (int base = 0; base < n; base += 32) {
        const float cur = nxt;
        if (base + 32 + lane < n) {
            nxt = in[base + 32 + lane]; /* the next group's symbols are requested while this group runs */
        }
        const int m = min(32, n - base);
        float my_min = 0.0f, my_max = 0.0f;
#pragma unroll 1
        for (int k = 0; k < m; k++) {
            const float sym = __shfl_sync(0xffffffffu, cur, k);
            last = sym;
            /* sbuf[sidx] = symbol (sidx stays 0 when ssize <= 0), then use_symbol's tracker: dsd_dibit.c:243-299 */
            tr.push(sym, sidx, midx, sum_window, min_sum, max_sum, vmin, vmax);
            if (tr.cap > 0) {
                sidx = (sidx >= tr.cap - 1) ? 0 : sidx + 1;
            }
            if (lane == k) {
                my_min = vmin, my_max = vmax;
            }
        }
        if (lane < m) {
            mm[base + lane] = make_float2(my_min, my_max);
        }
    }
    if (n > 0) {
        cq_thresholds(vmin, vmax, center, umid, lmid);
        maxref = __fmul_rn(vmax, 0.80f);
        minref = __fmul_rn(vmin, 0.80f);
    }
    tr.flush_chunk();
    tr.store_sbuf(p.sbuf + (size_t)ch * kSbuf);
    if (lane == 0) {
        p.sidx[ch] = sidx;
        p.midx[ch] = midx;
        p.sum_window[ch] = sum_window;
        p.minbuf_sum[ch] = min_sum;
        p.maxbuf_sum[ch] = max_sum;
        p.thr[0 * (size_t)n_ch + ch] = vmin;
        p.thr[1 * (size_t)n_ch + ch] = vmax;
        p.thr[2 * (size_t)n_ch + ch] = center;
        p.thr[3 * (size_t)n_ch + ch] = umid;
        p.thr[4 * (size_t)n_ch + ch] = lmid;
        p.thr[5 * (size_t)n_ch + ch] = minref;
        p.thr[6 * (size_t)n_ch + ch] = maxref;
        p.thr[7 * (size_t)n_ch + ch] = last;
    }
}

/* digitize + compute_dibit_soft_metric for every symbol of every channel, one thread per symbol (time-parallel: the only
 * state they read, {min, max} after use_symbol, was written per symbol by cqpsk_slice_kernel) */
__global__ void __launch_bounds__(256)
cqpsk_digitize_kernel(const CqSliceParams p) {
    const int ch = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_symbols[ch]) {
        return;
    }
    const float sym = p.symbols[(size_t)ch * p.sym_pitch + i];
    const float2 mmv = p.minmax[(size_t)ch * p.out_pitch + i];
    const float vmin = mmv.x, vmax = mmv.y;
    const int negative = p.negative[ch], p25 = p.p25_slice[ch];
    const int map_idx = p.map_idx[ch] < 5 ? p.map_idx[ch] : 0;
    /* OP25 orientation maps (include/dsd-neo/core/p25_cqpsk_dibit.h:28-52), 2 bits per entry, and their inverses */
    const unsigned fmap = (map_idx == 0) ? 0xE4u : (map_idx == 1) ? 0x4Eu : (map_idx == 2) ? 0x1Bu : (map_idx == 3) ? 0x8Du : 0x72u;
    unsigned inv = 0;
#pragma unroll
    for (int raw = 3; raw >= 0; raw--) { /* lowest raw dibit wins, like dsd_p25_cqpsk_raw_dibit_for_corrected */
        const unsigned c = (fmap >> (2 * raw)) & 3u;
        inv = (inv & ~(3u << (2 * c))) | ((unsigned)raw << (2 * c));
    }
    float center, umid, lmid;
    cq_thresholds(vmin, vmax, center, umid, lmid);
    int dibit;
    float ideal[4];
    const float sc = __fsub_rn(sym, center);
    if (p25) {
        const int raw = sc >= 2.0f ? 1 : (sc >= 0.0f ? 0 : (sc >= -2.0f ? 2 : 3));
        dibit = (int)((fmap >> (2 * raw)) & 3u);
        if (negative) {
            dibit = (dibit + 2) & 3;
        }
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const int corrected = negative ? ((d + 2) & 3) : d;
            const int mapped = (int)((inv >> (2 * corrected)) & 3u);
            /* base levels {+1, +3, -1, -3} for raw dibits 0..3 */
            const float level = (mapped == 0) ? 1.0f : ((mapped == 1) ? 3.0f : ((mapped == 2) ? -1.0f : -3.0f));
            ideal[d] = __fadd_rn(center, level);
        }
    } else {
        if (sym > center) {
            dibit = sym > umid ? (negative ? 3 : 1) : (negative ? 2 : 0);
        } else {
            dibit = sym < lmid ? (negative ? 1 : 3) : (negative ? 0 : 2);
        }
        const float plus_one = __fmul_rn(0.5f, __fadd_rn(center, umid)), minus_one = __fmul_rn(0.5f, __fadd_rn(lmid, center));
        if (negative) {
            ideal[0] = minus_one, ideal[1] = vmin, ideal[2] = plus_one, ideal[3] = vmax;
        } else {
            ideal[0] = plus_one, ideal[1] = vmax, ideal[2] = minus_one, ideal[3] = vmin;
        }
    }
    int mag0, mag1;
    bit_metrics(sym, ideal, mag0, mag1);
    /* dmr_compute_reliability, rf_mod == 1 */
    const float id = sc >= 2.0f ? 3.0f : (sc >= 0.0f ? 1.0f : (sc >= -2.0f ? -1.0f : -3.0f));
    float err = fabsf(__fsub_rn(sc, id));
    if (err > 1.0f) {
        err = 1.0f;
    }
    int rel = clamp255(__float2int_rz(__fadd_rn(__fmul_rn(__fsub_rn(1.0f, err), 255.0f), 0.5f)));
    if (p.snr_scale_num > 0) {
        rel = clamp255((rel * p.snr_scale_num) >> 8);
    }
    const int min_mag = mag0 < mag1 ? mag0 : mag1;
    if (min_mag > 0 && rel < min_mag) {
        mag0 = (mag0 * rel) / min_mag;
        mag1 = (mag1 * rel) / min_mag;
    }
    mag0 = clamp255(mag0);
    mag1 = clamp255(mag1);
    const int l0 = ((dibit >> 1) & 1) ? mag0 : -mag0, l1 = (dibit & 1) ? mag1 : -mag1;
    const size_t o = (size_t)ch * p.out_pitch + i;
    p.dibits[o] = (uint8_t)dibit;
    p.reliab[o] = (uint8_t)clamp255(min(abs(l0), abs(l1)));
    reinterpret_cast<short2*>(p.llr)[o] = make_short2((short)l0, (short)l1);
}

__global__ void
cqpsk_slicer_reset_kernel(float* sbuf, float* minbuf, float* maxbuf, int* sidx, int* midx, int* sum_window, double* a, double* b,
                          float* thr, int n_ch) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_ch) {
        return;
    }
    for (int k = 0; k < 128; k++) {
        sbuf[(size_t)ch * 128 + k] = 0.0f;
    }
    for (int k = 0; k < 1024; k++) { /* initState, src/core/util/dsd_init.c:519-592 */
        minbuf[(size_t)ch * 1024 + k] = -15000.0f;
        maxbuf[(size_t)ch * 1024 + k] = 15000.0f;
    }
    sidx[ch] = 0;
    midx[ch] = 0;
    sum_window[ch] = 0;
    a[ch] = 0.0;
    b[ch] = 0.0;
    const float init[8] = {-15000.0f, 15000.0f, 0.0f, 0.0f, 0.0f, -12000.0f, 12000.0f, 0.0f};
    for (int k = 0; k < 8; k++) {
        thr[(size_t)k * n_ch + ch] = init[k];
    }
}

}  // namespace

struct dsdneo_b200_symbolizer {
    int n_ch, rate, symrate, ssize, msize, use_cosine_filter, n_filters;
    int h_taps_len[DSDNEO_B200_SYM_MAX_FILTERS];
    void* arena; /* one allocation for all scalar arrays */
    SymScalars s;
    float *d_minbuf, *d_maxbuf, *d_sbuf, *d_carry, *d_hist, *d_taps;
    int* d_taps_len;
    float* d_filt[2]; /* matched-filter output, one buffer per pipeline slot */
    size_t filt_pitch[2];
    /* acquisition (getFrameSync on the device): configured by dsdneo_b200_symbolizer_set_acquire_patterns */
    int n_pat;
    AcqPattern pat[kAcqMaxPatterns];
    int *d_acquired, *d_hunt_since, *d_hunt_count, *d_lidx, *d_level_count, *d_hist_head, *d_start_off, *d_out_off;
    unsigned* d_hunt_bits;
    float *d_lbuf, *d_hist128;
    dsdneo_b200_acq_info* d_info;
};

extern "C" {

int
dsdneo_b200_sym_class_from_synctype(int synctype, int lastsynctype, int use_cosine_filter, dsdneo_b200_sym_class* out) {
    /* sync-type ids: include/dsd-neo/core/synctype_ids.h:30-123 */
    if (!out) {
        set_error("sym_class_from_synctype: NULL output");
        return DSDNEO_B200_EINVAL;
    }
    auto is_p25p1 = [](int s) { return s == 0 || s == 1; };
    auto is_dmr_bs = [](int s) { return s >= 10 && s <= 13; };
    auto is_dmr_ms = [](int s) { return s >= 32 && s <= 34; };
    auto is_ysf = [](int s) { return s == 30 || s == 31; };
    auto is_m17 = [](int s) { return s == 8 || s == 9 || s == 16 || s == 17 || s == 76 || s == 77 || s == 86 || s == 87 || (s >= 98 && s <= 101); };
    auto two_level = [](int s) { return s == 6 || s == 7 || s == 14 || s == 15 || s == 18 || s == 19 || s == 37 || s == 38; };
    if (two_level(synctype)) {
        set_error("sym_class_from_synctype: two-level sync types (D-STAR, ProVoice, EDACS) are not built yet");
        return DSDNEO_B200_EUNSUPPORTED;
    }
    if ((lastsynctype >= 20 && lastsynctype <= 29) || lastsynctype == 35 || lastsynctype == 36) {
        set_error("sym_class_from_synctype: dPMR / NXDN / P25p2 matched-filter selection depends on decoder options; pass the class explicitly");
        return DSDNEO_B200_EUNSUPPORTED;
    }
    /* matched filter (dsd_symbol.c:301-337) */
    int filter = DSDNEO_SYM_FILTER_NONE;
    if (use_cosine_filter) {
        if (is_dmr_bs(lastsynctype) || is_dmr_ms(lastsynctype) || is_ysf(lastsynctype)) {
            filter = DSDNEO_SYM_FILTER_DMR;
        } else if (is_m17(lastsynctype)) {
            filter = DSDNEO_SYM_FILTER_M17;
        } else if (is_p25p1(lastsynctype)) {
            filter = DSDNEO_SYM_FILTER_P25;
        }
    }
    out->filter = filter;
    out->rf_mod = 0;
    /* window (dsd_symbol.c:197-211): YSF by synctype, DMR BS / MS voice+data by lastsynctype */
    out->window_l = (is_ysf(synctype) || is_dmr_bs(lastsynctype) || lastsynctype == 32 || lastsynctype == 33) ? 1 : 2;
    out->track_minmax = is_p25p1(lastsynctype) ? 1 : 0; /* dsd_dibit.c:264 (rf_mod == 0) */
    /* is_four_level_neg_synctype (dsd_dibit.c:915-935) */
    const int s = synctype;
    out->negative = (s == 1 || s == 3 || s == 5 || s == 9 || s == 11 || s == 13 || s == 17 || s == 29 || s == 31 || s == 77 || s == 87
                     || s == 36 || s == 99 || s == 101)
                        ? 1
                        : 0;
    return 0;
}

void
dsdneo_b200_symbolizer_destroy(dsdneo_b200_symbolizer* y) {
    if (!y) {
        return;
    }
    cudaFree(y->arena);
    cudaFree(y->d_minbuf);
    cudaFree(y->d_maxbuf);
    cudaFree(y->d_sbuf);
    cudaFree(y->d_carry);
    cudaFree(y->d_hist);
    cudaFree(y->d_taps);
    cudaFree(y->d_taps_len);
    cudaFree(y->d_filt[0]);
    cudaFree(y->d_filt[1]);
    cudaFree(y->d_acquired);
    cudaFree(y->d_hunt_since);
    cudaFree(y->d_hunt_count);
    cudaFree(y->d_lidx);
    cudaFree(y->d_level_count);
    cudaFree(y->d_hist_head);
    cudaFree(y->d_start_off);
    cudaFree(y->d_out_off);
    cudaFree(y->d_hunt_bits);
    cudaFree(y->d_lbuf);
    cudaFree(y->d_hist128);
    cudaFree(y->d_info);
    free(y);
}

dsdneo_b200_symbolizer*
dsdneo_b200_symbolizer_create(const dsdneo_b200_symbolizer_config* cfg) {
    if (!cfg || cfg->n_channels <= 0 || cfg->output_rate_hz <= 0 || cfg->symbol_rate_hz <= 0 || cfg->n_filters < 0
        || cfg->n_filters > DSDNEO_B200_SYM_MAX_FILTERS) {
        set_error("symbolizer_create: bad config");
        return NULL;
    }
    for (int f = 0; f < cfg->n_filters; f++) {
        if (cfg->filter_len[f] < 0 || cfg->filter_len[f] > kMaxTaps || (cfg->filter_len[f] > 0 && !cfg->filter_taps[f])) {
            set_error("symbolizer_create: filter %d has %d taps (max %d)", f, cfg->filter_len[f], kMaxTaps);
            return NULL;
        }
    }
    if (ensure_device()) {
        return NULL;
    }
    dsdneo_b200_symbolizer* y = (dsdneo_b200_symbolizer*)calloc(1, sizeof(*y));
    if (!y) {
        set_error("symbolizer_create: out of host memory");
        return NULL;
    }
    const size_t n = (size_t)cfg->n_channels;
    y->n_ch = cfg->n_channels;
    y->rate = cfg->output_rate_hz;
    y->symrate = cfg->symbol_rate_hz;
    y->ssize = cfg->ssize > 0 ? cfg->ssize : 128;   /* opts->ssize default, src/core/util/dsd_init.c:169 */
    y->msize = cfg->msize > 0 ? cfg->msize : 1024;  /* opts->msize default, :170 */
    y->use_cosine_filter = cfg->use_cosine_filter;
    y->n_filters = cfg->n_filters;
    /* scalar arena: 24 x 4-byte arrays, 3 x 8-byte arrays */
    const size_t arena_bytes = n * (24 * 4 + 3 * 8) + 256;
    cudaError_t e = cudaMalloc(&y->arena, arena_bytes);
    if (e == cudaSuccess) {
        e = cudaMemset(y->arena, 0, arena_bytes);
    }
    if (e == cudaSuccess) {
        char* b = (char*)y->arena;
        double* d8 = (double*)b;
        y->s.minbuf_sum = d8;
        y->s.maxbuf_sum = d8 + n;
        y->s.symbolcnt = (long long*)(d8 + 2 * n);
        int* i4 = (int*)(d8 + 3 * n);
        int k = 0;
        y->s.filter = i4 + n * k++;
        y->s.window_l = i4 + n * k++;
        y->s.track = i4 + n * k++;
        y->s.negative = i4 + n * k++;
        y->s.rf_mod = i4 + n * k++;
        y->s.sps_num = i4 + n * k++;
        y->s.sps_den = i4 + n * k++;
        y->s.sps_accum = i4 + n * k++;
        y->s.sps = i4 + n * k++;
        y->s.center_idx = i4 + n * k++;
        y->s.jitter = i4 + n * k++;
        y->s.lastsample = (float*)(i4 + n * k++);
        y->s.vmin = (float*)(i4 + n * k++);
        y->s.vmax = (float*)(i4 + n * k++);
        y->s.center = (float*)(i4 + n * k++);
        y->s.umid = (float*)(i4 + n * k++);
        y->s.lmid = (float*)(i4 + n * k++);
        y->s.minref = (float*)(i4 + n * k++);
        y->s.maxref = (float*)(i4 + n * k++);
        y->s.sidx = i4 + n * k++;
        y->s.midx = i4 + n * k++;
        y->s.sum_window = i4 + n * k++;
        y->s.carry_n = i4 + n * k++;
        y->s.snr_num = i4 + n * k++;
    }
#define SYM_ALLOC(ptr, bytes)                                                                                          \
    if (e == cudaSuccess) {                                                                                            \
        e = cudaMalloc((void**)&(ptr), (bytes));                                                                       \
    }
    SYM_ALLOC(y->d_minbuf, n * kMinMax * sizeof(float));
    SYM_ALLOC(y->d_maxbuf, n * kMinMax * sizeof(float));
    SYM_ALLOC(y->d_sbuf, n * kSbuf * sizeof(float));
    SYM_ALLOC(y->d_carry, n * kCarry * sizeof(float));
    SYM_ALLOC(y->d_hist, n * kMaxTaps * sizeof(float));
    SYM_ALLOC(y->d_taps, (size_t)DSDNEO_B200_SYM_MAX_FILTERS * kMaxTaps * sizeof(float));
    SYM_ALLOC(y->d_taps_len, DSDNEO_B200_SYM_MAX_FILTERS * sizeof(int));
#undef SYM_ALLOC
    if (e == cudaSuccess) {
        float* h = (float*)calloc((size_t)DSDNEO_B200_SYM_MAX_FILTERS * kMaxTaps, sizeof(float));
        for (int f = 0; f < cfg->n_filters; f++) {
            y->h_taps_len[f] = cfg->filter_len[f];
            if (cfg->filter_len[f] > 0) {
                memcpy(h + (size_t)f * kMaxTaps, cfg->filter_taps[f], (size_t)cfg->filter_len[f] * sizeof(float));
            }
        }
        e = cudaMemcpy(y->d_taps, h, (size_t)DSDNEO_B200_SYM_MAX_FILTERS * kMaxTaps * sizeof(float), cudaMemcpyHostToDevice);
        free(h);
    }
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->d_taps_len, y->h_taps_len, sizeof(y->h_taps_len), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "symbolizer_create", __FILE__, __LINE__);
        dsdneo_b200_symbolizer_destroy(y);
        return NULL;
    }
    if (dsdneo_b200_symbolizer_reset(y, NULL) != 0) {
        dsdneo_b200_symbolizer_destroy(y);
        return NULL;
    }
    /* default class: no sync seen yet => no matched filter, window 2/2, no tracking, positive polarity */
    dsdneo_b200_sym_class* cls = (dsdneo_b200_sym_class*)malloc(n * sizeof(dsdneo_b200_sym_class));
    for (size_t i = 0; i < n; i++) {
        cls[i].filter = DSDNEO_SYM_FILTER_NONE, cls[i].window_l = 2, cls[i].track_minmax = 0, cls[i].negative = 0, cls[i].rf_mod = 0;
    }
    int rc = dsdneo_b200_symbolizer_set_class(y, cls);
    free(cls);
    if (rc == 0) {
        rc = dsdneo_b200_symbolizer_set_snr(y, NULL);
    }
    if (rc) {
        dsdneo_b200_symbolizer_destroy(y);
        return NULL;
    }
    return y;
}

int
dsdneo_b200_symbolizer_reset(dsdneo_b200_symbolizer* y, void* stream) {
    if (!y) {
        set_error("symbolizer_reset: NULL");
        return DSDNEO_B200_EINVAL;
    }
    sym_reset_kernel<<<y->n_ch, 128, 0, as_stream(stream)>>>(y->s, y->d_minbuf, y->d_maxbuf, y->d_sbuf, y->d_carry, y->d_hist, y->n_ch);
    DSDNEO_KERNEL_CHECK();
    count_launch();
    if (y->d_acquired) { /* every channel hunts again, with an empty window and an empty symbol history */
        const size_t n = (size_t)y->n_ch;
        cudaStream_t s = as_stream(stream);
        DSDNEO_CUDA(cudaMemsetAsync(y->d_acquired, 0, n * sizeof(int), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_hunt_since, 0, n * sizeof(int), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_hunt_count, 0, n * sizeof(int), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_lidx, 0, n * sizeof(int), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_level_count, 0, n * sizeof(int), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_hist_head, 0, n * sizeof(int), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_hunt_bits, 0, n * sizeof(unsigned), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_lbuf, 0, n * 24 * sizeof(float), s));
        DSDNEO_CUDA(cudaMemsetAsync(y->d_hist128, 0, n * 128 * sizeof(float), s));
    }
    return 0;
}

int
dsdneo_b200_symbolizer_set_class(dsdneo_b200_symbolizer* y, const dsdneo_b200_sym_class* per_channel) {
    if (!y || !per_channel) {
        set_error("symbolizer_set_class: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    const size_t n = (size_t)y->n_ch;
    int* h = (int*)malloc(5 * n * sizeof(int));
    if (!h) {
        set_error("symbolizer_set_class: out of host memory");
        return DSDNEO_B200_ENOMEM;
    }
    for (size_t i = 0; i < n; i++) {
        int f = per_channel[i].filter;
        if (f >= y->n_filters || (f >= 0 && y->h_taps_len[f] <= 0)) {
            free(h);
            set_error("symbolizer_set_class: channel %zu selects filter %d which was not supplied at create", i, f);
            return DSDNEO_B200_EINVAL;
        }
        h[i] = f < 0 ? -1 : f;
        h[n + i] = per_channel[i].window_l;
        h[2 * n + i] = per_channel[i].track_minmax ? 1 : 0;
        h[3 * n + i] = per_channel[i].negative ? 1 : 0;
        h[4 * n + i] = per_channel[i].rf_mod == 2 ? 2 : 0;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->s.filter, h, n * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->s.window_l, h + n, n * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->s.track, h + 2 * n, n * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->s.negative, h + 3 * n, n * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->s.rf_mod, h + 4 * n, n * sizeof(int), cudaMemcpyHostToDevice);
    }
    free(h);
    if (e != cudaSuccess) {
        return cuda_fail(e, "symbolizer_set_class", __FILE__, __LINE__);
    }
    return 0;
}

/* apply_c4fm_snr_weight (src/core/frames/dsd_dibit.c:504-546): the reliability of every dibit is scaled by
 * (204 + (w256 >> 2)) / 256, w256 from the C4FM SNR the metrics hook reports (sentinel -100 dB => w256 = 0). */
static int
snr_scale_num(double snr_db) {
    int w256 = 0;
    if (snr_db > -13.0) {
        if (snr_db >= 12.0) {
            w256 = 255;
        } else {
            double w = (snr_db + 13.0) / 25.0;
            w = w < 0.0 ? 0.0 : (w > 1.0 ? 1.0 : w);
            w256 = (int)(w * 255.0 + 0.5);
        }
    }
    return 204 + (w256 >> 2);
}

int
dsdneo_b200_symbolizer_set_snr(dsdneo_b200_symbolizer* y, const double* h_snr_c4fm_db) {
    if (!y) {
        set_error("symbolizer_set_snr: NULL symbolizer");
        return DSDNEO_B200_EINVAL;
    }
    const size_t n = (size_t)y->n_ch;
    int* h = (int*)malloc(n * sizeof(int));
    if (!h) {
        set_error("symbolizer_set_snr: out of host memory");
        return DSDNEO_B200_ENOMEM;
    }
    for (size_t i = 0; i < n; i++) {
        h[i] = snr_scale_num(h_snr_c4fm_db ? h_snr_c4fm_db[i] : -100.0);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) {
        e = cudaMemcpy(y->s.snr_num, h, n * sizeof(int), cudaMemcpyHostToDevice);
    }
    free(h);
    if (e != cudaSuccess) {
        return cuda_fail(e, "symbolizer_set_snr", __FILE__, __LINE__);
    }
    return 0;
}

int
dsdneo_b200_symbolizer_set_acquire_patterns(dsdneo_b200_symbolizer* y, const dsdneo_b200_acq_pattern* patterns, int n_patterns) {
    if (!y || !patterns || n_patterns < 1 || n_patterns > kAcqMaxPatterns) {
        set_error("symbolizer_set_acquire_patterns: 1..%d patterns", kAcqMaxPatterns);
        return DSDNEO_B200_EINVAL;
    }
    for (int k = 0; k < n_patterns; k++) {
        const char* sym = patterns[k].symbols;
        if (!sym || strlen(sym) != 24) {
            set_error("symbolizer_set_acquire_patterns: pattern %d is not a 24-symbol '1' / '3' string", k);
            return DSDNEO_B200_EUNSUPPORTED;
        }
        unsigned bits = 0;
        for (int i = 0; i < 24; i++) {
            if (sym[i] != '1' && sym[i] != '3') {
                set_error("symbolizer_set_acquire_patterns: pattern %d has a symbol other than '1' / '3'", k);
                return DSDNEO_B200_EINVAL;
            }
            bits = (bits << 1) | (sym[i] == '1' ? 1u : 0u);
        }
        const int f = patterns[k].cls.filter;
        if (f >= y->n_filters || (f >= 0 && y->h_taps_len[f] <= 0)) {
            set_error("symbolizer_set_acquire_patterns: pattern %d selects filter %d which was not supplied at create", k, f);
            return DSDNEO_B200_EINVAL;
        }
        {
            int whole = y->rate / y->symrate;
            whole = whole < 2 ? 2 : (whole > 64 ? 64 : whole);
            if (f >= 0 && y->h_taps_len[f] - 1 + 2 * (whole + 2) > kAcqMargin) {
                set_error("symbolizer_set_acquire_patterns: filter %d (%d taps) at %d samples per symbol does not start up within "
                          "the %d-sample launch margin", f, y->h_taps_len[f], whole, kAcqMargin);
                return DSDNEO_B200_EUNSUPPORTED;
            }
        }
        y->pat[k].bits = bits;
        y->pat[k].sync_type = patterns[k].sync_type;
        y->pat[k].kind = patterns[k].kind;
        y->pat[k].filter = f < 0 ? -1 : f;
        y->pat[k].window_l = patterns[k].cls.window_l;
        y->pat[k].track = patterns[k].cls.track_minmax ? 1 : 0;
        y->pat[k].negative = patterns[k].cls.negative ? 1 : 0;
    }
    y->n_pat = n_patterns;
    if (!y->d_acquired) {
        const size_t n = (size_t)y->n_ch;
        cudaError_t e = cudaSuccess;
#define ACQ_ALLOC(ptr, bytes)                                                                                          \
    if (e == cudaSuccess) {                                                                                            \
        e = cudaMalloc((void**)&(ptr), (bytes));                                                                       \
        if (e == cudaSuccess) {                                                                                        \
            e = cudaMemset((ptr), 0, (bytes));                                                                         \
        }                                                                                                              \
    }
        ACQ_ALLOC(y->d_acquired, n * sizeof(int));
        ACQ_ALLOC(y->d_hunt_since, n * sizeof(int));
        ACQ_ALLOC(y->d_hunt_count, n * sizeof(int));
        ACQ_ALLOC(y->d_lidx, n * sizeof(int));
        ACQ_ALLOC(y->d_level_count, n * sizeof(int));
        ACQ_ALLOC(y->d_hist_head, n * sizeof(int));
        ACQ_ALLOC(y->d_start_off, n * sizeof(int));
        ACQ_ALLOC(y->d_out_off, n * sizeof(int));
        ACQ_ALLOC(y->d_hunt_bits, n * sizeof(unsigned));
        ACQ_ALLOC(y->d_lbuf, n * 24 * sizeof(float));
        ACQ_ALLOC(y->d_hist128, n * 128 * sizeof(float));
        ACQ_ALLOC(y->d_info, n * sizeof(dsdneo_b200_acq_info));
#undef ACQ_ALLOC
        if (e != cudaSuccess) {
            return cuda_fail(e, "symbolizer_set_acquire_patterns", __FILE__, __LINE__);
        }
    }
    return 0;
}

int
dsdneo_b200_symbolizer_set_acquired(dsdneo_b200_symbolizer* y, const int* h_acquired) {
    if (!y || !y->d_acquired) {
        set_error("symbolizer_set_acquired: acquisition is not configured (set_acquire_patterns)");
        return DSDNEO_B200_EINVAL;
    }
    const size_t n = (size_t)y->n_ch;
    DSDNEO_CUDA(cudaDeviceSynchronize());
    if (h_acquired) {
        DSDNEO_CUDA(cudaMemcpy(y->d_acquired, h_acquired, n * sizeof(int), cudaMemcpyHostToDevice));
    } else {
        DSDNEO_CUDA(cudaMemset(y->d_acquired, 0, n * sizeof(int)));
    }
    /* a new hunt starts with an empty window */
    DSDNEO_CUDA(cudaMemset(y->d_hunt_since, 0, n * sizeof(int)));
    DSDNEO_CUDA(cudaMemset(y->d_hunt_count, 0, n * sizeof(int)));
    DSDNEO_CUDA(cudaMemset(y->d_lidx, 0, n * sizeof(int)));
    DSDNEO_CUDA(cudaMemset(y->d_level_count, 0, n * sizeof(int)));
    DSDNEO_CUDA(cudaMemset(y->d_hunt_bits, 0, n * sizeof(unsigned)));
    return 0;
}

int
dsdneo_b200_selftest_div5(unsigned long long* d_n_mismatch, unsigned* d_first_bad, void* stream) {
    if (!d_n_mismatch || !d_first_bad) {
        set_error("selftest_div5: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemsetAsync(d_n_mismatch, 0, sizeof(unsigned long long), as_stream(stream)));
    DSDNEO_CUDA(cudaMemsetAsync(d_first_bad, 0xff, sizeof(unsigned), as_stream(stream)));
    div5_selftest_kernel<<<4096, 256, 0, as_stream(stream)>>>(d_n_mismatch, d_first_bad); /* 2^20 threads x 4096 operands */
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

static int
symbolize_check(dsdneo_b200_symbolizer* y, int n_samples, int mode, const dsdneo_b200_symbol_out* out) {
    if (!y || !out || n_samples < 0 || !out->d_symbols || !out->d_count || out->pitch == 0) {
        set_error("symbolize_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (mode != DSDNEO_SYM_MODE_GET_SYMBOL && mode != DSDNEO_SYM_MODE_GET_DIBIT_SOFT) {
        set_error("symbolize_batch: unknown mode %d", mode);
        return DSDNEO_B200_EINVAL;
    }
    if (mode == DSDNEO_SYM_MODE_GET_DIBIT_SOFT && (!out->d_dibits || !out->d_reliability || !out->d_llr)) {
        set_error("symbolize_batch: GET_DIBIT_SOFT needs dibit, reliability and llr outputs");
        return DSDNEO_B200_EINVAL;
    }
    int whole = y->rate / y->symrate;
    whole = whole < 2 ? 2 : (whole > 64 ? 64 : whole);
    const size_t need = ((size_t)n_samples + kCarry) / (size_t)(whole - 1) + 2;
    if (out->pitch < need) {
        set_error("symbolize_batch: output pitch %zu is too small for %d samples (need >= %zu)", out->pitch, n_samples, need);
        return DSDNEO_B200_EINVAL;
    }
    return 0;
}

static SymParams
symbolize_params(dsdneo_b200_symbolizer* y, int n_samples, int mode, int have_sync, const dsdneo_b200_symbol_out* out, int slot) {
    SymParams sp;
    sp.s = y->s;
    sp.filt = y->d_filt[slot];
    sp.filt_pitch = y->filt_pitch[slot];
    sp.carry = y->d_carry;
    sp.sbuf = y->d_sbuf;
    sp.minbuf = y->d_minbuf;
    sp.maxbuf = y->d_maxbuf;
    sp.symbols = out->d_symbols;
    sp.dibits = out->d_dibits;
    sp.reliab = out->d_reliability;
    sp.llr = out->d_llr;
    sp.count = out->d_count;
    sp.out_pitch = out->pitch;
    sp.n_ch = y->n_ch;
    sp.n = n_samples;
    sp.mode = mode;
    sp.have_sync = have_sync ? 1 : 0;
    sp.rate = y->rate;
    sp.symrate = y->symrate;
    sp.ssize = y->ssize;
    sp.msize = y->msize;
    sp.snr_num = y->s.snr_num;
    sp.start_off = NULL;
    sp.out_off = NULL;
    sp.acquired = NULL;
    return sp;
}

} /* extern "C" */

/*
 * The two halves of dsdneo_b200_symbolize_batch as separate stages, so that a caller which pipelines consecutive launches
 * (csrc/p25p1_rx.cu) can run the matched filter of launch i+1 under the slicer of launch i: fir_stage writes the filter output
 * of one launch into buffer `slot` (0 / 1) and advances the filter history; sym_stage consumes that buffer.  Each stage keeps
 * its own carried state, so the only ordering a caller owes is fir(i) -> sym(i) and sym(i) -> fir(i + 2).
 */
int
dsdneo_symbolize_fir_stage(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int slot, cudaStream_t s) {
    if (!y || !d_disc || n_samples < 0 || disc_pitch < (size_t)n_samples || slot < 0 || slot > 1) {
        set_error("symbolize fir stage: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t pitch = (((size_t)n_samples + 31) & ~(size_t)31) + 32; /* rows start on 128-byte lines */
    if (!y->d_filt[slot] || y->filt_pitch[slot] < pitch) {
        DSDNEO_CUDA(cudaDeviceSynchronize());
        cudaFree(y->d_filt[slot]);
        y->d_filt[slot] = NULL;
        DSDNEO_CUDA(cudaMalloc((void**)&y->d_filt[slot], (size_t)y->n_ch * pitch * sizeof(float)));
        y->filt_pitch[slot] = pitch;
    }
    if (n_samples == 0) {
        return 0;
    }
    FirParams fp;
    fp.in = d_disc;
    fp.in_pitch = disc_pitch;
    fp.out = y->d_filt[slot];
    fp.out_pitch = y->filt_pitch[slot];
    fp.taps = y->d_taps;
    fp.taps_len = y->d_taps_len;
    fp.filter = y->s.filter;
    fp.hist = y->d_hist;
    fp.n = n_samples;
    /* the bulk-copy form needs 16-byte aligned rows on both sides */
    const bool aligned = (disc_pitch & 3) == 0 && (((uintptr_t)d_disc) & 15) == 0 && (y->filt_pitch[slot] & 3) == 0;
    if (aligned) {
        dim3 grid((unsigned)((n_samples + kFir8Tile - 1) / kFir8Tile), (unsigned)y->n_ch);
        KernelTimer kt("sps_fir_kernel", s);
        sps_fir8_kernel<<<grid, kFir8Threads, 0, s>>>(fp);
    } else {
        dim3 grid((unsigned)((n_samples + kFirTile - 1) / kFirTile), (unsigned)y->n_ch);
        KernelTimer kt("sps_fir_kernel", s);
        sps_fir_kernel<<<grid, kFirThreads, 0, s>>>(fp);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    {
        KernelTimer kt("sps_fir_hist_kernel", s);
        sps_fir_hist_kernel<<<(y->n_ch + 7) / 8, 256, 0, s>>>(d_disc, disc_pitch, y->d_hist, y->s.filter, y->d_taps_len, y->n_ch, n_samples);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_symbolize_sym_stage(dsdneo_b200_symbolizer* y, int n_samples, int mode, int have_sync, const dsdneo_b200_symbol_out* out, int slot,
                           cudaStream_t s) {
    int rc = symbolize_check(y, n_samples, mode, out);
    if (rc) {
        return rc;
    }
    if (slot < 0 || slot > 1 || !y->d_filt[slot]) {
        set_error("symbolize sym stage: no filter output in slot %d", slot);
        return DSDNEO_B200_EINVAL;
    }
    const SymParams sp = symbolize_params(y, n_samples, mode, have_sync, out, slot);
    {
        KernelTimer kt("symbolize_kernel", s);
        symbolize_kernel<<<(y->n_ch + kSymWarps - 1) / kSymWarps, kSymWarps * 32, 0, s>>>(sp);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

/* The acquisition form of the two stages as one step on one stream (hunting channels read the RAW samples, so the matched
 * filter of the launch cannot run ahead of them): sym_acquire_kernel, then the matched filter into buffer `slot`, then the
 * slicer for the synchronised part of every channel.  Used by dsdneo_b200_symbolize_acquire_batch and by the receive bank
 * while it acquires (csrc/p25p1_rx.cu). */
/* Channels flagged in drop[] leave the synchronised state: they hunt from an empty window, exactly as after
 * dsdneo_b200_symbolizer_set_acquired() with their flag at 0.  The flags are consumed. */
__global__ void
sym_drop_kernel(int* drop, int* acquired, int* hunt_since, int* hunt_count, int* lidx, int* level_count, unsigned* hunt_bits, int n_ch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch || !drop[c]) {
        return;
    }
    drop[c] = 0;
    if (acquired[c]) {
        acquired[c] = 0;
        hunt_since[c] = 0, hunt_count[c] = 0, lidx[c] = 0, level_count[c] = 0, hunt_bits[c] = 0u;
    }
}

/* host copy of the per-channel synchronised (1) / hunting (0) flags; synchronises the device */
int
dsdneo_symbolize_get_acquired(dsdneo_b200_symbolizer* y, int* h_acquired) {
    if (!y || !h_acquired) {
        set_error("symbolizer: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (!y->d_acquired) { /* acquisition never configured: every channel runs the synchronised rules */
        for (int c = 0; c < y->n_ch; c++) {
            h_acquired[c] = 1;
        }
        return 0;
    }
    DSDNEO_CUDA(cudaDeviceSynchronize());
    DSDNEO_CUDA(cudaMemcpy(h_acquired, y->d_acquired, (size_t)y->n_ch * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int
dsdneo_symbolize_drop_stage(dsdneo_b200_symbolizer* y, int* d_drop, cudaStream_t s) {
    if (!y || !y->d_acquired || !d_drop) {
        set_error("symbolize drop stage: acquisition is not configured");
        return DSDNEO_B200_EINVAL;
    }
    {
        KernelTimer kt("sym_drop_kernel", s);
        sym_drop_kernel<<<(y->n_ch + 127) / 128, 128, 0, s>>>(d_drop, y->d_acquired, y->d_hunt_since, y->d_hunt_count, y->d_lidx,
                                                              y->d_level_count, y->d_hunt_bits, y->n_ch);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_symbolize_acquire_stage(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int mode, int have_sync,
                               const dsdneo_b200_symbol_out* out, dsdneo_b200_acq_info* d_info, int slot, int hunt_filtered,
                               cudaStream_t s) {
    if (!y || !y->d_acquired || y->n_pat < 1 || slot < 0 || slot > 1) {
        set_error("symbolize acquire stage: acquisition is not configured (symbolizer_set_acquire_patterns)");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    if (hunt_filtered == 1) { /* the matched filter already runs (every hunt after a channel's first sync): hunt on its output
                               * (2: the caller has run the filter stage for this slot itself, on another stream) */
        rc = dsdneo_symbolize_fir_stage(y, d_disc, disc_pitch, n_samples, slot, s);
        if (rc) {
            return rc;
        }
    }
    /* hunting channels first: they read the samples and may switch their class for the stages below */
    AcqParams ap;
    ap.sp = symbolize_params(y, n_samples, mode, have_sync, out, slot);
    ap.sp.filt = hunt_filtered ? y->d_filt[slot] : d_disc;
    ap.sp.filt_pitch = hunt_filtered ? y->filt_pitch[slot] : disc_pitch;
    ap.filtered_input = hunt_filtered ? 1 : 0;
    ap.taps = y->d_taps;
    ap.taps_len = y->d_taps_len;
    for (int k = 0; k < y->n_pat; k++) {
        ap.pat[k] = y->pat[k];
    }
    ap.n_pat = y->n_pat;
    ap.acquired = y->d_acquired;
    ap.hunt_since = y->d_hunt_since;
    ap.hunt_bits = y->d_hunt_bits;
    ap.hunt_count = y->d_hunt_count;
    ap.lbuf = y->d_lbuf;
    ap.lidx = y->d_lidx;
    ap.level_count = y->d_level_count;
    ap.hist = y->d_hist128;
    ap.hist_head = y->d_hist_head;
    ap.start_off = y->d_start_off;
    ap.out_off = y->d_out_off;
    ap.info = d_info ? d_info : y->d_info;
    ap.fir_hist = y->d_hist;
    {
        KernelTimer kt("sym_acquire_kernel", s);
        sym_acquire_kernel<<<(y->n_ch + kSymWarps - 1) / kSymWarps, kSymWarps * 32, 0, s>>>(ap);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    if (!hunt_filtered) {
        rc = dsdneo_symbolize_fir_stage(y, d_disc, disc_pitch, n_samples, slot, s);
        if (rc) {
            return rc;
        }
    }
    SymParams sp = symbolize_params(y, n_samples, mode, have_sync, out, slot);
    sp.start_off = y->d_start_off;
    sp.out_off = y->d_out_off;
    sp.acquired = y->d_acquired;
    {
        KernelTimer kt("symbolize_kernel", s);
        symbolize_kernel<<<(y->n_ch + kSymWarps - 1) / kSymWarps, kSymWarps * 32, 0, s>>>(sp);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

extern "C" {

static int
symbolize_launch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int mode, int have_sync,
                 const dsdneo_b200_symbol_out* out, bool acquire, dsdneo_b200_acq_info* d_info, void* stream) {
    int rc = symbolize_check(y, n_samples, mode, out);
    if (rc) {
        return rc;
    }
    if (!d_disc || disc_pitch < (size_t)n_samples) {
        set_error("symbolize_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    cudaStream_t s = as_stream(stream);
    if (!acquire) {
        rc = dsdneo_symbolize_fir_stage(y, d_disc, disc_pitch, n_samples, 0, s);
        return rc ? rc : dsdneo_symbolize_sym_stage(y, n_samples, mode, have_sync, out, 0, s);
    }
    return dsdneo_symbolize_acquire_stage(y, d_disc, disc_pitch, n_samples, mode, have_sync, out, d_info, 0, 0, s);
}

int
dsdneo_b200_symbolize_reacquire_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples,
                                      const dsdneo_b200_symbol_out* out, dsdneo_b200_acq_info* d_info, void* stream) {
    if (!y || !y->d_acquired || y->n_pat < 1) {
        set_error("symbolize_reacquire_batch: acquisition is not configured (symbolizer_set_acquire_patterns)");
        return DSDNEO_B200_EINVAL;
    }
    if (n_samples < kAcqMargin) {
        set_error("symbolize_reacquire_batch: a launch must carry at least %d samples", kAcqMargin);
        return DSDNEO_B200_EINVAL;
    }
    int rc = symbolize_check(y, n_samples, DSDNEO_SYM_MODE_GET_DIBIT_SOFT, out);
    if (rc) {
        return rc;
    }
    if (!d_disc || disc_pitch < (size_t)n_samples) {
        set_error("symbolize_reacquire_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    return dsdneo_symbolize_acquire_stage(y, d_disc, disc_pitch, n_samples, DSDNEO_SYM_MODE_GET_DIBIT_SOFT, 1, out, d_info, 0, 1,
                                          as_stream(stream));
}

int
dsdneo_b200_symbolize_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int mode, int have_sync,
                            const dsdneo_b200_symbol_out* out, void* stream) {
    return symbolize_launch(y, d_disc, disc_pitch, n_samples, mode, have_sync, out, false, NULL, stream);
}

int
dsdneo_b200_symbolize_acquire_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples,
                                    const dsdneo_b200_symbol_out* out, dsdneo_b200_acq_info* d_info, void* stream) {
    if (!y || !y->d_acquired || y->n_pat < 1) {
        set_error("symbolize_acquire_batch: acquisition is not configured (symbolizer_set_acquire_patterns)");
        return DSDNEO_B200_EINVAL;
    }
    if (n_samples < kAcqMargin) {
        set_error("symbolize_acquire_batch: a launch must carry at least %d samples", kAcqMargin);
        return DSDNEO_B200_EINVAL;
    }
    return symbolize_launch(y, d_disc, disc_pitch, n_samples, DSDNEO_SYM_MODE_GET_DIBIT_SOFT, 1, out, true, d_info, stream);
}

} /* extern "C" */

/* ---- symbol-rate CQPSK slicer ------------------------------------------------------------------------------------------ */

struct dsdneo_b200_cqpsk_slicer {
    int n_ch, ssize, msize, snr_scale_num;
    uint8_t *d_negative, *d_p25, *d_map;
    float *d_sbuf, *d_minbuf, *d_maxbuf, *d_thr;
    int *d_sidx, *d_midx, *d_sum_window;
    double *d_min_sum, *d_max_sum;
    float2* d_minmax; /* per-symbol {min, max} between the tracker and the digitize kernel, grown on demand */
    size_t minmax_cap;
};

extern "C" {

dsdneo_b200_cqpsk_slicer*
dsdneo_b200_cqpsk_slicer_create(int n_channels, int ssize, int msize) {
    if (n_channels <= 0) {
        set_error("cqpsk_slicer_create: bad channel count");
        return NULL;
    }
    if (ensure_device()) {
        return NULL;
    }
    dsdneo_b200_cqpsk_slicer* q = (dsdneo_b200_cqpsk_slicer*)calloc(1, sizeof(*q));
    if (!q) {
        set_error("cqpsk_slicer_create: out of host memory");
        return NULL;
    }
    const size_t n = (size_t)n_channels;
    q->n_ch = n_channels;
    q->ssize = ssize > 0 ? ssize : 128;   /* opts->ssize / msize defaults, src/core/util/dsd_init.c:169-170 */
    q->msize = msize > 0 ? msize : 1024;
    cudaError_t e = cudaMalloc((void**)&q->d_negative, n);
#define CQS_ALLOC(ptr, bytes)                                                                                          \
    if (e == cudaSuccess) {                                                                                            \
        e = cudaMalloc((void**)&(ptr), (bytes));                                                                       \
    }
    CQS_ALLOC(q->d_p25, n);
    CQS_ALLOC(q->d_map, n);
    CQS_ALLOC(q->d_sbuf, n * 128 * sizeof(float));
    CQS_ALLOC(q->d_minbuf, n * 1024 * sizeof(float));
    CQS_ALLOC(q->d_maxbuf, n * 1024 * sizeof(float));
    CQS_ALLOC(q->d_thr, n * 8 * sizeof(float));
    CQS_ALLOC(q->d_sidx, n * sizeof(int));
    CQS_ALLOC(q->d_midx, n * sizeof(int));
    CQS_ALLOC(q->d_sum_window, n * sizeof(int));
    CQS_ALLOC(q->d_min_sum, n * sizeof(double));
    CQS_ALLOC(q->d_max_sum, n * sizeof(double));
#undef CQS_ALLOC
    if (e == cudaSuccess) {
        e = cudaMemset(q->d_negative, 0, n);
    }
    if (e == cudaSuccess) {
        e = cudaMemset(q->d_p25, 1, n);
    }
    if (e == cudaSuccess) {
        e = cudaMemset(q->d_map, 0, n);
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "cqpsk_slicer_create", __FILE__, __LINE__);
        dsdneo_b200_cqpsk_slicer_destroy(q);
        return NULL;
    }
    if (dsdneo_b200_cqpsk_slicer_reset(q, NULL) != 0 || cudaStreamSynchronize(0) != cudaSuccess) {
        dsdneo_b200_cqpsk_slicer_destroy(q);
        return NULL;
    }
    return q;
}

void
dsdneo_b200_cqpsk_slicer_destroy(dsdneo_b200_cqpsk_slicer* q) {
    if (!q) {
        return;
    }
    cudaFree(q->d_negative);
    cudaFree(q->d_p25);
    cudaFree(q->d_map);
    cudaFree(q->d_sbuf);
    cudaFree(q->d_minbuf);
    cudaFree(q->d_maxbuf);
    cudaFree(q->d_thr);
    cudaFree(q->d_sidx);
    cudaFree(q->d_midx);
    cudaFree(q->d_sum_window);
    cudaFree(q->d_min_sum);
    cudaFree(q->d_max_sum);
    cudaFree(q->d_minmax);
    free(q);
}

int
dsdneo_b200_cqpsk_slicer_reset(dsdneo_b200_cqpsk_slicer* q, void* stream) {
    if (!q) {
        set_error("cqpsk_slicer_reset: NULL slicer");
        return DSDNEO_B200_EINVAL;
    }
    cqpsk_slicer_reset_kernel<<<(q->n_ch + 127) / 128, 128, 0, as_stream(stream)>>>(q->d_sbuf, q->d_minbuf, q->d_maxbuf, q->d_sidx,
                                                                                  q->d_midx, q->d_sum_window, q->d_min_sum,
                                                                                  q->d_max_sum, q->d_thr, q->n_ch);
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_cqpsk_slicer_set_class(dsdneo_b200_cqpsk_slicer* q, const uint8_t* h_negative, const uint8_t* h_p25_slice,
                                   const uint8_t* h_map_idx, double snr_cqpsk_db) {
    if (!q) {
        set_error("cqpsk_slicer_set_class: NULL slicer");
        return DSDNEO_B200_EINVAL;
    }
    const size_t n = (size_t)q->n_ch;
    DSDNEO_CUDA(cudaDeviceSynchronize());
    if (h_negative) {
        DSDNEO_CUDA(cudaMemcpy(q->d_negative, h_negative, n, cudaMemcpyHostToDevice));
    }
    if (h_p25_slice) {
        DSDNEO_CUDA(cudaMemcpy(q->d_p25, h_p25_slice, n, cudaMemcpyHostToDevice));
    }
    if (h_map_idx) {
        DSDNEO_CUDA(cudaMemcpy(q->d_map, h_map_idx, n, cudaMemcpyHostToDevice));
    }
    /* apply_cqpsk_snr_weight, src/core/frames/dsd_dibit.c:404-427 */
    q->snr_scale_num = 0;
    if (!(snr_cqpsk_db <= -50.0)) {
        int w256 = 0;
        if (snr_cqpsk_db >= 25.0) {
            w256 = 255;
        } else if (snr_cqpsk_db > 0.0) {
            w256 = (int)((snr_cqpsk_db / 25.0) * 255.0 + 0.5);
        }
        q->snr_scale_num = 204 + (w256 >> 2);
    }
    return 0;
}

int
dsdneo_b200_cqpsk_slice_batch(dsdneo_b200_cqpsk_slicer* q, const float* d_symbols, size_t symbols_pitch, const int* d_n_symbols,
                              uint8_t* d_dibits, uint8_t* d_reliability, int16_t* d_llr, size_t out_pitch, void* stream) {
    if (!q || !d_symbols || !d_n_symbols || !d_dibits || !d_reliability || !d_llr) {
        set_error("cqpsk_slice_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t need = (size_t)q->n_ch * out_pitch;
    if (q->minmax_cap < need) {
        DSDNEO_CUDA(cudaDeviceSynchronize());
        cudaFree(q->d_minmax);
        q->d_minmax = NULL;
        q->minmax_cap = 0;
        DSDNEO_CUDA(cudaMalloc((void**)&q->d_minmax, need * sizeof(float2)));
        q->minmax_cap = need;
    }
    if (symbols_pitch > out_pitch) {
        set_error("cqpsk_slice_batch: out_pitch smaller than symbols_pitch");
        return DSDNEO_B200_EINVAL;
    }
    CqSliceParams p;
    p.minmax = q->d_minmax;
    p.symbols = d_symbols;
    p.sym_pitch = symbols_pitch;
    p.n_symbols = d_n_symbols;
    p.dibits = d_dibits;
    p.reliab = d_reliability;
    p.llr = d_llr;
    p.out_pitch = out_pitch;
    p.negative = q->d_negative;
    p.p25_slice = q->d_p25;
    p.map_idx = q->d_map;
    p.sbuf = q->d_sbuf;
    p.minbuf = q->d_minbuf;
    p.maxbuf = q->d_maxbuf;
    p.sidx = q->d_sidx;
    p.midx = q->d_midx;
    p.sum_window = q->d_sum_window;
    p.minbuf_sum = q->d_min_sum;
    p.maxbuf_sum = q->d_max_sum;
    p.thr = q->d_thr;
    p.n_ch = q->n_ch;
    p.ssize = q->ssize;
    p.msize = q->msize;
    p.snr_scale_num = q->snr_scale_num;
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("cqpsk_slice_kernel", s);
        cqpsk_slice_kernel<<<(q->n_ch + kCqWarps - 1) / kCqWarps, kCqWarps * 32, 0, s>>>(p);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    {
        KernelTimer kt("cqpsk_digitize_kernel", s);
        dim3 grid((unsigned)((symbols_pitch + 255) / 256), (unsigned)q->n_ch);
        cqpsk_digitize_kernel<<<grid, 256, 0, s>>>(p);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_cqpsk_slicer_get_state(dsdneo_b200_cqpsk_slicer* q, int channel, float* out8) {
    if (!q || !out8 || channel < 0 || channel >= q->n_ch) {
        set_error("cqpsk_slicer_get_state: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    DSDNEO_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < 8; k++) {
        DSDNEO_CUDA(cudaMemcpy(out8 + k, q->d_thr + (size_t)k * q->n_ch + channel, sizeof(float), cudaMemcpyDeviceToHost));
    }
    return 0;
}

} /* extern "C" */
