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
 *   symbolize_kernel    time-serial, one thread per channel (lane = channel, SoA state => coalesced): everything with a
 *                       loop-carried dependence -- the +-1-sample jitter nudge, clip, window mean, the 128-symbol
 *                       extrema scan (kept in shared memory), the 1024-entry running means, slicing and LLRs.
 * Bit-exact with the reference (floats included); the SNR-dependent reliability weight uses the reference's value for
 * "no SNR hook installed" (w256 = 0, dsd_dibit.c:520-537).
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

constexpr int kMaxTaps = DSDNEO_B200_SYM_MAX_TAPS; /* 256 */
constexpr int kCarry = 96;                         /* leftover samples carried between launches (< longest symbol) */
constexpr int kSbuf = 128;
constexpr int kMinMax = 1024;
constexpr int kSymThreads = 32;
constexpr int kScanBlk = 16;  /* use_symbol's 128-entry extrema scan is kept as 8 block summaries + the block being replaced */
constexpr int kWin = 128;     /* samples per channel staged in shared memory per refill */

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
    float* carry;        /* [kCarry][n_ch] */
    float* sbuf;         /* [kSbuf][n_ch] */
    float* minbuf;       /* [kMinMax][n_ch] */
    float* maxbuf;       /* [kMinMax][n_ch] */
    float* symbols;      /* outputs, [n_ch][out_pitch] */
    uint8_t* dibits;
    uint8_t* reliab;
    int16_t* llr;        /* [n_ch][out_pitch][2] */
    int* count;          /* [n_ch] */
    size_t out_pitch;
    int n_ch, n, mode, have_sync, rate, symrate, ssize, msize;
    float2* minmax;      /* [n_ch][out_pitch] scratch: {min, max} after use_symbol per symbol (symbolize -> digitize kernel) */
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

__global__ void __launch_bounds__(kSymThreads)
symbolize_kernel(const SymParams p) {
    __shared__ float s_sbuf[kSbuf][kSymThreads];
    __shared__ float s_blk[kSbuf / kScanBlk][4][kSymThreads]; /* per 16-entry block of sbuf: two smallest, two largest */
    __shared__ float s_win[kWin * (kSymThreads + 1)];          /* staged samples, [sample][channel], padded rows */
    const int lane = threadIdx.x;
    const int ch = blockIdx.x * kSymThreads + lane;
    const bool valid = ch < p.n_ch;
    const int c = valid ? ch : p.n_ch - 1; /* inactive lanes shadow the last channel but never store */
    const int N = p.n_ch;

    for (int k = 0; k < kSbuf; k++) {
        s_sbuf[k][lane] = p.sbuf[(size_t)k * N + c];
    }
    /* state -> registers */
    const int window_l = p.s.window_l[c], track = p.s.track[c], negative = p.s.negative[c];
    int sps_num = p.s.sps_num[c], sps_den = p.s.sps_den[c], sps_accum = p.s.sps_accum[c];
    int sps = p.s.sps[c], center_idx = p.s.center_idx[c], jitter = p.s.jitter[c];
    float lastsample = p.s.lastsample[c];
    float vmin = p.s.vmin[c], vmax = p.s.vmax[c], center = p.s.center[c], umid = p.s.umid[c], lmid = p.s.lmid[c];
    float minref = p.s.minref[c], maxref = p.s.maxref[c];
    int sidx = p.s.sidx[c], midx = p.s.midx[c], sum_window = p.s.sum_window[c];
    double minbuf_sum = p.s.minbuf_sum[c], maxbuf_sum = p.s.maxbuf_sum[c];
    const int carry_n = p.s.carry_n[c];
    long long symbolcnt = p.s.symbolcnt[c];

    const float* filt = p.filt + (size_t)c * p.filt_pitch;
    const long avail = (long)carry_n + p.n;
    long pos = 0;
    auto sample_at = [&](long k) -> float { return k < carry_n ? p.carry[(size_t)k * N + c] : filt[k - carry_n]; };

    const int whole0 = p.rate / p.symrate;
    const long reserve = (long)(whole0 < 2 ? 2 : (whole0 > 64 ? 64 : whole0)) + 2; /* longest possible symbol */
    float* o_sym = p.symbols + (size_t)c * p.out_pitch;
    uint8_t* o_dib = p.dibits ? p.dibits + (size_t)c * p.out_pitch : nullptr;
    uint8_t* o_rel = p.reliab ? p.reliab + (size_t)c * p.out_pitch : nullptr;
    float2* o_mm = p.minmax ? p.minmax + (size_t)c * p.out_pitch : nullptr;
    int16_t* o_llr = p.llr ? p.llr + (size_t)c * p.out_pitch * 2 : nullptr;
    long nsym = 0;
    const int have_sync = (p.mode == DSDNEO_SYM_MODE_GET_DIBIT_SOFT) ? 1 : p.have_sync;

    const int cap = p.ssize < 0 ? 0 : (p.ssize > kSbuf ? kSbuf : p.ssize);
    const int n_blk = (cap + kScanBlk - 1) / kScanBlk;
    auto rescan_block = [&](int b) {
        float mn1 = 3.4028234663852886e38f, mn2 = mn1, mx1 = -mn1, mx2 = -mn1;
        const int k1 = min(cap, (b + 1) * kScanBlk);
        for (int k = b * kScanBlk; k < k1; k++) {
            const float v = s_sbuf[k][lane];
            two_min_push(mn1, mn2, v);
            two_max_push(mx1, mx2, v);
        }
        s_blk[b][0][lane] = mn1, s_blk[b][1][lane] = mn2, s_blk[b][2][lane] = mx1, s_blk[b][3][lane] = mx2;
    };
    if (track) {
        for (int b = 0; b < n_blk; b++) {
            rescan_block(b);
        }
    }

    /* Samples are staged through shared memory: the warp loads kWin consecutive samples of its 32 channels with
     * coalesced 128-byte reads and each lane then walks its own column.  Lanes drift apart only by the +-1 timing
     * nudges; a lane that falls outside the staged window reads global memory directly for that symbol. */
    long wbase = 0, wend = 0;
    int pf_idx = -1; /* prefetched minbuf / maxbuf ring entry */
    float pf_min = 0.0f, pf_max = 0.0f;
    int cn_max = carry_n, cn_min = carry_n; /* carry lengths over the warp: bounds for the branch-free refill */
    for (int o = 16; o > 0; o >>= 1) {
        cn_max = max(cn_max, __shfl_xor_sync(0xffffffffu, cn_max, o));
        cn_min = min(cn_min, __shfl_xor_sync(0xffffffffu, cn_min, o));
    }
    bool active = nsym < (long)p.out_pitch && (avail - pos) >= reserve;
    while (__any_sync(0xffffffffu, active)) {
        if (__any_sync(0xffffffffu, active && (pos < wbase || pos + reserve > wend))) {
            long mp = active ? pos : 0x7fffffffffffffffL;
            for (int o = 16; o > 0; o >>= 1) {
                const long other = __shfl_xor_sync(0xffffffffu, mp, o);
                mp = other < mp ? other : mp;
            }
            wbase = mp;
            wend = wbase + kWin;
            __syncwarp();
            const bool interior = wbase >= cn_max && wbase + kWin - cn_min <= (long)p.n;
            if (interior) { /* the whole window lies inside this launch's filtered samples for every channel row */
                const unsigned d0 = (unsigned)__cvta_generic_to_shared(&s_win[lane * (kSymThreads + 1)]);
                for (int r = 0; r < kSymThreads; r++) {
                    const int cr = min(blockIdx.x * kSymThreads + r, p.n_ch - 1);
                    const int cn = __shfl_sync(0xffffffffu, carry_n, r);
                    const float* src = p.filt + (size_t)cr * p.filt_pitch + (wbase - cn) + lane;
#pragma unroll
                    for (int g = 0; g < kWin / 32; g++) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 4u * (unsigned)(g * 32 * (kSymThreads + 1) + r)),
                                     "l"(src + g * 32)
                                     : "memory");
                    }
                }
            }
            for (int r = 0; r < (interior ? 0 : kSymThreads); r++) {
                const int cr = __shfl_sync(0xffffffffu, c, r);
                const int cn = __shfl_sync(0xffffffffu, carry_n, r);
                const float* fr = p.filt + (size_t)cr * p.filt_pitch;
#pragma unroll
                for (int g = 0; g < kWin / 32; g++) {
                    const long k = wbase + g * 32 + lane;
                    float* dst = &s_win[(g * 32 + lane) * (kSymThreads + 1) + r];
                    const float* src = nullptr;
                    if (k < cn) {
                        src = &p.carry[(size_t)k * N + cr];
                    } else if (k - cn < p.n) {
                        src = &fr[k - cn];
                    }
                    if (src) { /* all 128 copies of the refill are in flight before the single wait below */
                        const unsigned d32 = (unsigned)__cvta_generic_to_shared(dst);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d32), "l"(src) : "memory");
                    } else {
                        *dst = 0.0f;
                    }
                }
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncwarp();
        }
        if (active) {
        /* ---- symbol_apply_rtl_fsk_discriminator_timing (dsd_symbol.c:1328-1387) ---- */
        if (sps_num != p.rate || sps_den != p.symrate) {
            sps_num = p.rate;
            sps_den = p.symrate;
            sps_accum = 0;
            jitter = -1;
            center = 0.0f, vmin = -30000.0f, vmax = 30000.0f, lmid = -20000.0f, umid = 20000.0f;
            minref = -24000.0f, maxref = 24000.0f;
            for (int k = 0; k < kMinMax; k++) {
                if (valid) {
                    p.minbuf[(size_t)k * N + c] = vmin;
                    p.maxbuf[(size_t)k * N + c] = vmax;
                }
            }
            midx = 0;
            sum_window = 0;
            pf_idx = -1;
        }
        {
            int whole = p.rate / p.symrate, rem = p.rate % p.symrate;
            if (whole < 2) {
                whole = 2, rem = 0;
            }
            if (whole > 64) {
                whole = 64, rem = 0;
            }
            if (rem > 0 && sps_den > 0) {
                int acc = sps_accum + rem;
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
        const bool in_win = pos >= wbase && pos + sps + 1 <= wend; /* a nudge reads at most one extra sample */
        const float* wp = s_win + (in_win ? (int)(pos - wbase) : 0) * (kSymThreads + 1) + lane;
        /* Synchronised steady state: no timing nudge (have_sync), and the zero-crossing detector is idle once it has
         * fired (jitter >= 0 is only cleared by the nudge), so only the window samples and the symbol's last sample
         * (which becomes lastsample) matter; same clip, same accumulation order. */
        const bool quick = in_win && have_sync == 1 && jitter >= 0 && sps != 20 && sps != 5;
        if (quick) {
            const int lo = max(center_idx - window_l, 0), hi = min(center_idx + 2, sps - 1);
            for (int i = lo; i <= hi; i++) {
                float s = wp[i * (kSymThreads + 1)];
                s = s > vmax ? vmax : (s < vmin ? vmin : s);
                sum = __fadd_rn(sum, s);
                cnt++;
            }
            float s = wp[(sps - 1) * (kSymThreads + 1)];
            lastsample = s > vmax ? vmax : (s < vmin ? vmin : s);
            pos += sps;
        }
        for (int i = quick ? sps : 0; i < sps; i++) {
            if (i == 0 && have_sync == 0 && jitter >= 0) { /* dsd_symbol.c:462-516 */
                if (sps == 20) {
                    if (jitter >= 7 && jitter <= 10) {
                        i--;
                    } else if (jitter >= 11 && jitter <= 14) {
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
            float s = in_win ? *wp : sample_at(pos);
            wp += kSymThreads + 1;
            pos++;
            if (have_sync == 1) { /* symbol_apply_sync_clip, rf_mod == 0 */
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
            } else if (i >= center_idx - window_l && i <= center_idx + 2) {
                sum = __fadd_rn(sum, s);
                cnt++;
            }
            lastsample = s;
        }
        const float sym = cnt > 0 ? __fdiv_rn(sum, (float)cnt) : 0.0f;
        symbolcnt++;
        if (valid) {
            o_sym[nsym] = sym;
        }
        if (p.mode == DSDNEO_SYM_MODE_GET_DIBIT_SOFT) {
            /* ---- get_dibit_and_analog_signal: sbuf, use_symbol (dsd_dibit.c:243-299) ---- */
            s_sbuf[sidx][lane] = sym;
            if (track) {
                float lmin = 0.0f, lmax = 0.0f;
                if (cap >= 2) {
                    /* avg of the two smallest / two largest entries of sbuf[0..cap) (dsd_dibit.c:264-289): only the
                     * 16-entry block that received the new symbol is rescanned, the rest comes from block summaries */
                    if (sidx >= 0 && sidx < cap) {
                        rescan_block(sidx / kScanBlk);
                    }
                    float mn1 = s_blk[0][0][lane], mn2 = s_blk[0][1][lane], mx1 = s_blk[0][2][lane], mx2 = s_blk[0][3][lane];
                    for (int b = 1; b < n_blk; b++) {
                        const float a1 = s_blk[b][0][lane], a2 = s_blk[b][1][lane], z1 = s_blk[b][2][lane], z2 = s_blk[b][3][lane];
                        mn2 = fminf(fmaxf(mn1, a1), fminf(mn2, a2));
                        mn1 = fminf(mn1, a1);
                        mx2 = fmaxf(fminf(mx1, z1), fmaxf(mx2, z2));
                        mx1 = fmaxf(mx1, z1);
                    }
                    lmin = __fmul_rn(__fadd_rn(mn1, mn2), 0.5f);
                    lmax = __fmul_rn(__fadd_rn(mx1, mx2), 0.5f);
                }
                const int window = p.msize < 1 ? 1 : (p.msize > kMinMax ? kMinMax : p.msize);
                if (sum_window != window) { /* dsd_state_recompute_minmax_sums */
                    double a = 0.0, b = 0.0;
                    for (int k = 0; k < window; k++) {
                        a += (double)p.minbuf[(size_t)k * N + c];
                        b += (double)p.maxbuf[(size_t)k * N + c];
                    }
                    minbuf_sum = a, maxbuf_sum = b, sum_window = window;
                    if (midx < 0 || midx >= window) {
                        midx = 0;
                    }
                }
                int idx = (midx < 0 || midx >= window) ? 0 : midx;
                /* the ring entry being replaced was requested one symbol ago (pf_*), so its latency is off the chain */
                const float old_min = (pf_idx == idx) ? pf_min : p.minbuf[(size_t)idx * N + c];
                const float old_max = (pf_idx == idx) ? pf_max : p.maxbuf[(size_t)idx * N + c];
                minbuf_sum += (double)lmin - (double)old_min;
                maxbuf_sum += (double)lmax - (double)old_max;
                if (valid) {
                    p.minbuf[(size_t)idx * N + c] = lmin;
                    p.maxbuf[(size_t)idx * N + c] = lmax;
                }
                const int written = idx;
                idx++;
                midx = idx >= window ? 0 : idx;
                pf_idx = midx;
                if (midx == written) {
                    pf_min = lmin, pf_max = lmax;
                } else {
                    pf_min = p.minbuf[(size_t)midx * N + c];
                    pf_max = p.maxbuf[(size_t)midx * N + c];
                }
                if ((window & (window - 1)) == 0) { /* power of two (default 1024): x / w == x * (1 / w) exactly */
                    const double inv = 1.0 / (double)window;
                    vmin = (float)(minbuf_sum * inv);
                    vmax = (float)(maxbuf_sum * inv);
                } else {
                    vmin = (float)(minbuf_sum / (double)window);
                    vmax = (float)(maxbuf_sum / (double)window);
                }
                center = __fmul_rn(__fadd_rn(vmax, vmin), 0.5f); /* x / 2.0f, exact */
                umid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmax, center), 5.0f), 0.125f), center); /* .. / 8.0f, exact */
                lmid = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(vmin, center), 5.0f), 0.125f), center);
                maxref = __fmul_rn(vmax, 0.80f);
                minref = __fmul_rn(vmin, 0.80f);
            } else {
                maxref = vmax;
                minref = vmin;
            }
            if (cap > 0) {
                sidx = (sidx >= cap - 1) ? 0 : sidx + 1;
            }
            /* digitize and the soft metric only read {sym, thresholds}: sym_digitize_kernel does them one thread per symbol.
             * Tracked channels hand over {min, max} per symbol (centre / mid thresholds follow from them); the others keep
             * the same thresholds for the whole launch, which the second kernel reads from the carried state. */
            if (valid && track) {
                o_mm[nsym] = make_float2(vmin, vmax);
            }
        }
        nsym++;
        } /* active */
        active = nsym < (long)p.out_pitch && (avail - pos) >= reserve;
    }

    /* leftover samples -> carry (read everything first: source and destination overlap in the carry array) */
    const int left = (int)(avail - pos);
    float keep[kCarry / 8];
    for (int base = 0; base < left; base += kCarry / 8) {
        const int m = min(kCarry / 8, left - base);
        for (int k = 0; k < m; k++) {
            keep[k] = sample_at(pos + base + k);
        }
        if (valid) {
            for (int k = 0; k < m; k++) {
                p.carry[(size_t)(base + k) * N + c] = keep[k]; /* base + k < pos + base + k: never overtakes the reads */
            }
        }
    }
    if (!valid) {
        return;
    }
    for (int k = 0; k < kSbuf; k++) {
        p.sbuf[(size_t)k * N + c] = s_sbuf[k][lane];
    }
    p.s.sps_num[c] = sps_num, p.s.sps_den[c] = sps_den, p.s.sps_accum[c] = sps_accum;
    p.s.sps[c] = sps, p.s.center_idx[c] = center_idx, p.s.jitter[c] = jitter;
    p.s.lastsample[c] = lastsample;
    p.s.vmin[c] = vmin, p.s.vmax[c] = vmax, p.s.center[c] = center, p.s.umid[c] = umid, p.s.lmid[c] = lmid;
    p.s.minref[c] = minref, p.s.maxref[c] = maxref;
    p.s.sidx[c] = sidx, p.s.midx[c] = midx, p.s.sum_window[c] = sum_window;
    p.s.minbuf_sum[c] = minbuf_sum, p.s.maxbuf_sum[c] = maxbuf_sum;
    p.s.carry_n[c] = left;
    p.s.symbolcnt[c] = symbolcnt;
    p.count[c] = (int)nsym;
}

/* digitize (dsd_dibit.c:963-976,1018-1041) + compute_dibit_soft_metric (:644-721) + c4fm_reliability_from_thresholds
 * (:455-502) for every symbol of every channel, one thread per symbol: they only read the symbol and the thresholds in
 * force after use_symbol -- per symbol {min, max} from symbolize_kernel for tracked channels (centre and mid thresholds
 * follow from them by the reference's own formulas), the carried thresholds for the others (constant over a launch:
 * the only other writer is the first-call reset, which precedes the launch's first symbol). */
__global__ void __launch_bounds__(256)
sym_digitize_kernel(const SymParams p) {
    const int c = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.count[c]) {
        return;
    }
    const float sym = p.symbols[(size_t)c * p.out_pitch + i];
    const int negative = p.s.negative[c];
    float vmin, vmax, center, umid, lmid;
    if (p.s.track[c]) {
        const float2 mm = p.minmax[(size_t)c * p.out_pitch + i];
        vmin = mm.x;
        vmax = mm.y;
        cq_thresholds(vmin, vmax, center, umid, lmid);
    } else {
        vmin = p.s.vmin[c], vmax = p.s.vmax[c], center = p.s.center[c], umid = p.s.umid[c], lmid = p.s.lmid[c];
    }
    /* ---- digitize (dsd_dibit.c:963-976,1018-1041) ---- */
    int dibit;
    if (sym > center) {
        dibit = sym > umid ? (negative ? 3 : 1) : (negative ? 2 : 0);
    } else {
        dibit = sym < lmid ? (negative ? 1 : 3) : (negative ? 0 : 2);
    }
    /* ---- compute_dibit_soft_metric (dsd_dibit.c:644-721) ---- */
    const float plus_one = __fmul_rn(0.5f, __fadd_rn(center, umid)), minus_one = __fmul_rn(0.5f, __fadd_rn(lmid, center));
    float ideal[4];
    if (negative) {
        ideal[0] = minus_one, ideal[1] = vmin, ideal[2] = plus_one, ideal[3] = vmax;
    } else {
        ideal[0] = plus_one, ideal[1] = vmax, ideal[2] = minus_one, ideal[3] = vmin;
    }
    int mag0, mag1;
    bit_metrics(sym, ideal, mag0, mag1);
    /* c4fm_reliability_from_thresholds (dsd_dibit.c:455-502) */
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
    rel = clamp255((rel * 204) >> 8); /* apply_c4fm_snr_weight with no SNR hook: w256 = 0 (dsd_dibit.c:520-537) */
    const int min_mag = mag0 < mag1 ? mag0 : mag1;
    if (min_mag > 0 && rel < min_mag) {
        mag0 = (mag0 * rel) / min_mag;
        mag1 = (mag1 * rel) / min_mag;
    }
    mag0 = clamp255(mag0);
    mag1 = clamp255(mag1);
    const int l0 = ((dibit >> 1) & 1) ? mag0 : -mag0, l1 = (dibit & 1) ? mag1 : -mag1;
    const size_t o = (size_t)c * p.out_pitch + i;
    p.dibits[o] = (uint8_t)dibit;
    p.reliab[o] = (uint8_t)clamp255(mag1 < mag0 ? mag1 : mag0);
    reinterpret_cast<short2*>(p.llr)[o] = make_short2((short)l0, (short)l1);
}

__global__ void
sym_reset_kernel(SymScalars s, float* minbuf, float* maxbuf, float* sbuf, float* carry, float* hist, int n_ch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) {
        return;
    }
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
    for (int k = 0; k < kMinMax; k++) {
        minbuf[(size_t)k * n_ch + c] = -15000.0f;
        maxbuf[(size_t)k * n_ch + c] = 15000.0f;
    }
    for (int k = 0; k < kSbuf; k++) {
        sbuf[(size_t)k * n_ch + c] = 0.0f;
    }
    for (int k = 0; k < kCarry; k++) {
        carry[(size_t)k * n_ch + c] = 0.0f;
    }
    for (int k = 0; k < kMaxTaps; k++) {
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
 * One lane per channel.  The symbol value does not feed back into anything but the tracker, so the per-channel chain is
 * the 128-entry extrema scan (kept in shared memory, [entry][lane]) and the two f64 running sums. */
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
    float* sbuf;          /* [128][n_ch] */
    float* minbuf;        /* [1024][n_ch] */
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

__global__ void __launch_bounds__(32)
cqpsk_slice_kernel(const CqSliceParams p) {
    __shared__ float s_sbuf[128 * 32];
    const int lane = threadIdx.x;
    const int ch = blockIdx.x * 32 + lane;
    const bool valid = ch < p.n_ch;
    const int n_ch = p.n_ch;
    int cap = p.ssize < 0 ? 0 : (p.ssize > 128 ? 128 : p.ssize);
    const int window = p.msize < 1 ? 1 : (p.msize > 1024 ? 1024 : p.msize);
    for (int k = 0; k < 128; k++) {
        s_sbuf[k * 32 + lane] = valid ? p.sbuf[(size_t)k * n_ch + ch] : 0.0f;
    }
    int n = 0, sidx = 0, midx = 0;
    double min_sum = 0.0, max_sum = 0.0;
    float vmin = 0.0f, vmax = 0.0f, center = 0.0f, umid = 0.0f, lmid = 0.0f, minref = 0.0f, maxref = 0.0f, last = 0.0f;
    if (valid) {
        n = p.n_symbols[ch];
        sidx = p.sidx[ch];
        midx = p.midx[ch];
        min_sum = p.minbuf_sum[ch];
        max_sum = p.maxbuf_sum[ch];
        vmin = p.thr[0 * (size_t)n_ch + ch];
        vmax = p.thr[1 * (size_t)n_ch + ch];
        center = p.thr[2 * (size_t)n_ch + ch];
        umid = p.thr[3 * (size_t)n_ch + ch];
        lmid = p.thr[4 * (size_t)n_ch + ch];
        minref = p.thr[5 * (size_t)n_ch + ch];
        maxref = p.thr[6 * (size_t)n_ch + ch];
        last = p.thr[7 * (size_t)n_ch + ch];
        if (n > 0 && p.sum_window[ch] != window) { /* dsd_state_sync_minmax_sums, core/state.h:1388-1428 */
            double a = 0.0, b = 0.0;
            for (int i = 0; i < window; i++) {
                a += (double)p.minbuf[(size_t)i * n_ch + ch];
                b += (double)p.maxbuf[(size_t)i * n_ch + ch];
            }
            min_sum = a;
            max_sum = b;
            p.sum_window[ch] = window;
            if (midx < 0 || midx >= window) {
                midx = 0;
            }
        }
    }
    const float* in = p.symbols + (size_t)(valid ? ch : 0) * p.sym_pitch;
    float2* mm = p.minmax + (size_t)(valid ? ch : 0) * p.out_pitch;
    float* my_sbuf = s_sbuf + lane;
    /* Fast path of the extrema scan (full 128-entry window): eight 16-entry block summaries {two smallest, two largest} in
     * registers; a new symbol only invalidates its own block, which is rescanned (two independent half chains), and the
     * eight summaries are merged pairwise (depth 3).  The two smallest / largest of a multiset do not depend on the order of
     * evaluation, so this is the reference's result exactly. */
    const bool blocked = (cap == 128);
    float bmn1[8], bmn2[8], bmx1[8], bmx2[8];
    auto scan_block = [&](int b, float& mn1, float& mn2, float& mx1, float& mx2) {
        const float* e = my_sbuf + b * 16 * 32;
        const float a0 = e[0], a1 = e[32], c0 = e[8 * 32], c1 = e[9 * 32];
        float p1 = fminf(a0, a1), p2 = fmaxf(a0, a1), q1 = fminf(c0, c1), q2 = fmaxf(c0, c1);
        float P1 = p2, P2 = p1, Q1 = q2, Q2 = q1;
#pragma unroll
        for (int k = 2; k < 8; k++) {
            const float v = e[k * 32], w = e[(8 + k) * 32];
            two_min_push(p1, p2, v);
            two_max_push(P1, P2, v);
            two_min_push(q1, q2, w);
            two_max_push(Q1, Q2, w);
        }
        mn1 = fminf(p1, q1);
        mn2 = fminf(fmaxf(p1, q1), fminf(p2, q2));
        mx1 = fmaxf(P1, Q1);
        mx2 = fmaxf(fminf(P1, Q1), fmaxf(P2, Q2));
    };
    if (blocked) {
#pragma unroll
        for (int b = 0; b < 8; b++) {
            scan_block(b, bmn1[b], bmn2[b], bmx1[b], bmx2[b]);
        }
    }
    const bool pow2_window = (window & (window - 1)) == 0;
    const double inv_window = 1.0 / (double)window; /* exact for a power of two: x / 2^k == x * 2^-k */
    /* the ring entries the next symbol replaces and the next symbol itself are requested one iteration early */
    const bool prefetch = window >= 2;
    float nxt_min = 0.0f, nxt_max = 0.0f, nxt_sym = 0.0f;
    if (valid && n > 0) {
        const int idx0 = (midx < 0 || midx >= window) ? 0 : midx;
        nxt_min = p.minbuf[(size_t)idx0 * n_ch + ch];
        nxt_max = p.maxbuf[(size_t)idx0 * n_ch + ch];
        nxt_sym = in[0];
    }
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        const float sym = nxt_sym;
        if (i + 1 < n) {
            nxt_sym = in[i + 1];
        }
        last = sym;
        if (cap > 0) {
            my_sbuf[sidx * 32] = sym; /* cap <= 0: the reference writes sbuf[sidx] with sidx stuck at its initial 0 */
        } else {
            my_sbuf[0] = sym;
        }
        /* use_symbol: average of the two smallest / two largest of sbuf[0..cap) (order independent, so exact) */
        float lmin = 0.0f, lmax = 0.0f;
        if (blocked) {
            const int b = sidx >> 4;
            float r1, r2, R1, R2;
            scan_block(b, r1, r2, R1, R2);
#pragma unroll
            for (int k = 0; k < 8; k++) { /* static indices keep the summaries in registers */
                if (k == b) {
                    bmn1[k] = r1;
                    bmn2[k] = r2;
                    bmx1[k] = R1;
                    bmx2[k] = R2;
                }
            }
            float t1[4], t2[4], T1[4], T2[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                t1[k] = fminf(bmn1[2 * k], bmn1[2 * k + 1]);
                t2[k] = fminf(fmaxf(bmn1[2 * k], bmn1[2 * k + 1]), fminf(bmn2[2 * k], bmn2[2 * k + 1]));
                T1[k] = fmaxf(bmx1[2 * k], bmx1[2 * k + 1]);
                T2[k] = fmaxf(fminf(bmx1[2 * k], bmx1[2 * k + 1]), fmaxf(bmx2[2 * k], bmx2[2 * k + 1]));
            }
            const float u1a = fminf(t1[0], t1[1]), u2a = fminf(fmaxf(t1[0], t1[1]), fminf(t2[0], t2[1]));
            const float u1b = fminf(t1[2], t1[3]), u2b = fminf(fmaxf(t1[2], t1[3]), fminf(t2[2], t2[3]));
            const float U1a = fmaxf(T1[0], T1[1]), U2a = fmaxf(fminf(T1[0], T1[1]), fmaxf(T2[0], T2[1]));
            const float U1b = fmaxf(T1[2], T1[3]), U2b = fmaxf(fminf(T1[2], T1[3]), fmaxf(T2[2], T2[3]));
            const float mn1 = fminf(u1a, u1b), mn2 = fminf(fmaxf(u1a, u1b), fminf(u2a, u2b));
            const float mx1 = fmaxf(U1a, U1b), mx2 = fmaxf(fminf(U1a, U1b), fmaxf(U2a, U2b));
            lmin = __fmul_rn(__fadd_rn(mn1, mn2), 0.5f);
            lmax = __fmul_rn(__fadd_rn(mx1, mx2), 0.5f);
        } else if (cap >= 2) {
            const float a = my_sbuf[0], b = my_sbuf[32];
            float mn1 = fminf(a, b), mn2 = fmaxf(a, b), mx1 = mn2, mx2 = mn1;
#pragma unroll 8
            for (int k = 2; k < cap; k++) {
                const float v = my_sbuf[k * 32];
                two_min_push(mn1, mn2, v);
                two_max_push(mx1, mx2, v);
            }
            lmin = __fmul_rn(__fadd_rn(mn1, mn2), 0.5f);
            lmax = __fmul_rn(__fadd_rn(mx1, mx2), 0.5f);
        }
        {
            int idx = midx;
            if (idx < 0 || idx >= window) {
                idx = 0;
            }
            float* mb = p.minbuf + (size_t)idx * n_ch + ch;
            float* xb = p.maxbuf + (size_t)idx * n_ch + ch;
            const float old_min = prefetch ? nxt_min : *mb, old_max = prefetch ? nxt_max : *xb;
            min_sum += (double)lmin - (double)old_min;
            max_sum += (double)lmax - (double)old_max;
            *mb = lmin;
            *xb = lmax;
            idx++;
            midx = idx >= window ? 0 : idx;
            if (prefetch && i + 1 < n) {
                nxt_min = p.minbuf[(size_t)midx * n_ch + ch];
                nxt_max = p.maxbuf[(size_t)midx * n_ch + ch];
            }
        }
        if (pow2_window) {
            vmin = (float)(min_sum * inv_window);
            vmax = (float)(max_sum * inv_window);
        } else {
            vmin = (float)(min_sum / (double)window);
            vmax = (float)(max_sum / (double)window);
        }
        if (cap > 0) {
            sidx = (sidx >= cap - 1) ? 0 : sidx + 1;
        }
        /* everything else of the symbol (thresholds, slicing, soft metric) only depends on {sym, min, max}: it is done by
         * cqpsk_digitize_kernel, one thread per symbol */
        mm[i] = make_float2(vmin, vmax);
    }
    if (n > 0) {
        cq_thresholds(vmin, vmax, center, umid, lmid);
        maxref = __fmul_rn(vmax, 0.80f);
        minref = __fmul_rn(vmin, 0.80f);
    }
    if (valid) {
        for (int k = 0; k < 128; k++) {
            p.sbuf[(size_t)k * n_ch + ch] = s_sbuf[k * 32 + lane];
        }
        p.sidx[ch] = sidx;
        p.midx[ch] = midx;
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
        sbuf[(size_t)k * n_ch + ch] = 0.0f;
    }
    for (int k = 0; k < 1024; k++) { /* initState, src/core/util/dsd_init.c:519-592 */
        minbuf[(size_t)k * n_ch + ch] = -15000.0f;
        maxbuf[(size_t)k * n_ch + ch] = 15000.0f;
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
    float* d_filt;
    size_t filt_pitch;
    float2* d_minmax; /* per-symbol {min, max} between symbolize_kernel and sym_digitize_kernel, grown on demand */
    size_t minmax_cap;
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
    cudaFree(y->d_filt);
    cudaFree(y->d_minmax);
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
    /* scalar arena: 22 x 4-byte arrays, 3 x 8-byte arrays */
    const size_t arena_bytes = n * (22 * 4 + 3 * 8) + 256;
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
        cls[i].filter = DSDNEO_SYM_FILTER_NONE, cls[i].window_l = 2, cls[i].track_minmax = 0, cls[i].negative = 0;
    }
    int rc = dsdneo_b200_symbolizer_set_class(y, cls);
    free(cls);
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
    sym_reset_kernel<<<(y->n_ch + 127) / 128, 128, 0, as_stream(stream)>>>(y->s, y->d_minbuf, y->d_maxbuf, y->d_sbuf, y->d_carry,
                                                                            y->d_hist, y->n_ch);
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_symbolizer_set_class(dsdneo_b200_symbolizer* y, const dsdneo_b200_sym_class* per_channel) {
    if (!y || !per_channel) {
        set_error("symbolizer_set_class: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    const size_t n = (size_t)y->n_ch;
    int* h = (int*)malloc(4 * n * sizeof(int));
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
    free(h);
    if (e != cudaSuccess) {
        return cuda_fail(e, "symbolizer_set_class", __FILE__, __LINE__);
    }
    return 0;
}

int
dsdneo_b200_symbolize_batch(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int mode, int have_sync,
                            const dsdneo_b200_symbol_out* out, void* stream) {
    if (!y || !d_disc || !out || n_samples < 0 || disc_pitch < (size_t)n_samples || !out->d_symbols || !out->d_count
        || out->pitch == 0) {
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
    {
        int whole = y->rate / y->symrate;
        whole = whole < 2 ? 2 : (whole > 64 ? 64 : whole);
        const size_t need = ((size_t)n_samples + kCarry) / (size_t)(whole - 1) + 2;
        if (out->pitch < need) {
            set_error("symbolize_batch: output pitch %zu is too small for %d samples (need >= %zu)", out->pitch, n_samples, need);
            return DSDNEO_B200_EINVAL;
        }
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    const size_t pitch = ((size_t)n_samples + 3) & ~(size_t)3;
    if (!y->d_filt || y->filt_pitch < pitch) {
        DSDNEO_CUDA(cudaDeviceSynchronize());
        cudaFree(y->d_filt);
        y->d_filt = NULL;
        DSDNEO_CUDA(cudaMalloc((void**)&y->d_filt, (size_t)y->n_ch * (pitch ? pitch : 4) * sizeof(float)));
        y->filt_pitch = pitch ? pitch : 4;
    }
    if (n_samples > 0) {
        FirParams fp;
        fp.in = d_disc;
        fp.in_pitch = disc_pitch;
        fp.out = y->d_filt;
        fp.out_pitch = y->filt_pitch;
        fp.taps = y->d_taps;
        fp.taps_len = y->d_taps_len;
        fp.filter = y->s.filter;
        fp.hist = y->d_hist;
        fp.n = n_samples;
        dim3 grid((unsigned)((n_samples + kFirTile - 1) / kFirTile), (unsigned)y->n_ch);
        {
            KernelTimer kt("sps_fir_kernel", s);
            sps_fir_kernel<<<grid, kFirThreads, 0, s>>>(fp);
        }
        DSDNEO_KERNEL_CHECK();
        count_launch();
        {
            KernelTimer kt("sps_fir_hist_kernel", s);
            sps_fir_hist_kernel<<<(y->n_ch + 7) / 8, 256, 0, s>>>(d_disc, disc_pitch, y->d_hist, y->s.filter, y->d_taps_len, y->n_ch,
                                                                 n_samples);
        }
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    SymParams sp;
    sp.s = y->s;
    sp.filt = y->d_filt;
    sp.filt_pitch = y->filt_pitch;
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
    sp.minmax = NULL;
    if (mode == DSDNEO_SYM_MODE_GET_DIBIT_SOFT) {
        const size_t need = (size_t)y->n_ch * out->pitch;
        if (y->minmax_cap < need) {
            DSDNEO_CUDA(cudaDeviceSynchronize());
            cudaFree(y->d_minmax);
            y->d_minmax = NULL;
            y->minmax_cap = 0;
            DSDNEO_CUDA(cudaMalloc((void**)&y->d_minmax, need * sizeof(float2)));
            y->minmax_cap = need;
        }
        sp.minmax = y->d_minmax;
    }
    {
        KernelTimer kt("symbolize_kernel", s);
        symbolize_kernel<<<(y->n_ch + kSymThreads - 1) / kSymThreads, kSymThreads, 0, s>>>(sp);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    if (mode == DSDNEO_SYM_MODE_GET_DIBIT_SOFT && out->pitch > 0) {
        KernelTimer kt("sym_digitize_kernel", s);
        dim3 grid((unsigned)((out->pitch + 255) / 256), (unsigned)y->n_ch);
        sym_digitize_kernel<<<grid, 256, 0, s>>>(sp);
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    return 0;
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
        cqpsk_slice_kernel<<<(q->n_ch + 31) / 32, 32, 0, s>>>(p);
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
