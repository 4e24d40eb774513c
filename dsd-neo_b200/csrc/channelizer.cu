// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * K2 (+K1): polyphase FIR channelizer, wideband complex IQ -> M narrowband channels.
 *
 * The reference has no channelizer (it tunes one channel with half-band decimators,
 * src/dsp/demod_pipeline.cpp:983-1001); this stage is the north-star's addition and its oracle is the
 * mathematical direct form (oracle/oracle_dsp.c:oracle_pfb_direct, float64):
 *
 *     y_k[n] = sum_{m=0}^{L-1} h[m] x[t_n - m] exp(-j 2 pi k (t_n - m) / M),   t_n = n M + M - 1,  L = T M
 *
 * i.e. mix channel k (centre k fs/M) to DC, low-pass with the prototype h, keep every M-th sample
 * (critically sampled).  Polyphase form used here: with x split into blocks of M,
 *     v_r[n] = sum_{q=0}^{T-1} h[q M + M-1-r] x[(n-q) M + r],      y_k[n] = sum_r v_r[n] exp(-j 2 pi k r / M)
 * so each input sample is touched by exactly one thread (r), T times, out of a register window, and the
 * channel outputs are an M-point forward DFT per output time.
 *
 * Optional fused K1: cu8 input is widened on load exactly like widen_u8_to_f32_bias127
 * (src/dsp/simd_widen.cpp:139-147): (float(u8) - 127.5f) * (1.0f / 127.5f).
 *
 * Kernel (M = 256): one CTA = 256 threads (thread r = branch r) walks a range of output times in chunks of
 * 16.  Per chunk: 16 coalesced float2 loads per thread (all issued before use), T-tap branch FIR from the
 * register window, 16 x 256-point FFT done as two radix-16 passes through padded shared memory (bank-conflict
 * free), then a shared-memory transpose so every channel receives 16 consecutive outputs = one full 128-byte
 * line per store.  HBM traffic is the algorithmic minimum: 8 B in + 8 B out per input sample.
 *
 * General kernel (M = 512 ... 8192, and bin-pruned outputs for multi-GPU channel sharding), pfbn_kernel:
 * a rank that owns the channels k = r0 (mod R) needs only those DFT bins.  One decimation-in-frequency step
 * taken analytically prunes the transform to N = M / R points:
 *     y_{R k' + r0}[n] = sum_{n'=0}^{N-1} W_N^{n' k'} * ( W_M^{n' r0} * sum_{j=0}^{R-1} W_R^{j r0} v_{n' + j N}[n] )
 * so every rank still reads the whole wideband tile and runs all M branch filters (8 + 8/R bytes per input
 * sample of HBM traffic), but its transform and its output shrink by R.  One CTA walks chunks of C output
 * times: branch FIRs + fold from a sliding register window straight into padded shared memory, an in-place
 * mixed-radix FFT (one radix-2/4/8 pass, then radix-16 passes, twiddles from a W_N table), and a transposed
 * store so each channel receives C consecutive outputs (C * 8 contiguous bytes).  R = 1 is the plain M-channel case.
 *
 * Compiled with -fmad=true (nothing here is bit-pinned to a CPU path; tolerance is stated in the tests).
 */
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

constexpr int kM = 256;
constexpr int kChunk = 16;                 /* output times per chunk */
constexpr int kFftPitch = kM + kM / 16;    /* 272 float2: element i lives at i + (i >> 4) */
constexpr int kOutPitch = kChunk + 1;      /* 17 float2 per channel row in the transpose buffer */
constexpr int kMaxT = 16;

struct PfbParams {
    const void* in;        /* cf32 (float2) or cu8 (uchar2), n_out * M samples */
    const float2* hist;    /* (T-1) * M samples preceding `in` */
    const float* proto;    /* L = T * M prototype taps */
    float2* out;           /* [M][out_pitch] */
    size_t out_pitch;
    long n_out;            /* output times in this launch */
    int chunks_per_cta;
};

__device__ __forceinline__ float2
cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

#ifndef PFB_PACKED_ADD
#define PFB_PACKED_ADD 1
#endif
__device__ __forceinline__ float2
cadd(float2 a, float2 b) {
#if PFB_PACKED_ADD
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}

__device__ __forceinline__ float2
csub(float2 a, float2 b) {
#if PFB_PACKED_ADD
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); /* exact: a - b */
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}

__device__ __forceinline__ float2
mul_mj(float2 a) { /* a * (-j) */
    return make_float2(a.y, -a.x);
}

/* forward 4-point DFT, in place: W4 = -j */
__device__ __forceinline__ void
dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s02 = cadd(x0, x2), d02 = csub(x0, x2);
    const float2 s13 = cadd(x1, x3), d13 = mul_mj(csub(x1, x3));
    x0 = cadd(s02, s13);
    x2 = csub(s02, s13);
    x1 = cadd(d02, d13);
    x3 = csub(d02, d13);
}

/* forward 16-point DFT: in natural order a[n], out natural order a[k] */
__device__ __forceinline__ void
dft16(float2 (&a)[16]) {
    /* n = 4 p + b, k = c + 4 d:  X[c+4d] = sum_b W4^{bd} ( W16^{bc} sum_p x[4p+b] W4^{pc} ) */
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        dft4(a[b], a[4 + b], a[8 + b], a[12 + b]); /* a[4c + b] = t[b][c] */
    }
    /* twiddles W16^{bc} = exp(-j 2 pi bc / 16) */
    a[4 * 1 + 1] = cmul(a[4 * 1 + 1], make_float2(c1, -s1)); /* bc = 1 */
    a[4 * 2 + 1] = cmul(a[4 * 2 + 1], make_float2(h, -h));   /* 2 */
    a[4 * 3 + 1] = cmul(a[4 * 3 + 1], make_float2(s1, -c1)); /* 3 */
    a[4 * 1 + 2] = cmul(a[4 * 1 + 2], make_float2(h, -h));   /* 2 */
    a[4 * 2 + 2] = mul_mj(a[4 * 2 + 2]);                     /* 4 */
    a[4 * 3 + 2] = cmul(a[4 * 3 + 2], make_float2(-h, -h));  /* 6 */
    a[4 * 1 + 3] = cmul(a[4 * 1 + 3], make_float2(s1, -c1)); /* 3 */
    a[4 * 2 + 3] = cmul(a[4 * 2 + 3], make_float2(-h, -h));  /* 6 */
    a[4 * 3 + 3] = cmul(a[4 * 3 + 3], make_float2(-c1, s1)); /* 9 */
#pragma unroll
    for (int c = 0; c < 4; c++) {
        dft4(a[4 * c + 0], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]); /* a[4c + d] = X[c + 4d] */
    }
    /* a[4c+d] holds X[c+4d]: transpose the 4x4 index to natural order */
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int d = c + 1; d < 4; d++) {
            const float2 t = a[4 * c + d];
            a[4 * c + d] = a[4 * d + c];
            a[4 * d + c] = t;
        }
    }
}

template <bool CU8>
__device__ __forceinline__ float2
load_sample(const void* base, long idx) {
    if (CU8) {
        const uchar2 u = __ldg(reinterpret_cast<const uchar2*>(base) + idx);
        const float inv = 1.0f / 127.5f;
        return make_float2(__fmul_rn(__fsub_rn((float)u.x, 127.5f), inv), __fmul_rn(__fsub_rn((float)u.y, 127.5f), inv));
    } else {
        return __ldg(reinterpret_cast<const float2*>(base) + idx);
    }
}

template <int T, bool CU8>
__global__ void __launch_bounds__(kM, 2)
pfb256_kernel(const PfbParams p) {
    extern __shared__ __align__(16) unsigned char pfb_smem[];
    float2* V = reinterpret_cast<float2*>(pfb_smem);       /* [kChunk][kFftPitch] */
    float2* O = V + kChunk * kFftPitch;                    /* [kM][kOutPitch] */

    const int r = threadIdx.x;
    const int j = r & 15;  /* index inside a radix-16 pass */
    const int f = r >> 4;  /* which of the 16 FFTs of the chunk */

    /* per-thread constants: branch taps and the pass-1 twiddles W256^{j b} */
    float g[T];
#pragma unroll
    for (int q = 0; q < T; q++) {
        g[q] = p.proto[q * kM + (kM - 1 - r)];
    }
    float2 tw[16];
#pragma unroll
    for (int b = 0; b < 16; b++) {
        float sn, cs;
        sincospif(-(float)(j * b) / 128.0f, &sn, &cs);
        tw[b] = make_float2(cs, sn);
    }

    const long chunk0 = (long)blockIdx.x * p.chunks_per_cta;
    const long n_begin = chunk0 * kChunk;
    if (n_begin >= p.n_out) {
        return;
    }

    /* register window: xs[T-1+i] = x[(n0+i) M + r]; xs[0..T-2] = the T-1 older blocks */
    float2 xs[T - 1 + kChunk];
#pragma unroll
    for (int q = 1; q < T; q++) {
        const long blk = n_begin - q;
        xs[T - 1 - q] = (blk >= 0) ? load_sample<CU8>(p.in, blk * kM + r) : p.hist[(long)(T - 1 + blk) * kM + r];
    }

    for (int c = 0; c < p.chunks_per_cta; c++) {
        const long n0 = n_begin + (long)c * kChunk;
        if (n0 >= p.n_out) {
            break;
        }
        const int nv = (int)min((long)kChunk, p.n_out - n0);
        /* ---- branch FIR: 16 independent loads, then T MACs per output ---- */
#pragma unroll
        for (int i = 0; i < kChunk; i++) {
            xs[T - 1 + i] = (i < nv) ? load_sample<CU8>(p.in, (n0 + i) * kM + r) : make_float2(0.0f, 0.0f);
        }
        const int pr = r + (r >> 4);
#pragma unroll
        for (int i = 0; i < kChunk; i++) {
            float2 v = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int q = 0; q < T; q++) {
                v.x = fmaf(g[q], xs[T - 1 + i - q].x, v.x);
                v.y = fmaf(g[q], xs[T - 1 + i - q].y, v.y);
            }
            V[i * kFftPitch + pr] = v;
        }
#pragma unroll
        for (int q = 0; q < T - 1; q++) {
            xs[q] = xs[kChunk + q];
        }
        __syncthreads();

        /* ---- FFT pass 1 (in place): thread (f, j) transforms elements j + 16 a, a = 0..15 ---- */
        float2 a[16];
        float2* Vf = V + f * kFftPitch;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a[k] = Vf[j + 17 * k]; /* (j + 16k) + ((j + 16k) >> 4) */
        }
        dft16(a);
#pragma unroll
        for (int b = 1; b < 16; b++) {
            a[b] = cmul(a[b], tw[b]);
        }
#pragma unroll
        for (int b = 0; b < 16; b++) {
            Vf[17 * b + j] = a[b]; /* Z[b][j] at (16 b + j) + b */
        }
        __syncthreads();

        /* ---- FFT pass 2: thread (f, b = j) transforms Z[b][0..15]; X[b + 16 c] ---- */
#pragma unroll
        for (int k = 0; k < 16; k++) {
            a[k] = Vf[17 * j + k];
        }
        dft16(a);
#pragma unroll
        for (int cidx = 0; cidx < 16; cidx++) {
            O[(j + 16 * cidx) * kOutPitch + f] = a[cidx]; /* channel k = b + 16 c, time f */
        }
        __syncthreads();

        /* ---- transpose out: half-warp = one channel, 16 consecutive times = one 128-byte line ---- */
        {
            const int t = r & 15;
            const int kbase = r >> 4; /* 0..15 */
            if (t < nv) {
#pragma unroll
                for (int it = 0; it < kM / 16; it++) {
                    const int k = kbase + 16 * it;
                    __stcs(&p.out[(size_t)k * p.out_pitch + n0 + t], O[k * kOutPitch + t]);
                }
            }
        }
        /* O is rewritten only after the next chunk's two barriers; V after this barrier-separated phase */
    }
}


/* ---------------------------------------------------------------------------------------------------------
 * General M, pruned bins.
 * ------------------------------------------------------------------------------------------------------- */
constexpr int kMaxFold = 32; /* R <= 32 */
constexpr int kTmaBox = 256;  /* inner box width of the tensor-map copies (elements; the hardware limit per dimension) */

__device__ __forceinline__ void
pfb_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void
pfb_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
pfb_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PFB_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni PFB_WAIT_DONE;\n"
        "bra.uni PFB_WAIT_LOOP;\n"
        "PFB_WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

/* one 2-D tensor-map copy (TMA): box {kTmaBox x rows} of 32-bit elements at element column c0, row c1 -> dense shared memory */
__device__ __forceinline__ void
pfb_tma_load_2d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}

struct PfbNParams {
    const void* in;       /* cf32 / cu8, n_out * M samples */
    const float2* hist;   /* (T-1) * M samples preceding `in` */
    const float* proto;   /* T * M prototype taps (cu8 input: pre-scaled by 1/127.5) */
    const float* bias;    /* cu8 input: -127.5 * sum_q proto[q M + M-1-b] per branch b (the widening's offset after the filter) */
    const float2* twN;    /* W_N^m = exp(-j 2 pi m / N), m < N */
    float2* out;          /* [N][out_pitch]: row k' = channel R k' + r0 */
    uchar2* out_u8;       /* instead of `out`: the same rows re-quantised to cu8, round(v * out_gain * 127.5 + 127.5) clamped */
    float out_gain;
    size_t out_pitch;
    long n_out;
    int M, N, R, r0;
    int C;                /* output times per chunk: 2, 4, 8 or 16 */
    int rho;              /* radix of the first pass: 1 (none), 2, 4 or 8; N / rho is a power of 16 */
    int lgN, lgC;
    int use_tma;          /* cu8 pair path: the step's window is ONE tensor-map copy per 256 pairs instead of 15 cp.async per thread */
    int chunks_per_cta;
    float2 wR[kMaxFold];  /* W_R^{j r0} */
};

__device__ __forceinline__ int
pad16(int i) {
    return i + (i >> 4);
}

/* a carried history sample (kept widened) back in the byte domain: (k - 127.5) / 127.5 re-expands to exactly k; the reset
 * state 0.0 (never a widened byte: k - 127.5 is a half-integer) stays 127.5 */
__device__ __forceinline__ float
hist_byte(float x) {
    return x == 0.0f ? 127.5f : rintf(fmaf(x, 127.5f, 127.5f));
}

template <int RHO>
__device__ __forceinline__ void
dft_small(float2 (&a)[8]) {
    if (RHO == 2) {
        const float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    } else if (RHO == 4) {
        dft4(a[0], a[1], a[2], a[3]);
    } else {
        constexpr float h = 0.70710678118654752f;
        dft4(a[0], a[2], a[4], a[6]); /* even part: a[0], a[2], a[4], a[6] = E[0..3] */
        dft4(a[1], a[3], a[5], a[7]); /* odd part:  a[1], a[3], a[5], a[7] = O[0..3] */
        const float2 o0 = a[1], o1 = cmul(a[3], make_float2(h, -h)), o2 = mul_mj(a[5]), o3 = cmul(a[7], make_float2(-h, -h));
        const float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
        a[0] = cadd(e0, o0);
        a[1] = cadd(e1, o1);
        a[2] = cadd(e2, o2);
        a[3] = cadd(e3, o3);
        a[4] = csub(e0, o0);
        a[5] = csub(e1, o1);
        a[6] = csub(e2, o2);
        a[7] = csub(e3, o3);
    }
}

/* first pass of the in-place DIF transform: radix RHO over the whole row (S = N).  A thread keeps one butterfly
 * column (its twiddles stay in registers) and walks the rows of the chunk. */
template <int RHO, int NT>
__device__ __forceinline__ void
pfbn_first_pass(float2* X, const PfbNParams& p, int pitchT) {
    const int sub = p.N / RHO;
    const int U = min(sub, NT), rows_par = NT / U;
    for (int j = threadIdx.x % U; j < sub; j += U) {
        float2 tw[RHO];
#pragma unroll
        for (int k = 1; k < RHO; k++) {
            tw[k] = __ldg(p.twN + j * k);
        }
        const int pstride = sub + (sub >> 4); /* sub is a multiple of 16 */
        for (int row = threadIdx.x / U; row < p.C; row += rows_par) {
            float2* Xr = X + row * pitchT + pad16(j);
            float2 a[8];
#pragma unroll
            for (int q = 0; q < RHO; q++) {
                a[q] = Xr[q * pstride];
            }
            dft_small<RHO>(a);
#pragma unroll
            for (int k = 1; k < RHO; k++) {
                a[k] = cmul(a[k], tw[k]);
            }
#pragma unroll
            for (int k = 0; k < RHO; k++) {
                Xr[k * pstride] = a[k];
            }
        }
    }
    __syncthreads();
}

/* SC = output times per register window (4 or 8): the T - 1 + SC samples of one branch are requested together, then
 * filtered out of registers. */
template <int T, bool CU8, int SC, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
pfbn_kernel(const PfbNParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(16) unsigned char pfb_smem[];
    float2* X = reinterpret_cast<float2*>(pfb_smem); /* [C][pitchT], element n of a row at pad16(n) */
    const int N = p.N, C = p.C, M = p.M, R = p.R;
    const int pitchT = N + (N >> 4) + 16 / C; /* = 16 / C (mod 16): the transposed read of the store phase is conflict-free */
    const int n16 = N >> 4;
    unsigned short* binOf = reinterpret_cast<unsigned short*>(X + C * pitchT); /* [N]: position -> output row */

    /* position -> bin: the digits of pos (most significant first: radix rho, then 16s) are the bin's digits, least
     * significant first */
    for (int pos = threadIdx.x; pos < N; pos += NT) {
        int rem = pos, bin = 0, mul = 1, size = N;
        if (p.rho > 1) {
            size = N / p.rho;
            bin = rem / size;
            rem -= bin * size;
            mul = p.rho;
        }
        while (size > 1) {
            size >>= 4;
            const int d = rem / size;
            rem -= d * size;
            bin += d * mul;
            mul <<= 4;
        }
        binOf[pos] = (unsigned short)bin;
    }
    __syncthreads();

    /* cu8 pair path: every thread stages the window of its NEXT (task, branch) step with 4-byte cp.async copies into its own
     * column of stage[2][T-1+SC][blockDim] while it filters the current one (no cross-thread sharing, so no barrier: only
     * cp.async.wait_group); the first step of a chunk is requested before the previous chunk's transform and store */
    constexpr bool kStaged = CU8 && (T <= 8);
    unsigned* stage = reinterpret_cast<unsigned*>(pfb_smem + (((size_t)C * pitchT * sizeof(float2) + (size_t)N * 2 + 127) & ~(size_t)127));
    unsigned qi = 0; /* steps taken by this thread: stage buffer = qi & 1 */
    /* TMA form of the staging (p.use_tma): thread 0 requests the whole CTA's window of a step as 2-D tensor-map copies
     * (rows x 256 pairs each) that complete on an mbarrier per stage buffer; a block barrier per step keeps a buffer from being
     * refilled while a warp still reads it.  Layout: stage[buf][half = t >> 8][row][t & 255]. */
    constexpr int kRows = T - 1 + SC, kHalves = NT / kTmaBox;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(stage + 2 * kRows * NT);
    unsigned tma_uses[2] = {0u, 0u};
    if (kStaged && p.use_tma) {
        if (threadIdx.x == 0) {
            pfb_mbar_init(bar0, 1);
            pfb_mbar_init(bar0 + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    auto stage_window = [&](long n0s, int nvs, int task, int j, unsigned buf) {
        const int s = (task >> (p.lgN - 1)) * SC, np = (task & ((N >> 1) - 1)) * 2;
        if (!((n0s + s >= T - 1) && (s + SC <= nvs))) {
            return; /* start-up / ragged windows are read directly */
        }
        if (p.use_tma) {
            if (threadIdx.x == 0) { /* its pair is the first of the step's NT pairs */
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const unsigned bar = bar0 + 8u * buf;
                pfb_mbar_expect_tx(bar, (unsigned)(kRows * NT * 4));
#pragma unroll
                for (int h = 0; h < kHalves; h++) {
                    pfb_tma_load_2d((unsigned)__cvta_generic_to_shared(stage + ((size_t)buf * kHalves + h) * kRows * kTmaBox), &tmap, bar,
                                    ((np + j * N) >> 1) + h * kTmaBox, (int)(n0s + s - (T - 1)));
                }
            }
            return;
        }
        const unsigned short* src = reinterpret_cast<const unsigned short*>(p.in) + (n0s + s - (T - 1)) * (long)M + np + j * N;
        const unsigned d32 = (unsigned)__cvta_generic_to_shared(stage + (size_t)buf * (T - 1 + SC) * NT + threadIdx.x);
#pragma unroll
        for (int i = 0; i < T - 1 + SC; i++) {
            /* shared address = base + constant; the global pointer advances by one row (a 64-bit add) per copy */
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d32 + i * NT * 4), "l"(src) : "memory");
            src += M;
        }
    };

    const long chunk0 = (long)blockIdx.x * p.chunks_per_cta;
    for (int c = 0; c < p.chunks_per_cta; c++) {
        const long n0 = (chunk0 + c) * C;
        if (n0 >= p.n_out) {
            break;
        }
        const int nv = (int)min((long)C, p.n_out - n0);
        const bool more = (c + 1 < p.chunks_per_cta) && (n0 + C < p.n_out);
        const int nv_next = more ? (int)min((long)C, p.n_out - (n0 + C)) : 0;

        /* ---- branch FIRs + fold: a task = (fold index n', window s); the R branches of n' accumulate in place.
         * cu8 input is filtered as raw byte values (exact in f32); the widening (u - 127.5) / 127.5 of
         * widen_u8_to_f32_bias127 is affine, so it is applied once per output: taps pre-scaled, bias as the
         * accumulator's start value. ---- */
        const int n_win = (C + SC - 1) / SC;
        if constexpr (T <= 8) {
            /* two adjacent branches per thread: one 4-byte (cu8) / 16-byte (cf32) request feeds both windows */
            const int n_tasks = (N >> 1) * n_win;
            if (kStaged && c == 0) {
                if ((int)threadIdx.x < n_tasks) {
                    stage_window(n0, nv, threadIdx.x, 0, qi & 1u);
                }
                if (!p.use_tma) {
                    asm volatile("cp.async.commit_group;" ::: "memory");
                }
            }
            for (int task = threadIdx.x; task < n_tasks; task += NT) {
                const int s = (task >> (p.lgN - 1)) * SC, np = (task & ((N >> 1) - 1)) * 2;
                float2 wMa = make_float2(1.0f, 0.0f), wMb = wMa;
                if (p.r0) {
                    float sn, cs;
                    sincospif(-2.0f * (float)(np * p.r0) / (float)M, &sn, &cs);
                    wMa = make_float2(cs, sn);
                    sincospif(-2.0f * (float)((np + 1) * p.r0) / (float)M, &sn, &cs);
                    wMb = make_float2(cs, sn);
                }
                float2* dst0 = X + s * pitchT + pad16(np); /* np even: np + 1 is the next padded slot */
                const bool interior = (n0 + s >= T - 1) && (s + SC <= nv);
                const long first = (n0 + s - (T - 1)) * (long)M;
                for (int j = 0; j < R; j++) {
                    const int b = np + j * N;
                    if (kStaged) {
                        /* request the next step's window, then wait for this step's */
                        if (p.use_tma) {
                            __syncthreads(); /* every warp is done with the buffer the next copy fills */
                        }
                        if (j + 1 < R) {
                            stage_window(n0, nv, task, j + 1, (qi + 1u) & 1u);
                        } else if (task + NT < n_tasks) {
                            stage_window(n0, nv, task + NT, 0, (qi + 1u) & 1u);
                        } else if (more) {
                            stage_window(n0 + C, nv_next, threadIdx.x, 0, (qi + 1u) & 1u);
                        }
                        if (!p.use_tma) {
                            asm volatile("cp.async.commit_group;" ::: "memory");
                        }
                    }
                    float2 ga[T], gb[T];
#pragma unroll
                    for (int q = 0; q < T; q++) {
                        const float2 t = __ldg(reinterpret_cast<const float2*>(p.proto + q * M + (M - 2 - b)));
                        ga[q] = make_float2(t.y, t.y); /* branch b:     proto[q M + M-1-b] */
                        gb[q] = make_float2(t.x, t.x); /* branch b + 1: proto[q M + M-2-b] */
                    }
                    const float2 wa = cmul(p.wR[j], wMa), wb = cmul(p.wR[j], wMb);
                    float2 va[SC], vb[SC];
                    {
                        float2 va0 = make_float2(0.0f, 0.0f), vb0 = va0;
                        if (CU8) {
                            const float2 bias = __ldg(reinterpret_cast<const float2*>(p.bias + b));
                            va0 = make_float2(bias.x, bias.x);
                            vb0 = make_float2(bias.y, bias.y);
                        }
#pragma unroll
                        for (int o = 0; o < SC; o++) {
                            va[o] = va0;
                            vb[o] = vb0;
                        }
                    }
                    if (kStaged) {
                        if (!p.use_tma) {
                            asm volatile("cp.async.wait_group 1;" ::: "memory");
                        } else if (interior) {
                            const unsigned bsel = qi & 1u;
                            pfb_mbar_wait(bar0 + 8u * bsel, tma_uses[bsel] & 1u);
                            tma_uses[bsel]++;
                        }
                    }
                    if (kStaged && interior) {
                        /* stream the staged rows through the accumulators: row i feeds outputs i-(T-1) .. i with taps
                         * T-1 .. 0 (every output sums its taps from the oldest sample to the newest, as the direct path) */
                        /* cp.async form: stage[buf][row][t]; TMA form: stage[buf][t >> 8][row][t & 255] */
                        const unsigned* mine = p.use_tma ? stage + ((size_t)(qi & 1u) * kHalves + (threadIdx.x / kTmaBox)) * kRows * kTmaBox
                                                               + (threadIdx.x % kTmaBox)
                                                         : stage + (size_t)(qi & 1u) * kRows * NT + threadIdx.x;
                        const int row_pitch = p.use_tma ? kTmaBox : NT;
#pragma unroll
                        for (int i = 0; i < T - 1 + SC; i++) {
                            const unsigned raw = mine[i * row_pitch];
                            float2 xa = make_float2(__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7440)),
                                                    __uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7441)));
                            float2 xb = make_float2(__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7442)),
                                                    __uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7443)));
                            xa = __fadd2_rn(xa, make_float2(-8388608.0f, -8388608.0f));
                            xb = __fadd2_rn(xb, make_float2(-8388608.0f, -8388608.0f));
#pragma unroll
                            for (int o = 0; o < SC; o++) {
                                const int q = T - 1 + o - i;
                                if (q >= 0 && q < T) {
                                    va[o] = __ffma2_rn(ga[q], xa, va[o]);
                                    vb[o] = __ffma2_rn(gb[q], xb, vb[o]);
                                }
                            }
                        }
                    } else {
                        float2 xa[T - 1 + SC], xb[T - 1 + SC];
                        if (interior) {
                            if (CU8) {
                                const unsigned short* src = reinterpret_cast<const unsigned short*>(p.in) + first;
#pragma unroll
                                for (int i = 0; i < T - 1 + SC; i++) {
                                    const unsigned raw = __ldg(reinterpret_cast<const unsigned*>(src + (b + i * M)));
                                    xa[i] = make_float2(__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7440)),
                                                        __uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7441)));
                                    xb[i] = make_float2(__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7442)),
                                                        __uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7443)));
                                }
#pragma unroll
                                for (int i = 0; i < T - 1 + SC; i++) {
                                    xa[i] = __fadd2_rn(xa[i], make_float2(-8388608.0f, -8388608.0f));
                                    xb[i] = __fadd2_rn(xb[i], make_float2(-8388608.0f, -8388608.0f));
                                }
                            } else {
                                const float2* src = reinterpret_cast<const float2*>(p.in) + first;
#pragma unroll
                                for (int i = 0; i < T - 1 + SC; i++) {
                                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (b + i * M)));
                                    xa[i] = make_float2(v.x, v.y);
                                    xb[i] = make_float2(v.z, v.w);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < T - 1 + SC; i++) {
                                const long blk = n0 + s + i - (T - 1);
                                float2 xva = make_float2(0.0f, 0.0f), xvb = xva;
                                if (blk < 0) {
                                    xva = p.hist[(long)(T - 1 + blk) * M + b]; /* kept widened */
                                    xvb = p.hist[(long)(T - 1 + blk) * M + b + 1];
                                    if (CU8) {
                                        xva = make_float2(hist_byte(xva.x), hist_byte(xva.y));
                                        xvb = make_float2(hist_byte(xvb.x), hist_byte(xvb.y));
                                    }
                                } else if (blk < p.n_out) {
                                    if (CU8) {
                                        const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const uchar2*>(p.in) + blk * M + b));
                                        xva = make_float2((float)u.x, (float)u.y);
                                        xvb = make_float2((float)u.z, (float)u.w);
                                    } else {
                                        const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(p.in) + blk * M + b));
                                        xva = make_float2(v.x, v.y);
                                        xvb = make_float2(v.z, v.w);
                                    }
                                }
                                xa[i] = xva;
                                xb[i] = xvb;
                            }
                        }
#pragma unroll
                        for (int o = 0; o < SC; o++) {
#pragma unroll
                            for (int q = T - 1; q >= 0; q--) {
                                va[o] = __ffma2_rn(ga[q], xa[T - 1 + o - q], va[o]);
                                vb[o] = __ffma2_rn(gb[q], xb[T - 1 + o - q], vb[o]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < SC; o++) {
                        if (s + o < C) {
                            float2* dst = dst0 + o * pitchT;
                            float2 ra = va[o], rb = vb[o];
                            if (R > 1) {
                                ra = cmul(ra, wa);
                                rb = cmul(rb, wb);
                                if (j > 0) {
                                    ra = __fadd2_rn(ra, dst[0]);
                                    rb = __fadd2_rn(rb, dst[1]);
                                }
                            }
                            dst[0] = ra;
                            dst[1] = rb;
                        }
                    }
                    qi++;
                }
            }
        } else {
        for (int task = threadIdx.x; task < N * n_win; task += NT) {
            const int s = (task >> p.lgN) * SC, np = task & (N - 1);
            float2 wM = make_float2(1.0f, 0.0f);
            if (p.r0) {
                float sn, cs;
                sincospif(-2.0f * (float)(np * p.r0) / (float)M, &sn, &cs);
                wM = make_float2(cs, sn);
            }
            float2* dst0 = X + s * pitchT + pad16(np);
            const bool interior = (n0 + s >= T - 1) && (s + SC <= nv);
            const long first = (n0 + s - (T - 1)) * (long)M; /* oldest sample of the window, branch 0 */
            for (int j = 0; j < R; j++) {
                const int b = np + j * N;
                float2 g2[T];
#pragma unroll
                for (int q = 0; q < T; q++) {
                    const float g = __ldg(p.proto + q * M + (M - 1 - b));
                    g2[q] = make_float2(g, g);
                }
                float2 xs[T - 1 + SC];
                if (interior) {
                    if (CU8) {
                        const unsigned short* src = reinterpret_cast<const unsigned short*>(p.in) + first;
#pragma unroll
                        for (int i = 0; i < T - 1 + SC; i++) {
                            const unsigned raw = __ldg(src + (b + i * M));
                            xs[i] = make_float2(__uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7440)),
                                                __uint_as_float(__byte_perm(raw, 0x4B000000u, 0x7441)));
                        }
#pragma unroll
                        for (int i = 0; i < T - 1 + SC; i++) {
                            xs[i] = __fadd2_rn(xs[i], make_float2(-8388608.0f, -8388608.0f));
                        }
                    } else {
                        const float2* src = reinterpret_cast<const float2*>(p.in) + first;
#pragma unroll
                        for (int i = 0; i < T - 1 + SC; i++) {
                            xs[i] = __ldg(src + (b + i * M));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < T - 1 + SC; i++) {
                        const long blk = n0 + s + i - (T - 1);
                        float2 v = make_float2(0.0f, 0.0f);
                        if (blk < 0) {
                            v = p.hist[(long)(T - 1 + blk) * M + b]; /* kept widened */
                            if (CU8) {
                                v = make_float2(hist_byte(v.x), hist_byte(v.y));
                            }
                        } else if (blk < p.n_out) {
                            if (CU8) {
                                const uchar2 u = __ldg(reinterpret_cast<const uchar2*>(p.in) + blk * M + b);
                                v = make_float2((float)u.x, (float)u.y);
                            } else {
                                v = __ldg(reinterpret_cast<const float2*>(p.in) + blk * M + b);
                            }
                        }
                        xs[i] = v;
                    }
                }
                const float2 w = cmul(p.wR[j], wM);
                float2 v0 = make_float2(0.0f, 0.0f);
                if (CU8) {
                    const float bias = __ldg(p.bias + b);
                    v0 = make_float2(bias, bias);
                }
#pragma unroll
                for (int i = 0; i < SC; i++) {
                    float2 v = v0;
#pragma unroll
                    for (int q = 0; q < T; q++) {
                        v = __ffma2_rn(g2[q], xs[T - 1 + i - q], v);
                    }
                    if (s + i < C) {
                        float2* dst = dst0 + i * pitchT;
                        if (R > 1) {
                            v = cmul(v, w);
                            if (j > 0) {
                                v = __fadd2_rn(v, *dst);
                            }
                        }
                        *dst = v;
                    }
                }
            }
        }
        }
        __syncthreads();

        /* ---- in-place decimation-in-frequency transform of the C rows ---- */
        if (p.rho == 2) {
            pfbn_first_pass<2, NT>(X, p, pitchT);
        } else if (p.rho == 4) {
            pfbn_first_pass<4, NT>(X, p, pitchT);
        } else if (p.rho == 8) {
            pfbn_first_pass<8, NT>(X, p, pitchT);
        }
        {
            const int U = min(n16, NT), rows_par = NT / U;
            for (int S = N / p.rho; S >= 16; S >>= 4) {
                const int sub = S >> 4; /* butterflies per block of size S */
                const int tw_step = N / S;
                for (int u = threadIdx.x % U; u < n16; u += U) {
                    const int g = u / sub, j = u - g * sub;
                    const int base = g * S + j;
                    float2 tw[16];
                    if (S > 16) {
#pragma unroll
                        for (int k = 1; k < 16; k++) {
                            tw[k] = __ldg(p.twN + j * k * tw_step);
                        }
                    }
                    /* sub is 1 (base a multiple of 16) or a multiple of 16: the padded index is linear in q */
                    const int pstride = sub + (sub >> 4);
                    for (int row = threadIdx.x / U; row < C; row += rows_par) {
                        float2* Xr = X + row * pitchT + pad16(base);
                        float2 a[16];
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            a[q] = Xr[q * pstride];
                        }
                        dft16(a);
                        if (S > 16) {
#pragma unroll
                            for (int k = 1; k < 16; k++) {
                                a[k] = cmul(a[k], tw[k]);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            Xr[k * pstride] = a[k];
                        }
                    }
                }
                __syncthreads();
            }
        }

        /* ---- transposed store: C consecutive lanes = C consecutive times of one channel ---- */
        if (p.out_u8) {
            /* cu8 rows (the receive bank's native input format: 2 B per sample instead of 8): a lane packs four consecutive
             * times of one channel into one 8-byte store; a chunk that is cut short by the end of the launch stores bytes */
            const int G = C >> 2, lgG = p.lgC - 2; /* C >= 4 */
            const float sc = p.out_gain * 127.5f;
            for (int t = threadIdx.x; t < N * G; t += NT) {
                const int g = t & (G - 1), pos = t >> lgG, i0 = 4 * g;
                if (i0 >= nv) {
                    continue;
                }
                const float2* src = X + i0 * pitchT + pad16(pos);
                unsigned w[2] = {0u, 0u};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float2 v = src[k * pitchT];
                    const unsigned qx = (unsigned)__float2int_rn(fminf(fmaxf(fmaf(v.x, sc, 127.5f), 0.0f), 255.0f));
                    const unsigned qy = (unsigned)__float2int_rn(fminf(fmaxf(fmaf(v.y, sc, 127.5f), 0.0f), 255.0f));
                    w[k >> 1] |= (qx | (qy << 8)) << (16 * (k & 1));
                }
                uchar2* dst = p.out_u8 + (size_t)binOf[pos] * p.out_pitch + n0 + i0;
                if (i0 + 4 <= nv) {
                    __stcs(reinterpret_cast<uint2*>(dst), make_uint2(w[0], w[1]));
                } else {
                    for (int k = 0; i0 + k < nv; k++) {
                        const unsigned q = w[k >> 1] >> (16 * (k & 1));
                        dst[k] = make_uchar2((unsigned char)(q & 0xffu), (unsigned char)((q >> 8) & 0xffu));
                    }
                }
            }
        } else {
            const int i = threadIdx.x & (C - 1); /* blockDim is a multiple of C */
            const int pstep = NT >> p.lgC;
            float2* dst = p.out + n0 + i;
            const float2* src = X + i * pitchT;
            if (i < nv) {
#pragma unroll 4
                for (int pos = threadIdx.x >> p.lgC; pos < N; pos += pstep) {
                    __stcs(dst + (size_t)binOf[pos] * p.out_pitch, src[pad16(pos)]);
                }
            }
        }
        __syncthreads();
    }
}

/* hist_new := the last hist_len samples of (hist_old ++ in) */
template <bool CU8>
__global__ void __launch_bounds__(256)
pfb_hist_kernel(const void* in, long n_in, const float2* hist_old, float2* hist_new, int hist_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hist_len) {
        const long src = n_in - hist_len + i;
        hist_new[i] = (src >= 0) ? load_sample<CU8>(in, src) : hist_old[i + n_in];
    }
}

}  // namespace

struct dsdneo_b200_channelizer {
    int M, T, cu8;
    float* d_proto;
    float2* d_hist;      /* current history: the (T-1) * M samples before the next tile */
    float2* d_hist_alt;  /* written by the next history update, then the two swap */
    float2* d_tw[6];     /* W_N tables for N = M >> i (bin strides 1, 2, ... 32), made on first use */
    float* d_proto_u8;   /* cu8 input: prototype * (1 / 127.5) ... */
    float* d_bias_u8;    /* ... and -127.5 * (branch tap sum of it), per branch */
    float* h_proto;
    void* d_stage_in;
    size_t stage_in_cap;
    float* d_stage_out;
    size_t stage_out_cap;
};

template <int T, bool CU8>
static int
launch_pfb(const PfbParams& p, int grid, cudaStream_t s) {
    const size_t smem = (size_t)(kChunk * kFftPitch + kM * kOutPitch) * sizeof(float2);
    /* the attribute is per device: one process may drive several GPUs, so the cache is keyed by device ordinal */
    static bool attr_done[64] = {};
    int dev = 0;
    DSDNEO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        DSDNEO_CUDA(cudaFuncSetAttribute(pfb256_kernel<T, CU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) {
            attr_done[dev] = true;
        }
    }
    {
        KernelTimer kt("pfb256_kernel", s);
        pfb256_kernel<T, CU8><<<grid, kM, smem, s>>>(p);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

static int
pfbn_chunk_times(int N) {
    /* C * N elements of shared memory per CTA: 8192 (70 KB, two or three CTAs per SM) up to N = 1024, 16384 above */
    int c = (N <= 1024 ? 8192 : 16384) / N;
    return c > 16 ? 16 : (c < 2 ? 2 : c);
}

template <int T, bool CU8, int SC, int NT>
static int
launch_pfbn_nt(const PfbNParams& p, const CUtensorMap& tmap, int grid, size_t smem, cudaStream_t s) {
    static bool attr_done[64] = {};
    int dev = 0;
    DSDNEO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        DSDNEO_CUDA(cudaFuncSetAttribute(pfbn_kernel<T, CU8, SC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        if (dev >= 0 && dev < 64) {
            attr_done[dev] = true;
        }
    }
    {
        KernelTimer kt("pfbn_kernel", s);
        pfbn_kernel<T, CU8, SC, NT><<<grid, NT, smem, s>>>(p, tmap);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

template <int T, bool CU8, int SC>
static int
launch_pfbn(const PfbNParams& p, const CUtensorMap& tmap, int grid, int threads, size_t smem, cudaStream_t s) {
    return threads == 256 ? launch_pfbn_nt<T, CU8, SC, 256>(p, tmap, grid, smem, s) : launch_pfbn_nt<T, CU8, SC, 512>(p, tmap, grid, smem, s);
}

/* Tensor map over a cu8 tile seen as [rows][M / 2] 32-bit elements (two adjacent branches per element), box = 256 elements x
 * `box_rows` rows.  The driver entry point is looked up once (no link against libcuda).  Returns false when the map cannot be
 * made (alignment, old driver): the caller keeps the cp.async staging. */
static bool
make_tile_map(CUtensorMap* map, const void* d_in, int M, long n_rows, int box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = NULL;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* ptr = NULL;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) {
            fn = (encode_fn)ptr;
        } else {
            (void)cudaGetLastError();
        }
    }
    if (!fn || (((uintptr_t)d_in) & 15) != 0 || ((size_t)M * 2) % 16 != 0 || n_rows < box_rows) {
        return false;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)(M / 2), (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)M * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTmaBox, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(d_in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* W_N table for bin stride R = 1 << lg (N = M >> lg), built in float64 on first use */
static int
ensure_twiddles(dsdneo_b200_channelizer* c, int lg) {
    if (c->d_tw[lg]) {
        return 0;
    }
    const int N = c->M >> lg;
    float2* h = (float2*)malloc(sizeof(float2) * (size_t)N);
    if (!h) {
        set_error("channelize: out of host memory");
        return DSDNEO_B200_ENOMEM;
    }
    const double pi = 3.14159265358979323846;
    for (int m = 0; m < N; m++) {
        const double a = -2.0 * pi * (double)m / (double)N;
        h[m] = make_float2((float)cos(a), (float)sin(a));
    }
    cudaError_t e = cudaMalloc((void**)&c->d_tw[lg], sizeof(float2) * (size_t)N);
    if (e == cudaSuccess) {
        e = cudaMemcpy(c->d_tw[lg], h, sizeof(float2) * (size_t)N, cudaMemcpyHostToDevice);
    }
    free(h);
    if (e != cudaSuccess) {
        return cuda_fail(e, "channelize (twiddle table)", __FILE__, __LINE__);
    }
    return 0;
}

extern "C" {

int
dsdneo_b200_channelizer_design_prototype(int n_channels, int taps_per_branch, double cutoff_rel, float* h_out) {
    /* Blackman-windowed sinc, L = M*T taps, cutoff = cutoff_rel * (fs / (2M)) (1.0 = half the channel spacing),
     * unit DC gain.  Designed in float64, stored as f32. */
    if (n_channels < 2 || taps_per_branch < 1 || !h_out || !(cutoff_rel > 0.0)) {
        set_error("channelizer_design_prototype: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    const int L = n_channels * taps_per_branch;
    const double fc = cutoff_rel * 0.5 / (double)n_channels; /* cycles per input sample */
    const double mid = 0.5 * (double)(L - 1);
    const double pi = 3.14159265358979323846;
    double sum = 0.0;
    double* tmp = (double*)malloc(sizeof(double) * (size_t)L);
    if (!tmp) {
        set_error("channelizer_design_prototype: out of memory");
        return DSDNEO_B200_ENOMEM;
    }
    for (int m = 0; m < L; m++) {
        const double t = (double)m - mid;
        const double x = 2.0 * fc * t;
        const double sinc = (fabs(x) < 1e-12) ? 1.0 : sin(pi * x) / (pi * x);
        const double w = 0.42 - 0.5 * cos(2.0 * pi * (double)m / (double)(L - 1)) + 0.08 * cos(4.0 * pi * (double)m / (double)(L - 1));
        tmp[m] = 2.0 * fc * sinc * w;
        sum += tmp[m];
    }
    for (int m = 0; m < L; m++) {
        h_out[m] = (float)(tmp[m] / sum);
    }
    free(tmp);
    return L;
}

dsdneo_b200_channelizer*
dsdneo_b200_channelizer_create(int n_channels, int taps_per_branch, int input_is_cu8, const float* prototype) {
    if (ensure_device()) {
        return NULL;
    }
    if (n_channels < 256 || n_channels > 8192 || (n_channels & (n_channels - 1)) != 0) {
        set_error("channelizer_create: n_channels must be a power of two in 256 ... 8192 (got %d)", n_channels);
        return NULL;
    }
    if (taps_per_branch != 4 && taps_per_branch != 8 && taps_per_branch != 12 && taps_per_branch != 16) {
        set_error("channelizer_create: taps_per_branch must be 4, 8, 12 or 16");
        return NULL;
    }
    dsdneo_b200_channelizer* c = (dsdneo_b200_channelizer*)calloc(1, sizeof(*c));
    if (!c) {
        set_error("channelizer_create: out of host memory");
        return NULL;
    }
    c->M = n_channels;
    c->T = taps_per_branch;
    c->cu8 = input_is_cu8 ? 1 : 0;
    const int L = c->M * c->T;
    c->h_proto = (float*)malloc(sizeof(float) * (size_t)L);
    if (!c->h_proto) {
        free(c);
        set_error("channelizer_create: out of host memory");
        return NULL;
    }
    if (prototype) {
        memcpy(c->h_proto, prototype, sizeof(float) * (size_t)L);
    } else if (dsdneo_b200_channelizer_design_prototype(c->M, c->T, 1.0, c->h_proto) < 0) {
        free(c->h_proto);
        free(c);
        return NULL;
    }
    cudaError_t e = cudaMalloc((void**)&c->d_proto, sizeof(float) * (size_t)L);
    if (e == cudaSuccess) {
        e = cudaMalloc((void**)&c->d_hist, sizeof(float2) * (size_t)(c->T - 1) * c->M);
    }
    if (e == cudaSuccess) {
        e = cudaMemcpy(c->d_proto, c->h_proto, sizeof(float) * (size_t)L, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        e = cudaMalloc((void**)&c->d_hist_alt, sizeof(float2) * (size_t)(c->T - 1) * c->M);
    }
    if (e == cudaSuccess && c->cu8) {
        /* the general kernel filters raw bytes: fold widen_u8_to_f32_bias127's scale into the taps and its offset into a
         * per-branch start value (float64 here, f32 on the device) */
        float* hs = (float*)malloc(sizeof(float) * (size_t)(L + c->M));
        if (!hs) {
            e = cudaErrorMemoryAllocation;
        } else {
            for (int m = 0; m < L; m++) {
                hs[m] = (float)((double)c->h_proto[m] / 127.5);
            }
            for (int b = 0; b < c->M; b++) {
                double acc = 0.0;
                for (int q = 0; q < c->T; q++) {
                    acc += (double)hs[q * c->M + (c->M - 1 - b)];
                }
                hs[L + b] = (float)(-127.5 * acc);
            }
            e = cudaMalloc((void**)&c->d_proto_u8, sizeof(float) * (size_t)L);
            if (e == cudaSuccess) {
                e = cudaMalloc((void**)&c->d_bias_u8, sizeof(float) * (size_t)c->M);
            }
            if (e == cudaSuccess) {
                e = cudaMemcpy(c->d_proto_u8, hs, sizeof(float) * (size_t)L, cudaMemcpyHostToDevice);
            }
            if (e == cudaSuccess) {
                e = cudaMemcpy(c->d_bias_u8, hs + L, sizeof(float) * (size_t)c->M, cudaMemcpyHostToDevice);
            }
            free(hs);
        }
    }
    if (e == cudaSuccess) {
        e = cudaMemset(c->d_hist, 0, sizeof(float2) * (size_t)(c->T - 1) * c->M);
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "channelizer_create", __FILE__, __LINE__);
        dsdneo_b200_channelizer_destroy(c);
        return NULL;
    }
    return c;
}

void
dsdneo_b200_channelizer_destroy(dsdneo_b200_channelizer* c) {
    if (!c) {
        return;
    }
    cudaFree(c->d_proto);
    cudaFree(c->d_hist);
    cudaFree(c->d_hist_alt);
    cudaFree(c->d_proto_u8);
    cudaFree(c->d_bias_u8);
    for (int i = 0; i < 6; i++) {
        cudaFree(c->d_tw[i]);
    }
    cudaFree(c->d_stage_in);
    cudaFree(c->d_stage_out);
    free(c->h_proto);
    free(c);
}

int
dsdneo_b200_channelizer_reset(dsdneo_b200_channelizer* c, void* stream) {
    if (!c) {
        set_error("channelizer_reset: NULL");
        return DSDNEO_B200_EINVAL;
    }
    DSDNEO_CUDA(cudaMemsetAsync(c->d_hist, 0, sizeof(float2) * (size_t)(c->T - 1) * c->M, as_stream(stream)));
    return 0;
}

int
dsdneo_b200_channelizer_prime(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, void* stream) {
    if (!c || !d_in || n_in_samples == 0) {
        set_error("channelizer_prime: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    const int hist_len = (c->T - 1) * c->M;
    const int grid_h = (hist_len + 255) / 256;
    if (c->cu8) {
        pfb_hist_kernel<true><<<grid_h, 256, 0, s>>>(d_in, (long)n_in_samples, c->d_hist, c->d_hist_alt, hist_len);
    } else {
        pfb_hist_kernel<false><<<grid_h, 256, 0, s>>>(d_in, (long)n_in_samples, c->d_hist, c->d_hist_alt, hist_len);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    float2* t = c->d_hist;
    c->d_hist = c->d_hist_alt;
    c->d_hist_alt = t;
    return 0;
}

int
dsdneo_b200_channelizer_get_prototype(dsdneo_b200_channelizer* c, float* h_out, int max_taps) {
    if (!c || !h_out || max_taps < c->M * c->T) {
        set_error("channelizer_get_prototype: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    memcpy(h_out, c->h_proto, sizeof(float) * (size_t)(c->M * c->T));
    return c->M * c->T;
}

static int
channelize_impl(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, int bin_stride, int bin_first, int advance,
                float* d_out, uint8_t* d_out_u8, float out_gain, size_t out_pitch_pairs, void* stream) {
    if (!c || !d_in || (!d_out && !d_out_u8) || n_in_samples == 0 || (n_in_samples % (size_t)c->M) != 0) {
        set_error("channelize: bad argument (n_in_samples must be a positive multiple of n_channels)");
        return DSDNEO_B200_EINVAL;
    }
    int lg = 0;
    while ((1 << lg) < bin_stride) {
        lg++;
    }
    if (bin_stride < 1 || bin_stride > kMaxFold || (1 << lg) != bin_stride || c->M / bin_stride < 256 || bin_first < 0 ||
        bin_first >= bin_stride) {
        set_error("channelize_bins: bin_stride must be a power of two <= %d that leaves >= 256 channels, 0 <= bin_first < "
                  "bin_stride (got %d, %d; M = %d)",
                  kMaxFold, bin_stride, bin_first, c->M);
        return DSDNEO_B200_EINVAL;
    }
    const long n_out = (long)(n_in_samples / (size_t)c->M);
    if (out_pitch_pairs < (size_t)n_out) {
        set_error("channelize: out_pitch_pairs smaller than the number of outputs per channel");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
    }
    if (d_out_u8 && (c->M / bin_stride > 4096 || (out_pitch_pairs & 3) != 0 || !(out_gain > 0.0f))) {
        set_error("channelize_bins_cu8: needs at most 4096 output channels, out_pitch_pairs a multiple of 4 and a positive gain");
        return DSDNEO_B200_EINVAL;
    }
    if (c->M == kM && bin_stride == 1 && !d_out_u8) {
        PfbParams p;
        p.in = d_in;
        p.hist = c->d_hist;
        p.proto = c->d_proto;
        p.out = reinterpret_cast<float2*>(d_out);
        p.out_pitch = out_pitch_pairs;
        p.n_out = n_out;
        const long n_chunks = (n_out + kChunk - 1) / kChunk;
        /* one resident wave (2 CTAs per SM), at least 4 chunks each so the (T-1)-block window preload is amortised */
        long cpc = (n_chunks + sms * 2 - 1) / (sms * 2);
        if (cpc < 4) {
            cpc = 4;
        }
        p.chunks_per_cta = (int)cpc;
        const int grid = (int)((n_chunks + cpc - 1) / cpc);
#define PFB_CASE(TT)                                                                                                   \
    case TT: rc = c->cu8 ? launch_pfb<TT, true>(p, grid, s) : launch_pfb<TT, false>(p, grid, s); break;
        switch (c->T) {
            PFB_CASE(4)
            PFB_CASE(8)
            PFB_CASE(12)
            PFB_CASE(16)
            default: set_error("channelize: unsupported taps_per_branch"); return DSDNEO_B200_EUNSUPPORTED;
        }
#undef PFB_CASE
    } else {
        rc = ensure_twiddles(c, lg);
        if (rc) {
            return rc;
        }
        PfbNParams p;
        memset(&p, 0, sizeof(p));
        p.in = d_in;
        p.hist = c->d_hist;
        p.proto = c->cu8 ? c->d_proto_u8 : c->d_proto;
        p.bias = c->d_bias_u8;
        p.twN = c->d_tw[lg];
        p.out = reinterpret_cast<float2*>(d_out);
        p.out_u8 = reinterpret_cast<uchar2*>(d_out_u8);
        p.out_gain = out_gain;
        p.out_pitch = out_pitch_pairs;
        p.n_out = n_out;
        p.M = c->M;
        p.R = bin_stride;
        p.r0 = bin_first;
        p.N = c->M / bin_stride;
        p.C = pfbn_chunk_times(p.N);
        int lgN = 0;
        while ((1 << lgN) < p.N) {
            lgN++;
        }
        p.rho = 1 << (lgN & 3);
        p.lgN = lgN;
        p.lgC = 0;
        while ((1 << p.lgC) < p.C) {
            p.lgC++;
        }
        const double pi = 3.14159265358979323846;
        for (int j = 0; j < bin_stride; j++) {
            const double a = -2.0 * pi * (double)((j * bin_first) % bin_stride) / (double)bin_stride;
            p.wR[j] = make_float2((float)cos(a), (float)sin(a));
        }
        const int pitchT = p.N + p.N / 16 + 16 / p.C;
        const int threads = (p.C * p.N <= 8192) ? 256 : 512;
        size_t smem = (((size_t)p.C * pitchT * sizeof(float2) + (size_t)p.N * sizeof(unsigned short)) + 127) & ~(size_t)127;
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        if (c->cu8 && c->T <= 8) {
            const int rows = c->T - 1 + (p.C >= 8 ? 8 : 4);
            smem += (size_t)2 * rows * threads * sizeof(unsigned) + 16; /* window staging (two buffers) + two mbarriers */
            /* tensor-map staging when a step's pairs are one contiguous run of the row: N / 2 a multiple of the block size */
            static int tma_env = -1;
            if (tma_env < 0) {
                const char* e = getenv("DSDNEO_B200_PFB_TMA");
                tma_env = e ? atoi(e) : 1;
            }
            p.use_tma = tma_env && ((p.N / 2) % threads == 0) && make_tile_map(&tmap, d_in, c->M, n_out, rows) ? 1 : 0;
        }
        const int per_sm = (p.C * p.N <= 8192) ? 2 : 1;
        const long n_chunks = (n_out + p.C - 1) / p.C;
        long cpc = (n_chunks + (long)sms * per_sm - 1) / ((long)sms * per_sm);
        if (cpc < 1) {
            cpc = 1;
        }
        p.chunks_per_cta = (int)cpc;
        const int grid = (int)((n_chunks + cpc - 1) / cpc);
#define PFBN_CASE(TT)                                                                                                  \
    case TT:                                                                                                           \
        if (p.C >= 8) {                                                                                                \
            rc = c->cu8 ? launch_pfbn<TT, true, 8>(p, tmap, grid, threads, smem, s)                                          \
                        : launch_pfbn<TT, false, 8>(p, tmap, grid, threads, smem, s);                                        \
        } else {                                                                                                       \
            rc = c->cu8 ? launch_pfbn<TT, true, 4>(p, tmap, grid, threads, smem, s)                                          \
                        : launch_pfbn<TT, false, 4>(p, tmap, grid, threads, smem, s);                                        \
        }                                                                                                              \
        break;
        switch (c->T) {
            PFBN_CASE(4)
            PFBN_CASE(8)
            PFBN_CASE(12)
            PFBN_CASE(16)
            default: set_error("channelize: unsupported taps_per_branch"); return DSDNEO_B200_EUNSUPPORTED;
        }
#undef PFBN_CASE
    }
    if (rc) {
        return rc;
    }
    if (advance) {
        const int hist_len = (c->T - 1) * c->M;
        const int grid_h = (hist_len + 255) / 256;
        if (c->cu8) {
            pfb_hist_kernel<true><<<grid_h, 256, 0, s>>>(d_in, (long)n_in_samples, c->d_hist, c->d_hist_alt, hist_len);
        } else {
            pfb_hist_kernel<false><<<grid_h, 256, 0, s>>>(d_in, (long)n_in_samples, c->d_hist, c->d_hist_alt, hist_len);
        }
        DSDNEO_KERNEL_CHECK();
        count_launch();
        float2* t = c->d_hist;
        c->d_hist = c->d_hist_alt;
        c->d_hist_alt = t;
    }
    return 0;
}

int
dsdneo_b200_channelize_bins(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, int bin_stride, int bin_first,
                            int advance, float* d_out, size_t out_pitch_pairs, void* stream) {
    return channelize_impl(c, d_in, n_in_samples, bin_stride, bin_first, advance, d_out, NULL, 0.0f, out_pitch_pairs, stream);
}

int
dsdneo_b200_channelize_bins_cu8(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, int bin_stride, int bin_first,
                                int advance, float gain, uint8_t* d_out, size_t out_pitch_pairs, void* stream) {
    if (!d_out) {
        set_error("channelize_bins_cu8: NULL output");
        return DSDNEO_B200_EINVAL;
    }
    return channelize_impl(c, d_in, n_in_samples, bin_stride, bin_first, advance, NULL, d_out, gain, out_pitch_pairs, stream);
}

int
dsdneo_b200_channelize(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, float* d_out,
                       size_t out_pitch_pairs, void* stream) {
    return dsdneo_b200_channelize_bins(c, d_in, n_in_samples, 1, 0, 1, d_out, out_pitch_pairs, stream);
}

int
dsdneo_b200_channelize_host(dsdneo_b200_channelizer* c, const void* h_in, size_t n_in_samples, float* h_out,
                            size_t out_pitch_pairs) {
    if (!c || !h_in || !h_out) {
        set_error("channelize_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t in_bytes = n_in_samples * (c->cu8 ? 2 : 8);
    const size_t out_floats = (size_t)c->M * out_pitch_pairs * 2;
    if (c->stage_in_cap < in_bytes) {
        cudaFree(c->d_stage_in);
        c->d_stage_in = NULL;
        c->stage_in_cap = 0;
        DSDNEO_CUDA(cudaMalloc(&c->d_stage_in, in_bytes));
        c->stage_in_cap = in_bytes;
    }
    if (c->stage_out_cap < out_floats) {
        cudaFree(c->d_stage_out);
        c->d_stage_out = NULL;
        c->stage_out_cap = 0;
        DSDNEO_CUDA(cudaMalloc((void**)&c->d_stage_out, out_floats * sizeof(float)));
        c->stage_out_cap = out_floats;
    }
    DSDNEO_CUDA(cudaMemcpyAsync(c->d_stage_in, h_in, in_bytes, cudaMemcpyHostToDevice, 0));
    rc = dsdneo_b200_channelize(c, c->d_stage_in, n_in_samples, c->d_stage_out, out_pitch_pairs, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpyAsync(h_out, c->d_stage_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost, 0));
    DSDNEO_CUDA(cudaStreamSynchronize(0));
    return 0;
}

} /* extern "C" */
