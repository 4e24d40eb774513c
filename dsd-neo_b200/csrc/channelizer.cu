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
 * Compiled with -fmad=true (nothing here is bit-pinned to a CPU path; tolerance is stated in the tests).
 */
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

__device__ __forceinline__ float2
cadd(float2 a, float2 b) {
    return make_float2(a.x + b.x, a.y + b.y);
}

__device__ __forceinline__ float2
csub(float2 a, float2 b) {
    return make_float2(a.x - b.x, a.y - b.y);
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

template <bool CU8>
__global__ void __launch_bounds__(1024)
pfb_save_hist_kernel(const void* in, long n_in, float2* hist, int hist_len) {
    /* hist := last hist_len input samples (older entries shift down when n_in < hist_len).
     * Single CTA; read everything, barrier, then write, so the in-place shift is race-free. */
    constexpr int kPer = (kMaxT - 1) * kM / 1024 + 1;
    float2 v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        const int i = threadIdx.x + 1024 * k;
        if (i < hist_len) {
            const long src = n_in - hist_len + i;
            v[k] = (src >= 0) ? load_sample<CU8>(in, src) : hist[i + n_in];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; k++) {
        const int i = threadIdx.x + 1024 * k;
        if (i < hist_len) {
            hist[i] = v[k];
        }
    }
}

}  // namespace

struct dsdneo_b200_channelizer {
    int M, T, cu8;
    float* d_proto;
    float2* d_hist;
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
    if (n_channels != kM) {
        set_error("channelizer_create: only M = %d channels is built in this round (got %d)", kM, n_channels);
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
dsdneo_b200_channelizer_get_prototype(dsdneo_b200_channelizer* c, float* h_out, int max_taps) {
    if (!c || !h_out || max_taps < c->M * c->T) {
        set_error("channelizer_get_prototype: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    memcpy(h_out, c->h_proto, sizeof(float) * (size_t)(c->M * c->T));
    return c->M * c->T;
}

int
dsdneo_b200_channelize(dsdneo_b200_channelizer* c, const void* d_in, size_t n_in_samples, float* d_out,
                       size_t out_pitch_pairs, void* stream) {
    if (!c || !d_in || !d_out || n_in_samples == 0 || (n_in_samples % (size_t)c->M) != 0) {
        set_error("channelize: bad argument (n_in_samples must be a positive multiple of n_channels)");
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
    PfbParams p;
    p.in = d_in;
    p.hist = c->d_hist;
    p.proto = c->d_proto;
    p.out = reinterpret_cast<float2*>(d_out);
    p.out_pitch = out_pitch_pairs;
    p.n_out = n_out;
    const long n_chunks = (n_out + kChunk - 1) / kChunk;
    /* one resident wave (2 CTAs per SM), at least 4 chunks each so the (T-1)-block window preload is amortised */
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
    }
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
    if (rc) {
        return rc;
    }
    const int hist_len = (c->T - 1) * c->M;
    if (c->cu8) {
        pfb_save_hist_kernel<true><<<1, 1024, 0, s>>>(d_in, (long)n_in_samples, c->d_hist, hist_len);
    } else {
        pfb_save_hist_kernel<false><<<1, 1024, 0, s>>>(d_in, (long)n_in_samples, c->d_hist, hist_len);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
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
