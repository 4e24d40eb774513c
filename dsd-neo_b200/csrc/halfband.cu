// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * Batched complex half-band decimator cascade (K3).
 *
 * Reference being replaced (arancormonk/dsd-neo @ 4d06905), per channel and per block:
 *   full_demod_apply_halfband_decimation   src/dsp/demod_pipeline.cpp:983-1001   (stage 0: 31 taps, others: 15)
 *     simd_hb_decim2_complex               src/dsp/simd_fir.cpp:363-373 (dispatch, short blocks -> scalar :302-305)
 *       _scalar                            src/dsp/simd_fir.cpp:139-222
 *       _avx2 (fused multiply-add)         src/dsp/simd_fir_avx2.cpp:312-390,457-516
 *   taps                                   src/dsp/halfband.cpp:35-74
 *
 * The reference uses this to bring ONE tuned channel from the capture rate down to 48 kS/s; behind a polyphase
 * channelizer it is only needed when the channelizer's output rate is a power-of-two multiple of the demod rate.
 * One thread = one decimated output; every stage is a stream over [channel][time] with 8 B in per input pair and 8 B out
 * per output pair (HBM-bound, the taps' overlap is served by L1).  Arithmetic order per output is the reference's
 * (centre tap, then even taps outwards-in, (x- + x+) pre-add; FMA or mul+add), so results are bit-identical.
 * Block edges behave as in the reference: beyond a block's end the block's last sample repeats; before its start the
 * stream's previous samples (carried history across launches) are used.
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

constexpr int kHbMaxPasses = 10; /* demod_state.h:150 hb_hist_i[10][30] */
constexpr int kHbHist = 30;      /* history slots per stage (31-tap stage uses all, 15-tap stages the last 14) */

__constant__ float c_hb15[15] = {-108.0f / 32768.0f, 0.0f, 1800.0f / 32768.0f, 0.0f, -500.0f / 32768.0f, 0.0f,
                                 7000.0f / 32768.0f, 0.5f, 7000.0f / 32768.0f, 0.0f, -500.0f / 32768.0f, 0.0f,
                                 1800.0f / 32768.0f, 0.0f, -108.0f / 32768.0f};
__constant__ float c_hb31[31] = {0.0f, 0.0f, 13.0f / 32768.0f, 0.0f, -73.0f / 32768.0f, 0.0f, 233.0f / 32768.0f, 0.0f,
                                 -587.0f / 32768.0f, 0.0f, 1314.0f / 32768.0f, 0.0f, -2953.0f / 32768.0f, 0.0f,
                                 10244.0f / 32768.0f, 16386.0f / 32768.0f, 10244.0f / 32768.0f, 0.0f,
                                 -2953.0f / 32768.0f, 0.0f, 1314.0f / 32768.0f, 0.0f, -587.0f / 32768.0f, 0.0f,
                                 233.0f / 32768.0f, 0.0f, -73.0f / 32768.0f, 0.0f, 13.0f / 32768.0f, 0.0f, 0.0f};

template <int TAPS, bool FMA>
__global__ void __launch_bounds__(256)
hb_decim2_kernel(const float2* __restrict__ in, size_t in_pitch, float2* __restrict__ out, size_t out_pitch,
                 const float2* __restrict__ hist_all, int blk_in, int n_blocks) {
    constexpr int H = TAPS - 1, C = H / 2;
    const float* taps = (TAPS == 31) ? c_hb31 : c_hb15;
    const int ch = blockIdx.y;
    const int blk_out = blk_in >> 1;
    const long total_out = (long)blk_out * n_blocks;
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total_out) {
        return;
    }
    const int b = (int)(o / blk_out);
    const int n = (int)(o - (long)b * blk_out);
    const long base = (long)b * blk_in;
    const float2* x = in + (size_t)ch * in_pitch;
    const float2* hist = hist_all + (size_t)ch * kHbHist + (kHbHist - H); /* the stage's last H inputs */
    auto sample = [&](int rel) -> float2 {
        if (rel >= blk_in) {
            rel = blk_in - 1; /* simd_fir.cpp:156-157,166-169: the block's last sample repeats */
        }
        const long g = base + rel;
        return (g >= 0) ? __ldg(&x[g]) : hist[H + g];
    };
    const int mid = 2 * n;
    const float2 xc = sample(mid);
    float ai, aq;
    if (FMA) {
        ai = __fmul_rn(taps[C], xc.x);
        aq = __fmul_rn(taps[C], xc.y);
    } else {
        ai = __fadd_rn(0.0f, __fmul_rn(taps[C], xc.x));
        aq = __fadd_rn(0.0f, __fmul_rn(taps[C], xc.y));
    }
#pragma unroll
    for (int e = 0; e < C; e += 2) {
        const float t = taps[e];
        if (t == 0.0f) {
            continue;
        }
        const int d = C - e;
        const float2 xm = sample(mid - d), xp = sample(mid + d);
        const float si = __fadd_rn(xm.x, xp.x), sq = __fadd_rn(xm.y, xp.y);
        if (FMA) {
            ai = __fmaf_rn(t, si, ai);
            aq = __fmaf_rn(t, sq, aq);
        } else {
            ai = __fadd_rn(ai, __fmul_rn(t, si));
            aq = __fadd_rn(aq, __fmul_rn(t, sq));
        }
    }
    out[(size_t)ch * out_pitch + o] = make_float2(ai, aq);
}

/* history := last 30 inputs of the stream (old history, then the launch's N inputs), one warp per channel */
__global__ void __launch_bounds__(256)
hb_state_update_kernel(const float2* in, size_t in_pitch, float2* hist_all, int n_channels, long N) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * (blockDim.x >> 5) + warp;
    if (ch >= n_channels) {
        return;
    }
    float2* hist = hist_all + (size_t)ch * kHbHist;
    const float2* x = in + (size_t)ch * in_pitch;
    float2 v = make_float2(0.0f, 0.0f);
    if (lane < kHbHist) {
        const long idx = N + lane; /* position in [hist(30) | in(N)] */
        v = (idx < kHbHist) ? hist[idx] : x[idx - kHbHist];
    }
    __syncwarp();
    if (lane < kHbHist) {
        hist[lane] = v;
    }
}

}  // namespace

struct dsdneo_b200_hb_cascade {
    int n_channels, passes, fir_arith;
    float2* d_hist;   /* [passes][n_channels][30] */
    float2* d_work[2];
    size_t work_pitch; /* pairs per channel in each scratch buffer */
    float* d_stage_in;
    float* d_stage_out;
    size_t stage_in_cap, stage_out_cap;
};

extern "C" {

dsdneo_b200_hb_cascade*
dsdneo_b200_hb_cascade_create(int n_channels, int passes, int fir_arith) {
    if (n_channels <= 0 || n_channels > 65535 || passes < 1 || passes > kHbMaxPasses) {
        set_error("hb_cascade_create: need 1..65535 channels and 1..%d passes", kHbMaxPasses);
        return NULL;
    }
    if (ensure_device()) {
        return NULL;
    }
    dsdneo_b200_hb_cascade* h = (dsdneo_b200_hb_cascade*)calloc(1, sizeof(*h));
    if (!h) {
        set_error("hb_cascade_create: out of host memory");
        return NULL;
    }
    h->n_channels = n_channels;
    h->passes = passes;
    h->fir_arith = (fir_arith == DSDNEO_FIR_ARITH_NOFMA) ? DSDNEO_FIR_ARITH_NOFMA : DSDNEO_FIR_ARITH_FMA;
    const size_t bytes = (size_t)passes * n_channels * kHbHist * sizeof(float2);
    cudaError_t e = cudaMalloc((void**)&h->d_hist, bytes);
    if (e == cudaSuccess) {
        e = cudaMemset(h->d_hist, 0, bytes);
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "hb_cascade_create", __FILE__, __LINE__);
        dsdneo_b200_hb_cascade_destroy(h);
        return NULL;
    }
    return h;
}

void
dsdneo_b200_hb_cascade_destroy(dsdneo_b200_hb_cascade* h) {
    if (!h) {
        return;
    }
    cudaFree(h->d_hist);
    cudaFree(h->d_work[0]);
    cudaFree(h->d_work[1]);
    cudaFree(h->d_stage_in);
    cudaFree(h->d_stage_out);
    free(h);
}

int
dsdneo_b200_hb_cascade_reset(dsdneo_b200_hb_cascade* h, void* stream) {
    if (!h) {
        set_error("hb_cascade_reset: NULL");
        return DSDNEO_B200_EINVAL;
    }
    DSDNEO_CUDA(cudaMemsetAsync(h->d_hist, 0, (size_t)h->passes * h->n_channels * kHbHist * sizeof(float2), as_stream(stream)));
    return 0;
}

int
dsdneo_b200_hb_cascade_decim_batch(dsdneo_b200_hb_cascade* h, const float* d_in, size_t in_pitch_pairs, int block_pairs,
                                   int n_blocks, float* d_out, size_t out_pitch_pairs, void* stream) {
    if (!h || !d_in || !d_out || block_pairs < 1 || n_blocks < 1) {
        set_error("hb_cascade_decim_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (block_pairs % (1 << h->passes) != 0) {
        set_error("hb_cascade_decim_batch: block_pairs %d is not a multiple of 2^passes (%d)", block_pairs, 1 << h->passes);
        return DSDNEO_B200_EUNSUPPORTED;
    }
    const size_t n_in = (size_t)block_pairs * n_blocks;
    if (in_pitch_pairs < n_in || out_pitch_pairs < (n_in >> h->passes)) {
        set_error("hb_cascade_decim_batch: pitch smaller than the data");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    if (h->passes > 1 && h->work_pitch < n_in / 2) {
        DSDNEO_CUDA(cudaDeviceSynchronize());
        for (int i = 0; i < 2; i++) {
            cudaFree(h->d_work[i]);
            h->d_work[i] = NULL;
            DSDNEO_CUDA(cudaMalloc((void**)&h->d_work[i], (size_t)h->n_channels * (n_in / 2) * sizeof(float2)));
        }
        h->work_pitch = n_in / 2;
    }
    const float2* src = reinterpret_cast<const float2*>(d_in);
    size_t src_pitch = in_pitch_pairs;
    int blk = block_pairs;
    for (int i = 0; i < h->passes; i++) {
        const bool last = (i == h->passes - 1);
        float2* dst = last ? reinterpret_cast<float2*>(d_out) : h->d_work[i & 1];
        const size_t dst_pitch = last ? out_pitch_pairs : h->work_pitch;
        float2* hist = h->d_hist + (size_t)i * h->n_channels * kHbHist;
        const int taps_len = (i == 0) ? 31 : 15;
        /* short blocks take the scalar kernel even on AVX2 hosts (simd_fir.cpp:302-305) */
        const bool fma = (h->fir_arith == DSDNEO_FIR_ARITH_FMA) && (blk >= taps_len);
        const long total_out = (long)(blk >> 1) * n_blocks;
        dim3 grid((unsigned)((total_out + 255) / 256), (unsigned)h->n_channels);
        {
            KernelTimer kt("hb_decim2_kernel", s);
            if (i == 0) {
                if (fma) {
                    hb_decim2_kernel<31, true><<<grid, 256, 0, s>>>(src, src_pitch, dst, dst_pitch, hist, blk, n_blocks);
                } else {
                    hb_decim2_kernel<31, false><<<grid, 256, 0, s>>>(src, src_pitch, dst, dst_pitch, hist, blk, n_blocks);
                }
            } else {
                if (fma) {
                    hb_decim2_kernel<15, true><<<grid, 256, 0, s>>>(src, src_pitch, dst, dst_pitch, hist, blk, n_blocks);
                } else {
                    hb_decim2_kernel<15, false><<<grid, 256, 0, s>>>(src, src_pitch, dst, dst_pitch, hist, blk, n_blocks);
                }
            }
        }
        DSDNEO_KERNEL_CHECK();
        hb_state_update_kernel<<<(h->n_channels + 7) / 8, 256, 0, s>>>(src, src_pitch, hist, h->n_channels, (long)blk * n_blocks);
        DSDNEO_KERNEL_CHECK();
        count_launch(2);
        src = dst;
        src_pitch = dst_pitch;
        blk >>= 1;
    }
    return 0;
}

int
dsdneo_b200_hb_cascade_decim_batch_host(dsdneo_b200_hb_cascade* h, const float* h_in, size_t in_pitch_pairs, int block_pairs,
                                        int n_blocks, float* h_out, size_t out_pitch_pairs) {
    if (!h || !h_in || !h_out) {
        set_error("hb_cascade_decim_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t in_floats = (size_t)h->n_channels * in_pitch_pairs * 2;
    const size_t out_floats = (size_t)h->n_channels * out_pitch_pairs * 2;
    if (h->stage_in_cap < in_floats) {
        cudaFree(h->d_stage_in);
        h->d_stage_in = NULL;
        h->stage_in_cap = 0;
        DSDNEO_CUDA(cudaMalloc((void**)&h->d_stage_in, in_floats * sizeof(float)));
        h->stage_in_cap = in_floats;
    }
    if (h->stage_out_cap < out_floats) {
        cudaFree(h->d_stage_out);
        h->d_stage_out = NULL;
        h->stage_out_cap = 0;
        DSDNEO_CUDA(cudaMalloc((void**)&h->d_stage_out, out_floats * sizeof(float)));
        h->stage_out_cap = out_floats;
    }
    DSDNEO_CUDA(cudaMemcpyAsync(h->d_stage_in, h_in, in_floats * sizeof(float), cudaMemcpyHostToDevice, 0));
    rc = dsdneo_b200_hb_cascade_decim_batch(h, h->d_stage_in, in_pitch_pairs, block_pairs, n_blocks, h->d_stage_out,
                                            out_pitch_pairs, NULL);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpyAsync(h_out, h->d_stage_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost, 0));
    DSDNEO_CUDA(cudaStreamSynchronize(0));
    return 0;
}

} /* extern "C" */
