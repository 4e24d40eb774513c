// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * MBE speech synthesis stage (K21), batched: decoded frame parameters (w0, L, Vl, Ml) -> 160 PCM samples, one warp per
 * frame.
 *
 * PARITY UNPINNED.  dsd-neo reaches this stage through mbelib-neo 2.x (mbe_processImbe4400Dataf / mbe_processAmbe2450Dataf,
 * call sites src/core/vocoder/dsd_mbe.c:268,296,581,617,685); that library is not vendored in the reference tree and is
 * absent here, so there is nothing to compare bits against.  This kernel and oracle/oracle_mbe.c both restate the
 * published algorithm of its ancestor mbelib 1.3.0 (mbe_spectralAmpEnhance, mbe_synthesizeSpeechf, mbe_floattoshort;
 * TIA-102.BABA eq. 105-111, 127-141): see the oracle's header for the scope and the one deliberate difference (random
 * phases / noise are a counter-based hash of (frame key, band, sample, index) instead of libc rand()).
 * The bit-level parameter decode needs the codec's quantiser tables and is not part of this stage.
 *
 * Mapping: the frame's two parameter sets live in shared memory; bands are walked in order (their branch -- voiced /
 * unvoiced transitions -- is warp-uniform), every lane owns samples n = lane + 32 j (j = 0..4) and accumulates them in
 * registers in the reference's band order, so CPU and GPU differ only through cosf/powf/logf (tests: +-1 LSB of int16).
 * Algorithmic bytes: 2 x 1.2 kB parameter sets in, 640 B float (+ 320 B int16) PCM out, parameter sets written back.
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

constexpr int kMbeWarps = 4; /* frames per CTA */
constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float
mbe_ws(int idx) {
    int n = idx - 160;
    n = n < 0 ? -n : n;
    return n > 105 ? 0.0f : (n <= 55 ? 1.0f : (float)(105 - n) * 0.02f);
}

__device__ __forceinline__ float
mbe_uniform(unsigned long long key, int band, int sample, int index, int stream) {
    unsigned long long z = key + 0x9E3779B97F4A7C15ull * (unsigned long long)(1 + band)
                           + 0xBF58476D1CE4E5B9ull * (unsigned long long)(1 + sample)
                           + 0x94D049BB133111EBull * (unsigned long long)(1 + index)
                           + 0xD6E8FEB86659FD93ull * (unsigned long long)(1 + stream);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float
mbe_rand_phase(unsigned long long key, int band, int index, int stream) {
    return mbe_uniform(key, band, -1, index, stream) * 6.2831853071795864769f - kPi;
}

__device__ __forceinline__ float
warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}

/* one unvoiced band component: `uvq`-tone multisine around harmonic l of w0 plus noise above 2700 Hz */
__device__ __forceinline__ float
mbe_unvoiced(unsigned long long key, int l, int n, float w0, float w0l, int uvq, float uvstep, float uvoffset, float uvthreshold,
             int phase_stream, int noise_stream) {
    float c = 0.0f;
    for (int i = 0; i < uvq; i++) {
        c = c + cosf((w0 * (float)n * ((float)l + ((float)i * uvstep) - uvoffset)) + mbe_rand_phase(key, l, i, phase_stream));
        if (w0l > uvthreshold) {
            c = c + ((w0l - uvthreshold) * 2.0f * mbe_uniform(key, l, n, i, noise_stream));
        }
    }
    return c;
}

__global__ void __launch_bounds__(32 * kMbeWarps)
mbe_synth_kernel(dsdneo_b200_mbe_parms* cur_all, dsdneo_b200_mbe_parms* prev_all, const unsigned long long* keys, int uvquality,
                 float* pcm_f, int16_t* pcm_s, int n_frames) {
    __shared__ dsdneo_b200_mbe_parms s_cur[kMbeWarps], s_prev[kMbeWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kMbeWarps + warp;
    if (f >= n_frames) {
        return;
    }
    dsdneo_b200_mbe_parms& cur = s_cur[warp];
    dsdneo_b200_mbe_parms& prev = s_prev[warp];
    {
        const int words = (int)(sizeof(dsdneo_b200_mbe_parms) / 4);
        const unsigned* gc = reinterpret_cast<const unsigned*>(cur_all + f);
        const unsigned* gp = reinterpret_cast<const unsigned*>(prev_all + f);
        unsigned* sc = reinterpret_cast<unsigned*>(&cur);
        unsigned* sp = reinterpret_cast<unsigned*>(&prev);
        for (int i = lane; i < words; i += 32) {
            sc[i] = gc[i];
            sp[i] = gp[i];
        }
    }
    __syncwarp();
    const unsigned long long key = keys ? keys[f] : (unsigned long long)f;
    int L = cur.L;
    L = L < 0 ? 0 : (L > 56 ? 56 : L);

    /* ---- mbe_spectralAmpEnhance ---- */
    {
        float r0 = 0.0f, r1 = 0.0f;
        for (int l = lane + 1; l <= L; l += 32) {
            const float m2 = cur.Ml[l] * cur.Ml[l];
            r0 += m2;
            r1 += m2 * cosf(cur.w0 * (float)l);
        }
        const float Rm0 = warp_sum(r0), Rm1 = warp_sum(r1);
        const float R2m0 = Rm0 * Rm0, R2m1 = Rm1 * Rm1;
        float part = 0.0f;
        for (int l = lane + 1; l <= L; l += 32) {
            float M = cur.Ml[l];
            if (M != 0.0f) {
                const float W = sqrtf(M)
                                * powf(((0.96f * kPi * ((R2m0 + R2m1) - (2.0f * Rm0 * Rm1 * cosf(cur.w0 * (float)l))))
                                        / (cur.w0 * Rm0 * (R2m0 - R2m1))),
                                       0.25f);
                if ((8 * l) <= L) {
                } else if (W > 1.2f) {
                    M = 1.2f * M;
                } else if (W < 0.5f) {
                    M = 0.5f * M;
                } else {
                    M = W * M;
                }
            }
            cur.Ml[l] = M;
            part += M * M;
        }
        const float sum = warp_sum(part);
        const float gamma = (sum == 0.0f) ? 1.0f : sqrtf(Rm0 / sum);
        __syncwarp();
        for (int l = lane + 1; l <= L; l += 32) {
            cur.Ml[l] = gamma * cur.Ml[l];
        }
        __syncwarp();
    }

    /* ---- mbe_synthesizeSpeechf ---- */
    constexpr int N = 160;
    const float uvthreshold = (2700.0f * kPi) / 4000.0f;
    const float uvsine = 1.3591409f * 2.7182818284590452354f;
    const int uvq = (uvquality < 1 || uvquality > 64) ? 3 : uvquality;
    const float qfactor = (uvq == 1) ? (1.0f / 2.7182818284590452354f) : (logf((float)uvq) / (float)uvq);
    const float uvstep = 1.0f / (float)uvq;
    const float uvoffset = (uvstep * (float)(uvq - 1)) / 2.0f;
    int nu = 0;
    for (int l = lane + 1; l <= L; l += 32) {
        nu += cur.Vl[l] == 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        nu += __shfl_xor_sync(0xffffffffu, nu, o);
    }
    const float cw0 = cur.w0, pw0 = prev.w0;
    int pL = prev.L;
    pL = pL < 0 ? 0 : (pL > 56 ? 56 : pL);
    const int maxl = L > pL ? L : pL;
    for (int l = lane + 1; l <= maxl; l += 32) { /* eq. 128, 129 */
        if (l > pL) {
            prev.Ml[l] = 0.0f;
            prev.Vl[l] = 1;
        }
        if (l > L) {
            cur.Ml[l] = 0.0f;
            cur.Vl[l] = 1;
        }
    }
    for (int l = lane + 1; l <= 56; l += 32) { /* eq. 139, 140 */
        const float psi = prev.PSIl[l] + ((pw0 + cw0) * ((float)(l * N) / 2.0f));
        cur.PSIl[l] = psi;
        cur.PHIl[l] = (l <= (int)(L / 4)) ? psi : psi + (((float)nu * mbe_rand_phase(key, l, 0, 0)) / (float)cur.L);
    }
    __syncwarp();
    float acc[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int l = 1; l <= maxl; l++) {
        const float cw0l = cw0 * (float)l, pw0l = pw0 * (float)l;
        const int cv = cur.Vl[l], pv = prev.Vl[l];
        const float cM = cur.Ml[l], pM = prev.Ml[l], cPHI = cur.PHIl[l], pPHI = prev.PHIl[l];
        if (cv == 0 && pv == 1) {
#pragma unroll 1
            for (int j = 0; j < 5; j++) {
                const int n = lane + 32 * j;
                const float C1 = mbe_ws(n + N) * pM * cosf((pw0l * (float)n) + pPHI); /* eq. 131 */
                float C3 = mbe_unvoiced(key, l, n, cw0, cw0l, uvq, uvstep, uvoffset, uvthreshold, 1, 3);
                C3 = C3 * uvsine * mbe_ws(n) * cM * qfactor;
                acc[j] = acc[j] + C1 + C3;
            }
        } else if (cv == 1 && pv == 0) {
#pragma unroll 1
            for (int j = 0; j < 5; j++) {
                const int n = lane + 32 * j;
                const float C1 = mbe_ws(n) * cM * cosf((cw0l * (float)(n - N)) + cPHI); /* eq. 132 */
                float C3 = mbe_unvoiced(key, l, n, pw0, pw0l, uvq, uvstep, uvoffset, uvthreshold, 1, 3);
                C3 = C3 * uvsine * mbe_ws(n + N) * pM * qfactor;
                acc[j] = acc[j] + C1 + C3;
            }
        } else if (cv == 1 || pv == 1) {
#pragma unroll
            for (int j = 0; j < 5; j++) { /* eq. 133 */
                const int n = lane + 32 * j;
                const float C1 = mbe_ws(n + N) * pM * cosf((pw0l * (float)n) + pPHI);
                const float C2 = mbe_ws(n) * cM * cosf((cw0l * (float)(n - N)) + cPHI);
                acc[j] = acc[j] + C1 + C2;
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 5; j++) {
                const int n = lane + 32 * j;
                float C3 = mbe_unvoiced(key, l, n, pw0, pw0l, uvq, uvstep, uvoffset, uvthreshold, 1, 3);
                C3 = C3 * uvsine * mbe_ws(n + N) * pM * qfactor;
                float C4 = mbe_unvoiced(key, l, n, cw0, cw0l, uvq, uvstep, uvoffset, uvthreshold, 2, 4);
                C4 = C4 * uvsine * mbe_ws(n) * cM * qfactor;
                acc[j] = acc[j] + C3 + C4;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const int n = lane + 32 * j;
        pcm_f[(size_t)f * N + n] = acc[j];
        if (pcm_s) { /* mbe_floattoshort */
            float v = acc[j] * 7.0f;
            v = v > 32760.0f ? 32760.0f : (v < -32760.0f ? -32760.0f : v);
            pcm_s[(size_t)f * N + n] = (int16_t)v;
        }
    }
    /* mbe_moveMbeParms(cur, prev_enhanced): both arrays hold the enhanced, phase-updated current frame */
    __syncwarp();
    {
        const int words = (int)(sizeof(dsdneo_b200_mbe_parms) / 4);
        unsigned* gc = reinterpret_cast<unsigned*>(cur_all + f);
        unsigned* gp = reinterpret_cast<unsigned*>(prev_all + f);
        const unsigned* sc = reinterpret_cast<const unsigned*>(&cur);
        for (int i = lane; i < words; i += 32) {
            gc[i] = sc[i];
            gp[i] = sc[i];
        }
    }
}

}  // namespace

extern "C" {

int
dsdneo_b200_mbe_synth_batch(dsdneo_b200_mbe_parms* d_cur, dsdneo_b200_mbe_parms* d_prev_enhanced, const uint64_t* d_keys,
                            int uvquality, float* d_pcm_f, int16_t* d_pcm_s, int n_frames, void* stream) {
    if (!d_cur || !d_prev_enhanced || !d_pcm_f || n_frames < 0) {
        set_error("mbe_synth_batch: bad argument");
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
        KernelTimer kt("mbe_synth_kernel", s);
        mbe_synth_kernel<<<(n_frames + kMbeWarps - 1) / kMbeWarps, 32 * kMbeWarps, 0, s>>>(
            d_cur, d_prev_enhanced, reinterpret_cast<const unsigned long long*>(d_keys), uvquality, d_pcm_f, d_pcm_s, n_frames);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_mbe_synth_batch_host(dsdneo_b200_mbe_parms* h_cur, dsdneo_b200_mbe_parms* h_prev_enhanced, const uint64_t* h_keys,
                                 int uvquality, float* h_pcm_f, int16_t* h_pcm_s, int n_frames) {
    if (!h_cur || !h_prev_enhanced || !h_pcm_f || n_frames < 0) {
        set_error("mbe_synth_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const size_t n = (size_t)n_frames, pb = n * sizeof(dsdneo_b200_mbe_parms);
    void *dc = NULL, *dp = NULL, *dk = NULL, *df = NULL, *ds = NULL;
    cudaError_t e = cudaMalloc(&dc, pb);
    if (e == cudaSuccess) e = cudaMalloc(&dp, pb);
    if (e == cudaSuccess && h_keys) e = cudaMalloc(&dk, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&df, n * 160 * 4);
    if (e == cudaSuccess && h_pcm_s) e = cudaMalloc(&ds, n * 160 * 2);
    if (e == cudaSuccess) e = cudaMemcpy(dc, h_cur, pb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dp, h_prev_enhanced, pb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && h_keys) e = cudaMemcpy(dk, h_keys, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = dsdneo_b200_mbe_synth_batch((dsdneo_b200_mbe_parms*)dc, (dsdneo_b200_mbe_parms*)dp, (const uint64_t*)dk, uvquality,
                                         (float*)df, (int16_t*)ds, n_frames, NULL);
    }
    if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_cur, dc, pb, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_prev_enhanced, dp, pb, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == 0) e = cudaMemcpy(h_pcm_f, df, n * 160 * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == 0 && h_pcm_s) e = cudaMemcpy(h_pcm_s, ds, n * 160 * 2, cudaMemcpyDeviceToHost);
    cudaFree(dc);
    cudaFree(dp);
    cudaFree(dk);
    cudaFree(df);
    cudaFree(ds);
    if (e != cudaSuccess) {
        return cuda_fail(e, "mbe_synth_batch_host", __FILE__, __LINE__);
    }
    return rc;
}

} /* extern "C" */
