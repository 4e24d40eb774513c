// Developer micro-benchmark: cycles per sample of the dc_est chain (FADD -> FMUL -> FADD, fsk_modem.c:96-103) run by ONE
// warp, with and without the shared-memory traffic of the real kernel.  nvcc -arch=sm_100a -fmad=false -O3.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kPitch = 132, kChunk = 128;

template <int MODE>  // 0: registers only, 1: + LDS.128, 2: + LDS.128 + STS.128, 3: scalar LDS/STS
__global__ void __launch_bounds__(256) chain(float* out, long long* cyc, int iters, int extra_warps_spin) {
    __shared__ __align__(16) float in[32 * kPitch];
    __shared__ __align__(16) float cb[32 * kPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * kPitch; i += blockDim.x) in[i] = 0.01f * (float)((i * 37) % 101) - 0.5f;
    __syncthreads();
    if (warp != 0) {
        if (extra_warps_spin) {  // other warps keep their schedulers busy
            float a = (float)threadIdx.x;
            for (int i = 0; i < iters * 400; i++) a = a * 1.0001f + 0.5f;
            if (a == 123.0f) out[threadIdx.x] = a;
        }
        return;
    }
    float dc = 0.1f * (float)lane;
    const float4* ib4 = reinterpret_cast<const float4*>(in + lane * kPitch);
    float4* cb4 = reinterpret_cast<float4*>(cb + lane * kPitch);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            float f = 0.25f;
#pragma unroll 16
            for (int q = 0; q < kChunk; q++) {
                dc = dc + 0.00025f * (f - dc);
                f = f + 0.001f;  // independent of the chain
            }
        } else if (MODE == 3) {
            const float* ib = in + lane * kPitch;
            float* cbp = cb + lane * kPitch;
#pragma unroll 16
            for (int q = 0; q < kChunk; q++) {
                const float f = ib[q];
                dc = dc + 0.00025f * (f - dc);
                cbp[q] = f - dc;
            }
        } else {
            float4 a0 = ib4[0], a1 = ib4[1];
            for (int q = 0; q < kChunk / 4; q += 4) {
                const float4 b0 = ib4[q + 2], b1 = ib4[q + 3];
                float4 c0, c1;
#define STEP(dst, src) dc = dc + 0.00025f * ((src) - dc); dst = (src) - dc;
                STEP(c0.x, a0.x) STEP(c0.y, a0.y) STEP(c0.z, a0.z) STEP(c0.w, a0.w)
                STEP(c1.x, a1.x) STEP(c1.y, a1.y) STEP(c1.z, a1.z) STEP(c1.w, a1.w)
                if (MODE == 2) { cb4[q] = c0; cb4[q + 1] = c1; } else { dc += 1e-30f * (c0.x + c1.w); }
                a0 = ib4[q + 4]; a1 = ib4[q + 5];
                STEP(c0.x, b0.x) STEP(c0.y, b0.y) STEP(c0.z, b0.z) STEP(c0.w, b0.w)
                STEP(c1.x, b1.x) STEP(c1.y, b1.y) STEP(c1.z, b1.z) STEP(c1.w, b1.w)
                if (MODE == 2) { cb4[q + 2] = c0; cb4[q + 3] = c1; } else { dc += 1e-30f * (c0.x + c1.w); }
            }
        }
    }
    const long long t1 = clock64();
    out[lane] = dc + cb[lane];
    if (lane == 0) cyc[0] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    const int iters = 200;
    for (int spin = 0; spin < 2; spin++) {
        for (int mode = 0; mode < 4; mode++) {
            for (int nthreads : {32, 256}) {
                long long h = 0;
                for (int rep = 0; rep < 2; rep++) {
                    if (mode == 0) chain<0><<<1, nthreads>>>(out, cyc, iters, spin);
                    if (mode == 1) chain<1><<<1, nthreads>>>(out, cyc, iters, spin);
                    if (mode == 2) chain<2><<<1, nthreads>>>(out, cyc, iters, spin);
                    if (mode == 3) chain<3><<<1, nthreads>>>(out, cyc, iters, spin);
                    cudaDeviceSynchronize();
                    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                }
                printf("mode %d threads %3d spin %d: %.2f cycles/sample\n", mode, nthreads, spin, (double)h / (iters * kChunk));
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
