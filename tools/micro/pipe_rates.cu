// Developer microbenchmark: issue rates of scalar vs packed FP32 ops and LDS.64 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int iters, long long* cyc) {
    float2 a[8];
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.001f, -0.001f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __ffma2_rn(a[i], m, c);
            if (MODE == 1) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }
            if (MODE == 2) a[i] = __fadd2_rn(a[i], c);
            if (MODE == 3) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); }
            if (MODE == 4) { a[i] = __fadd2_rn(a[i], c); a[i] = __ffma2_rn(a[i], m, c); }
            if (MODE == 5) { a[i].x = __fadd_rn(a[i].x, c.x); a[i].y = __fadd_rn(a[i].y, c.y); a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float2* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const char* names[] = {"FFMA2", "2xFFMA", "FADD2", "2xFADD", "FADD2+FFMA2", "2xFADD+2xFFMA"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 6; mode++) {
            int iters = 2000; long long h = 0;
            dim3 g(148), b(warps * 32);
            for (int rep = 0; rep < 2; rep++) {
                switch (mode) {
                    case 0: k<0><<<g, b>>>(out, iters, cyc); break; case 1: k<1><<<g, b>>>(out, iters, cyc); break;
                    case 2: k<2><<<g, b>>>(out, iters, cyc); break; case 3: k<3><<<g, b>>>(out, iters, cyc); break;
                    case 4: k<4><<<g, b>>>(out, iters, cyc); break; case 5: k<5><<<g, b>>>(out, iters, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double ops_per_iter = 8.0 * ((mode >= 4) ? 2 : 1);  // packed-equivalent ops (each = 2 flop-lanes)
            printf("warps/SM=%2d %-14s cycles/iter=%.1f  -> %.2f packed-equiv ops/clk/SM (float lanes/clk/SM=%.1f)\n", warps,
                   names[mode], (double)h / iters, ops_per_iter * warps / ((double)h / iters),
                   ops_per_iter * warps * 64 / ((double)h / iters));
        }
    }
    return 0;
}
