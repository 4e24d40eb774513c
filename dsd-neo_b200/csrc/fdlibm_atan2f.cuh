// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * atan2f()/atanf() with the exact operation sequence of the fdlibm-derived single-precision
 * routines shipped by glibc 2.39 (the libm the reference links on this image), so the device
 * discriminator's large-angle branch (src/dsp/fsk_modem.c:23-35 -> atan2f) returns the same
 * bits as the CPU reference.  Pure IEEE-754 binary32 add/mul/div in a fixed order: this file
 * must be compiled with FMA contraction off (nvcc -fmad=false, gcc -ffp-contract=off).
 *
 * Pinned by tests/test_oracle_dsp.py::test_atan2f_matches_libm (host build vs libm, 2e7 random
 * inputs incl. raw bit patterns; 2e8 checked during development) and
 * tests/test_gpu_demod.py::test_device_atan2f (device build vs libm).
 */
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DSDNEO_HD __host__ __device__ __forceinline__
#else
#define DSDNEO_HD static inline
#endif

namespace dsdneo {

DSDNEO_HD uint32_t
f32_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

DSDNEO_HD float
bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

DSDNEO_HD float
fd_atanf(float x) {
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const int32_t hx = (int32_t)f32_bits(x);
    const int32_t ix = hx & 0x7fffffff;
    float hi = 0.0f, lo = 0.0f;
    bool reduced = true;
    if (ix >= 0x4c000000) { /* |x| >= 2^25 (or NaN) */
        if (ix > 0x7f800000) {
            return x + x;
        }
        return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    }
    if (ix < 0x3ee00000) { /* |x| < 7/16: no reduction */
        if (ix < 0x31000000) {
            return x;
        }
        reduced = false;
    } else {
        x = bits_f32((uint32_t)ix);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) {
                hi = hi0, lo = lo0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                hi = hi1, lo = lo1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else if (ix < 0x401c0000) {
            hi = hi2, lo = lo2;
            x = (x - 1.5f) / (1.0f + 1.5f * x);
        } else {
            hi = hi3, lo = lo3;
            x = -1.0f / x;
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 =
        z
        * (3.3333334327e-01f
           + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f + w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
    const float s2 =
        w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f + w * -3.6531571299e-02f))));
    if (!reduced) {
        return x - x * (s1 + s2);
    }
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -r : r;
}

DSDNEO_HD float
fd_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)f32_bits(x), hy = (int32_t)f32_bits(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) {
        return x + y;
    }
    if (hx == 0x3f800000) {
        return fd_atanf(y);
    }
    const int quad = ((hy >> 31) & 1) | ((hx >> 30) & 2); /* 2*sign(x) + sign(y) */
    if (iy == 0) {
        return quad < 2 ? y : (quad == 2 ? pi + tiny : -pi - tiny);
    }
    if (ix == 0) {
        return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    }
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            return quad == 0 ? pi_o_4 + tiny : quad == 1 ? -pi_o_4 - tiny : quad == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny;
        }
        return quad == 0 ? 0.0f : quad == 1 ? -0.0f : quad == 2 ? pi + tiny : -pi - tiny;
    }
    if (iy == 0x7f800000) {
        return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    }
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) {
        z = pi_o_2 + 0.5f * pi_lo;
    } else if (hx < 0 && k < -60) {
        z = 0.0f;
    } else {
        z = fd_atanf(bits_f32(f32_bits(y / x) & 0x7fffffffu));
    }
    if (quad == 0) {
        return z;
    }
    if (quad == 1) {
        return bits_f32(f32_bits(z) ^ 0x80000000u);
    }
    if (quad == 2) {
        return pi - (z - pi_lo);
    }
    return (z - pi_lo) - pi;
}

}  // namespace dsdneo
