// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * Vocoder frame ECC and the DMR base-station voice burst cutter (BASELINE config C4).
 *
 * 1. dmr_voice_cut_kernel -- the collection phase of dmrBSBootstrap / dmrBS (src/protocol/dmr/dmr_bs.c:137-148 unpack through
 *    dsd_ambe_2450_dibit_map, :150-170 sync segment, :182-187 CACH, :711-722 inversion of the 90 buffered dibits of the first
 *    burst, :745-746 and :838-848 segment offsets): from every BS VOICE sync hit of every channel, `n_bursts` consecutive
 *    144-dibit bursts (the two TDMA slots alternate) are cut into CACH bits, three ambe_fr[4][24] frames and the 48 sync / EMB
 *    bits.  The interleave schedule is generated (it is regular: even dibits walk C0 then C1 downwards, odd dibits C1 / C2
 *    then C2 / C3); tests compare it with the reference's table.
 * 2. ambe3600x2450_decode_kernel / imbe7200x4400_decode_kernel -- ambe_fr[4][24] -> ambe_d[49] and imbe_fr[8][23] ->
 *    imbe_d[88] with the error counts the reference stores in state->errs / errs2: what mbe_decodeAmbe3600x2450Frame /
 *    mbe_decodeImbe7200x4400Frame do (call sites src/core/vocoder/dsd_mbe.c:168,188).  PARITY UNPINNED: that code lives in the
 *    un-vendored mbelib-neo; this follows the published algorithm of mbelib 1.3.0 / TIA-102.BABA section 7 ([23,12] Golay on
 *    C0, pseudo-random demodulation seeded with C0's data, Golay / [15,11] Hamming on the protected words, bit
 *    prioritisation order of the output vector).  Both codes are perfect, so their syndrome decoders are implementation
 *    independent; the tables are derived here (x^(11+i) mod 0xC75, coset leaders, check-mask columns), not recalled.
 *    One thread per frame; tables in global memory (initialised once per device).
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

__device__ uint16_t g_golay_gen[12];
__device__ uint16_t g_golay_fix[2048];
__device__ uint16_t g_hamming_fix[16];
__constant__ uint16_t kHammingCheck[4] = {0x7f08, 0x78e4, 0x66d2, 0x55b1};
const uint16_t kHammingCheckHost[4] = {0x7f08, 0x78e4, 0x66d2, 0x55b1};

__device__ __forceinline__ unsigned
golay_syndrome(unsigned block23) {
    unsigned ecc = 0;
    unsigned data = block23 >> 11;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        ecc ^= ((data >> (11 - i)) & 1u) ? (unsigned)g_golay_gen[i] : 0u;
    }
    return ecc ^ (block23 & 0x7ffu);
}

/* corrected 12 data bits of a 23-bit word; *errs += number of data bits changed (mbe_golay2312's return value) */
__device__ __forceinline__ unsigned
golay2312(unsigned block23, int* errs) {
    const unsigned fix = g_golay_fix[golay_syndrome(block23)];
    *errs += __popc(fix);
    return (block23 >> 11) ^ fix;
}

__device__ __forceinline__ unsigned
hamming1511(unsigned block15, int* errs) {
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        s = (s << 1) | (__popc(block15 & kHammingCheck[i]) & 1u);
    }
    if (s) {
        *errs += 1;
        block15 ^= g_hamming_fix[s];
    }
    return block15;
}

/* word of `n` bits from a byte row: bit j of the result = row[j] & 1 */
__device__ __forceinline__ unsigned
pack_row(const uint8_t* row, int n) {
    unsigned w = 0;
    for (int j = 0; j < n; j++) {
        w |= (unsigned)(row[j] & 1u) << j;
    }
    return w;
}

/* pseudo-random modulation word: bit (n-1-k) of the result = m(k+1), k = 0..n-1, i.e. the first PN bit meets the word's
 * top bit (the words are demodulated from index n-1 downwards) */
__device__ __forceinline__ unsigned
pn_word(unsigned& pr, int n) {
    unsigned w = 0;
    for (int k = 0; k < n; k++) {
        pr = (173u * pr + 13849u) & 0xffffu;
        w = (w << 1) | (pr >> 15);
    }
    return w;
}

__global__ void __launch_bounds__(128)
ambe3600x2450_decode_kernel(const uint8_t* __restrict__ ambe_fr, uint8_t* __restrict__ ambe_d, int32_t* __restrict__ c0_errs,
                            int32_t* __restrict__ total_errs, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) {
        return;
    }
    const uint8_t* fr = ambe_fr + (size_t)f * 96;
    const unsigned c0 = pack_row(fr, 24) >> 1; /* the [23,12] part of the [24,12] word: columns 1..23 */
    const unsigned c1 = pack_row(fr + 24, 23);
    const unsigned c2 = pack_row(fr + 48, 11);
    const unsigned c3 = pack_row(fr + 72, 14);
    int errs = 0;
    const unsigned u0 = golay2312(c0, &errs);
    int errs2 = errs;
    unsigned pr = 16u * u0;
    const unsigned u1 = golay2312(c1 ^ pn_word(pr, 23), &errs2);
    uint8_t* o = ambe_d + (size_t)f * 49;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        o[j] = (uint8_t)((u0 >> (11 - j)) & 1u);
        o[12 + j] = (uint8_t)((u1 >> (11 - j)) & 1u);
    }
#pragma unroll
    for (int j = 0; j < 11; j++) {
        o[24 + j] = (uint8_t)((c2 >> (10 - j)) & 1u);
    }
#pragma unroll
    for (int j = 0; j < 14; j++) {
        o[35 + j] = (uint8_t)((c3 >> (13 - j)) & 1u);
    }
    c0_errs[f] = errs;
    total_errs[f] = errs2;
}

__global__ void __launch_bounds__(128)
imbe7200x4400_decode_kernel(const uint8_t* __restrict__ imbe_fr, uint8_t* __restrict__ imbe_d, int32_t* __restrict__ c0_errs,
                            int32_t* __restrict__ total_errs, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) {
        return;
    }
    const uint8_t* fr = imbe_fr + (size_t)f * 184;
    int errs = 0;
    const unsigned u0 = golay2312(pack_row(fr, 23), &errs);
    int errs2 = errs;
    unsigned pr = 16u * u0;
    uint8_t* o = imbe_d + (size_t)f * 88;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        o[j] = (uint8_t)((u0 >> (11 - j)) & 1u);
    }
    for (int i = 1; i < 4; i++) {
        const unsigned u = golay2312(pack_row(fr + 23 * i, 23) ^ pn_word(pr, 23), &errs2);
#pragma unroll
        for (int j = 0; j < 12; j++) {
            o[12 * i + j] = (uint8_t)((u >> (11 - j)) & 1u);
        }
    }
    for (int i = 4; i < 7; i++) {
        const unsigned w = hamming1511(pack_row(fr + 23 * i, 15) ^ pn_word(pr, 15), &errs2);
#pragma unroll
        for (int j = 0; j < 11; j++) {
            o[48 + 11 * (i - 4) + j] = (uint8_t)((w >> (14 - j)) & 1u);
        }
    }
    const unsigned c7 = pack_row(fr + 23 * 7, 7);
#pragma unroll
    for (int j = 0; j < 7; j++) {
        o[81 + j] = (uint8_t)((c7 >> (6 - j)) & 1u);
    }
    c0_errs[f] = errs;
    total_errs[f] = errs2;
}

/* the same on the receive bank's voice records (dsdneo_b200_p25p1_voice: nine frames, row r of frame v = the low 23 bits of
 * bits[v][r]); one thread per (record, frame) */
__global__ void __launch_bounds__(128)
imbe7200x4400_decode_packed_kernel(const uint32_t* __restrict__ voices, int words_per_record, uint8_t* __restrict__ imbe_d,
                                   int32_t* __restrict__ c0_errs, int32_t* __restrict__ total_errs, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) {
        return;
    }
    const uint32_t* w = voices + (size_t)(f / 9) * words_per_record + (f % 9) * 8;
    int errs = 0;
    const unsigned u0 = golay2312(w[0] & 0x7fffffu, &errs);
    int errs2 = errs;
    unsigned pr = 16u * u0;
    uint8_t* o = imbe_d + (size_t)f * 88;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        o[j] = (uint8_t)((u0 >> (11 - j)) & 1u);
    }
    for (int i = 1; i < 4; i++) {
        const unsigned u = golay2312((w[i] & 0x7fffffu) ^ pn_word(pr, 23), &errs2);
#pragma unroll
        for (int j = 0; j < 12; j++) {
            o[12 * i + j] = (uint8_t)((u >> (11 - j)) & 1u);
        }
    }
    for (int i = 4; i < 7; i++) {
        const unsigned h = hamming1511((w[i] & 0x7fffu) ^ pn_word(pr, 15), &errs2);
#pragma unroll
        for (int j = 0; j < 11; j++) {
            o[48 + 11 * (i - 4) + j] = (uint8_t)((h >> (14 - j)) & 1u);
        }
    }
#pragma unroll
    for (int j = 0; j < 7; j++) {
        o[81 + j] = (uint8_t)((w[7] >> (6 - j)) & 1u);
    }
    c0_errs[f] = errs;
    total_errs[f] = errs2;
}

/* dsd_ambe_2450_dibit_map[i] (include/dsd-neo/core/ambe_interleave.h:25-32), generated: packed (row << 5 | col) for the
 * dibit's high bit (low 8 bits of the result) and low bit (next 8 bits) */
__host__ __device__ __forceinline__ unsigned
ambe_map(int i) {
    int hr, hc, lr, lc;
    if ((i & 1) == 0) {
        const int e = i >> 1;
        hr = 0;
        hc = 23 - e;
        if (e <= 5) {
            lr = 0;
            lc = 5 - e;
        } else {
            lr = 1;
            lc = 28 - e;
        }
    } else {
        const int o = i >> 1;
        if (o <= 10) {
            hr = 1;
            hc = 10 - o;
        } else {
            hr = 2;
            hc = 21 - o;
        }
        if (o <= 3) {
            lr = 2;
            lc = 3 - o;
        } else {
            lr = 3;
            lc = 17 - o;
        }
    }
    return (unsigned)((hr << 5) | hc) | ((unsigned)((lr << 5) | lc) << 8);
}

/* one block per (channel, hit, burst): 144 threads, one per dibit */
__global__ void __launch_bounds__(160)
dmr_voice_cut_kernel(const uint8_t* dibits, size_t dibit_pitch, const int32_t* counts, const int32_t* hits, const int32_t* n_hits,
                     int n_channels, int max_hits, int n_bursts, int inverted, uint8_t* cach24, uint8_t* ambe_fr, uint8_t* sync48,
                     uint8_t* valid_out) {
    const int slot = blockIdx.x / n_bursts, j = blockIdx.x - slot * n_bursts;
    const int ch = slot / max_hits, h = slot - ch * max_hits;
    if (ch >= n_channels) {
        return;
    }
    const int k = threadIdx.x;
    const bool present = h < min(n_hits[ch], max_hits);
    const long live = present ? (long)hits[((size_t)ch * max_hits + h) * 2] + 1 : 0; /* dibit after the sync's last dibit */
    const long start = live - 90 + 144L * j;
    const bool ok = present && start >= 0 && start + 144 <= counts[ch];
    const size_t rec = (size_t)slot * n_bursts + j;
    if (k == 0) {
        valid_out[rec] = ok ? 1 : 0;
    }
    if (!present) {
        return; /* unused hit slot: its record stays as the caller initialised it */
    }
    uint8_t* fr = ambe_fr + rec * 288;
    /* the positions no dibit reaches stay 0, as after the reference's memset (dmr_bs.c:121-123) */
    for (int z = k; z < 288; z += blockDim.x) {
        const int col = z % 24, row = (z / 24) & 3;
        if ((row == 1 && col == 23) || (row == 2 && col > 10) || (row == 3 && col > 13)) {
            fr[z] = 0;
        }
    }
    if (k >= 144) {
        return;
    }
    int d = 0;
    if (ok) {
        d = dibits[(size_t)ch * dibit_pitch + start + k] & 3;
        if (inverted && j == 0 && k < 90) {
            d ^= 2; /* the buffered part of the first burst (dmr_bs.c:711-722) */
        }
    }
    const uint8_t b1 = (uint8_t)((d >> 1) & 1), b0 = (uint8_t)(d & 1);
    if (k < 12) {
        /* dmr_cach_interleave {0,7,8,9,1,10,11,12,2,13,14,15,3,16,4,17,18,19,5,20,21,22,6,23}, 5 bits per entry */
        const unsigned long long lo = 0x07B9A262D414A0E0ull, hi = 0x0B9AD5A167289203ull;
        const int e0 = 2 * k, e1 = 2 * k + 1;
        cach24[rec * 24 + (int)(((e0 < 12 ? lo : hi) >> (5 * (e0 % 12))) & 31ull)] = b1;
        cach24[rec * 24 + (int)(((e1 < 12 ? lo : hi) >> (5 * (e1 % 12))) & 31ull)] = b0;
    } else if (k >= 66 && k < 90) {
        sync48[rec * 48 + 2 * (k - 66)] = b1;
        sync48[rec * 48 + 2 * (k - 66) + 1] = b0;
    } else {
        int frame, idx;
        if (k < 48) {
            frame = 0;
            idx = k - 12;
        } else if (k < 66) {
            frame = 1;
            idx = k - 48;
        } else if (k < 108) {
            frame = 1;
            idx = 18 + (k - 90);
        } else {
            frame = 2;
            idx = k - 108;
        }
        const unsigned m = ambe_map(idx);
        uint8_t* f = fr + frame * 96;
        f[((m >> 5) & 7) * 24 + (m & 31)] = b1;
        f[((m >> 13) & 7) * 24 + ((m >> 8) & 31)] = b0;
    }
}

int
grid_for(int n, int threads) {
    return (n + threads - 1) / threads;
}

int
parity_host(unsigned v) {
    int p = 0;
    while (v) {
        p ^= 1;
        v &= v - 1;
    }
    return p;
}

/* derive the code tables (header comment) and load them on the current device, once per device */
int
ensure_tables() {
    static bool ready[64] = {};
    int dev = 0;
    DSDNEO_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && ready[dev]) {
        return 0;
    }
    uint16_t gen[12], hfix[16];
    uint16_t* fix = (uint16_t*)calloc(2048, sizeof(uint16_t));
    if (!fix) {
        set_error("mbe_ecc: out of host memory");
        return DSDNEO_B200_ENOMEM;
    }
    unsigned r = 0x475; /* x^11 mod g(x), g = 0xC75 */
    for (int i = 11; i >= 0; i--) {
        gen[i] = (uint16_t)r;
        r <<= 1;
        if (r & 0x800u) {
            r ^= 0xC75u;
        }
    }
    auto syn = [&](unsigned block) {
        unsigned ecc = 0;
        for (int i = 0; i < 12; i++) {
            if (block & (0x400000u >> i)) {
                ecc ^= gen[i];
            }
        }
        return ecc ^ (block & 0x7ffu);
    };
    for (int a = 0; a < 23; a++) {
        const unsigned ea = 1u << a;
        fix[syn(ea)] = (uint16_t)(ea >> 11);
        for (int b = a + 1; b < 23; b++) {
            const unsigned eb = ea | (1u << b);
            fix[syn(eb)] = (uint16_t)(eb >> 11);
            for (int c = b + 1; c < 23; c++) {
                const unsigned ec = eb | (1u << c);
                fix[syn(ec)] = (uint16_t)(ec >> 11);
            }
        }
    }
    memset(hfix, 0, sizeof(hfix));
    for (int b = 0; b < 15; b++) {
        unsigned s = 0;
        for (int i = 0; i < 4; i++) {
            s = (s << 1) | (unsigned)parity_host((1u << b) & kHammingCheckHost[i]);
        }
        hfix[s] = (uint16_t)(1u << b);
    }
    cudaError_t e = cudaMemcpyToSymbol(g_golay_gen, gen, sizeof(gen));
    if (e == cudaSuccess) {
        e = cudaMemcpyToSymbol(g_golay_fix, fix, 2048 * sizeof(uint16_t));
    }
    if (e == cudaSuccess) {
        e = cudaMemcpyToSymbol(g_hamming_fix, hfix, sizeof(hfix));
    }
    free(fix);
    if (e != cudaSuccess) {
        return cuda_fail(e, "mbe_ecc tables", __FILE__, __LINE__);
    }
    if (dev >= 0 && dev < 64) {
        ready[dev] = true;
    }
    return 0;
}

struct Tmp {
    void* p = nullptr;
    cudaError_t err;
    explicit Tmp(size_t bytes) { err = cudaMalloc(&p, bytes ? bytes : 1); }
    ~Tmp() { cudaFree(p); }
};

template <int IN, int OUT, typename Launch>
int
decode_host(const uint8_t* h_fr, uint8_t* h_d, int32_t* h_c0, int32_t* h_tot, int n, Launch launch) {
    const size_t nn = (size_t)n;
    Tmp in(nn * IN), out(nn * OUT), c0(nn * 4), tot(nn * 4);
    DSDNEO_CUDA(in.err);
    DSDNEO_CUDA(out.err);
    DSDNEO_CUDA(c0.err);
    DSDNEO_CUDA(tot.err);
    DSDNEO_CUDA(cudaMemcpy(in.p, h_fr, nn * IN, cudaMemcpyHostToDevice));
    int rc = launch((const uint8_t*)in.p, (uint8_t*)out.p, (int32_t*)c0.p, (int32_t*)tot.p);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaMemcpy(h_d, out.p, nn * OUT, cudaMemcpyDeviceToHost));
    if (h_c0) {
        DSDNEO_CUDA(cudaMemcpy(h_c0, c0.p, nn * 4, cudaMemcpyDeviceToHost));
    }
    if (h_tot) {
        DSDNEO_CUDA(cudaMemcpy(h_tot, tot.p, nn * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

}  // namespace

extern "C" {

int
dsdneo_b200_ambe3600x2450_decode_batch(const uint8_t* d_ambe_fr, uint8_t* d_ambe_d, int32_t* d_c0_errors, int32_t* d_total_errors,
                                       int n_frames, void* stream) {
    if (!d_ambe_fr || !d_ambe_d || !d_c0_errors || !d_total_errors || n_frames < 0) {
        set_error("ambe3600x2450_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (!rc) {
        rc = ensure_tables();
    }
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("ambe3600x2450_decode_kernel", s);
        ambe3600x2450_decode_kernel<<<grid_for(n_frames, 128), 128, 0, s>>>(d_ambe_fr, d_ambe_d, d_c0_errors, d_total_errors, n_frames);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_imbe7200x4400_decode_batch(const uint8_t* d_imbe_fr, uint8_t* d_imbe_d, int32_t* d_c0_errors, int32_t* d_total_errors,
                                       int n_frames, void* stream) {
    if (!d_imbe_fr || !d_imbe_d || !d_c0_errors || !d_total_errors || n_frames < 0) {
        set_error("imbe7200x4400_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (!rc) {
        rc = ensure_tables();
    }
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("imbe7200x4400_decode_kernel", s);
        imbe7200x4400_decode_kernel<<<grid_for(n_frames, 128), 128, 0, s>>>(d_imbe_fr, d_imbe_d, d_c0_errors, d_total_errors, n_frames);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_p25p1_voice_imbe_decode_batch(const dsdneo_b200_p25p1_voice* d_voices, int n_records, uint8_t* d_imbe_d,
                                          int32_t* d_c0_errors, int32_t* d_total_errors, void* stream) {
    if (!d_voices || !d_imbe_d || !d_c0_errors || !d_total_errors || n_records < 0) {
        set_error("p25p1_voice_imbe_decode_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_records == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (!rc) {
        rc = ensure_tables();
    }
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    const int n = n_records * 9;
    {
        KernelTimer kt("imbe7200x4400_decode_packed_kernel", s);
        imbe7200x4400_decode_packed_kernel<<<grid_for(n, 128), 128, 0, s>>>(
            reinterpret_cast<const uint32_t*>(d_voices), (int)(sizeof(dsdneo_b200_p25p1_voice) / 4), d_imbe_d, d_c0_errors, d_total_errors, n);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_ambe3600x2450_decode_batch_host(const uint8_t* h_ambe_fr, uint8_t* h_ambe_d, int32_t* h_c0_errors,
                                            int32_t* h_total_errors, int n_frames) {
    if (!h_ambe_fr || !h_ambe_d || n_frames < 0) {
        set_error("ambe3600x2450_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    return decode_host<96, 49>(h_ambe_fr, h_ambe_d, h_c0_errors, h_total_errors, n_frames,
                               [&](const uint8_t* i, uint8_t* o, int32_t* a, int32_t* b) {
                                   return dsdneo_b200_ambe3600x2450_decode_batch(i, o, a, b, n_frames, NULL);
                               });
}

int
dsdneo_b200_imbe7200x4400_decode_batch_host(const uint8_t* h_imbe_fr, uint8_t* h_imbe_d, int32_t* h_c0_errors,
                                            int32_t* h_total_errors, int n_frames) {
    if (!h_imbe_fr || !h_imbe_d || n_frames < 0) {
        set_error("imbe7200x4400_decode_batch_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_frames == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    return decode_host<184, 88>(h_imbe_fr, h_imbe_d, h_c0_errors, h_total_errors, n_frames,
                                [&](const uint8_t* i, uint8_t* o, int32_t* a, int32_t* b) {
                                    return dsdneo_b200_imbe7200x4400_decode_batch(i, o, a, b, n_frames, NULL);
                                });
}

int
dsdneo_b200_dmr_voice_cut_batch(const uint8_t* d_dibits, size_t dibit_pitch, const int32_t* d_counts, const void* d_hits,
                                const int32_t* d_n_hits, int n_channels, int max_hits, int n_bursts, int inverted_dmr,
                                uint8_t* d_cach24, uint8_t* d_ambe_fr, uint8_t* d_sync48, uint8_t* d_valid, void* stream) {
    if (!d_dibits || !d_counts || !d_hits || !d_n_hits || !d_cach24 || !d_ambe_fr || !d_sync48 || !d_valid || n_channels < 0 ||
        max_hits < 1 || n_bursts < 1 || n_bursts > 64) {
        set_error("dmr_voice_cut_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (n_channels == 0) {
        return 0;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("dmr_voice_cut_kernel", s);
        dmr_voice_cut_kernel<<<(unsigned)((size_t)n_channels * max_hits * n_bursts), 160, 0, s>>>(
            d_dibits, dibit_pitch, d_counts, (const int32_t*)d_hits, d_n_hits, n_channels, max_hits, n_bursts, inverted_dmr ? 1 : 0,
            d_cach24, d_ambe_fr, d_sync48, d_valid);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

/* dsd_ambe_2450_dibit_map as this library generates it: out[i] = {high_row, high_col, low_row, low_col} (tests compare it with
 * the reference's table) */
int
dsdneo_b200_ambe_2450_dibit_map(uint8_t* out36x4) {
    if (!out36x4) {
        set_error("ambe_2450_dibit_map: NULL");
        return DSDNEO_B200_EINVAL;
    }
    for (int i = 0; i < 36; i++) {
        const unsigned m = ambe_map(i);
        out36x4[4 * i] = (uint8_t)((m >> 5) & 7);
        out36x4[4 * i + 1] = (uint8_t)(m & 31);
        out36x4[4 * i + 2] = (uint8_t)((m >> 13) & 7);
        out36x4[4 * i + 3] = (uint8_t)((m >> 8) & 31);
    }
    return 0;
}

} /* extern "C" */
