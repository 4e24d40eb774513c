// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * Batched frame-sync correlator (K12): the per-symbol part of getFrameSync() for FSK inputs.
 *
 * Reference being replaced (arancormonk/dsd-neo @ 4d06905), per channel and per symbol while hunting:
 *   getFrameSync main loop                      src/dsp/dsd_frame_sync.c:3098-3148
 *     frame_sync_symbol_to_dibit                :2110-2127   hunt-time slice: symbol > 0 -> '1', else '3'
 *     frame_sync_history_push / materialize     rolling window of the last 8..48 characters
 *     frame_sync_try_protocol_matches           strcmp of the window against include/dsd-neo/core/sync_patterns.h:33-67
 * The reference returns at the first match and lets the protocol handler consume the frame; its side effects
 * (threshold warm start, protocol gating by opts->frame_*, no-sync timeouts) are control flow and stay on the host.
 * The batched twin reports, per channel, every position where a table pattern completes (first pattern in table
 * order wins at a position), in stream order, so the host can cut frames without touching symbols.
 *
 * One warp per channel: 32 positions per step; sign bits are gathered with a ballot, so a 32-symbol window is two
 * registers; every pattern is one masked compare.  4 B in per symbol, 8 B out per hit: HBM-bound and tiny.
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

struct SyncTable {
    int n;
    uint32_t bits[DSDNEO_B200_SYNC_MAX_PATTERNS]; /* bit j = (pattern[j] == '1'), j = 0 oldest */
    uint32_t mask[DSDNEO_B200_SYNC_MAX_PATTERNS];
    int len[DSDNEO_B200_SYNC_MAX_PATTERNS];
    int id[DSDNEO_B200_SYNC_MAX_PATTERNS];
};

__global__ void __launch_bounds__(256)
frame_sync_search_kernel(const float* __restrict__ symbols, size_t pitch, const int* __restrict__ n_symbols, SyncTable tab,
                         uint32_t* hist_bits, int* hist_count, dsdneo_b200_sync_hit* hits, int max_hits, int* n_hits,
                         int n_channels) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ch = blockIdx.x * (blockDim.x >> 5) + warp;
    if (ch >= n_channels) {
        return;
    }
    const float* x = symbols + (size_t)ch * pitch;
    const int n = n_symbols[ch];
    uint32_t prev = hist_bits[ch]; /* bit 31 = newest symbol before this launch */
    int seen = hist_count[ch];     /* symbols since reset, saturating at 32 */
    dsdneo_b200_sync_hit* out = hits + (size_t)ch * max_hits;
    int count = 0;
    for (int base = 0; base < n; base += 32) {
        const int p = base + lane;
        const bool one = (p < n) && (x[p] > 0.0f); /* dsd_frame_sync.c:2119-2126 */
        const uint32_t cur = __ballot_sync(0xffffffffu, one);
        const unsigned long long w = ((unsigned long long)cur << 32) | prev; /* bit 32 + l = symbol base + l */
        int hit_id = -1;
        if (p < n) {
            const int avail = min(32, seen + lane + 1);
            for (int k = 0; k < tab.n; k++) {
                const int L = tab.len[k];
                if (avail >= L) {
                    const uint32_t win = (uint32_t)(w >> (33 + lane - L)) & tab.mask[k]; /* oldest symbol at bit 0 */
                    if (win == tab.bits[k]) {
                        hit_id = tab.id[k];
                        break;
                    }
                }
            }
        }
        const uint32_t hm = __ballot_sync(0xffffffffu, hit_id >= 0);
        if (hit_id >= 0) {
            const int slot = count + __popc(hm & ((1u << lane) - 1u));
            if (slot < max_hits) {
                out[slot].position = p;
                out[slot].sync_type = hit_id;
            }
        }
        count += __popc(hm);
        const int valid = min(32, n - base);
        if (valid == 32) {
            prev = cur;
        } else {
            prev = (uint32_t)(w >> valid); /* keep bit 31 = newest */
        }
        seen = min(32, seen + valid);
    }
    if (lane == 0) {
        hist_bits[ch] = prev;
        hist_count[ch] = seen;
        n_hits[ch] = count; /* may exceed max_hits: the caller sees how many were dropped */
    }
}

}  // namespace

struct dsdneo_b200_frame_sync {
    int n_channels;
    SyncTable tab;
    uint32_t* d_hist;
    int* d_count;
};

extern "C" {

dsdneo_b200_frame_sync*
dsdneo_b200_frame_sync_create(int n_channels, const dsdneo_b200_sync_pattern* patterns, int n_patterns) {
    if (n_channels <= 0 || !patterns || n_patterns < 1 || n_patterns > DSDNEO_B200_SYNC_MAX_PATTERNS) {
        set_error("frame_sync_create: need 1..%d patterns", DSDNEO_B200_SYNC_MAX_PATTERNS);
        return NULL;
    }
    if (ensure_device()) {
        return NULL;
    }
    dsdneo_b200_frame_sync* fs = (dsdneo_b200_frame_sync*)calloc(1, sizeof(*fs));
    if (!fs) {
        set_error("frame_sync_create: out of host memory");
        return NULL;
    }
    fs->n_channels = n_channels;
    fs->tab.n = n_patterns;
    for (int k = 0; k < n_patterns; k++) {
        const char* s = patterns[k].symbols;
        const int L = s ? (int)strnlen(s, 33) : 0;
        if (L < 8 || L > 32) {
            set_error("frame_sync_create: pattern %d has %d symbols (8..32 supported)", k, L);
            free(fs);
            return NULL;
        }
        uint32_t bits = 0;
        for (int j = 0; j < L; j++) {
            if (s[j] != '1' && s[j] != '3') {
                set_error("frame_sync_create: pattern %d: only '1' and '3' occur in hunt-time windows", k);
                free(fs);
                return NULL;
            }
            bits |= (uint32_t)(s[j] == '1') << j;
        }
        fs->tab.bits[k] = bits;
        fs->tab.mask[k] = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
        fs->tab.len[k] = L;
        fs->tab.id[k] = patterns[k].sync_type;
    }
    cudaError_t e = cudaMalloc((void**)&fs->d_hist, (size_t)n_channels * sizeof(uint32_t));
    if (e == cudaSuccess) {
        e = cudaMalloc((void**)&fs->d_count, (size_t)n_channels * sizeof(int));
    }
    if (e == cudaSuccess) {
        e = cudaMemset(fs->d_hist, 0, (size_t)n_channels * sizeof(uint32_t));
    }
    if (e == cudaSuccess) {
        e = cudaMemset(fs->d_count, 0, (size_t)n_channels * sizeof(int));
    }
    if (e != cudaSuccess) {
        cuda_fail(e, "frame_sync_create", __FILE__, __LINE__);
        dsdneo_b200_frame_sync_destroy(fs);
        return NULL;
    }
    return fs;
}

void
dsdneo_b200_frame_sync_destroy(dsdneo_b200_frame_sync* fs) {
    if (!fs) {
        return;
    }
    cudaFree(fs->d_hist);
    cudaFree(fs->d_count);
    free(fs);
}

int
dsdneo_b200_frame_sync_reset(dsdneo_b200_frame_sync* fs, void* stream) {
    if (!fs) {
        set_error("frame_sync_reset: NULL");
        return DSDNEO_B200_EINVAL;
    }
    DSDNEO_CUDA(cudaMemsetAsync(fs->d_hist, 0, (size_t)fs->n_channels * sizeof(uint32_t), as_stream(stream)));
    DSDNEO_CUDA(cudaMemsetAsync(fs->d_count, 0, (size_t)fs->n_channels * sizeof(int), as_stream(stream)));
    return 0;
}

int
dsdneo_b200_frame_sync_search_batch(dsdneo_b200_frame_sync* fs, const float* d_symbols, size_t pitch, const int* d_n_symbols,
                                    dsdneo_b200_sync_hit* d_hits, int max_hits, int* d_n_hits, void* stream) {
    if (!fs || !d_symbols || !d_n_symbols || !d_hits || !d_n_hits || max_hits < 1) {
        set_error("frame_sync_search_batch: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("frame_sync_search_kernel", s);
        frame_sync_search_kernel<<<(fs->n_channels + 7) / 8, 256, 0, s>>>(d_symbols, pitch, d_n_symbols, fs->tab, fs->d_hist,
                                                                        fs->d_count, d_hits, max_hits, d_n_hits, fs->n_channels);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

} /* extern "C" */
