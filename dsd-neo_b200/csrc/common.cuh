// SPDX-License-Identifier: GPL-3.0-or-later
// Shared helpers for libdsdneo_b200 (sm_100a only; no CPU fallback anywhere in this library).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dsdneo_b200.h"

namespace dsdneo {

// Thread-local sticky error text returned by dsdneo_b200_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define DSDNEO_CUDA(call)                                                                                              \
    do {                                                                                                               \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) {                                                                                      \
            return dsdneo::cuda_fail(e__, #call, __FILE__, __LINE__);                                                  \
        }                                                                                                              \
    } while (0)

#define DSDNEO_KERNEL_CHECK()                                                                                          \
    do {                                                                                                               \
        cudaError_t e__ = cudaGetLastError();                                                                          \
        if (e__ != cudaSuccess) {                                                                                      \
            return dsdneo::cuda_fail(e__, "kernel launch", __FILE__, __LINE__);                                        \
        }                                                                                                              \
    } while (0)

static inline cudaStream_t
as_stream(void* s) {
    return reinterpret_cast<cudaStream_t>(s);
}

// Every kernel launched by this library bumps this (bench.py reports it as gpu_launches).
extern std::atomic<unsigned long long> g_launch_count;
static inline void
count_launch(int n = 1) {
    g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed);
}

int ensure_device();  // 0 when a usable sm_100 device is current, else error code

// Optional per-kernel device timing (dsdneo_b200_timing_enable): CUDA events recorded on the launching
// stream around each kernel, accumulated per kernel name.  Off by default (no events, no overhead).
extern bool g_timing_on;
void timing_begin(const char* name, cudaStream_t s);
void timing_end(cudaStream_t s);
struct KernelTimer {
    cudaStream_t s;
    bool on;
    KernelTimer(const char* name, cudaStream_t stream) : s(stream), on(g_timing_on) {
        if (on) {
            timing_begin(name, s);
        }
    }
    ~KernelTimer() {
        if (on) {
            timing_end(s);
        }
    }
};

}  // namespace dsdneo

/* library-internal stage entry points of the demod bank (demod_bank.cu), used by frontend.cu */
struct dsdneo_b200_demod_bank;
int dsdneo_demod_bank_channels(const dsdneo_b200_demod_bank* b);
int dsdneo_demod_fir_stage(dsdneo_b200_demod_bank* b, const float* d_iq, size_t iq_pitch_pairs, int block_pairs,
                           int n_blocks, int slot, cudaStream_t s, int want_y = 0, int input_cu8 = 0);
int dsdneo_demod_rec_stage(dsdneo_b200_demod_bank* b, int block_pairs, int n_blocks, float* d_result,
                           size_t result_pitch, int slot, cudaStream_t s);

/* library-internal stage entry points of the symbolizer (symbolizer.cu), used by p25p1_rx.cu */
struct dsdneo_b200_symbolizer;
struct dsdneo_b200_symbol_out;
int dsdneo_symbolize_fir_stage(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int slot, cudaStream_t s);
int dsdneo_symbolize_sym_stage(dsdneo_b200_symbolizer* y, int n_samples, int mode, int have_sync, const dsdneo_b200_symbol_out* out,
                               int slot, cudaStream_t s);
int dsdneo_symbolize_acquire_stage(dsdneo_b200_symbolizer* y, const float* d_disc, size_t disc_pitch, int n_samples, int mode,
                                   int have_sync, const dsdneo_b200_symbol_out* out, dsdneo_b200_acq_info* d_info, int slot,
                                   int hunt_filtered, cudaStream_t s);
int dsdneo_symbolize_drop_stage(dsdneo_b200_symbolizer* y, int* d_drop, cudaStream_t s);
int dsdneo_symbolize_get_acquired(dsdneo_b200_symbolizer* y, int* h_acquired);

/* library-internal: the CQPSK chain (cqpsk.cu) behind the channel LPF of the demod bank */
struct dsdneo_b200_cqpsk_bank;
int dsdneo_cqpsk_bank_channels(const dsdneo_b200_cqpsk_bank* q);
int dsdneo_cqpsk_stage(dsdneo_b200_cqpsk_bank* q, int n_channels, const float2* d_y, size_t y_pitch, const float* d_pwr,
                       const float* d_squelch_level, float* d_channel_pwr, int* d_squelched, int block_pairs,
                       int n_blocks, float* d_symbols, size_t symbols_pitch, int* d_counts, cudaStream_t s);

/* library-internal: dsdneo_b200_p25p1_frame_cut_batch with hit positions relative to buffer index region_off (fec.cu) */
extern "C" int dsdneo_p25p1_frame_cut_region(const uint8_t* d_dibits, size_t dibit_pitch, const int16_t* d_llr, size_t llr_pitch,
                                             const int32_t* d_counts, const void* d_hits, const int32_t* d_n_hits, int n_channels,
                                             int max_hits, int n_payload, uint8_t* d_nid_code63, uint8_t* d_nid_reliab63,
                                             uint8_t* d_nid_parity, uint8_t* d_nid_parity_reliab, uint8_t* d_nid_valid,
                                             uint8_t* d_payload_dibits, int16_t* d_payload_llr, uint8_t* d_payload_valid,
                                             int region_off, void* stream);
