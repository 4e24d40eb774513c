// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * P25 Phase 1 C4FM receiver bank: the whole hot path of BASELINE.json's metric for N already-channelised streams as ONE
 * C-ABI object, device-resident from IQ to frames:
 *
 *   cu8 | cf32 IQ (48 kS/s per channel)                       reference, per channel (one process each)
 *     -> widen_u8_to_f32_bias127                               src/dsp/simd_widen.cpp:139-147
 *     -> full_demod (channel LPF, squelch, FSK discriminator)  src/dsp/demod_pipeline.cpp:1330-1350
 *     -> p25_filter + getDibitSoft (getSymbol, use_symbol,     src/dsp/dsd_symbol.c:1853-1880, src/core/frames/dsd_dibit.c:1043-1089
 *        digitize, soft metrics)
 *     -> frame sync                                            src/dsp/dsd_frame_sync.c:3098-3148
 *     -> NID read + p25p1_nid_decode                           src/engine/dispatch/dispatch_p25p1.c:121-143,203-223
 *     -> processTSBK / processHDU / processLDU1 / processLDU2  src/protocol/p25/phase1/
 *   -> frame records (NAC, DUID, TSBK octets, link-control / encryption-sync hex words after RS, LSD) and voice records
 *      (the imbe_fr[8][23] + reliabilities the reference hands to processMbeFrameSoft), plus the dibit stream.
 *
 * Streaming: every channel keeps the last kKeep symbols of its sliced stream on the device, and frames are decoded
 * kDelay = 864 symbols (one LDU) behind the slicer, so a frame that straddles two process() calls is decoded exactly once,
 * from a contiguous stream, on the call that completes it.
 *
 * This file only composes the batched stages (demod_bank.cu, symbolizer.cu, framesync.cu, fec.cu) on the device; every stage
 * fails loudly without a CUDA device, and there is no CPU path.
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

namespace {

constexpr int kTraceTiles = 32;
constexpr int kTraceEvents = 10; /* A0 A1 B0 B1 C0 C1 D0 D1 + host path: H2D start, H2D end */
constexpr int kHostSlots = 6; /* host-buffer tiles in flight: H2D + the four pipeline stages + D2H */
constexpr int kKeep = 1024;  /* symbols of history kept per channel (>= kDelay + the 90 dibits a DMR burst looks back) */
constexpr int kDelay = 864;  /* frames are decoded this many symbols behind the slicer: the longest frame (LDU) */
static const char kP25Sync[] = "111113113311333313133333"; /* P25P1_SYNC, include/dsd-neo/core/sync_patterns.h:34 */

/* the last kKeep entries of the previous call's stream become the head of this call's stream (ping-pong buffers) */
__global__ void
stream_tail_kernel(const uint8_t* dib_prev, const uint8_t* rel_prev, const short2* llr_prev, const float* sym_prev, uint8_t* dib,
                   uint8_t* rel, short2* llr, float* sym, const int* count_prev, size_t pitch) {
    const int ch = blockIdx.x;
    const int src0 = count_prev[ch]; /* previous valid length was kKeep + count_prev */
    const size_t row = (size_t)ch * pitch;
    for (int i = threadIdx.x; i < kKeep; i += blockDim.x) {
        dib[row + i] = dib_prev[row + src0 + i];
        rel[row + i] = rel_prev[row + src0 + i];
        llr[row + i] = llr_prev[row + src0 + i];
        sym[row + i] = sym_prev[row + src0 + i];
    }
}

__global__ void
stream_account_kernel(const int* count_new, int* valid, long long* stream_base, long long* total, int n_ch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) {
        return;
    }
    /* buffer index 0 of this call = stream index (symbols before this call) - kKeep */
    stream_base[c] = total[c] - kKeep;
    total[c] += count_new[c];
    valid[c] = kKeep + count_new[c];
}

/* per-slot observed NAC for the known-NAC retry of p25p1_nid_decode = the channel's last decoded NAC */
__global__ void
expand_nac_kernel(const int* chan_nac, int* slot_nac, int n_ch, int max_hits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ch * max_hits) {
        slot_nac[i] = chan_nac[i / max_hits];
    }
}

/* p25p1_apply_nac_update (dispatch_p25p1.c:145-157): the last valid decoded NAC of the call becomes the channel's NAC */
__global__ void
update_nac_kernel(const dsdneo_b200_p25p1_frame* frames, const int* frame_off, const int* n_hits, int max_hits, int* chan_nac,
                  int n_ch, int frame_capacity) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) {
        return;
    }
    const int n = min(n_hits[c], max_hits);
    int nac = chan_nac[c];
    for (int h = 0; h < n; h++) {
        const int r = frame_off[c] + h;
        if (r >= frame_capacity) {
            break;
        }
        const int v = frames[r].nac;
        if (frames[r].nid_status > 0 && v != 0 && v != 0xFFF) {
            nac = v;
        }
    }
    chan_nac[c] = nac;
}

/* Loss-of-sync watch (cfg.auto_reacquire_tiles = k): a channel none of whose sync hits of the last k tiles decoded to a valid
 * NID is flagged; the slicer stage two tiles later sends it back to the sync hunt (the reference returns to getFrameSync()
 * whenever a frame's NID fails; here the decision is taken per tile, on the device, in stream order -> deterministic). */
__global__ void
sync_watch_kernel(const int* n_hits, const int8_t* nid_status, const uint8_t* nid_valid, int max_hits, int* idle, int* drop, int k,
                  int n_ch) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) {
        return;
    }
    const int n = min(n_hits[c], max_hits);
    bool ok = false;
    for (int h = 0; h < n; h++) {
        ok = ok || (nid_valid[c * max_hits + h] && nid_status[c * max_hits + h] > 0);
    }
    int v = ok ? 0 : idle[c] + 1;
    if (v >= k) {
        drop[c] = 1;
        v = -2; /* the drop takes effect two tiles from here; from then on k tiles to lock again before the next one */
    }
    idle[c] = v;
}

}  // namespace

struct dsdneo_b200_p25p1_rx {
    dsdneo_b200_p25p1_rx_config cfg;
    int n_ch, max_hits, cap_pairs, cap_new; /* cap_new = symbols one call can add per channel */
    size_t pitch;                            /* stream row pitch = kKeep + cap_new */
    dsdneo_b200_demod_bank* bank;
    dsdneo_b200_symbolizer* sym;
    dsdneo_b200_frame_sync* fs;
    float* d_disc;
    int phase;         /* which of the two stream buffer sets receives this call */
    int acq_left;      /* tiles still to run through the acquisition form (cfg.acquire_tiles at start, or after _reacquire) */
    int acq_ready;     /* hunt state allocated, sync pattern set */
    int *d_idle, *d_drop[2]; /* auto re-acquisition: tiles without a valid NID per channel, drop flags per pipeline slot */
    uint8_t *d_dib[2], *d_rel[2];
    int16_t* d_llr[2];
    float* d_symv[2];
    int *d_count[2], *d_valid[2];
    long long *d_stream_base[2], *d_total;
    int *d_hits, *d_n_hits;
    uint8_t *d_code63, *d_rel63, *d_par, *d_prel, *d_nid_valid, *d_pay_dummy, *d_pay_valid;
    int16_t* d_payllr_dummy;
    int8_t* d_nid_status;
    int *d_nid_nac, *d_nid_errs, *d_slot_nac, *d_chan_nac;
    uint8_t* d_nid_duid;
    int *d_frame_off, *d_voice_off;
    /* the three-stage tile pipeline */
    int pipe_ready;
    cudaStream_t s_a, s_b, s_c, s_d;
    cudaEvent_t ev_fork, ev_a[2], ev_b[2], ev_c[2], ev_d[2];
    cudaEvent_t c_gate; /* optional extra dependency of the next tile's stage C (host path) */
    int c_gate_set;
    unsigned long long tiles;
    cudaEvent_t* trace; /* DSDNEO_B200_RX_TRACE=1: [kTraceTiles][8] stage start / end events of the first tiles after trace_reset */
    unsigned long long trace_base;
    /* host path */
    int host_ready;
    cudaStream_t s_h2d, s_d2h, s_rec; /* s_rec: the exact-size record copies of wait_host */
    cudaEvent_t ev_small[kHostSlots], ev_out_free[kHostSlots];
    void* d_in[kHostSlots];
    size_t in_cap;
    dsdneo_b200_p25p1_frame* d_frames[kHostSlots];
    dsdneo_b200_p25p1_voice* d_voices[kHostSlots];
    int* d_totals[kHostSlots];
    int* h_totals; /* pinned, [kHostSlots][2] */
    unsigned long long tickets, waited;
    dsdneo_b200_p25p1_rx_host_out pending[kHostSlots];
};

extern "C" {

void
dsdneo_b200_p25p1_rx_destroy(dsdneo_b200_p25p1_rx* rx) {
    if (!rx) {
        return;
    }
    cudaDeviceSynchronize(); /* tiles may still be in flight on the pipeline streams */
    if (rx->pipe_ready) {
        cudaStreamDestroy(rx->s_a);
        cudaStreamDestroy(rx->s_b);
        cudaStreamDestroy(rx->s_c);
        cudaStreamDestroy(rx->s_d);
        cudaEventDestroy(rx->ev_fork);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(rx->ev_a[i]);
            cudaEventDestroy(rx->ev_b[i]);
            cudaEventDestroy(rx->ev_c[i]);
            cudaEventDestroy(rx->ev_d[i]);
        }
    }
    dsdneo_b200_demod_bank_destroy(rx->bank);
    dsdneo_b200_symbolizer_destroy(rx->sym);
    dsdneo_b200_frame_sync_destroy(rx->fs);
    cudaFree(rx->d_disc);
    for (int i = 0; i < 2; i++) {
        cudaFree(rx->d_dib[i]);
        cudaFree(rx->d_rel[i]);
        cudaFree(rx->d_llr[i]);
        cudaFree(rx->d_symv[i]);
        cudaFree(rx->d_count[i]);
    }
    for (int i = 0; i < kHostSlots; i++) {
        cudaFree(rx->d_in[i]);
        cudaFree(rx->d_frames[i]);
        cudaFree(rx->d_voices[i]);
        cudaFree(rx->d_totals[i]);
    }
    for (int i = 0; i < 2; i++) {
        cudaFree(rx->d_valid[i]);
        cudaFree(rx->d_stream_base[i]);
    }
    cudaFree(rx->d_total);
    cudaFree(rx->d_hits);
    cudaFree(rx->d_n_hits);
    cudaFree(rx->d_idle);
    cudaFree(rx->d_drop[0]);
    cudaFree(rx->d_drop[1]);
    cudaFree(rx->d_code63);
    cudaFree(rx->d_rel63);
    cudaFree(rx->d_par);
    cudaFree(rx->d_prel);
    cudaFree(rx->d_nid_valid);
    cudaFree(rx->d_pay_dummy);
    cudaFree(rx->d_pay_valid);
    cudaFree(rx->d_payllr_dummy);
    cudaFree(rx->d_nid_status);
    cudaFree(rx->d_nid_nac);
    cudaFree(rx->d_nid_errs);
    cudaFree(rx->d_slot_nac);
    cudaFree(rx->d_chan_nac);
    cudaFree(rx->d_nid_duid);
    cudaFree(rx->d_frame_off);
    cudaFree(rx->d_voice_off);
    if (rx->host_ready) {
        cudaStreamDestroy(rx->s_h2d);
        cudaStreamDestroy(rx->s_d2h);
        cudaStreamDestroy(rx->s_rec);
        for (int i = 0; i < kHostSlots; i++) {
            cudaEventDestroy(rx->ev_small[i]);
            cudaEventDestroy(rx->ev_out_free[i]);
        }
        cudaFreeHost(rx->h_totals);
    }
    free(rx);
}

int
dsdneo_b200_p25p1_rx_frame_capacity(const dsdneo_b200_p25p1_rx* rx) {
    return rx ? rx->n_ch * rx->max_hits : 0;
}

int
dsdneo_b200_p25p1_rx_voice_capacity(const dsdneo_b200_p25p1_rx* rx) {
    /* at most one LDU per 864 symbols of new stream, plus one that was pending */
    return rx ? rx->n_ch * (rx->cap_new / kDelay + 2) : 0;
}

size_t
dsdneo_b200_p25p1_rx_dibit_pitch(const dsdneo_b200_p25p1_rx* rx) {
    return rx ? (size_t)rx->cap_new : 0;
}

/* hunt state + the P25 Phase 1 sync pattern for getFrameSync's acquisition on the device (once per bank) */
static int
rx_enable_acquire(dsdneo_b200_p25p1_rx* rx) {
    if (rx->acq_ready) {
        return 0;
    }
    dsdneo_b200_acq_pattern ap;
    ap.symbols = kP25Sync;
    ap.sync_type = 0; /* DSD_SYNC_P25P1_POS */
    ap.kind = 0;
    if (dsdneo_b200_sym_class_from_synctype(0, 0, 1, &ap.cls) != 0) {
        return DSDNEO_B200_EINVAL;
    }
    const int rc = dsdneo_b200_symbolizer_set_acquire_patterns(rx->sym, &ap, 1);
    if (!rc) {
        rx->acq_ready = 1;
    }
    return rc;
}

dsdneo_b200_p25p1_rx*
dsdneo_b200_p25p1_rx_create(const dsdneo_b200_p25p1_rx_config* cfg) {
    if (!cfg || cfg->n_channels <= 0 || cfg->rate_hz <= 0 || cfg->block_pairs <= 0 || cfg->max_pairs_per_call < cfg->block_pairs
        || cfg->max_pairs_per_call % cfg->block_pairs != 0 || !cfg->p25_filter_taps || cfg->p25_filter_len <= 0) {
        set_error("p25p1_rx_create: bad config (max_pairs_per_call must be a positive multiple of block_pairs; the p25_filter taps "
                  "for this sample rate are required)");
        return NULL;
    }
    if (ensure_device()) {
        return NULL;
    }
    dsdneo_b200_p25p1_rx* rx = (dsdneo_b200_p25p1_rx*)calloc(1, sizeof(*rx));
    if (!rx) {
        set_error("p25p1_rx_create: out of host memory");
        return NULL;
    }
    rx->cfg = *cfg;
    rx->n_ch = cfg->n_channels;
    rx->max_hits = cfg->max_hits > 0 ? (cfg->max_hits > 32 ? 32 : cfg->max_hits) : 32;
    rx->cap_pairs = cfg->max_pairs_per_call;
    const int symrate = 4800;
    int whole = cfg->rate_hz / symrate;
    whole = whole < 2 ? 2 : (whole > 64 ? 64 : whole);
    rx->cap_new = (int)(((size_t)rx->cap_pairs + 256) / (size_t)(whole - 1) + 2);
    rx->cap_new = (rx->cap_new + 31) & ~31;
    rx->pitch = (size_t)kKeep + (size_t)rx->cap_new;
    const size_t n = (size_t)rx->n_ch, slots = n * (size_t)rx->max_hits;

    dsdneo_b200_demod_bank_config bc;
    memset(&bc, 0, sizeof(bc));
    bc.n_channels = rx->n_ch;
    bc.rate_out_hz = cfg->rate_hz;
    bc.channel_lpf_enable = 1;
    bc.channel_lpf_profile = NULL; /* P25_C4FM */
    bc.channel_squelch_level = cfg->channel_squelch_level;
    bc.fir_arith = cfg->fir_arith;
    rx->bank = dsdneo_b200_demod_bank_create(&bc);
    dsdneo_b200_symbolizer_config sc;
    memset(&sc, 0, sizeof(sc));
    sc.n_channels = rx->n_ch;
    sc.output_rate_hz = cfg->rate_hz;
    sc.symbol_rate_hz = symrate;
    sc.use_cosine_filter = 1;
    sc.n_filters = 1;
    sc.filter_taps[0] = cfg->p25_filter_taps;
    sc.filter_len[0] = cfg->p25_filter_len;
    rx->sym = rx->bank ? dsdneo_b200_symbolizer_create(&sc) : NULL;
    dsdneo_b200_sync_pattern pat = {kP25Sync, 0 /* DSD_SYNC_P25P1_POS */};
    rx->fs = rx->sym ? dsdneo_b200_frame_sync_create(rx->n_ch, &pat, 1) : NULL;
    if (!rx->bank || !rx->sym || !rx->fs) {
        dsdneo_b200_p25p1_rx_destroy(rx);
        return NULL;
    }
    /* every channel: P25 Phase 1 positive sync class (p25_filter, window 2/2, min / max tracking) */
    {
        dsdneo_b200_sym_class one;
        dsdneo_b200_sym_class* cls = (dsdneo_b200_sym_class*)malloc(n * sizeof(*cls));
        if (!cls || dsdneo_b200_sym_class_from_synctype(0, 0, 1, &one) != 0) {
            free(cls);
            dsdneo_b200_p25p1_rx_destroy(rx);
            return NULL;
        }
        for (size_t i = 0; i < n; i++) {
            cls[i] = one;
        }
        int rc = dsdneo_b200_symbolizer_set_class(rx->sym, cls);
        free(cls);
        if (!rc && cfg->auto_reacquire_tiles > 0) {
            /* every tile runs the hunt kernel in front of the slicer (it returns at once for synchronised channels) */
            rc = rx_enable_acquire(rx);
            if (!rc && cfg->acquire_tiles <= 0) {
                int* ones = (int*)malloc(n * sizeof(int));
                if (!ones) {
                    rc = DSDNEO_B200_ENOMEM;
                } else {
                    for (size_t i = 0; i < n; i++) {
                        ones[i] = 1;
                    }
                    rc = dsdneo_b200_symbolizer_set_acquired(rx->sym, ones); /* the stream starts symbol-aligned */
                    free(ones);
                }
            }
        }
        if (!rc && cfg->acquire_tiles > 0) {
            /* start never-synchronised: getFrameSync's hunt for the P25 Phase 1 sync (hunting rules, timing nudges, basic lock,
             * sync warm start, matched-filter start-up) runs on the device for the first acquire_tiles tiles */
            rc = rx_enable_acquire(rx);
            if (!rc) {
                rc = dsdneo_b200_symbolizer_set_acquired(rx->sym, NULL);
            }
            rx->acq_left = cfg->acquire_tiles;
        }
        if (rc) {
            dsdneo_b200_p25p1_rx_destroy(rx);
            return NULL;
        }
    }
    cudaError_t e = cudaSuccess;
#define RX_ALLOC(ptr, bytes)                                                                                           \
    if (e == cudaSuccess) {                                                                                            \
        e = cudaMalloc((void**)&(ptr), (bytes));                                                                       \
        if (e == cudaSuccess) {                                                                                        \
            e = cudaMemset((ptr), 0, (bytes));                                                                         \
        }                                                                                                              \
    }
    RX_ALLOC(rx->d_disc, n * (size_t)rx->cap_pairs * sizeof(float));
    for (int i = 0; i < 2; i++) {
        RX_ALLOC(rx->d_dib[i], n * rx->pitch);
        RX_ALLOC(rx->d_rel[i], n * rx->pitch);
        RX_ALLOC(rx->d_llr[i], n * rx->pitch * 2 * sizeof(int16_t));
        RX_ALLOC(rx->d_symv[i], n * rx->pitch * sizeof(float));
        RX_ALLOC(rx->d_count[i], n * sizeof(int));
    }
    for (int i = 0; i < 2; i++) {
        RX_ALLOC(rx->d_valid[i], n * sizeof(int));
        RX_ALLOC(rx->d_stream_base[i], n * sizeof(long long));
    }
    RX_ALLOC(rx->d_total, n * sizeof(long long));
    RX_ALLOC(rx->d_hits, slots * 2 * sizeof(int));
    RX_ALLOC(rx->d_n_hits, n * sizeof(int));
    if (cfg->auto_reacquire_tiles > 0) {
        RX_ALLOC(rx->d_idle, n * sizeof(int));
        RX_ALLOC(rx->d_drop[0], n * sizeof(int));
        RX_ALLOC(rx->d_drop[1], n * sizeof(int));
    }
    RX_ALLOC(rx->d_code63, slots * 63);
    RX_ALLOC(rx->d_rel63, slots * 63);
    RX_ALLOC(rx->d_par, slots);
    RX_ALLOC(rx->d_prel, slots);
    RX_ALLOC(rx->d_nid_valid, slots);
    RX_ALLOC(rx->d_pay_dummy, slots);
    RX_ALLOC(rx->d_pay_valid, slots);
    RX_ALLOC(rx->d_payllr_dummy, slots * 2 * sizeof(int16_t));
    RX_ALLOC(rx->d_nid_status, slots);
    RX_ALLOC(rx->d_nid_nac, slots * sizeof(int));
    RX_ALLOC(rx->d_nid_errs, slots * sizeof(int));
    RX_ALLOC(rx->d_slot_nac, slots * sizeof(int));
    RX_ALLOC(rx->d_chan_nac, n * sizeof(int));
    RX_ALLOC(rx->d_nid_duid, slots);
    RX_ALLOC(rx->d_frame_off, n * sizeof(int));
    RX_ALLOC(rx->d_voice_off, n * sizeof(int));
#undef RX_ALLOC
    if (e != cudaSuccess) {
        cuda_fail(e, "p25p1_rx_create", __FILE__, __LINE__);
        dsdneo_b200_p25p1_rx_destroy(rx);
        return NULL;
    }
    return rx;
}

/*
 * One tile through the chain as four pipeline stages on the bank's own streams, so that consecutive tiles overlap on the
 * device: the latency-bound per-channel recurrences (discriminator, slicer) of one tile run under the throughput-bound
 * filters of the next.
 *   stage A (s_a)  channel LPF + phase (lpf_phase_kernel; cu8 widened in its staging), FIR state -> bank phase buffer [slot]
 *   stage B (s_b)  discriminator recurrences -> d_disc -> matched filter (sps_fir_kernel)    -> symbolizer filter buffer [slot]
 *   stage C (s_c)  stream tail, slicer (symbolize_kernel), stream accounting                   -> stream buffers [tile & 1]
 *   stage D (s_d)  sync hunt, frame cut, NID, frame decode, NAC tracking, output copies
 * Every stage owns its carried state, so the order inside a stream is the tile order and the only cross-stream edges are
 * A(i) -> B(i) -> C(i) -> D(i) and the buffer-reuse edges B(i) -> A(i + 2), C(i) -> B(i + 2), D(i) -> C(i + 2).
 */
static int
rx_pipeline_init(dsdneo_b200_p25p1_rx* rx) {
    if (rx->pipe_ready) {
        return 0;
    }
    int prio_least = 0, prio_greatest = 0;
    DSDNEO_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    /* the slicer is the longest serial chain of a tile: its CTAs (and the short frame kernels behind it) get SM slots first,
     * the wide filter grids fill what is left */
    const int p_hi = prio_greatest, p_mid = prio_greatest < prio_least ? prio_greatest + 1 : prio_least, p_lo = prio_least;
    DSDNEO_CUDA(cudaStreamCreateWithPriority(&rx->s_a, cudaStreamNonBlocking, p_lo));
    DSDNEO_CUDA(cudaStreamCreateWithPriority(&rx->s_b, cudaStreamNonBlocking, p_mid));
    DSDNEO_CUDA(cudaStreamCreateWithPriority(&rx->s_c, cudaStreamNonBlocking, p_hi));
    DSDNEO_CUDA(cudaStreamCreateWithPriority(&rx->s_d, cudaStreamNonBlocking, p_hi));
    DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_a[i], cudaEventDisableTiming));
        DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_b[i], cudaEventDisableTiming));
        DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_c[i], cudaEventDisableTiming));
        DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_d[i], cudaEventDisableTiming));
    }
    const char* tr = getenv("DSDNEO_B200_RX_TRACE");
    if (tr && tr[0] == '1') {
        rx->trace = (cudaEvent_t*)calloc((size_t)kTraceTiles * kTraceEvents, sizeof(cudaEvent_t));
        for (int i = 0; rx->trace && i < kTraceTiles * kTraceEvents; i++) {
            DSDNEO_CUDA(cudaEventCreate(&rx->trace[i]));
        }
        rx->trace_base = ~0ull;
    }
    rx->pipe_ready = 1;
    return 0;
}

static void
rx_trace(dsdneo_b200_p25p1_rx* rx, unsigned long long tile, int what, cudaStream_t s) {
    if (rx->trace && rx->trace_base != ~0ull && tile >= rx->trace_base && tile < rx->trace_base + kTraceTiles) {
        cudaEventRecord(rx->trace[(tile - rx->trace_base) * kTraceEvents + what], s);
    }
}

long long
dsdneo_b200_p25p1_rx_submit(dsdneo_b200_p25p1_rx* rx, const void* d_iq, size_t iq_pitch_pairs, int n_pairs,
                            const dsdneo_b200_p25p1_rx_out* out, void* stream) {
    if (!rx || !d_iq || !out || !out->d_frames || !out->d_totals || n_pairs <= 0 || n_pairs > rx->cap_pairs
        || n_pairs % rx->cfg.block_pairs != 0 || iq_pitch_pairs < (size_t)n_pairs || out->frame_capacity <= 0
        || (out->voice_capacity > 0 && !out->d_voices)) {
        set_error("p25p1_rx_process: bad argument (n_pairs must be a multiple of block_pairs, at most max_pairs_per_call)");
        return DSDNEO_B200_EINVAL;
    }
    int rc = rx_pipeline_init(rx);
    if (rc) {
        return rc;
    }
    const unsigned long long tile = rx->tiles;
    const int slot = (int)(tile & 1);
    const int n_ch = rx->n_ch, cur = rx->phase, prev = rx->phase ^ 1;
    /* An acquiring tile runs its four stages one after the other on ONE stream, behind everything the previous tile queued:
     * hunting channels read the raw discriminator samples, so the matched filter cannot run ahead of the slicer.  Only the
     * first cfg.acquire_tiles tiles of a stream pay for that. */
    const bool acq = rx->acq_left > 0;
    const cudaStream_t st_a = acq ? rx->s_d : rx->s_a, st_b = acq ? rx->s_d : rx->s_b, st_c = acq ? rx->s_d : rx->s_c, st_d = rx->s_d;
    if (acq && tile >= 1) {
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_d, rx->ev_a[slot ^ 1], 0));
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_d, rx->ev_b[slot ^ 1], 0));
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_d, rx->ev_c[slot ^ 1], 0));
    }
    /* ---- stage A ---- */
    cudaStream_t s = st_a;
    DSDNEO_CUDA(cudaEventRecord(rx->ev_fork, as_stream(stream))); /* the caller's input is ready on its stream */
    DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_fork, 0));
    if (tile >= 2) {
        DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_b[slot], 0)); /* the recurrences two tiles back have read phase buffer [slot] */
    }
    rx_trace(rx, tile, 0, s);
    /* cu8 input is widened inside the channel filter's staging (widen_u8_to_f32_bias127 fused into lpf_phase_kernel) */
    const int n_blocks = n_pairs / rx->cfg.block_pairs;
    rc = dsdneo_demod_fir_stage(rx->bank, (const float*)d_iq, iq_pitch_pairs, rx->cfg.block_pairs, n_blocks, slot, s, 0,
                                rx->cfg.input_cu8 ? 1 : 0);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaEventRecord(rx->ev_a[slot], s)); /* also: the caller's input buffer is consumed */
    rx_trace(rx, tile, 1, s);
    /* ---- stage B ---- */
    s = st_b;
    DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_a[slot], 0));
    if (tile >= 2) {
        DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_c[slot], 0)); /* the slicer two tiles back has read filter buffer [slot] */
    }
    rx_trace(rx, tile, 2, s);
    rc = dsdneo_demod_rec_stage(rx->bank, rx->cfg.block_pairs, n_blocks, rx->d_disc, (size_t)rx->cap_pairs, slot, s);
    if (rc) {
        return rc;
    }
    if (!acq) {
        rc = dsdneo_symbolize_fir_stage(rx->sym, rx->d_disc, (size_t)rx->cap_pairs, n_pairs, slot, s);
        if (rc) {
            return rc;
        }
    }
    DSDNEO_CUDA(cudaEventRecord(rx->ev_b[slot], s));
    rx_trace(rx, tile, 3, s);
    /* ---- stage C ---- */
    s = st_c;
    DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_b[slot], 0));
    if (tile >= 2) {
        DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_d[slot], 0)); /* the frame stage two tiles back has read stream buffers [cur] */
    }
    if (rx->c_gate_set) { /* host path: the stream buffers this tile writes were last read by a device-to-host copy */
        DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->c_gate, 0));
        rx->c_gate_set = 0;
    }
    rx_trace(rx, tile, 4, s);
    {
        KernelTimer kt("stream_tail_kernel", s);
        stream_tail_kernel<<<n_ch, 256, 0, s>>>(rx->d_dib[prev], rx->d_rel[prev], (const short2*)rx->d_llr[prev], rx->d_symv[prev],
                                               rx->d_dib[cur], rx->d_rel[cur], (short2*)rx->d_llr[cur], rx->d_symv[cur],
                                               rx->d_count[prev], rx->pitch);
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    dsdneo_b200_symbol_out so;
    so.d_symbols = rx->d_symv[cur] + kKeep;
    so.d_dibits = rx->d_dib[cur] + kKeep;
    so.d_reliability = rx->d_rel[cur] + kKeep;
    so.d_llr = rx->d_llr[cur] + 2 * kKeep;
    so.d_count = rx->d_count[cur];
    so.pitch = rx->pitch;
    const bool watch = rx->cfg.auto_reacquire_tiles > 0;
    if (watch && tile >= 2) { /* decisions of the frame stage two tiles back (stage C already waits for it) */
        rc = dsdneo_symbolize_drop_stage(rx->sym, rx->d_drop[slot], s);
        if (rc) {
            return rc;
        }
    }
    if (watch && !acq) {
        /* the pipelined form of the acquisition tile: the matched filter has run in stage B, the hunt kernel takes the channels
         * that are hunting, the slicer the others (neither touches state of the stages in front) */
        rc = dsdneo_symbolize_acquire_stage(rx->sym, rx->d_disc, (size_t)rx->cap_pairs, n_pairs, DSDNEO_SYM_MODE_GET_DIBIT_SOFT, 1, &so, NULL,
                                            slot, 2, s);
    } else if (acq) {
        /* getFrameSync's hunt + getDibitSoft on this tile.  The hunt runs on the matched filter's output, as every hunt of the
         * reference after a channel's first sync does (lastsynctype known): a cold hunt on raw samples locks half a symbol off
         * once the 91-tap p25_filter (45 samples = 4.5 symbols of delay) switches on, and the reference only recovers from
         * that by losing the frame and hunting again -- with the filter on. */
        rc = dsdneo_symbolize_acquire_stage(rx->sym, rx->d_disc, (size_t)rx->cap_pairs, n_pairs, DSDNEO_SYM_MODE_GET_DIBIT_SOFT, 1, &so, NULL,
                                            slot, 1, s);
    } else {
        rc = dsdneo_symbolize_sym_stage(rx->sym, n_pairs, DSDNEO_SYM_MODE_GET_DIBIT_SOFT, 1, &so, slot, s);
    }
    if (rc) {
        return rc;
    }
    {
        KernelTimer kt("stream_account_kernel", s);
        stream_account_kernel<<<(n_ch + 127) / 128, 128, 0, s>>>(rx->d_count[cur], rx->d_valid[cur], rx->d_stream_base[cur], rx->d_total,
                                                                 n_ch);
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    DSDNEO_CUDA(cudaEventRecord(rx->ev_c[slot], s));
    rx_trace(rx, tile, 5, s);
    /* ---- stage D ---- */
    s = st_d;
    DSDNEO_CUDA(cudaStreamWaitEvent(s, rx->ev_c[slot], 0));
    rx_trace(rx, tile, 6, s);
    const int region = kKeep - kDelay;
    rc = dsdneo_b200_frame_sync_search_batch(rx->fs, rx->d_symv[cur] + region, rx->pitch, rx->d_count[cur],
                                             (dsdneo_b200_sync_hit*)rx->d_hits, rx->max_hits, rx->d_n_hits, s);
    if (rc) {
        return rc;
    }
    rc = dsdneo_p25p1_frame_cut_region(rx->d_dib[cur], rx->pitch, rx->d_llr[cur], rx->pitch, rx->d_valid[cur], rx->d_hits, rx->d_n_hits, n_ch,
                                       rx->max_hits, 0, rx->d_code63, rx->d_rel63, rx->d_par, rx->d_prel, rx->d_nid_valid,
                                       rx->d_pay_dummy, rx->d_payllr_dummy, rx->d_pay_valid, region, s);
    if (rc) {
        return rc;
    }
    const int slots = n_ch * rx->max_hits;
    {
        KernelTimer kt("expand_nac_kernel", s);
        expand_nac_kernel<<<(slots + 255) / 256, 256, 0, s>>>(rx->d_chan_nac, rx->d_slot_nac, n_ch, rx->max_hits);
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    const int threshold = rx->cfg.erasure_threshold > 0 ? rx->cfg.erasure_threshold : 64;
    rc = dsdneo_b200_p25p1_nid_decode_batch(rx->d_code63, rx->d_rel63, rx->cfg.track_nac ? rx->d_slot_nac : NULL, rx->d_par, rx->d_prel,
                                            threshold, rx->d_nid_status, rx->d_nid_nac, rx->d_nid_duid, rx->d_nid_errs, slots, s);
    if (rc) {
        return rc;
    }
    rc = dsdneo_b200_p25p1_frames_decode_batch(rx->d_dib[cur], rx->pitch, rx->d_llr[cur], rx->pitch, rx->d_valid[cur],
                                               (const dsdneo_b200_sync_hit*)rx->d_hits, rx->d_n_hits, n_ch, rx->max_hits, region,
                                               rx->d_stream_base[cur], rx->d_nid_status, rx->d_nid_valid, rx->d_nid_nac, rx->d_nid_duid,
                                               rx->d_nid_errs, threshold, rx->cfg.hard_override_disabled ? 0 : 1, rx->d_frame_off,
                                               rx->d_voice_off, out->d_totals, out->d_frames, out->frame_capacity, out->d_voices,
                                               out->voice_capacity, s);
    if (rc) {
        return rc;
    }
    if (watch) {
        KernelTimer kt("sync_watch_kernel", s);
        sync_watch_kernel<<<(n_ch + 127) / 128, 128, 0, s>>>(rx->d_n_hits, rx->d_nid_status, rx->d_nid_valid, rx->max_hits, rx->d_idle,
                                                            rx->d_drop[slot], rx->cfg.auto_reacquire_tiles, n_ch);
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    if (rx->cfg.track_nac) {
        KernelTimer kt("update_nac_kernel", s);
        update_nac_kernel<<<(n_ch + 127) / 128, 128, 0, s>>>(out->d_frames, rx->d_frame_off, rx->d_n_hits, rx->max_hits, rx->d_chan_nac,
                                                            n_ch, out->frame_capacity);
        DSDNEO_KERNEL_CHECK();
        count_launch();
    }
    if (out->d_dibits) { /* the call's new dibits, [n_channels][dibit_pitch], + counts */
        DSDNEO_CUDA(cudaMemcpy2DAsync(out->d_dibits, out->dibit_pitch, rx->d_dib[cur] + kKeep, rx->pitch,
                                      (size_t)min((size_t)rx->cap_new, out->dibit_pitch), (size_t)n_ch, cudaMemcpyDeviceToDevice, s));
    }
    if (out->d_counts) {
        DSDNEO_CUDA(cudaMemcpyAsync(out->d_counts, rx->d_count[cur], (size_t)n_ch * sizeof(int), cudaMemcpyDeviceToDevice, s));
    }
    DSDNEO_CUDA(cudaEventRecord(rx->ev_d[slot], s));
    rx_trace(rx, tile, 7, s);
    if (acq) { /* the stage streams pick up their carried state behind this tile */
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_a, rx->ev_d[slot], 0));
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_b, rx->ev_d[slot], 0));
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_c, rx->ev_d[slot], 0));
        rx->acq_left--;
    }
    rx->phase ^= 1;
    return (long long)rx->tiles++;
}

int
dsdneo_b200_p25p1_rx_reacquire(dsdneo_b200_p25p1_rx* rx, const int* h_synchronised, int tiles) {
    if (!rx || tiles < 1) {
        set_error("p25p1_rx_reacquire: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = rx_enable_acquire(rx);
    if (rc) {
        return rc;
    }
    /* drains the pipeline (device synchronise), then the flagged channels hunt from an empty window */
    rc = dsdneo_b200_symbolizer_set_acquired(rx->sym, h_synchronised);
    if (rc) {
        return rc;
    }
    rx->acq_left = tiles;
    return 0;
}

int
dsdneo_b200_p25p1_rx_channel_status(dsdneo_b200_p25p1_rx* rx, int* h_synchronised, int* h_idle_tiles) {
    if (!rx || (!h_synchronised && !h_idle_tiles)) {
        set_error("p25p1_rx_channel_status: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    if (h_synchronised) {
        const int rc = dsdneo_symbolize_get_acquired(rx->sym, h_synchronised);
        if (rc) {
            return rc;
        }
    }
    if (h_idle_tiles) {
        if (rx->d_idle) {
            DSDNEO_CUDA(cudaDeviceSynchronize());
            DSDNEO_CUDA(cudaMemcpy(h_idle_tiles, rx->d_idle, (size_t)rx->n_ch * sizeof(int), cudaMemcpyDeviceToHost));
        } else {
            memset(h_idle_tiles, 0, (size_t)rx->n_ch * sizeof(int));
        }
    }
    return 0;
}

int
dsdneo_b200_p25p1_rx_wait(dsdneo_b200_p25p1_rx* rx, long long ticket, void* stream) {
    if (!rx || !rx->pipe_ready || ticket < 0 || (unsigned long long)ticket >= rx->tiles) {
        set_error("p25p1_rx_wait: unknown ticket");
        return DSDNEO_B200_EINVAL;
    }
    /* stage C runs the tiles in order: the event of a later tile in the same slot covers this one as well */
    DSDNEO_CUDA(cudaStreamWaitEvent(as_stream(stream), rx->ev_d[(int)(ticket & 1)], 0));
    return 0;
}

int
dsdneo_b200_p25p1_rx_input_consumed(dsdneo_b200_p25p1_rx* rx, long long ticket, void* stream) {
    if (!rx || !rx->pipe_ready || ticket < 0 || (unsigned long long)ticket >= rx->tiles) {
        set_error("p25p1_rx_input_consumed: unknown ticket");
        return DSDNEO_B200_EINVAL;
    }
    DSDNEO_CUDA(cudaStreamWaitEvent(as_stream(stream), rx->ev_a[(int)(ticket & 1)], 0));
    return 0;
}

/* Debug aid (DSDNEO_B200_RX_TRACE=1): start tracing at the next tile / read the stage start and end times (ms, relative to the
 * first traced tile's stage A start) of the traced tiles: ms[tile][10] = {A0, A1, B0, B1, C0, C1, D0, D1, H2D0, H2D1} (the last
 * two only on the host path).  Synchronises. */
int
dsdneo_b200_p25p1_rx_trace(dsdneo_b200_p25p1_rx* rx, float* ms, int max_tiles) {
    if (!rx || !rx->trace) {
        return 0;
    }
    if (!ms) {
        rx->trace_base = rx->tiles;
        return 0;
    }
    cudaDeviceSynchronize();
    const unsigned long long done = rx->tiles - rx->trace_base;
    const int n = (int)(done < (unsigned long long)kTraceTiles ? done : kTraceTiles);
    int k = 0;
    for (; k < n && k < max_tiles; k++) {
        for (int w = 0; w < kTraceEvents; w++) {
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, rx->trace[0], rx->trace[k * kTraceEvents + w]) != cudaSuccess) {
                (void)cudaGetLastError(); /* an event that was never recorded (device path: no H2D) */
                t = -1.0f;
            }
            ms[k * kTraceEvents + w] = t;
        }
    }
    return k;
}

int
dsdneo_b200_p25p1_rx_process(dsdneo_b200_p25p1_rx* rx, const void* d_iq, size_t iq_pitch_pairs, int n_pairs,
                             const dsdneo_b200_p25p1_rx_out* out, void* stream) {
    const long long t = dsdneo_b200_p25p1_rx_submit(rx, d_iq, iq_pitch_pairs, n_pairs, out, stream);
    if (t < 0) {
        return (int)t;
    }
    return dsdneo_b200_p25p1_rx_wait(rx, t, stream);
}

/*
 * Host-buffer streaming form (the reference's demod thread consumes its input ring the same way, src/io/radio/rtl_sdr_fm.cpp:
 * 3458-3512): submit(tile i) queues H2D, the whole chain and the D2H of the dibit stream on three internal streams and
 * returns a ticket; wait(ticket) blocks until the tile's results are in the caller's buffers.  The record counts are only
 * known after the tile ran, so wait() copies exactly totals[0] frame and totals[1] voice records.  At most kHostSlots (6) tiles
 * may be in flight -- the depth of the four-stage pipeline plus the copies either side: submit() first completes the tile
 * submitted six calls earlier if the caller has not waited for it yet.
 */
static int
rx_host_init(dsdneo_b200_p25p1_rx* rx) {
    if (rx->host_ready) {
        return 0;
    }
    DSDNEO_CUDA(cudaStreamCreateWithFlags(&rx->s_h2d, cudaStreamNonBlocking));
    DSDNEO_CUDA(cudaStreamCreateWithFlags(&rx->s_d2h, cudaStreamNonBlocking));
    DSDNEO_CUDA(cudaStreamCreateWithFlags(&rx->s_rec, cudaStreamNonBlocking));
    for (int i = 0; i < kHostSlots; i++) {
        DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_small[i], cudaEventDisableTiming));
        DSDNEO_CUDA(cudaEventCreateWithFlags(&rx->ev_out_free[i], cudaEventDisableTiming));
        DSDNEO_CUDA(cudaMalloc((void**)&rx->d_frames[i], (size_t)dsdneo_b200_p25p1_rx_frame_capacity(rx) * sizeof(dsdneo_b200_p25p1_frame)));
        DSDNEO_CUDA(cudaMalloc((void**)&rx->d_voices[i], (size_t)dsdneo_b200_p25p1_rx_voice_capacity(rx) * sizeof(dsdneo_b200_p25p1_voice)));
        DSDNEO_CUDA(cudaMalloc((void**)&rx->d_totals[i], 2 * sizeof(int)));
    }
    DSDNEO_CUDA(cudaMallocHost((void**)&rx->h_totals, 2 * kHostSlots * sizeof(int)));
    rx->host_ready = 1;
    return 0;
}

int
dsdneo_b200_p25p1_rx_wait_host(dsdneo_b200_p25p1_rx* rx, long long ticket) {
    if (!rx || !rx->host_ready || ticket < 0 || (unsigned long long)ticket >= rx->tickets) {
        set_error("p25p1_rx_wait_host: unknown ticket");
        return DSDNEO_B200_EINVAL;
    }
    if ((unsigned long long)ticket < rx->waited) {
        return 0; /* already completed (by an earlier wait or by a later submit) */
    }
    for (unsigned long long t = rx->waited; t <= (unsigned long long)ticket; t++) {
        const int slot = (int)(t % kHostSlots);
        const dsdneo_b200_p25p1_rx_host_out* o = &rx->pending[slot];
        DSDNEO_CUDA(cudaEventSynchronize(rx->ev_small[slot])); /* totals, dibits and counts are in host memory */
        int nf = rx->h_totals[2 * slot], nv = rx->h_totals[2 * slot + 1];
        nf = nf > o->frame_capacity ? o->frame_capacity : nf;
        nv = nv > o->voice_capacity ? o->voice_capacity : nv;
        /* on their own stream: s_d2h may already hold the next tile's copies, which wait for that tile's kernels */
        if (nf > 0) {
            DSDNEO_CUDA(cudaMemcpyAsync(o->h_frames, rx->d_frames[slot], (size_t)nf * sizeof(dsdneo_b200_p25p1_frame), cudaMemcpyDeviceToHost,
                                        rx->s_rec));
        }
        if (nv > 0 && o->h_voices) {
            DSDNEO_CUDA(cudaMemcpyAsync(o->h_voices, rx->d_voices[slot], (size_t)nv * sizeof(dsdneo_b200_p25p1_voice), cudaMemcpyDeviceToHost,
                                        rx->s_rec));
        }
        DSDNEO_CUDA(cudaEventRecord(rx->ev_out_free[slot], rx->s_rec));
        DSDNEO_CUDA(cudaEventSynchronize(rx->ev_out_free[slot]));
        if (o->h_totals) {
            o->h_totals[0] = nf;
            o->h_totals[1] = nv;
        }
        rx->waited = t + 1;
    }
    return 0;
}

long long
dsdneo_b200_p25p1_rx_submit_host(dsdneo_b200_p25p1_rx* rx, const void* h_iq, size_t iq_pitch_pairs, int n_pairs,
                                 const dsdneo_b200_p25p1_rx_host_out* out) {
    if (!rx || !h_iq || !out || !out->h_frames || out->frame_capacity <= 0 || n_pairs <= 0 || n_pairs > rx->cap_pairs
        || iq_pitch_pairs < (size_t)n_pairs) {
        set_error("p25p1_rx_submit_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = rx_host_init(rx);
    if (rc) {
        return rc;
    }
    if (rx->tickets >= kHostSlots && rx->waited + kHostSlots <= rx->tickets) { /* the slot about to be reused still holds an unread tile */
        rc = dsdneo_b200_p25p1_rx_wait_host(rx, (long long)(rx->tickets - kHostSlots));
        if (rc) {
            return rc;
        }
    }
    const int slot = (int)(rx->tickets % kHostSlots);
    const size_t elt = rx->cfg.input_cu8 ? 2 : 8;
    const size_t in_bytes = (size_t)rx->n_ch * iq_pitch_pairs * elt;
    if (rx->in_cap < in_bytes) {
        DSDNEO_CUDA(cudaDeviceSynchronize());
        for (int i = 0; i < kHostSlots; i++) {
            cudaFree(rx->d_in[i]);
            rx->d_in[i] = NULL;
            DSDNEO_CUDA(cudaMalloc(&rx->d_in[i], in_bytes));
        }
        rx->in_cap = in_bytes;
    }
    if (rx->tickets >= kHostSlots) {
        /* d_in[slot] was consumed by stage A of the tile kHostSlots back; stage A runs the tiles in order, so the event of the
         * newest tile that shares its pipeline slot covers it */
        DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_h2d, rx->ev_a[(int)(rx->tiles & 1)], 0));
    }
    if (rx->pipe_ready) {
        rx_trace(rx, rx->tiles, 8, rx->s_h2d);
    }
    DSDNEO_CUDA(cudaMemcpyAsync(rx->d_in[slot], h_iq, in_bytes, cudaMemcpyHostToDevice, rx->s_h2d));
    if (rx->pipe_ready) {
        rx_trace(rx, rx->tiles, 9, rx->s_h2d);
    }
    if (rx->tickets >= 2) {
        /* the stream buffers this tile writes were last read by the dibit D2H of the tile two calls back (the tile in
         * between only read them for its tail and wrote the other set) */
        rx->c_gate = rx->ev_small[(int)((rx->tickets - 2) % kHostSlots)];
        rx->c_gate_set = 1;
    }
    dsdneo_b200_p25p1_rx_out dev;
    memset(&dev, 0, sizeof(dev));
    dev.d_frames = rx->d_frames[slot];
    dev.frame_capacity = dsdneo_b200_p25p1_rx_frame_capacity(rx);
    dev.d_voices = rx->d_voices[slot];
    dev.voice_capacity = dsdneo_b200_p25p1_rx_voice_capacity(rx);
    dev.d_totals = rx->d_totals[slot];
    const long long tile = dsdneo_b200_p25p1_rx_submit(rx, rx->d_in[slot], iq_pitch_pairs, n_pairs, &dev, rx->s_h2d);
    if (tile < 0) {
        return tile;
    }
    const int cur = rx->phase ^ 1; /* the stream buffers this tile was written to */
    DSDNEO_CUDA(cudaStreamWaitEvent(rx->s_d2h, rx->ev_d[(int)(tile & 1)], 0));
    DSDNEO_CUDA(cudaMemcpyAsync(rx->h_totals + 2 * slot, rx->d_totals[slot], 2 * sizeof(int), cudaMemcpyDeviceToHost, rx->s_d2h));
    if (out->h_dibits) {
        DSDNEO_CUDA(cudaMemcpy2DAsync(out->h_dibits, out->dibit_pitch, rx->d_dib[cur] + kKeep, rx->pitch,
                                      (size_t)min((size_t)rx->cap_new, out->dibit_pitch), (size_t)rx->n_ch, cudaMemcpyDeviceToHost, rx->s_d2h));
    }
    if (out->h_counts) {
        DSDNEO_CUDA(cudaMemcpyAsync(out->h_counts, rx->d_count[cur], (size_t)rx->n_ch * sizeof(int), cudaMemcpyDeviceToHost, rx->s_d2h));
    }
    DSDNEO_CUDA(cudaEventRecord(rx->ev_small[slot], rx->s_d2h));
    rx->pending[slot] = *out;
    return (long long)rx->tickets++;
}

int
dsdneo_b200_p25p1_rx_process_host(dsdneo_b200_p25p1_rx* rx, const void* h_iq, size_t iq_pitch_pairs, int n_pairs,
                                  const dsdneo_b200_p25p1_rx_host_out* out) {
    const long long t = dsdneo_b200_p25p1_rx_submit_host(rx, h_iq, iq_pitch_pairs, n_pairs, out);
    if (t < 0) {
        return (int)t;
    }
    return dsdneo_b200_p25p1_rx_wait_host(rx, t);
}

} /* extern "C" */
