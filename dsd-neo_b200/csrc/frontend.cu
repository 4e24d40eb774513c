// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * FSK front end = channelizer (K2, optional fused K1) -> batched full_demod (K4+K5+K6): the whole block
 * side of the hot path behind one call, with device-resident intermediates.
 *
 * Replaces, for N channels at once, what the reference does per process in its demod thread
 * (src/io/radio/rtl_sdr_fm.cpp:3458-3512: read block -> full_demod(d) -> output ring), including the
 * tuner/half-band channel selection that a polyphase bank makes unnecessary.
 *
 * The *_host entry point is the reference-facing call for host buffers: it pipelines
 * H2D copy / kernels / D2H copy per reference block on three streams so PCIe transfers overlap compute.
 */
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace dsdneo;

struct dsdneo_b200_frontend {
    dsdneo_b200_channelizer* cz;
    dsdneo_b200_demod_bank* bank;
    int M, cu8, block_pairs;
    float* d_chan;       /* [M][chan_pitch] cf32 */
    size_t chan_pitch;
    /* host-pipeline resources */
    cudaStream_t s_h2d, s_comp, s_d2h;
    void* d_in[2];
    float* d_out[2];
    size_t in_cap, out_cap;
    cudaEvent_t ev_h2d[2], ev_comp[2], ev_d2h[2], ev_in_free[2];
    cudaEvent_t ev_ticket[4];          /* completion of the last four submit_host calls */
    unsigned long long host_blocks;    /* blocks queued by submit_host so far (slot = host_blocks & 1) */
    unsigned long long host_tickets;   /* submit_host calls so far */
    int streams_ready;
    /* device-side two-stage pipeline (process_async): FIR-side stream, recurrence stream */
    cudaStream_t s_fir, s_rec;
    cudaEvent_t ev_fork, ev_fir_done[2], ev_rec_done[2];
    int async_ready;
    unsigned long long async_seq;
};

static int
ensure_async(dsdneo_b200_frontend* fe) {
    if (fe->async_ready) {
        return 0;
    }
    DSDNEO_CUDA(cudaStreamCreateWithFlags(&fe->s_fir, cudaStreamNonBlocking));
    DSDNEO_CUDA(cudaStreamCreateWithFlags(&fe->s_rec, cudaStreamNonBlocking));
    DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_fir_done[i], cudaEventDisableTiming));
        DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_rec_done[i], cudaEventDisableTiming));
    }
    fe->async_ready = 1;
    fe->async_seq = 0;
    return 0;
}

static int
ensure_chan(dsdneo_b200_frontend* fe, size_t n_out, cudaStream_t s) {
    if (fe->d_chan && fe->chan_pitch >= n_out) {
        return 0;
    }
    DSDNEO_CUDA(cudaStreamSynchronize(s));
    cudaFree(fe->d_chan);
    fe->d_chan = NULL;
    DSDNEO_CUDA(cudaMalloc((void**)&fe->d_chan, (size_t)fe->M * n_out * 2 * sizeof(float)));
    fe->chan_pitch = n_out;
    return 0;
}

extern "C" {

dsdneo_b200_frontend*
dsdneo_b200_frontend_create(const dsdneo_b200_frontend_config* cfg) {
    if (!cfg || cfg->n_channels <= 0 || cfg->wideband_rate_hz <= 0 || cfg->block_pairs <= 0) {
        set_error("frontend_create: bad config");
        return NULL;
    }
    if (cfg->wideband_rate_hz % cfg->n_channels != 0) {
        set_error("frontend_create: wideband rate %d is not a multiple of %d channels", cfg->wideband_rate_hz, cfg->n_channels);
        return NULL;
    }
    dsdneo_b200_frontend* fe = (dsdneo_b200_frontend*)calloc(1, sizeof(*fe));
    if (!fe) {
        set_error("frontend_create: out of host memory");
        return NULL;
    }
    fe->M = cfg->n_channels;
    fe->cu8 = cfg->input_is_cu8 ? 1 : 0;
    fe->block_pairs = cfg->block_pairs;
    fe->cz = dsdneo_b200_channelizer_create(cfg->n_channels, cfg->taps_per_branch, cfg->input_is_cu8, cfg->prototype);
    if (!fe->cz) {
        free(fe);
        return NULL;
    }
    dsdneo_b200_demod_bank_config bc;
    memset(&bc, 0, sizeof(bc));
    bc.n_channels = cfg->n_channels;
    bc.rate_out_hz = cfg->wideband_rate_hz / cfg->n_channels;
    bc.channel_lpf_enable = cfg->channel_lpf_enable;
    bc.channel_lpf_profile = cfg->channel_lpf_profile;
    bc.channel_squelch_level = cfg->channel_squelch_level;
    bc.fir_arith = cfg->fir_arith;
    fe->bank = dsdneo_b200_demod_bank_create(&bc);
    if (!fe->bank) {
        dsdneo_b200_channelizer_destroy(fe->cz);
        free(fe);
        return NULL;
    }
    return fe;
}

void
dsdneo_b200_frontend_destroy(dsdneo_b200_frontend* fe) {
    if (!fe) {
        return;
    }
    if (fe->async_ready) {
        cudaStreamDestroy(fe->s_fir);
        cudaStreamDestroy(fe->s_rec);
        cudaEventDestroy(fe->ev_fork);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(fe->ev_fir_done[i]);
            cudaEventDestroy(fe->ev_rec_done[i]);
        }
    }
    if (fe->streams_ready) {
        cudaStreamDestroy(fe->s_h2d);
        cudaStreamDestroy(fe->s_comp);
        cudaStreamDestroy(fe->s_d2h);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(fe->ev_h2d[i]);
            cudaEventDestroy(fe->ev_comp[i]);
            cudaEventDestroy(fe->ev_d2h[i]);
            cudaEventDestroy(fe->ev_in_free[i]);
        }
        for (int i = 0; i < 4; i++) {
            cudaEventDestroy(fe->ev_ticket[i]);
        }
    }
    for (int i = 0; i < 2; i++) {
        cudaFree(fe->d_in[i]);
        cudaFree(fe->d_out[i]);
    }
    cudaFree(fe->d_chan);
    dsdneo_b200_channelizer_destroy(fe->cz);
    dsdneo_b200_demod_bank_destroy(fe->bank);
    free(fe);
}

int
dsdneo_b200_frontend_reset(dsdneo_b200_frontend* fe, void* stream) {
    if (!fe) {
        set_error("frontend_reset: NULL");
        return DSDNEO_B200_EINVAL;
    }
    int rc = dsdneo_b200_channelizer_reset(fe->cz, stream);
    if (rc) {
        return rc;
    }
    return dsdneo_b200_demod_bank_reset(fe->bank, stream);
}

dsdneo_b200_demod_bank*
dsdneo_b200_frontend_bank(dsdneo_b200_frontend* fe) {
    return fe ? fe->bank : NULL;
}

int
dsdneo_b200_frontend_process(dsdneo_b200_frontend* fe, const void* d_wideband, size_t n_in_samples, float* d_result,
                             size_t result_pitch, void* stream) {
    if (!fe || !d_wideband || !d_result) {
        set_error("frontend_process: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    const size_t per_block = (size_t)fe->M * (size_t)fe->block_pairs;
    if (n_in_samples == 0 || n_in_samples % per_block != 0) {
        set_error("frontend_process: n_in_samples must be a positive multiple of n_channels*block_pairs (%zu)", per_block);
        return DSDNEO_B200_EINVAL;
    }
    const size_t n_out = n_in_samples / (size_t)fe->M;
    const int n_blocks = (int)(n_out / (size_t)fe->block_pairs);
    cudaStream_t s = as_stream(stream);
    int rc = ensure_chan(fe, n_out, s);
    if (rc) {
        return rc;
    }
    rc = dsdneo_b200_channelize(fe->cz, d_wideband, n_in_samples, fe->d_chan, fe->chan_pitch, stream);
    if (rc) {
        return rc;
    }
    return dsdneo_b200_full_demod_batch(fe->bank, fe->d_chan, fe->chan_pitch, fe->block_pairs, n_blocks, d_result,
                                        result_pitch, stream);
}

/*
 * Pipelined form: the time-parallel stages (channelizer, channel LPF + phase discriminator) of call i+1 run on
 * one internal stream while the time-serial recurrence stage of call i runs on another (it only needs 1 CTA per
 * 32 channels, so the rest of the GPU is free).  Inputs must be ready on `stream` at the time of the call; outputs
 * are complete once dsdneo_b200_frontend_join() has been ordered into a stream.
 */
int
dsdneo_b200_frontend_process_async(dsdneo_b200_frontend* fe, const void* d_wideband, size_t n_in_samples,
                                   float* d_result, size_t result_pitch, void* stream) {
    if (!fe || !d_wideband || !d_result) {
        set_error("frontend_process_async: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    const size_t per_block = (size_t)fe->M * (size_t)fe->block_pairs;
    if (n_in_samples == 0 || n_in_samples % per_block != 0) {
        set_error("frontend_process_async: n_in_samples must be a positive multiple of n_channels*block_pairs (%zu)", per_block);
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    rc = ensure_async(fe);
    if (rc) {
        return rc;
    }
    const size_t n_out = n_in_samples / (size_t)fe->M;
    const int n_blocks = (int)(n_out / (size_t)fe->block_pairs);
    if (result_pitch < n_out) { /* the recurrence stage writes result + ch * result_pitch + n, n < n_out */
        set_error("frontend_process_async: result_pitch %zu is smaller than the %zu samples produced per channel", result_pitch, n_out);
        return DSDNEO_B200_EINVAL;
    }
    const int slot = (int)(fe->async_seq & 1);
    rc = ensure_chan(fe, n_out, fe->s_fir);
    if (rc) {
        return rc;
    }
    /* inputs produced on the caller's stream are visible to the FIR stream */
    DSDNEO_CUDA(cudaEventRecord(fe->ev_fork, as_stream(stream)));
    DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_fir, fe->ev_fork, 0));
    if (fe->async_seq >= 2) {
        /* scratch slot is free once the recurrence stage of call i-2 has drained it */
        DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_fir, fe->ev_rec_done[slot], 0));
    }
    rc = dsdneo_b200_channelize(fe->cz, d_wideband, n_in_samples, fe->d_chan, fe->chan_pitch, fe->s_fir);
    if (rc) {
        return rc;
    }
    rc = dsdneo_demod_fir_stage(fe->bank, fe->d_chan, fe->chan_pitch, fe->block_pairs, n_blocks, slot, fe->s_fir);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaEventRecord(fe->ev_fir_done[slot], fe->s_fir));
    DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_rec, fe->ev_fir_done[slot], 0));
    rc = dsdneo_demod_rec_stage(fe->bank, fe->block_pairs, n_blocks, d_result, result_pitch, slot, fe->s_rec);
    if (rc) {
        return rc;
    }
    DSDNEO_CUDA(cudaEventRecord(fe->ev_rec_done[slot], fe->s_rec));
    fe->async_seq++;
    return 0;
}

/** Orders everything queued by process_async before subsequent work on `stream`. */
int
dsdneo_b200_frontend_join(dsdneo_b200_frontend* fe, void* stream) {
    if (!fe) {
        set_error("frontend_join: NULL");
        return DSDNEO_B200_EINVAL;
    }
    if (!fe->async_ready || fe->async_seq == 0) {
        return 0;
    }
    const int last = (int)((fe->async_seq - 1) & 1);
    DSDNEO_CUDA(cudaStreamWaitEvent(as_stream(stream), fe->ev_rec_done[last], 0));
    DSDNEO_CUDA(cudaStreamWaitEvent(as_stream(stream), fe->ev_fir_done[last], 0));
    return 0;
}

/*
 * Streaming form for host buffers: queues the per-block H2D / kernels / D2H pipeline of one tile on the three internal
 * streams and returns without waiting, so the transfers of consecutive tiles overlap (the reference's own demod thread
 * streams blocks through a ring in the same way, src/io/radio/rtl_sdr_fm.cpp:3458-3512).  Returns a ticket >= 0; the
 * caller's buffers must stay valid and untouched until dsdneo_b200_frontend_wait_host(ticket) has returned.  At most
 * four tickets may be outstanding.
 */
long long
dsdneo_b200_frontend_submit_host(dsdneo_b200_frontend* fe, const void* h_wideband, size_t n_in_samples, float* h_result,
                                 size_t result_pitch) {
    if (!fe || !h_wideband || !h_result) {
        set_error("frontend_submit_host: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    const size_t per_block = (size_t)fe->M * (size_t)fe->block_pairs;
    if (n_in_samples == 0 || n_in_samples % per_block != 0) {
        set_error("frontend_submit_host: n_in_samples must be a positive multiple of n_channels*block_pairs (%zu)", per_block);
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    const int n_blocks = (int)(n_in_samples / per_block);
    if (result_pitch < (size_t)n_blocks * fe->block_pairs) {
        set_error("frontend_submit_host: result_pitch too small");
        return DSDNEO_B200_EINVAL;
    }
    if (!fe->streams_ready) {
        DSDNEO_CUDA(cudaStreamCreateWithFlags(&fe->s_h2d, cudaStreamNonBlocking));
        DSDNEO_CUDA(cudaStreamCreateWithFlags(&fe->s_comp, cudaStreamNonBlocking));
        DSDNEO_CUDA(cudaStreamCreateWithFlags(&fe->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_h2d[i], cudaEventDisableTiming));
            DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_comp[i], cudaEventDisableTiming));
            DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_d2h[i], cudaEventDisableTiming));
            DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_in_free[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 4; i++) {
            DSDNEO_CUDA(cudaEventCreateWithFlags(&fe->ev_ticket[i], cudaEventDisableTiming));
        }
        fe->streams_ready = 1;
        fe->host_blocks = 0;
        fe->host_tickets = 0;
    }
    const size_t in_bytes = per_block * (fe->cu8 ? 2 : 8);
    const size_t out_floats = (size_t)fe->M * fe->block_pairs;
    if (fe->in_cap < in_bytes || fe->out_cap < out_floats || !fe->d_chan || fe->chan_pitch < (size_t)fe->block_pairs) {
        DSDNEO_CUDA(cudaDeviceSynchronize()); /* (re)allocation: drain anything still queued on the old buffers */
    }
    if (fe->in_cap < in_bytes) {
        for (int i = 0; i < 2; i++) {
            cudaFree(fe->d_in[i]);
            fe->d_in[i] = NULL;
            DSDNEO_CUDA(cudaMalloc(&fe->d_in[i], in_bytes));
        }
        fe->in_cap = in_bytes;
    }
    if (fe->out_cap < out_floats) {
        for (int i = 0; i < 2; i++) {
            cudaFree(fe->d_out[i]);
            fe->d_out[i] = NULL;
            DSDNEO_CUDA(cudaMalloc((void**)&fe->d_out[i], out_floats * sizeof(float)));
        }
        fe->out_cap = out_floats;
    }
    rc = ensure_chan(fe, (size_t)fe->block_pairs, fe->s_comp);
    if (rc) {
        return rc;
    }
    const unsigned char* src = (const unsigned char*)h_wideband;
    for (int b = 0; b < n_blocks; b++, fe->host_blocks++) {
        const int slot = (int)(fe->host_blocks & 1);
        if (fe->host_blocks >= 2) {
            DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_h2d, fe->ev_in_free[slot], 0)); /* kernels two blocks back consumed d_in[slot] */
        }
        DSDNEO_CUDA(cudaMemcpyAsync(fe->d_in[slot], src + (size_t)b * in_bytes, in_bytes, cudaMemcpyHostToDevice, fe->s_h2d));
        DSDNEO_CUDA(cudaEventRecord(fe->ev_h2d[slot], fe->s_h2d));

        DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_comp, fe->ev_h2d[slot], 0));
        if (fe->host_blocks >= 2) {
            DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_comp, fe->ev_d2h[slot], 0)); /* d_out[slot] drained */
        }
        rc = dsdneo_b200_channelize(fe->cz, fe->d_in[slot], per_block, fe->d_chan, fe->chan_pitch, fe->s_comp);
        if (rc) {
            return rc;
        }
        DSDNEO_CUDA(cudaEventRecord(fe->ev_in_free[slot], fe->s_comp));
        rc = dsdneo_b200_full_demod_batch(fe->bank, fe->d_chan, fe->chan_pitch, fe->block_pairs, 1, fe->d_out[slot],
                                          (size_t)fe->block_pairs, fe->s_comp);
        if (rc) {
            return rc;
        }
        DSDNEO_CUDA(cudaEventRecord(fe->ev_comp[slot], fe->s_comp));

        DSDNEO_CUDA(cudaStreamWaitEvent(fe->s_d2h, fe->ev_comp[slot], 0));
        DSDNEO_CUDA(cudaMemcpy2DAsync(h_result + (size_t)b * fe->block_pairs, result_pitch * sizeof(float), fe->d_out[slot],
                                      (size_t)fe->block_pairs * sizeof(float), (size_t)fe->block_pairs * sizeof(float),
                                      (size_t)fe->M, cudaMemcpyDeviceToHost, fe->s_d2h));
        DSDNEO_CUDA(cudaEventRecord(fe->ev_d2h[slot], fe->s_d2h));
    }
    const long long ticket = (long long)fe->host_tickets++;
    DSDNEO_CUDA(cudaEventRecord(fe->ev_ticket[ticket & 3], fe->s_d2h)); /* the tile's last D2H copy */
    return ticket;
}

/** Blocks until the tile queued under `ticket` is complete in the caller's result buffer. */
int
dsdneo_b200_frontend_wait_host(dsdneo_b200_frontend* fe, long long ticket) {
    if (!fe || !fe->streams_ready || ticket < 0 || (unsigned long long)ticket >= fe->host_tickets) {
        set_error("frontend_wait_host: unknown ticket");
        return DSDNEO_B200_EINVAL;
    }
    /* If the ticket's event slot was reused (> 4 tiles later), the slot now holds a LATER tile recorded on the same
     * in-order D2H stream: waiting on it is conservative and still correct. */
    DSDNEO_CUDA(cudaEventSynchronize(fe->ev_ticket[ticket & 3]));
    return 0;
}

int
dsdneo_b200_frontend_process_host(dsdneo_b200_frontend* fe, const void* h_wideband, size_t n_in_samples,
                                  float* h_result, size_t result_pitch) {
    const long long t = dsdneo_b200_frontend_submit_host(fe, h_wideband, n_in_samples, h_result, result_pitch);
    if (t < 0) {
        return (int)t;
    }
    return dsdneo_b200_frontend_wait_host(fe, t);
}

} /* extern "C" */
