// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * Per-channel symbol stream history on the device (include/dsdneo_b200.h, "symbol stream" section).
 *
 * The reference decodes a channel as one sequential stream: a frame reader looks BACK into dibits it has already taken (the
 * DMR burst readers keep a rolling buffer and unpack 90 dibits from it at every sync: src/protocol/dmr/dmr_data.c:118-157,
 * src/protocol/dmr/dmr_bs.c:137-148) and keeps reading past the end of whatever block of samples happened to arrive.  The
 * batched slicer produces one launch's symbols at a time, so something has to join the launches: this object keeps the last
 * `keep` symbols / dibits / reliabilities / LLRs of every channel in front of the next launch's outputs (ping-pong rows, the
 * slicer writes its outputs straight behind the history, no extra copy of the new symbols), and counts stream positions.
 * A caller runs the sync hunt `delay` symbols behind the slicer, so that every sync is found exactly once and with its whole
 * frame present, and the frame cutters on the joined rows.  The P25 Phase 1 receive bank (p25p1_rx.cu) has the same layout
 * built in; this is the stand-alone form for the other cutters (DMR data / voice bursts).
 */
#include "common.cuh"

using namespace dsdneo;

namespace {

/* the last `keep` entries of the previous rows (valid length keep + count_prev) become the head of the current rows */
__global__ void
symstream_tail_kernel(const uint8_t* dib_prev, const uint8_t* rel_prev, const short2* llr_prev, const float* sym_prev, uint8_t* dib,
                      uint8_t* rel, short2* llr, float* sym, const int* count_prev, size_t pitch, int keep) {
    const int ch = blockIdx.x;
    const int src0 = count_prev[ch];
    const size_t row = (size_t)ch * pitch;
    for (int i = threadIdx.x; i < keep; i += blockDim.x) {
        dib[row + i] = dib_prev[row + src0 + i];
        rel[row + i] = rel_prev[row + src0 + i];
        llr[row + i] = llr_prev[row + src0 + i];
        sym[row + i] = sym_prev[row + src0 + i];
    }
}

__global__ void
symstream_account_kernel(int* count_new, int* valid, long long* stream_base, long long* total, int n_ch, int keep, int max_new) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) {
        return;
    }
    const int n = min(max(count_new[c], 0), max_new);
    count_new[c] = n;
    stream_base[c] = total[c] - keep; /* row index 0 = stream index (symbols before this launch) - keep */
    total[c] += n;
    valid[c] = keep + n;
}

__global__ void
symstream_rebase_kernel(int32_t* hits, const int32_t* n_hits, int n_ch, int max_hits, int offset) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ch * max_hits) {
        return;
    }
    const int c = i / max_hits, h = i - c * max_hits;
    if (h < min(n_hits[c], max_hits)) {
        hits[(size_t)i * 2] += offset;
    }
}

/* one thread per channel: the hits of one sync type, in stream order */
__global__ void
sync_hits_select_kernel(const int32_t* hits, const int32_t* n_hits, int n_ch, int max_hits, int sync_type, int32_t* out, int out_max,
                        int32_t* n_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) {
        return;
    }
    const int n = min(n_hits[c], max_hits);
    int k = 0;
    for (int h = 0; h < n; h++) {
        const int32_t pos = hits[((size_t)c * max_hits + h) * 2], typ = hits[((size_t)c * max_hits + h) * 2 + 1];
        if (typ == sync_type) {
            if (k < out_max) {
                out[((size_t)c * out_max + k) * 2] = pos;
                out[((size_t)c * out_max + k) * 2 + 1] = typ;
            }
            k++;
        }
    }
    n_out[c] = min(k, out_max);
}

}  // namespace

struct dsdneo_b200_symbol_stream {
    int n_ch, keep, max_new, phase, open;
    size_t pitch;
    uint8_t *d_dib[2], *d_rel[2];
    int16_t* d_llr[2];
    float* d_sym[2];
    int *d_count[2], *d_valid;
    long long *d_base, *d_total;
};

extern "C" {

void
dsdneo_b200_symbol_stream_destroy(dsdneo_b200_symbol_stream* ss) {
    if (!ss) {
        return;
    }
    for (int b = 0; b < 2; b++) {
        cudaFree(ss->d_dib[b]);
        cudaFree(ss->d_rel[b]);
        cudaFree(ss->d_llr[b]);
        cudaFree(ss->d_sym[b]);
        cudaFree(ss->d_count[b]);
    }
    cudaFree(ss->d_valid);
    cudaFree(ss->d_base);
    cudaFree(ss->d_total);
    free(ss);
}

dsdneo_b200_symbol_stream*
dsdneo_b200_symbol_stream_create(int n_channels, int keep, int max_new) {
    if (n_channels <= 0 || keep <= 0 || max_new <= 0 || (keep & 31) != 0) {
        set_error("symbol_stream_create: bad argument (keep must be a positive multiple of 32)");
        return NULL;
    }
    if (ensure_device()) {
        return NULL;
    }
    dsdneo_b200_symbol_stream* ss = (dsdneo_b200_symbol_stream*)calloc(1, sizeof(*ss));
    if (!ss) {
        return NULL;
    }
    ss->n_ch = n_channels, ss->keep = keep, ss->max_new = max_new;
    ss->pitch = ((size_t)keep + (size_t)max_new + 31) & ~(size_t)31; /* rows start on 128-byte lines in every array */
    const size_t n = (size_t)n_channels, cells = n * ss->pitch;
    cudaError_t e = cudaSuccess;
#define SS_ALLOC(ptr, bytes)                                                                                           \
    if (e == cudaSuccess) {                                                                                            \
        e = cudaMalloc((void**)&(ptr), (bytes));                                                                       \
        if (e == cudaSuccess) {                                                                                        \
            e = cudaMemset((ptr), 0, (bytes));                                                                         \
        }                                                                                                              \
    }
    for (int b = 0; b < 2; b++) {
        SS_ALLOC(ss->d_dib[b], cells);
        SS_ALLOC(ss->d_rel[b], cells);
        SS_ALLOC(ss->d_llr[b], cells * 2 * sizeof(int16_t));
        SS_ALLOC(ss->d_sym[b], cells * sizeof(float));
        SS_ALLOC(ss->d_count[b], n * sizeof(int));
    }
    SS_ALLOC(ss->d_valid, n * sizeof(int));
    SS_ALLOC(ss->d_base, n * sizeof(long long));
    SS_ALLOC(ss->d_total, n * sizeof(long long));
#undef SS_ALLOC
    if (e != cudaSuccess) {
        cuda_fail(e, "symbol_stream_create", __FILE__, __LINE__);
        dsdneo_b200_symbol_stream_destroy(ss);
        return NULL;
    }
    return ss;
}

int
dsdneo_b200_symbol_stream_reset(dsdneo_b200_symbol_stream* ss, void* stream) {
    if (!ss) {
        set_error("symbol_stream_reset: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    cudaStream_t s = as_stream(stream);
    const size_t n = (size_t)ss->n_ch, cells = n * ss->pitch;
    for (int b = 0; b < 2; b++) {
        DSDNEO_CUDA(cudaMemsetAsync(ss->d_dib[b], 0, cells, s));
        DSDNEO_CUDA(cudaMemsetAsync(ss->d_rel[b], 0, cells, s));
        DSDNEO_CUDA(cudaMemsetAsync(ss->d_llr[b], 0, cells * 2 * sizeof(int16_t), s));
        DSDNEO_CUDA(cudaMemsetAsync(ss->d_sym[b], 0, cells * sizeof(float), s));
        DSDNEO_CUDA(cudaMemsetAsync(ss->d_count[b], 0, n * sizeof(int), s));
    }
    DSDNEO_CUDA(cudaMemsetAsync(ss->d_valid, 0, n * sizeof(int), s));
    DSDNEO_CUDA(cudaMemsetAsync(ss->d_base, 0, n * sizeof(long long), s));
    DSDNEO_CUDA(cudaMemsetAsync(ss->d_total, 0, n * sizeof(long long), s));
    ss->phase = 0, ss->open = 0;
    return 0;
}

int
dsdneo_b200_symbol_stream_begin(dsdneo_b200_symbol_stream* ss, dsdneo_b200_symbol_out* out) {
    if (!ss || !out || ss->open) {
        set_error("symbol_stream_begin: bad argument, or the previous launch was not committed");
        return DSDNEO_B200_EINVAL;
    }
    const int cur = ss->phase;
    out->d_symbols = ss->d_sym[cur] + ss->keep;
    out->d_dibits = ss->d_dib[cur] + ss->keep;
    out->d_reliability = ss->d_rel[cur] + ss->keep;
    out->d_llr = ss->d_llr[cur] + 2 * (size_t)ss->keep;
    out->d_count = ss->d_count[cur];
    out->pitch = ss->pitch; /* row pitch of the arrays; the slicer may add at most max_new symbols per channel */
    ss->open = 1;
    return 0;
}

int
dsdneo_b200_symbol_stream_commit(dsdneo_b200_symbol_stream* ss, dsdneo_b200_symbol_stream_view* view, void* stream) {
    if (!ss || !view || !ss->open) {
        set_error("symbol_stream_commit: bad argument, or no launch is open (symbol_stream_begin)");
        return DSDNEO_B200_EINVAL;
    }
    cudaStream_t s = as_stream(stream);
    const int cur = ss->phase, prev = cur ^ 1;
    {
        KernelTimer kt("symstream_tail_kernel", s);
        symstream_tail_kernel<<<ss->n_ch, 256, 0, s>>>(ss->d_dib[prev], ss->d_rel[prev], (const short2*)ss->d_llr[prev], ss->d_sym[prev],
                                                      ss->d_dib[cur], ss->d_rel[cur], (short2*)ss->d_llr[cur], ss->d_sym[cur],
                                                      ss->d_count[prev], ss->pitch, ss->keep);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    {
        KernelTimer kt("symstream_account_kernel", s);
        symstream_account_kernel<<<(ss->n_ch + 127) / 128, 128, 0, s>>>(ss->d_count[cur], ss->d_valid, ss->d_base, ss->d_total, ss->n_ch,
                                                                        ss->keep, ss->max_new);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    view->d_symbols = ss->d_sym[cur];
    view->d_dibits = ss->d_dib[cur];
    view->d_reliability = ss->d_rel[cur];
    view->d_llr = ss->d_llr[cur];
    view->pitch = ss->pitch;
    view->d_valid = ss->d_valid;
    view->d_new = ss->d_count[cur];
    view->d_stream_base = ss->d_base;
    view->keep = ss->keep;
    ss->phase ^= 1;
    ss->open = 0;
    return 0;
}

int
dsdneo_b200_sync_hits_rebase(void* d_hits, const int32_t* d_n_hits, int n_channels, int max_hits, int offset, void* stream) {
    if (!d_hits || !d_n_hits || n_channels <= 0 || max_hits <= 0) {
        set_error("sync_hits_rebase: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("symstream_rebase_kernel", s);
        symstream_rebase_kernel<<<(n_channels * max_hits + 255) / 256, 256, 0, s>>>((int32_t*)d_hits, d_n_hits, n_channels, max_hits, offset);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

int
dsdneo_b200_sync_hits_select(const void* d_hits, const int32_t* d_n_hits, int n_channels, int max_hits, int sync_type, void* d_hits_out,
                             int out_max_hits, int32_t* d_n_out, void* stream) {
    if (!d_hits || !d_n_hits || !d_hits_out || !d_n_out || n_channels <= 0 || max_hits <= 0 || out_max_hits <= 0) {
        set_error("sync_hits_select: bad argument");
        return DSDNEO_B200_EINVAL;
    }
    int rc = ensure_device();
    if (rc) {
        return rc;
    }
    cudaStream_t s = as_stream(stream);
    {
        KernelTimer kt("sync_hits_select_kernel", s);
        sync_hits_select_kernel<<<(n_channels + 127) / 128, 128, 0, s>>>((const int32_t*)d_hits, d_n_hits, n_channels, max_hits, sync_type,
                                                                        (int32_t*)d_hits_out, out_max_hits, d_n_out);
    }
    DSDNEO_KERNEL_CHECK();
    count_launch();
    return 0;
}

} /* extern "C" */
