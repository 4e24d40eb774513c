/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference sample side -- getSymbol() (src/dsp/dsd_symbol.c:1853-1880)
 * and getDibitSoft() (src/core/frames/dsd_dibit.c:1078-1089) -- the way the reference's own tests do
 * (tests/dsp/test_rtl_symbol_cache_generation.c): memset opts/state, install dsd_rtl_stream_io_hooks.read and
 * dsd_rtl_stream_metrics_hooks that report an FSK-discriminator stream, then pull symbols.  The reference keeps the
 * matched-filter state in process globals (src/dsp/dsd_filters.c:325-345), so only ONE handle may be live at a time.
 */
#include <dsd-neo/core/dibit.h>
#include <dsd-neo/core/opts.h>
#include <dsd-neo/core/state.h>
#include <dsd-neo/core/synctype_ids.h>
#include <dsd-neo/dsp/sps_filters.h>
#include <dsd-neo/dsp/symbol.h>
#include <dsd-neo/dsp/sync_calibration.h>
#include <dsd-neo/io/rtl_stream_c.h>
#include <dsd-neo/runtime/rtl_stream_io_hooks.h>
#include <dsd-neo/runtime/rtl_stream_metrics_hooks.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

extern int g_oracle_shutdown_requested;

typedef struct ref_sym {
    dsd_opts* opts;
    dsd_state* state;
    const float* src;
    long n_src, pos;
    int out_rate, sym_rate, profile;
} ref_sym;

static ref_sym* g_live = NULL;
static int g_fake_ctx;

static int
hook_read(void* ctx, float* out, size_t count, int* out_got) {
    (void)ctx;
    ref_sym* h = g_live;
    long left = h->n_src - h->pos;
    long n = (long)count < left ? (long)count : left;
    if (n <= 0) {
        extern volatile uint8_t exitflag;
        exitflag = 1; /* getFrameSync() polls this and returns instead of hunting forever */
        *out_got = 0;
        return -1;
    }
    memcpy(out, h->src + h->pos, (size_t)n * sizeof(float));
    h->pos += n;
    *out_got = (int)n;
    return 0;
}

static double
hook_pwr(const void* ctx) {
    (void)ctx;
    return 0.0;
}

static int g_kind = RTL_STREAM_OUTPUT_FSK_DISCRIMINATOR, g_cqpsk_active = 0;
static double g_snr_cqpsk = -100.0;
static int hook_output_kind(void) { return g_kind; }
static int
hook_cqpsk_status(int* cqpsk, int* timing) {
    if (cqpsk) {
        *cqpsk = g_cqpsk_active;
    }
    if (timing) {
        *timing = g_cqpsk_active;
    }
    return 0;
}
static double hook_snr_cqpsk(void) { return g_snr_cqpsk; }
static unsigned int hook_output_rate(void) { return (unsigned int)g_live->out_rate; }
static uint32_t hook_generation(void) { return 1U; }

static int
hook_symbol_profile(int* rate, int* levels, int* profile) {
    if (rate) {
        *rate = g_live->sym_rate;
    }
    if (levels) {
        *levels = 4;
    }
    if (profile) {
        *profile = g_live->profile;
    }
    return 0;
}

void*
ref_sym_create(int output_rate_hz, int symbol_rate_hz, int synctype, int lastsynctype, int use_cosine_filter, int ssize,
               int msize) {
    ref_sym* h = (ref_sym*)calloc(1, sizeof(*h));
    h->opts = (dsd_opts*)calloc(1, sizeof(dsd_opts));
    h->state = (dsd_state*)calloc(1, sizeof(dsd_state));
    h->out_rate = output_rate_hz;
    h->sym_rate = symbol_rate_hz;
    h->profile = RTL_STREAM_CHANNEL_PROFILE_P25_C4FM;
    dsd_opts* o = h->opts;
    dsd_state* s = h->state;
    /* the subset of initOpts()/initState() that the sample side reads (src/core/util/dsd_init.c:169-177,519-592) */
    o->audio_in_type = AUDIO_IN_RTL;
    o->ssize = ssize;
    o->msize = msize;
    o->use_cosine_filter = use_cosine_filter;
    s->rf_mod = 0;
    s->jitter = -1;
    s->synctype = synctype;
    s->lastsynctype = lastsynctype;
    s->min = -15000;
    s->max = 15000;
    s->minref = -12000;
    s->maxref = 12000;
    for (int i = 0; i < 1024; i++) {
        s->maxbuf[i] = 15000;
        s->minbuf[i] = -15000;
    }
    s->samplesPerSymbol = 10;
    s->symbolCenter = 4;
    s->rtl_ctx = (struct RtlSdrContext*)&g_fake_ctx;
    s->dibit_buf = (int*)calloc(1000000, sizeof(int));
    s->dibit_buf_p = s->dibit_buf + 200;
    s->dmr_payload_buf = (int*)calloc(1000000, sizeof(int));
    s->dmr_payload_p = s->dmr_payload_buf + 200;
    s->dmr_soft_buf = (dsd_dibit_soft_t*)calloc(1000000, sizeof(dsd_dibit_soft_t));
    s->dmr_soft_p = s->dmr_soft_buf + 200;
    g_live = h;
    init_rrc_filter_memory();
    dsd_rtl_stream_io_hooks io;
    memset(&io, 0, sizeof(io));
    io.read = hook_read;
    io.return_pwr = hook_pwr;
    dsd_rtl_stream_io_hooks_set(io);
    dsd_rtl_stream_metrics_hooks mh;
    memset(&mh, 0, sizeof(mh));
    mh.output_kind = hook_output_kind;
    mh.output_rate_hz = hook_output_rate;
    mh.symbol_profile = hook_symbol_profile;
    mh.stream_generation = hook_generation;
    g_kind = RTL_STREAM_OUTPUT_FSK_DISCRIMINATOR;
    g_cqpsk_active = 0;
    g_snr_cqpsk = -100.0;
    dsd_rtl_stream_metrics_hooks_set(&mh);
    g_oracle_shutdown_requested = 0;
    return h;
}

/* Symbol-rate CQPSK stream (output kind 2, src/dsp/dsd_symbol.c:1583-1625): the hooks report what the CQPSK block side
 * reports (kind 2, profile P25_CQPSK, cqpsk_status, the CQPSK SNR), state->rf_mod = 1 as the modulation detector / -mq set it. */
void*
ref_sym_create_cqpsk(int symbol_rate_hz, int synctype, int lastsynctype, int ssize, int msize, int map_idx, int cqpsk_active,
                     double snr_db) {
    ref_sym* h = (ref_sym*)ref_sym_create(symbol_rate_hz, symbol_rate_hz, synctype, lastsynctype, 0, ssize, msize);
    h->profile = RTL_STREAM_CHANNEL_PROFILE_P25_CQPSK;
    h->state->rf_mod = 1;
    h->state->p25_cqpsk_dibit_map_idx = (uint8_t)map_idx;
    g_kind = RTL_STREAM_OUTPUT_SYMBOL_CQPSK;
    g_cqpsk_active = cqpsk_active;
    g_snr_cqpsk = snr_db;
    dsd_rtl_stream_metrics_hooks mh;
    memset(&mh, 0, sizeof(mh));
    mh.output_kind = hook_output_kind;
    mh.output_rate_hz = hook_output_rate;
    mh.symbol_profile = hook_symbol_profile;
    mh.stream_generation = hook_generation;
    mh.cqpsk_status = hook_cqpsk_status;
    mh.snr_cqpsk_db = hook_snr_cqpsk;
    dsd_rtl_stream_metrics_hooks_set(&mh);
    return h;
}

void
ref_sym_destroy(void* hv) {
    ref_sym* h = (ref_sym*)hv;
    if (!h) {
        return;
    }
    if (g_live == h) {
        g_live = NULL;
    }
    free(h->state->dibit_buf);
    free(h->state->dmr_payload_buf);
    free(h->state->dmr_soft_buf);
    free(h->state);
    free(h->opts);
    free(h);
}

void
ref_sym_feed(void* hv, const float* samples, long n) {
    ref_sym* h = (ref_sym*)hv;
    h->src = samples;
    h->n_src = n;
    h->pos = 0;
}

/* samples the reference has not consumed yet = unread source + what sits in its 512-float symbol cache */
static long
samples_left(const ref_sym* h) {
    long cached = h->state->rtl_symbol_cache_len - h->state->rtl_symbol_cache_pos;
    if (cached < 0) {
        cached = 0;
    }
    return (h->n_src - h->pos) + cached;
}

void
ref_sym_set_sync(void* hv, int synctype, int lastsynctype) {
    ref_sym* h = (ref_sym*)hv;
    h->state->synctype = synctype;
    h->state->lastsynctype = lastsynctype;
}

/* getSymbol(opts, state, have_sync) until fewer than `reserve` samples remain.  Returns symbols produced. */
long
ref_sym_get_symbols(void* hv, int have_sync, long max_symbols, long reserve, float* out) {
    ref_sym* h = (ref_sym*)hv;
    g_live = h;
    long n = 0;
    while (n < max_symbols && samples_left(h) >= reserve) {
        out[n++] = getSymbol(h->opts, h->state, have_sync);
        if (g_oracle_shutdown_requested) {
            return -1;
        }
    }
    return n;
}

/* getDibitSoft() loop (have_sync = 1 inside the reference).  soft5: {reliability, llr0 lo, llr0 hi, llr1 lo, llr1 hi}. */
long
ref_sym_get_dibits(void* hv, long max_symbols, long reserve, uint8_t* dibits, uint8_t* reliab, int16_t* llr2, float* symbols) {
    ref_sym* h = (ref_sym*)hv;
    g_live = h;
    long n = 0;
    while (n < max_symbols && samples_left(h) >= reserve) {
        int sidx_before = h->state->sidx;
        dsd_dibit_soft_t soft;
        int d = getDibitSoft(h->opts, h->state, &soft);
        if (g_oracle_shutdown_requested) {
            return -1;
        }
        dibits[n] = (uint8_t)d;
        reliab[n] = soft.reliability;
        llr2[2 * n] = soft.llr[0];
        llr2[2 * n + 1] = soft.llr[1];
        symbols[n] = h->state->sbuf[sidx_before]; /* get_dibit_and_analog_signal stores the symbol at sbuf[sidx] first */
        n++;
    }
    return n;
}

/* {min, max, center, umid, lmid, minref, maxref, lastsample} + {samplesPerSymbol, symbolCenter, jitter, sidx, midx} */
void
ref_sym_get_state(void* hv, float* f8, int* i5) {
    const dsd_state* s = ((ref_sym*)hv)->state;
    f8[0] = s->min, f8[1] = s->max, f8[2] = s->center, f8[3] = s->umid, f8[4] = s->lmid, f8[5] = s->minref, f8[6] = s->maxref;
    f8[7] = s->lastsample;
    i5[0] = s->samplesPerSymbol, i5[1] = s->symbolCenter, i5[2] = s->jitter, i5[3] = s->sidx, i5[4] = s->midx;
}

long
ref_sym_consumed(void* hv) {
    ref_sym* h = (ref_sym*)hv;
    return h->n_src - samples_left(h);
}

/* Normalised taps of one of the reference's matched filters at `sps`, read out as an impulse response
 * (apply_sps_fir, src/dsp/dsd_filters.c:172-201).  which: 0 p25, 1 dmr, 2 nxdn, 3 dpmr, 4 m17.  Returns the tap count. */
int
ref_sps_fir_taps(int which, int sps, float* taps_out, int cap) {
    float (*fn)(float, int) = which == 0 ? p25_filter : which == 1 ? dmr_filter : which == 2 ? nxdn_filter : which == 3 ? dpmr_filter : m17_filter;
    init_rrc_filter_memory();
    float resp[2048];
    int last_nz = -1;
    for (int n = 0; n < 2048; n++) {
        resp[n] = fn(n == 0 ? 1.0f : 0.0f, sps);
        if (resp[n] != 0.0f) {
            last_nz = n;
        }
    }
    init_rrc_filter_memory();
    int len = last_nz + 1;
    /* output n of the impulse response = taps[len-1-n] (newest sample meets the LAST tap) */
    if (len > cap) {
        return -len;
    }
    for (int n = 0; n < len; n++) {
        taps_out[len - 1 - n] = resp[n];
    }
    return len;
}


/* Drives the reference's own capture writer (write_symbol_capture_record, src/core/frames/dsd_dibit.c:794-818) after the
 * 16-byte header openSymbolOutFile emits (src/core/file/dsd_file.c:876-888; that TU needs libsndfile, so its header
 * constant is rebuilt here from the same macros of include/dsd-neo/core/dibit.h:35-37). */
int
ref_symbol_capture_write(const char* path, const unsigned char* dibits, const unsigned char* reliab, const short* llr,
                         const float* symbols, int n) {
    dsd_opts* o = (dsd_opts*)calloc(1, sizeof(dsd_opts));
    dsd_state* s = (dsd_state*)calloc(1, sizeof(dsd_state));
    if (!o || !s) {
        return -1;
    }
    o->symbol_out_f = fopen(path, "wb");
    if (!o->symbol_out_f) {
        return -1;
    }
    unsigned char header[DSD_SYMBOL_CAPTURE_SOFT_HEADER_SIZE];
    memset(header, 0, sizeof(header));
    memcpy(header, DSD_SYMBOL_CAPTURE_SOFT_MAGIC, 8);
    header[8] = 2;
    header[9] = DSD_SYMBOL_CAPTURE_SOFT_RECORD_SIZE;
    fwrite(header, 1, sizeof(header), o->symbol_out_f);
    for (int i = 0; i < n; i++) {
        dsd_dibit_soft_t soft;
        soft.reliability = reliab[i];
        soft.llr[0] = llr[2 * i];
        soft.llr[1] = llr[2 * i + 1];
        write_symbol_capture_record(o, s, dibits[i], symbols[i], &soft);
    }
    fclose(o->symbol_out_f);
    int written = (int)s->symbol_capture_soft_records;
    free(o);
    free(s);
    return written;
}


/* External stream server: installs caller-supplied functions into the reference's two runtime hook tables
 * (include/dsd-neo/runtime/rtl_stream_io_hooks.h:25-32, rtl_stream_metrics_hooks.h:28-50) exactly as
 * src/engine/rtl_stream_io_hooks_install.c does, and hands `ctx` to the reader through state->rtl_ctx. */
void
ref_sym_use_external_hooks(void* hv, void* read_fn, void* pwr_fn, void* ctx, void* rate_fn, void* kind_fn, void* profile_fn,
                           void* generation_fn) {
    ref_sym* h = (ref_sym*)hv;
    h->state->rtl_ctx = (struct RtlSdrContext*)ctx;
    dsd_rtl_stream_io_hooks io;
    memset(&io, 0, sizeof(io));
    io.read = (int (*)(void*, float*, size_t, int*))read_fn;
    io.return_pwr = (double (*)(const void*))pwr_fn;
    dsd_rtl_stream_io_hooks_set(io);
    dsd_rtl_stream_metrics_hooks mh;
    memset(&mh, 0, sizeof(mh));
    mh.output_rate_hz = (unsigned int (*)(void))rate_fn;
    mh.output_kind = (int (*)(void))kind_fn;
    mh.symbol_profile = (int (*)(int*, int*, int*))profile_fn;
    mh.stream_generation = (uint32_t (*)(void))generation_fn;
    dsd_rtl_stream_metrics_hooks_set(&mh);
}

/* Same, for a symbol-rate CQPSK stream: also installs the cqpsk_status and snr_cqpsk_db hooks and puts the decoder state
 * where the modulation detector / -mq leave it (rf_mod = 1, orientation map). */
void
ref_sym_use_external_hooks_cqpsk(void* hv, void* read_fn, void* pwr_fn, void* ctx, void* rate_fn, void* kind_fn, void* profile_fn,
                                 void* generation_fn, void* cqpsk_status_fn, void* snr_cqpsk_fn, int map_idx) {
    ref_sym* h = (ref_sym*)hv;
    h->state->rtl_ctx = (struct RtlSdrContext*)ctx;
    h->state->rf_mod = 1;
    h->state->p25_cqpsk_dibit_map_idx = (uint8_t)map_idx;
    dsd_rtl_stream_io_hooks io;
    memset(&io, 0, sizeof(io));
    io.read = (int (*)(void*, float*, size_t, int*))read_fn;
    io.return_pwr = (double (*)(const void*))pwr_fn;
    dsd_rtl_stream_io_hooks_set(io);
    dsd_rtl_stream_metrics_hooks mh;
    memset(&mh, 0, sizeof(mh));
    mh.output_rate_hz = (unsigned int (*)(void))rate_fn;
    mh.output_kind = (int (*)(void))kind_fn;
    mh.symbol_profile = (int (*)(int*, int*, int*))profile_fn;
    mh.stream_generation = (uint32_t (*)(void))generation_fn;
    mh.cqpsk_status = (int (*)(int*, int*))cqpsk_status_fn;
    mh.snr_cqpsk_db = (double (*)(void))snr_cqpsk_fn;
    dsd_rtl_stream_metrics_hooks_set(&mh);
}

/* exactly n getDibitSoft() calls (the caller knows the stream holds enough samples) */
long
ref_sym_get_dibits_n(void* hv, long n_symbols, uint8_t* dibits, uint8_t* reliab, int16_t* llr2, float* symbols) {
    ref_sym* h = (ref_sym*)hv;
    g_live = h;
    for (long n = 0; n < n_symbols; n++) {
        int sidx_before = h->state->sidx;
        dsd_dibit_soft_t soft;
        int d = getDibitSoft(h->opts, h->state, &soft);
        if (g_oracle_shutdown_requested) {
            return -1;
        }
        dibits[n] = (uint8_t)d;
        reliab[n] = soft.reliability;
        llr2[2 * n] = soft.llr[0];
        llr2[2 * n + 1] = soft.llr[1];
        symbols[n] = h->state->sbuf[sidx_before];
    }
    return n_symbols;
}


/* ---- acquisition: the UNMODIFIED getFrameSync() (src/dsp/dsd_frame_sync.c:3098-3148) on the hook-fed stream ------------------
 * One call hunts symbol by symbol (getSymbol with have_sync = 0, hunt-time slicing, the rolling DMR payload buffer, pattern
 * matching) and, on a match, runs the reference's own acquisition side effects: sync warm start of the slicer thresholds
 * (dsd_sync_warm_start_thresholds_outer_only, src/dsp/sync_calibration.c:155-226) and, for DMR, the re-digitisation of the
 * 66 dibits in front of the sync (dmr_resample_on_sync, src/dsp/dmr_sync.c:108-126). */
extern volatile uint8_t exitflag;
int getFrameSync(dsd_opts* opts, dsd_state* state);

/* frame_mask: bit 0 P25 Phase 1, bit 1 DMR.  rf_mod: 0 C4FM, 2 GFSK (what the reference's -fs preset selects for DMR,
 * src/runtime/decode_mode.c:242-266); the modulation is locked like -mc / -mg do, so the auto-detector stays out. */
void
ref_sym_configure_acquire(void* hv, int frame_mask, int rf_mod) {
    ref_sym* h = (ref_sym*)hv;
    dsd_opts* o = h->opts;
    dsd_state* s = h->state;
    o->frame_p25p1 = (frame_mask & 1) ? 1 : 0;
    o->frame_dmr = (frame_mask & 2) ? 1 : 0;
    o->inverted_dmr = (frame_mask & 4) ? 1 : 0; /* the -xr switch */
    o->mod_cli_lock = 1;
    o->mod_c4fm = rf_mod == 0;
    o->mod_gfsk = rf_mod == 2;
    o->mod_qpsk = 0;
    s->rf_mod = rf_mod;
    s->synctype = DSD_SYNC_NONE;
    s->lastsynctype = DSD_SYNC_NONE;
    if (!s->symbol_history) {
        s->symbol_history_size = DSD_SYMBOL_HISTORY_SIZE;
        s->symbol_history = (float*)calloc((size_t)s->symbol_history_size, sizeof(float));
        s->symbol_history_head = 0;
        s->symbol_history_count = 0;
    }
}

/* Returns the sync type getFrameSync() reported (DSD_SYNC_NONE / -1 when the samples ran out first). */
int
ref_sym_frame_sync(void* hv) {
    ref_sym* h = (ref_sym*)hv;
    g_live = h;
    exitflag = 0;
    const int st = getFrameSync(h->opts, h->state);
    if (st >= 0) {
        h->state->synctype = st;
    }
    exitflag = 0;
    return st;
}

/* the most recent `n` entries of the reference's symbol history (oldest first) and of its rolling DMR payload / soft buffers */
void
ref_sym_recent(void* hv, int n, float* symbols, int32_t* payload_dibits, uint8_t* payload_reliab) {
    ref_sym* h = (ref_sym*)hv;
    const dsd_state* s = h->state;
    for (int i = 0; i < n; i++) {
        const int back = n - 1 - i;
        float v = 0.0f;
        if (s->symbol_history && back < s->symbol_history_count) {
            int idx = s->symbol_history_head - 1 - back;
            while (idx < 0) {
                idx += s->symbol_history_size;
            }
            v = s->symbol_history[idx % s->symbol_history_size];
        }
        symbols[i] = v;
        payload_dibits[i] = *(s->dmr_payload_p - n + i);
        payload_reliab[i] = s->dmr_soft_p ? (s->dmr_soft_p - n + i)->reliability : 0;
    }
}

long
ref_sym_symbol_count(void* hv) {
    return (long)((ref_sym*)hv)->state->symbol_history_count;
}
