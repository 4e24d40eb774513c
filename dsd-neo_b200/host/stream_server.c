// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * GPU-backed stream server for the reference's one real runtime seam: the unmodified
 * decoder thread pulls discriminator floats through
 *     dsd_rtl_stream_io_hooks      { read, return_pwr }                  include/dsd-neo/runtime/rtl_stream_io_hooks.h:25-28
 *     dsd_rtl_stream_metrics_hooks { output_rate_hz, output_kind, symbol_profile, stream_generation, ... }
 *                                                                       include/dsd-neo/runtime/rtl_stream_metrics_hooks.h:28-47
 * (installed by src/engine/rtl_stream_io_hooks_install.c:26-34; consumer src/dsp/dsd_symbol.c:889-943,1412-1435: one float
 * or a refill of up to 512 floats per call, blocking, < 0 on end of stream).  One server = one monitored channel: the
 * ingest loop pushes that channel's rows of the front end's output, the reference drains them.  Plain host C (a mutex +
 * condition-variable float ring); nothing here touches the device.
 *
 * `read` receives the server handle through rtl_ctx (the engine stores it in state->rtl_ctx).  The metrics hooks have no
 * context argument in the reference, so they answer for the server made current with ..._make_current().
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dsdneo_b200.h"

struct dsdneo_b200_stream_server {
    float* ring;
    size_t cap, head, count; /* head = index of the oldest float */
    int closed;
    unsigned int output_rate_hz;
    int symbol_rate_hz, levels, channel_profile;
    int output_kind;   /* 1 = FSK discriminator samples, 2 = symbol-rate CQPSK symbols */
    int cqpsk_active;  /* what cqpsk_status reports for {cqpsk_enable, timing_active} */
    double snr_cqpsk_db;
    uint32_t generation;
    double pwr;
    pthread_mutex_t mu;
    pthread_cond_t can_read, can_write;
};

/* One current server per process, as the reference has one decoder per process.  The pointer is atomic so the ingest thread
 * can switch it while the decoder thread's hooks read it; a hook loads it once per call. */
static _Atomic(dsdneo_b200_stream_server*) g_current_srv = NULL;
static inline dsdneo_b200_stream_server*
current_srv(void) {
    return atomic_load_explicit(&g_current_srv, memory_order_acquire);
}

dsdneo_b200_stream_server*
dsdneo_b200_stream_server_create(size_t ring_floats, unsigned int output_rate_hz, int symbol_rate_hz, int levels, int channel_profile) {
    if (ring_floats < 1024 || output_rate_hz == 0) {
        return NULL;
    }
    dsdneo_b200_stream_server* s = (dsdneo_b200_stream_server*)calloc(1, sizeof(*s));
    if (!s) {
        return NULL;
    }
    s->ring = (float*)malloc(ring_floats * sizeof(float));
    if (!s->ring) {
        free(s);
        return NULL;
    }
    s->cap = ring_floats;
    s->output_rate_hz = output_rate_hz;
    s->symbol_rate_hz = symbol_rate_hz;
    s->levels = levels;
    s->channel_profile = channel_profile;
    s->generation = 1;
    s->output_kind = 1;
    s->cqpsk_active = 0;
    s->snr_cqpsk_db = -100.0;
    pthread_mutex_init(&s->mu, NULL);
    pthread_cond_init(&s->can_read, NULL);
    pthread_cond_init(&s->can_write, NULL);
    return s;
}

void
dsdneo_b200_stream_server_destroy(dsdneo_b200_stream_server* s) {
    if (!s) {
        return;
    }
    {
        dsdneo_b200_stream_server* expect = s;
        atomic_compare_exchange_strong(&g_current_srv, &expect, NULL);
    }
    pthread_mutex_destroy(&s->mu);
    pthread_cond_destroy(&s->can_read);
    pthread_cond_destroy(&s->can_write);
    free(s->ring);
    free(s);
}

void
dsdneo_b200_stream_server_make_current(dsdneo_b200_stream_server* s) {
    atomic_store_explicit(&g_current_srv, s, memory_order_release);
}

/* producer side: blocks while the ring is full (back-pressure on the ingest loop) unless `block` is 0 */
size_t
dsdneo_b200_stream_server_push(dsdneo_b200_stream_server* s, const float* samples, size_t n, int block) {
    if (!s || !samples) {
        return 0;
    }
    size_t done = 0;
    pthread_mutex_lock(&s->mu);
    while (done < n && !s->closed) {
        if (s->count == s->cap) {
            if (!block) {
                break;
            }
            pthread_cond_wait(&s->can_write, &s->mu);
            continue;
        }
        const size_t tail = (s->head + s->count) % s->cap;
        size_t m = s->cap - s->count;
        if (m > n - done) {
            m = n - done;
        }
        if (m > s->cap - tail) {
            m = s->cap - tail;
        }
        memcpy(s->ring + tail, samples + done, m * sizeof(float));
        s->count += m;
        done += m;
        pthread_cond_signal(&s->can_read);
    }
    pthread_mutex_unlock(&s->mu);
    return done;
}

void
dsdneo_b200_stream_server_close(dsdneo_b200_stream_server* s) {
    if (!s) {
        return;
    }
    pthread_mutex_lock(&s->mu);
    s->closed = 1;
    pthread_cond_broadcast(&s->can_read);
    pthread_cond_broadcast(&s->can_write);
    pthread_mutex_unlock(&s->mu);
}

/* retune / reconfigure: buffered samples belong to the old channel; the consumer drops its own cache when it sees the
 * generation change (src/dsp/dsd_symbol.c:793-810) */
void
dsdneo_b200_stream_server_bump_generation(dsdneo_b200_stream_server* s) {
    if (!s) {
        return;
    }
    pthread_mutex_lock(&s->mu);
    s->head = s->count = 0;
    s->generation++;
    pthread_cond_broadcast(&s->can_write);
    pthread_mutex_unlock(&s->mu);
}

void
dsdneo_b200_stream_server_set_power(dsdneo_b200_stream_server* s, double pwr) {
    if (s) {
        pthread_mutex_lock(&s->mu);
        s->pwr = pwr;
        pthread_mutex_unlock(&s->mu);
    }
}

/* ---- dsd_rtl_stream_io_hooks ---- */

int
dsdneo_b200_stream_hook_read(void* rtl_ctx, float* out, size_t count, int* out_got) {
    dsdneo_b200_stream_server* s = (dsdneo_b200_stream_server*)rtl_ctx;
    if (out_got) {
        *out_got = 0;
    }
    if (!s || !out || count == 0) {
        return -1;
    }
    pthread_mutex_lock(&s->mu);
    while (s->count == 0 && !s->closed) {
        pthread_cond_wait(&s->can_read, &s->mu);
    }
    if (s->count == 0) { /* closed and drained */
        pthread_mutex_unlock(&s->mu);
        return -1;
    }
    size_t got = 0;
    while (got < count && s->count > 0) {
        size_t m = s->count;
        if (m > count - got) {
            m = count - got;
        }
        if (m > s->cap - s->head) {
            m = s->cap - s->head;
        }
        memcpy(out + got, s->ring + s->head, m * sizeof(float));
        s->head = (s->head + m) % s->cap;
        s->count -= m;
        got += m;
    }
    pthread_cond_signal(&s->can_write);
    pthread_mutex_unlock(&s->mu);
    if (out_got) {
        *out_got = (int)got;
    }
    return 0;
}

double
dsdneo_b200_stream_hook_return_pwr(const void* rtl_ctx) {
    dsdneo_b200_stream_server* s = (dsdneo_b200_stream_server*)rtl_ctx;
    if (!s) {
        return 0.0;
    }
    pthread_mutex_lock(&s->mu);
    const double p = s->pwr;
    pthread_mutex_unlock(&s->mu);
    return p;
}

/* ---- dsd_rtl_stream_metrics_hooks (context-free in the reference) ---- */

unsigned int
dsdneo_b200_stream_hook_output_rate_hz(void) {
    dsdneo_b200_stream_server* const cur = current_srv();
    return cur ? cur->output_rate_hz : 0u;
}

int
dsdneo_b200_stream_hook_output_kind(void) {
    dsdneo_b200_stream_server* const cur = current_srv();
    /* RTL_STREAM_OUTPUT_FSK_DISCRIMINATOR = 1, RTL_STREAM_OUTPUT_SYMBOL_CQPSK = 2 (include/dsd-neo/io/rtl_stream_c.h:31-33) */
    return (cur && cur->output_kind == 2) ? 2 : 1;
}

/* dsd_rtl_stream_metrics_hooks.cqpsk_status (rtl_stream_metrics_hooks.h:33): the sample side slices with cqpsk_slice() only
 * when both flags are set (src/core/frames/dsd_dibit.c:829-844) */
int
dsdneo_b200_stream_hook_cqpsk_status(int* out_cqpsk_enable, int* out_cqpsk_timing_active) {
    dsdneo_b200_stream_server* const cur = current_srv();
    const int on = (cur && cur->output_kind == 2 && cur->cqpsk_active) ? 1 : 0;
    if (out_cqpsk_enable) {
        *out_cqpsk_enable = on;
    }
    if (out_cqpsk_timing_active) {
        *out_cqpsk_timing_active = on;
    }
    return 0;
}

/* dsd_rtl_stream_metrics_hooks.snr_cqpsk_db: <= -50 means "no estimate" to the reliability weighting (dsd_dibit.c:404-427) */
double
dsdneo_b200_stream_hook_snr_cqpsk_db(void) {
    dsdneo_b200_stream_server* const cur = current_srv();
    return cur ? cur->snr_cqpsk_db : -100.0;
}

void
dsdneo_b200_stream_server_set_output_kind(dsdneo_b200_stream_server* s, int output_kind, int cqpsk_active, double snr_cqpsk_db) {
    if (!s) {
        return;
    }
    pthread_mutex_lock(&s->mu);
    s->output_kind = (output_kind == 2) ? 2 : 1;
    s->cqpsk_active = cqpsk_active ? 1 : 0;
    s->snr_cqpsk_db = snr_cqpsk_db;
    pthread_mutex_unlock(&s->mu);
}

int
dsdneo_b200_stream_hook_symbol_profile(int* out_symbol_rate_hz, int* out_levels, int* out_channel_profile) {
    dsdneo_b200_stream_server* const cur = current_srv();
    if (!cur) {
        return -1;
    }
    if (out_symbol_rate_hz) {
        *out_symbol_rate_hz = cur->symbol_rate_hz;
    }
    if (out_levels) {
        *out_levels = cur->levels;
    }
    if (out_channel_profile) {
        *out_channel_profile = cur->channel_profile;
    }
    return 0;
}

uint32_t
dsdneo_b200_stream_hook_stream_generation(void) {
    dsdneo_b200_stream_server* const cur = current_srv();
    return cur ? cur->generation : 0u;
}
