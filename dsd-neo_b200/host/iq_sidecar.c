// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * The reference's IQ capture sidecar ("format": "dsd-neo-iq", `--iq-capture` / `--iq-replay`, docs/iq-capture-replay.md:34-75),
 * as harness I/O for the block side: the metadata a many-channel ingest needs to feed a capture
 * to the channelizer / receive bank -- sample format and rate, tuned centres, the replay rate chain, the byte count -- read the
 * way dsd_iq_replay_read_metadata does (src/io/iq/iq_replay.c:1520-1790: flat JSON object, required fields, value checks
 * :675-720, the optional v2 "events" array) and the replayable byte count of dsd_iq_replay_compute_effective_bytes
 * (:1859-1882).  Plain host C, no device work.  Not restated: the event timeline's contents and its validation (:1066-1220);
 * events are counted and a capture that needs them (contains_retunes) is flagged for the caller.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/dsdneo_b200.h"

typedef struct {
    const char* p;
    const char* end;
} cur_t;

static void
ws(cur_t* c) {
    while (c->p < c->end && (*c->p == ' ' || *c->p == '\t' || *c->p == '\n' || *c->p == '\r')) {
        c->p++;
    }
}

/* JSON string at the cursor -> out (ASCII; the escapes the reference accepts); returns 0 on success */
static int
str_tok(cur_t* c, char* out, size_t cap) {
    size_t n = 0;
    if (c->p >= c->end || *c->p != '"') {
        return -1;
    }
    c->p++;
    while (c->p < c->end && *c->p != '"') {
        unsigned char ch = (unsigned char)*c->p++;
        if (ch == '\\') {
            if (c->p >= c->end) {
                return -1;
            }
            const char e = *c->p++;
            switch (e) {
                case '"': ch = '"'; break;
                case '\\': ch = '\\'; break;
                case '/': ch = '/'; break;
                case 'b': ch = '\b'; break;
                case 'f': ch = '\f'; break;
                case 'n': ch = '\n'; break;
                case 'r': ch = '\r'; break;
                case 't': ch = '\t'; break;
                case 'u': {
                    unsigned v = 0;
                    for (int i = 0; i < 4; i++) {
                        if (c->p >= c->end) {
                            return -1;
                        }
                        const char h = *c->p++;
                        v = v * 16 + (unsigned)(h >= '0' && h <= '9'   ? h - '0'
                                                : h >= 'a' && h <= 'f' ? h - 'a' + 10
                                                : h >= 'A' && h <= 'F' ? h - 'A' + 10
                                                                       : 99);
                        if (v > 0xFFFF) {
                            return -1;
                        }
                    }
                    if (v > 0x7f) {
                        return -1; /* iq_replay.c:317 */
                    }
                    ch = (unsigned char)v;
                    break;
                }
                default: return -1;
            }
        } else if (ch < 0x20) {
            return -1;
        }
        if (out && n + 1 < cap) {
            out[n++] = (char)ch;
        }
    }
    if (c->p >= c->end) {
        return -1;
    }
    c->p++;
    if (out && cap) {
        out[n] = 0;
    }
    return 0;
}

/* scalar value: string / number / true / false / null.  kind: 's', 'n' (integer), 'f' (non-integer number), 't', 'F', '0' */
static int
scalar_tok(cur_t* c, char* text, size_t cap, char* kind) {
    ws(c);
    if (c->p >= c->end) {
        return -1;
    }
    if (*c->p == '"') {
        *kind = 's';
        return str_tok(c, text, cap);
    }
    const char* s = c->p;
    while (c->p < c->end && *c->p != ',' && *c->p != '}' && *c->p != ']' && *c->p != ' ' && *c->p != '\n' && *c->p != '\r' && *c->p != '\t') {
        c->p++;
    }
    const size_t n = (size_t)(c->p - s);
    if (n == 0 || n + 1 > cap) {
        return -1;
    }
    memcpy(text, s, n);
    text[n] = 0;
    if (!strcmp(text, "true")) {
        *kind = 't';
    } else if (!strcmp(text, "false")) {
        *kind = 'F';
    } else if (!strcmp(text, "null")) {
        *kind = '0';
    } else {
        *kind = 'n';
        size_t i = text[0] == '-' ? 1 : 0;
        if (i >= n) {
            return -1;
        }
        for (; i < n; i++) {
            if (text[i] < '0' || text[i] > '9') {
                if (text[i] == '.' || text[i] == 'e' || text[i] == 'E' || text[i] == '+' || text[i] == '-') {
                    *kind = 'f';
                } else {
                    return -1;
                }
            }
        }
    }
    return 0;
}

static int
skip_events(cur_t* c, uint32_t* count) { /* '[' { flat objects } ']' */
    *count = 0;
    c->p++;
    ws(c);
    if (c->p < c->end && *c->p == ']') {
        c->p++;
        return 0;
    }
    for (;;) {
        ws(c);
        if (c->p >= c->end || *c->p != '{') {
            return -1;
        }
        c->p++;
        ws(c);
        if (c->p < c->end && *c->p == '}') {
            c->p++;
        } else {
            for (;;) {
                char text[128], kind;
                ws(c);
                if (str_tok(c, NULL, 0)) {
                    return -1;
                }
                ws(c);
                if (c->p >= c->end || *c->p++ != ':') {
                    return -1;
                }
                if (scalar_tok(c, text, sizeof(text), &kind)) {
                    return -1; /* nested event fields are unsupported (iq_replay.c:817) */
                }
                ws(c);
                if (c->p < c->end && *c->p == ',') {
                    c->p++;
                    continue;
                }
                if (c->p < c->end && *c->p == '}') {
                    c->p++;
                    break;
                }
                return -1;
            }
        }
        (*count)++;
        ws(c);
        if (c->p < c->end && *c->p == ',') {
            c->p++;
            continue;
        }
        if (c->p < c->end && *c->p == ']') {
            c->p++;
            return 0;
        }
        return -1;
    }
}

static int
to_u64(const char* t, char kind, uint64_t* out) {
    if (kind != 'n' || t[0] == '-') {
        return -1;
    }
    char* e = NULL;
    *out = strtoull(t, &e, 10);
    return (e && *e == 0) ? 0 : -1;
}

int
dsdneo_b200_iq_sidecar_parse(const char* json, size_t len, dsdneo_b200_iq_info* out) {
    if (!json || !out) {
        return DSDNEO_B200_EINVAL;
    }
    memset(out, 0, sizeof(*out));
    cur_t c = {json, json + len};
    char format[32] = "", sfmt[16] = "";
    enum { F_FORMAT = 1, F_VERSION = 2, F_SFMT = 4, F_RATE = 8, F_CENTER = 16, F_CAPCENTER = 32, F_BASEDEC = 64, F_POST = 128, F_DEMOD = 256,
           F_DATABYTES = 512 };
    unsigned seen = 0;
    int combine_seen = 0, combine = 1, events_seen = 0;
    char iq_order[8] = "", endianness[16] = "";
    /* metadata_require_required_fields (iq_replay.c:1683-1727): every one of these keys must be present */
    static const char* const k_required[] = {
        "format", "version", "sample_format", "iq_order", "endianness", "capture_stage", "sample_rate_hz", "center_frequency_hz",
        "capture_center_frequency_hz", "ppm", "tuner_gain_tenth_db", "rtl_dsp_bw_khz", "base_decimation", "post_downsample",
        "demod_rate_hz", "offset_tuning_enabled", "fs4_shift_enabled", "combine_rotate_enabled", "muted_bytes_excluded",
        "contains_retunes", "capture_retune_count", "source_backend", "source_args", "capture_started_utc", "data_file",
        "data_bytes", "capture_drops", "capture_drop_blocks", "input_ring_drops", "notes"};
    enum { N_REQUIRED = (int)(sizeof(k_required) / sizeof(k_required[0])) };
    unsigned char have[N_REQUIRED];
    memset(have, 0, sizeof(have));
    ws(&c);
    if (c.p >= c.end || *c.p++ != '{') {
        return DSDNEO_B200_EINVAL;
    }
    ws(&c);
    if (c.p < c.end && *c.p == '}') {
        return DSDNEO_B200_EINVAL;
    }
    for (;;) {
        char key[64], text[2048], kind = 0;
        ws(&c);
        if (str_tok(&c, key, sizeof(key))) {
            return DSDNEO_B200_EINVAL;
        }
        ws(&c);
        if (c.p >= c.end || *c.p++ != ':') {
            return DSDNEO_B200_EINVAL;
        }
        ws(&c);
        for (int r = 0; r < N_REQUIRED; r++) {
            if (!strcmp(key, k_required[r])) {
                have[r] = 1;
            }
        }
        if (c.p < c.end && *c.p == '[') {
            if (strcmp(key, "events") || skip_events(&c, &out->event_count)) {
                return DSDNEO_B200_EINVAL; /* nested structures are unsupported in metadata (iq_replay.c:1629) */
            }
            events_seen = 1;
        } else if (c.p < c.end && *c.p == '{') {
            return DSDNEO_B200_EINVAL;
        } else {
            if (scalar_tok(&c, text, sizeof(text), &kind)) {
                return DSDNEO_B200_EINVAL;
            }
            uint64_t v = 0;
#define U64_FIELD(name, dst, flag)                                                                                     \
    if (!strcmp(key, name)) {                                                                                          \
        if (to_u64(text, kind, &v)) {                                                                                  \
            return DSDNEO_B200_EINVAL;                                                                                 \
        }                                                                                                              \
        dst = v;                                                                                                       \
        seen |= (flag);                                                                                                \
    }
#define U32_FIELD(name, dst, flag)                                                                                     \
    if (!strcmp(key, name)) {                                                                                          \
        if (to_u64(text, kind, &v) || v > 0xFFFFFFFFull) {                                                             \
            return DSDNEO_B200_EINVAL;                                                                                 \
        }                                                                                                              \
        dst = (uint32_t)v;                                                                                             \
        seen |= (flag);                                                                                                \
    }
#define BOOL_FIELD(name, dst)                                                                                          \
    if (!strcmp(key, name)) {                                                                                          \
        if (kind != 't' && kind != 'F') {                                                                              \
            return DSDNEO_B200_EINVAL;                                                                                 \
        }                                                                                                              \
        dst = kind == 't';                                                                                             \
    }
            if (!strcmp(key, "format")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
                strncpy(format, text, sizeof(format) - 1);
                seen |= F_FORMAT;
            } else if (!strcmp(key, "sample_format")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
                strncpy(sfmt, text, sizeof(sfmt) - 1);
                seen |= F_SFMT;
            } else if (!strcmp(key, "data_file")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
                strncpy(out->data_file, text, sizeof(out->data_file) - 1);
            } else if (!strcmp(key, "capture_stage")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
                strncpy(out->capture_stage, text, sizeof(out->capture_stage) - 1);
            } else if (!strcmp(key, "iq_order")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
                strncpy(iq_order, text, sizeof(iq_order) - 1);
            } else if (!strcmp(key, "endianness")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
                strncpy(endianness, text, sizeof(endianness) - 1);
            } else if (!strcmp(key, "source_backend") || !strcmp(key, "source_args") || !strcmp(key, "capture_started_utc") ||
                       !strcmp(key, "notes")) {
                if (kind != 's') {
                    return DSDNEO_B200_EINVAL;
                }
            } else if (!strcmp(key, "ppm") || !strcmp(key, "tuner_gain_tenth_db") || !strcmp(key, "rtl_dsp_bw_khz")) {
                if (kind != 'n') {
                    return DSDNEO_B200_EINVAL; /* signed 32-bit integers (token_to_i32) */
                }
            } else if (!strcmp(key, "capture_drops") || !strcmp(key, "capture_drop_blocks") || !strcmp(key, "input_ring_drops")) {
                uint64_t t64;
                if (to_u64(text, kind, &t64)) {
                    return DSDNEO_B200_EINVAL;
                }
            }
            U32_FIELD("version", out->version, F_VERSION)
            U32_FIELD("sample_rate_hz", out->sample_rate_hz, F_RATE)
            U64_FIELD("center_frequency_hz", out->center_frequency_hz, F_CENTER)
            U64_FIELD("capture_center_frequency_hz", out->capture_center_frequency_hz, F_CAPCENTER)
            U32_FIELD("base_decimation", out->base_decimation, F_BASEDEC)
            U32_FIELD("post_downsample", out->post_downsample, F_POST)
            U32_FIELD("demod_rate_hz", out->demod_rate_hz, F_DEMOD)
            U64_FIELD("data_bytes", out->data_bytes, F_DATABYTES)
            U32_FIELD("capture_retune_count", out->capture_retune_count, 0)
            BOOL_FIELD("offset_tuning_enabled", out->offset_tuning_enabled)
            BOOL_FIELD("fs4_shift_enabled", out->fs4_shift_enabled)
            BOOL_FIELD("muted_bytes_excluded", out->muted_bytes_excluded)
            BOOL_FIELD("contains_retunes", out->contains_retunes)
            BOOL_FIELD("size_limit_reached", out->size_limit_reached)
            if (!strcmp(key, "combine_rotate_enabled")) {
                if (kind != 't' && kind != 'F') {
                    return DSDNEO_B200_EINVAL;
                }
                combine = kind == 't';
                combine_seen = 1;
            }
#undef U64_FIELD
#undef U32_FIELD
#undef BOOL_FIELD
        }
        ws(&c);
        if (c.p < c.end && *c.p == ',') {
            c.p++;
            continue;
        }
        if (c.p < c.end && *c.p == '}') {
            c.p++;
            break;
        }
        return DSDNEO_B200_EINVAL;
    }
    ws(&c);
    if (c.p != c.end) {
        return DSDNEO_B200_EINVAL; /* trailing data */
    }
    (void)seen;
    for (int r = 0; r < N_REQUIRED; r++) {
        if (!have[r]) {
            return DSDNEO_B200_EINVAL;
        }
    }
    /* metadata_finalize (iq_replay.c:1757-1800) */
    if (strcmp(format, "dsd-neo-iq") || (out->version != 1 && out->version != 2) || (events_seen && out->version != 2) ||
        strcmp(iq_order, "IQ")) {
        return DSDNEO_B200_EINVAL;
    }
    if (!strcmp(sfmt, "cu8")) {
        out->sample_format = DSDNEO_B200_IQ_CU8;
        if (strcmp(endianness, "none")) {
            return DSDNEO_B200_EINVAL;
        }
    } else if (!strcmp(sfmt, "cf32")) {
        out->sample_format = DSDNEO_B200_IQ_CF32;
        if (strcmp(endianness, "little")) {
            return DSDNEO_B200_EINVAL;
        }
    } else if (!strcmp(sfmt, "cs16")) {
        return DSDNEO_B200_EUNSUPPORTED; /* a format of the reference no kernel here takes */
    } else {
        return DSDNEO_B200_EINVAL;
    }
    /* validate_replay_semantics (iq_replay.c:675-720): rate chain and capture stage */
    if (out->sample_rate_hz == 0 || out->post_downsample == 0 || out->demod_rate_hz == 0 || out->base_decimation == 0 ||
        (out->base_decimation & (out->base_decimation - 1)) != 0 || out->base_decimation > 1024u ||
        (uint64_t)out->sample_rate_hz / out->base_decimation / out->post_downsample != out->demod_rate_hz ||
        (strcmp(out->capture_stage, "post_mute_pre_widen") && strcmp(out->capture_stage, "post_driver_cf32_pre_ring"))) {
        return DSDNEO_B200_EINVAL;
    }
    out->historical_cu8_two_pass = combine_seen && !combine;
    return 0;
}

long long
dsdneo_b200_iq_effective_bytes(const dsdneo_b200_iq_info* info, uint64_t actual_file_size, int* size_mismatch) {
    if (!info || (info->sample_format != DSDNEO_B200_IQ_CU8 && info->sample_format != DSDNEO_B200_IQ_CF32)) {
        return DSDNEO_B200_EINVAL;
    }
    const uint64_t align = info->sample_format == DSDNEO_B200_IQ_CU8 ? 2 : 8;
    uint64_t raw = actual_file_size;
    int mismatch = 0;
    if (info->data_bytes > 0) {
        raw = info->data_bytes < actual_file_size ? info->data_bytes : actual_file_size;
        mismatch = info->data_bytes != actual_file_size;
    }
    if (size_mismatch) {
        *size_mismatch = mismatch;
    }
    return (long long)(raw - raw % align);
}
