/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY.  The UNMODIFIED reference IQ-replay metadata reader (src/io/iq/iq_replay.c:
 * dsd_iq_replay_read_metadata, dsd_iq_replay_compute_effective_bytes) behind a flat record, to pin
 * dsdneo_b200_iq_sidecar_parse / dsdneo_b200_iq_effective_bytes.  The four helpers iq_replay.c takes from other translation
 * units (sample-format table of iq_capture.c, the POSIX file wrappers of src/platform) are supplied here.
 * Built into oracle/_ref/libdsdneo_ref_iq.so by oracle/Makefile.  No reference source is copied.
 */
#include <dsd-neo/io/iq_replay.h>
#include <dsd-neo/io/iq_types.h>
#include <dsd-neo/platform/file_compat.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

size_t
dsd_iq_sample_format_alignment_bytes(dsd_iq_sample_format format) {
    switch (format) {
        case DSD_IQ_FORMAT_CU8: return 2;
        case DSD_IQ_FORMAT_CF32: return 8;
        case DSD_IQ_FORMAT_CS16: return 4;
        default: return 0;
    }
}

const char*
dsd_iq_sample_format_name(dsd_iq_sample_format format) {
    switch (format) {
        case DSD_IQ_FORMAT_CU8: return "cu8";
        case DSD_IQ_FORMAT_CF32: return "cf32";
        case DSD_IQ_FORMAT_CS16: return "cs16";
        default: return "unknown";
    }
}

int
dsd_stat_path(const char* path, dsd_stat_t* st) {
    return stat(path, (struct stat*)st);
}

FILE*
dsd_fopen_existing_regular_file(const char* path, const char* mode) {
    struct stat st;
    if (stat(path, &st) != 0 || !S_ISREG(st.st_mode)) {
        return NULL;
    }
    return fopen(path, mode);
}

typedef struct ref_iq_info { /* same layout as dsdneo_b200_iq_info */
    uint32_t version;
    int32_t sample_format;
    uint32_t sample_rate_hz;
    uint64_t center_frequency_hz, capture_center_frequency_hz, data_bytes;
    uint32_t base_decimation, post_downsample, demod_rate_hz;
    int32_t offset_tuning_enabled, fs4_shift_enabled, historical_cu8_two_pass;
    int32_t muted_bytes_excluded, contains_retunes, size_limit_reached;
    uint32_t capture_retune_count, event_count;
    char data_file[256];
    char capture_stage[64];
} ref_iq_info;

int
ref_iq_read_metadata(const char* path, ref_iq_info* out, char* err, size_t err_size) {
    dsd_iq_replay_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    memset(out, 0, sizeof(*out));
    const int rc = dsd_iq_replay_read_metadata(path, &cfg, err, err_size);
    if (rc != DSD_IQ_OK) {
        return rc;
    }
    out->version = cfg.metadata_version;
    out->sample_format = (int32_t)cfg.format;
    out->sample_rate_hz = cfg.sample_rate_hz;
    out->center_frequency_hz = cfg.center_frequency_hz;
    out->capture_center_frequency_hz = cfg.capture_center_frequency_hz;
    out->data_bytes = cfg.data_bytes;
    out->base_decimation = cfg.base_decimation;
    out->post_downsample = cfg.post_downsample;
    out->demod_rate_hz = cfg.demod_rate_hz;
    out->offset_tuning_enabled = cfg.offset_tuning_enabled;
    out->fs4_shift_enabled = cfg.fs4_shift_enabled;
    out->historical_cu8_two_pass = cfg.historical_cu8_two_pass;
    out->muted_bytes_excluded = cfg.muted_bytes_excluded;
    out->contains_retunes = cfg.contains_retunes;
    out->size_limit_reached = cfg.size_limit_reached;
    out->capture_retune_count = cfg.capture_retune_count;
    out->event_count = cfg.event_count;
    strncpy(out->capture_stage, cfg.capture_stage, sizeof(out->capture_stage) - 1);
    dsd_iq_replay_config_clear(&cfg);
    return 0;
}

long long
ref_iq_effective_bytes(uint64_t data_bytes, uint64_t file_size, int format, int* mismatch) {
    uint64_t eff = 0;
    if (dsd_iq_replay_compute_effective_bytes(data_bytes, file_size, (dsd_iq_sample_format)format, &eff, mismatch) != DSD_IQ_OK) {
        return -1;
    }
    return (long long)eff;
}
