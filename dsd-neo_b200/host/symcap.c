// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * The reference's soft symbol-capture format ("DSDNSYM2", `dsd-neo -c file.bin` / `-i file.bin`), as harness I/O for the
 * batched sample side: the symbolizer's per-channel outputs (dibit, reliability, LLR pair,
 * float symbol) packed into the 10-byte records the unmodified CLI replays.
 *   header  16 B  "DSDNSYM2", 2, 10, 0 x6                 openSymbolOutFile, src/core/file/dsd_file.c:876-888
 *   record  10 B  dibit & 3 | reliability | llr[0] LE i16 | llr[1] LE i16 | symbol f32 bits LE
 *                                                          write_symbol_capture_record, src/core/frames/dsd_dibit.c:794-818
 *   reader                                                 read_soft_symbol_record, src/dsp/dsd_symbol.c:120-173
 * Plain host C: byte shuffling, no device work.
 */
#include <stdio.h>
#include <string.h>

#include "../../include/dsdneo_b200.h"

static const unsigned char k_header[DSDNEO_B200_SYMCAP_HEADER_SIZE] = {'D', 'S', 'D', 'N', 'S', 'Y', 'M', '2', 2,
                                                                       DSDNEO_B200_SYMCAP_RECORD_SIZE, 0, 0, 0, 0, 0, 0};

size_t
dsdneo_b200_symbol_capture_size(size_t n_records, int with_header) {
    return (with_header ? DSDNEO_B200_SYMCAP_HEADER_SIZE : 0) + n_records * DSDNEO_B200_SYMCAP_RECORD_SIZE;
}

int
dsdneo_b200_symbol_capture_pack(const uint8_t* dibits, const uint8_t* reliability, const int16_t* llr, const float* symbols,
                                size_t n_records, int with_header, uint8_t* out) {
    if (!out || (n_records && (!dibits || !reliability || !llr || !symbols))) {
        return DSDNEO_B200_EINVAL;
    }
    if (with_header) {
        memcpy(out, k_header, sizeof(k_header));
        out += sizeof(k_header);
    }
    for (size_t i = 0; i < n_records; i++, out += DSDNEO_B200_SYMCAP_RECORD_SIZE) {
        const uint16_t l0 = (uint16_t)llr[2 * i], l1 = (uint16_t)llr[2 * i + 1];
        uint32_t raw;
        memcpy(&raw, &symbols[i], sizeof(raw));
        out[0] = (uint8_t)(dibits[i] & 3u);
        out[1] = reliability[i];
        out[2] = (uint8_t)(l0 & 0xFFu);
        out[3] = (uint8_t)(l0 >> 8);
        out[4] = (uint8_t)(l1 & 0xFFu);
        out[5] = (uint8_t)(l1 >> 8);
        out[6] = (uint8_t)(raw & 0xFFu);
        out[7] = (uint8_t)((raw >> 8) & 0xFFu);
        out[8] = (uint8_t)((raw >> 16) & 0xFFu);
        out[9] = (uint8_t)(raw >> 24);
    }
    return 0;
}

long long
dsdneo_b200_symbol_capture_unpack(const uint8_t* in, size_t len, uint8_t* dibits, uint8_t* reliability, int16_t* llr,
                                  float* symbols, size_t max_records) {
    if (!in) {
        return DSDNEO_B200_EINVAL;
    }
    if (len >= 8 && memcmp(in, k_header, 8) == 0) {
        /* probe_symbol_replay_format: a soft file must carry version 2 and the 10-byte record size */
        if (len < sizeof(k_header) || in[8] != 2 || in[9] != DSDNEO_B200_SYMCAP_RECORD_SIZE) {
            return DSDNEO_B200_EINVAL;
        }
        in += sizeof(k_header);
        len -= sizeof(k_header);
    }
    size_t n = len / DSDNEO_B200_SYMCAP_RECORD_SIZE;
    if (n > max_records) {
        n = max_records;
    }
    for (size_t i = 0; i < n; i++, in += DSDNEO_B200_SYMCAP_RECORD_SIZE) {
        if (dibits) {
            dibits[i] = in[0] & 3u;
        }
        if (reliability) {
            reliability[i] = in[1];
        }
        if (llr) {
            llr[2 * i] = (int16_t)((uint16_t)in[2] | ((uint16_t)in[3] << 8));
            llr[2 * i + 1] = (int16_t)((uint16_t)in[4] | ((uint16_t)in[5] << 8));
        }
        if (symbols) {
            const uint32_t raw = (uint32_t)in[6] | ((uint32_t)in[7] << 8) | ((uint32_t)in[8] << 16) | ((uint32_t)in[9] << 24);
            memcpy(&symbols[i], &raw, sizeof(raw));
        }
    }
    return (long long)n;
}

int
dsdneo_b200_symbol_capture_write_file(const char* path, int append, const uint8_t* dibits, const uint8_t* reliability,
                                      const int16_t* llr, const float* symbols, size_t n_records) {
    if (!path) {
        return DSDNEO_B200_EINVAL;
    }
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        return DSDNEO_B200_EINVAL;
    }
    int rc = 0;
    if (!append && fwrite(k_header, 1, sizeof(k_header), f) != sizeof(k_header)) {
        rc = DSDNEO_B200_EINVAL;
    }
    uint8_t rec[64 * DSDNEO_B200_SYMCAP_RECORD_SIZE];
    for (size_t i = 0; rc == 0 && i < n_records; i += 64) {
        const size_t m = (n_records - i < 64) ? n_records - i : 64;
        rc = dsdneo_b200_symbol_capture_pack(dibits + i, reliability + i, llr + 2 * i, symbols + i, m, 0, rec);
        if (rc == 0 && fwrite(rec, DSDNEO_B200_SYMCAP_RECORD_SIZE, m, f) != m) {
            rc = DSDNEO_B200_EINVAL;
        }
    }
    fclose(f);
    return rc;
}
