/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Reference-signature FEC entry points (include/dsd-neo/fec/block_codes.h:29-56, include/dsd-neo/fec/bptc.h,
 * include/dsd-neo/protocol/p25/p25_12.h, p25p1_check_hdu.h / p25p1_check_ldu.h) implemented as batch-of-1 calls into
 * libdsdneo_b200.so.  Link this INSTEAD of src/fec/fec.c, src/fec/bptc.c (196x96 part), src/protocol/p25/p25_12.c and the
 * RS wrappers when replacing the reference's CPU decoders one-for-one; throughput comes from the `*_batch` entry points.
 * Semantics (in-place correction, return values) are the reference's; see include/dsdneo_b200.h.
 */
#include <stdbool.h>
#include <stdint.h>
#include <string.h>

#include "dsdneo_b200.h"

void
InitAllFecFunction(void) { /* tables are built on first use inside the library */
}

static bool
block1(int code, unsigned char* rx, unsigned char* decoded) {
    uint8_t ok = 0;
    if (dsdneo_b200_fec_block_decode_batch_host(code, rx, decoded, &ok, 1) != 0) {
        return false;
    }
    return ok != 0;
}

bool Hamming_7_4_decode(unsigned char* rxBits) { return block1(DSDNEO_FEC_HAMMING_7_4, rxBits, NULL); }
bool Hamming_12_8_decode(unsigned char* rx, unsigned char* dec, int n) { (void)n; return block1(DSDNEO_FEC_HAMMING_12_8, rx, dec); }
bool Hamming_13_9_decode(unsigned char* rx, unsigned char* dec, int n) { (void)n; return block1(DSDNEO_FEC_HAMMING_13_9, rx, dec); }
bool Hamming_15_11_decode(unsigned char* rx, unsigned char* dec, int n) { (void)n; return block1(DSDNEO_FEC_HAMMING_15_11, rx, dec); }
bool Hamming_16_11_4_decode(unsigned char* rx, unsigned char* dec, int n) { (void)n; return block1(DSDNEO_FEC_HAMMING_16_11_4, rx, dec); }
bool Golay_20_8_decode(unsigned char* rxBits) { return block1(DSDNEO_FEC_GOLAY_20_8, rxBits, NULL); }
bool Golay_24_12_decode(unsigned char* rxBits) { return block1(DSDNEO_FEC_GOLAY_24_12, rxBits, NULL); }
bool QR_16_7_6_decode(unsigned char* rxBits) { return block1(DSDNEO_FEC_QR_16_7_6, rxBits, NULL); }

uint32_t
BPTC_196x96_Extract_Data(uint8_t InputDeInteleavedData[196], uint8_t DMRDataExtracted[96], uint8_t R[3]) {
    uint32_t errs = 0;
    if (dsdneo_b200_bptc_196x96_batch_host(InputDeInteleavedData, 0, DMRDataExtracted, R, &errs, 1) != 0) {
        return 24; /* every line irrecoverable */
    }
    return errs;
}

int
p25_12_soft_llr(const uint8_t* input, const int16_t* bit_llr196, uint8_t treturn[12]) {
    (void)input;
    int32_t metric = 0;
    if (dsdneo_b200_p25_12_soft_llr_batch_host(bit_llr196, treturn, &metric, 1) != 0) {
        return 0x7fffffff;
    }
    return (int)metric;
}

int
check_and_fix_redsolomon_36_20_17(char* data, const char* parity) {
    uint8_t st = 1;
    if (dsdneo_b200_p25_rs_decode_batch_host(DSDNEO_P25_RS_36_20_17, (uint8_t*)data, (const uint8_t*)parity, &st, 1) != 0) {
        return 1;
    }
    return st;
}

int
check_and_fix_reedsolomon_24_12_13(char* data, const char* parity) {
    uint8_t st = 1;
    if (dsdneo_b200_p25_rs_decode_batch_host(DSDNEO_P25_RS_24_12_13, (uint8_t*)data, (const uint8_t*)parity, &st, 1) != 0) {
        return 1;
    }
    return st;
}

int
check_and_fix_reedsolomon_24_16_9(char* data, const char* parity) {
    uint8_t st = 1;
    if (dsdneo_b200_p25_rs_decode_batch_host(DSDNEO_P25_RS_24_16_9, (uint8_t*)data, (const uint8_t*)parity, &st, 1) != 0) {
        return 1;
    }
    return st;
}
