/* SPDX-License-Identifier: GPL-3.0-or-later */
/* Decode tables shared between host/fec_tables.c (builder) and csrc/fec.cu (device copy). */
#ifndef DSDNEO_FEC_TABLES_H_
#define DSDNEO_FEC_TABLES_H_

#include "../../include/dsdneo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsdneo_hamming_table {
    int n, r, k, stop_on_fail;
    unsigned row_mask[5];       /* syndrome bit `row` (row 0 = MSB) = parity(word & row_mask[row]); word bit j = position j */
    unsigned char pos_of[32];   /* syndrome -> position to flip, 0xFF = uncorrectable */
} dsdneo_hamming_table;

typedef struct dsdneo_gq_table {
    int n, k, r, maxw;
    unsigned row_mask[12];
    unsigned char corr[4096][3];
} dsdneo_gq_table;

typedef struct dsdneo_fec_tables {
    dsdneo_hamming_table ham[5];
    dsdneo_gq_table gq[3];      /* 0 = Golay(20,8), 1 = Golay(24,12), 2 = QR(16,7,6) */
    signed char gf_exp[64], gf_log[64];
} dsdneo_fec_tables;

void dsdneo_fec_build_tables(dsdneo_fec_tables* t);

#ifdef __cplusplus
}
#endif

#endif
