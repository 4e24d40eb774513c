// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * extern "C" access to the reference's header-only C++ FEC templates so tests can call the
 * raw 63-symbol decoder.  The P25 wrappers are reached through the reference's own C symbols.
 */
#include <dsd-neo/fec/BCH_63_16.hpp>
#include <dsd-neo/fec/ReedSolomon.hpp>
#include <dsd-neo/protocol/p25/p25p1_check_nid.h>
#include <stdint.h>
#include <string.h>
#include <time.h>
#include <dsd-neo/fec/block_codes.h>
#include <dsd-neo/fec/bptc.h>
#include <dsd-neo/protocol/p25/p25_12.h>
#include <dsd-neo/protocol/p25/p25p1_check_hdu.h>

extern "C" {

/* ReedSolomon_63<TT>::decode (include/dsd-neo/fec/ReedSolomon.hpp:738-771) on raw 6-bit symbols. */
int
ref_rs63_decode(int tt, const int* in63, int* out63) {
    switch (tt) {
        case 8: { ReedSolomon_63<8> rs; return rs.decode(in63, out63); }
        case 6: { ReedSolomon_63<6> rs; return rs.decode(in63, out63); }
        case 4: { ReedSolomon_63<4> rs; return rs.decode(in63, out63); }
        default: return -1;
    }
}

/* BCH_63_16_11::decode_with_result (include/dsd-neo/fec/BCH_63_16.hpp:288-329), the P25 NID code. */
int
ref_bch_63_16_decode(const char* in63, char* out16, int* error_count) {
    static const BCH_63_16_11 bch;
    const BCH_63_16_Result r = bch.decode_with_result(in63, out16);
    if (error_count) {
        *error_count = r.error_count;
    }
    return r.success ? 1 : 0;
}

/* p25p1_nid_decode (include/dsd-neo/protocol/p25/p25p1_check_nid.h:76), flat arguments for ctypes */
int
ref_p25p1_nid_decode(const char* code63, const uint8_t* reliab63, int observed_nac, int parity, int parity_reliab, int* nac,
                     int* duid, int* errs) {
    const struct p25p1_nid_result r = p25p1_nid_decode(code63, reliab63, observed_nac, (unsigned char)parity, (uint8_t)parity_reliab);
    *nac = r.nac;
    *duid = r.duid;
    *errs = r.error_count;
    return (int)r.status;
}

/* Native timing loops over the reference's own FEC entry points for bench.py's cpu_baseline leg (a Python loop would
 * measure ctypes call overhead, not the decoders).  `items` holds n_items inputs back to back in the reference's own
 * byte-per-bit / dibit + LLR layouts; every call works on a scratch copy (the decoders correct in place).  Returns seconds
 * for `repeats` passes over all items; *checksum keeps the calls observable.
 *   kind 0  Golay_24_12_decode                      item = 24 bytes
 *   kind 1  BPTCDeInterleaveDMRData + BPTC_196x96_Extract_Data   item = 196 bytes
 *   kind 2  p25_12_soft_llr                         item = 98 dibit bytes + 196 int16 LLRs (490 bytes)
 *   kind 3  check_and_fix_redsolomon_36_20_17       item = 120 data bits + 96 parity bits (216 bytes)
 *   kind 4  p25p1_nid_decode                        item = 63 bits + 63 reliabilities + parity + parity reliability (128 bytes) */
double
ref_fec_loop(int kind, const uint8_t* items, int n_items, int repeats, long* checksum) {
    struct timespec t0, t1;
    long acc = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < repeats; r++) {
        for (int i = 0; i < n_items; i++) {
            switch (kind) {
                case 0: {
                    unsigned char w[24];
                    memcpy(w, items + (size_t)i * 24, 24);
                    acc += Golay_24_12_decode(w) ? 1 : 0;
                    acc += w[3];
                } break;
                case 1: {
                    uint8_t de[196], out[96], R[3];
                    BPTCDeInterleaveDMRData(items + (size_t)i * 196, de);
                    acc += (long)BPTC_196x96_Extract_Data(de, out, R);
                    acc += out[5];
                } break;
                case 2: {
                    const uint8_t* it = items + (size_t)i * 490;
                    int16_t llr[196];
                    uint8_t out12[12];
                    memcpy(llr, it + 98, sizeof(llr));
                    acc += p25_12_soft_llr(it, llr, out12);
                    acc += out12[3];
                } break;
                case 3: {
                    char d[120], p[96];
                    memcpy(d, items + (size_t)i * 216, 120);
                    memcpy(p, items + (size_t)i * 216 + 120, 96);
                    acc += check_and_fix_redsolomon_36_20_17(d, p);
                    acc += d[7];
                } break;
                case 4: {
                    const uint8_t* it = items + (size_t)i * 128;
                    const struct p25p1_nid_result nr = p25p1_nid_decode((const char*)it, it + 63, 0, it[126], it[127]);
                    acc += nr.nac + (int)nr.status;
                } break;
                default: return -1.0;
            }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) {
        *checksum = acc;
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

} /* extern "C" */
