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

} /* extern "C" */
