// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * extern "C" access to the reference's header-only C++ FEC templates so tests can call the
 * raw 63-symbol decoder.  The P25 wrappers are reached through the reference's own C symbols.
 */
#include <dsd-neo/fec/BCH_63_16.hpp>
#include <dsd-neo/fec/ReedSolomon.hpp>

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

} /* extern "C" */
