// SPDX-License-Identifier: GPL-3.0-or-later
/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 * extern "C" access to the reference's header-only C++ FEC templates so tests can call the
 * raw 63-symbol decoder.  The P25 wrappers are reached through the reference's own C symbols.
 */
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

} /* extern "C" */
