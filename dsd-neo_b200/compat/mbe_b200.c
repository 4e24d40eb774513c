/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * Reference-side glue (compiled inside the dsd-neo tree, NOT part of libdsdneo_b200.so): the two frame-ECC entry points of
 * the mbelib-neo API that dsd-neo calls for every vocoder frame (src/core/vocoder/dsd_mbe.c:168 and :188, API contract
 * CMakeLists.txt:622-655), routed to the batched device kernels with a batch of one.
 *
 *     int mbe_decodeImbe7200x4400Frame(const char imbe_fr[8][23], char imbe_d[88], mbe_process_result* result);
 *     int mbe_decodeAmbe3600x2450Frame(const char ambe_fr[4][24], char ambe_d[49], mbe_process_result* result);
 *
 * PARITY UNPINNED: mbelib-neo's source is not vendored in dsd-neo; the kernels follow the published mbelib 1.3.0 /
 * TIA-102.BABA algorithm.  `mbe_process_result` is mbelib-neo's type: this file needs its header and
 * fills the two counters dsd-neo reads back (store_mbe_result, dsd_mbe.c:114-118: c0_errors -> state->errs, total_errors ->
 * state->errs2).  Build with -DDSDNEO_B200_MBE_SHIM and link the wrappers in place of the library's symbols, e.g.
 *     target_link_options(dsd-neo PRIVATE -Wl,--wrap=mbe_decodeImbe7200x4400Frame -Wl,--wrap=mbe_decodeAmbe3600x2450Frame)
 * A production integration would not call a GPU per frame: the receive bank already returns every LDU's nine IMBE frames, and
 * dsdneo_b200_p25p1_voice_imbe_decode_batch / dsdneo_b200_ambe3600x2450_decode_batch decode them for all channels at once
 * (INTEGRATION.md section 7); this shim is the drop-in form of the same kernels.
 */
#ifdef DSDNEO_B200_MBE_SHIM
#include <mbelib-neo/mbelib.h>
#include <stdint.h>
#include <string.h>

#include "dsdneo_b200.h"

int
__wrap_mbe_decodeImbe7200x4400Frame(const char imbe_fr[8][23], char imbe_d[88], mbe_process_result* result) {
    int32_t c0 = 0, total = 0;
    uint8_t out[88];
    if (!imbe_fr || !imbe_d) {
        return -1;
    }
    if (dsdneo_b200_imbe7200x4400_decode_batch_host((const uint8_t*)imbe_fr, out, &c0, &total, 1) != 0) {
        return -1; /* no silent CPU fallback: the caller's error path (mbe_synthesizeSilencef, dsd_mbe.c:137-150) takes over */
    }
    memcpy(imbe_d, out, sizeof(out));
    if (result) {
        result->c0_errors = c0;
        result->total_errors = total;
    }
    return 0;
}

int
__wrap_mbe_decodeAmbe3600x2450Frame(const char ambe_fr[4][24], char ambe_d[49], mbe_process_result* result) {
    int32_t c0 = 0, total = 0;
    uint8_t out[49];
    if (!ambe_fr || !ambe_d) {
        return -1;
    }
    if (dsdneo_b200_ambe3600x2450_decode_batch_host((const uint8_t*)ambe_fr, out, &c0, &total, 1) != 0) {
        return -1;
    }
    memcpy(ambe_d, out, sizeof(out));
    if (result) {
        result->c0_errors = c0;
        result->total_errors = total;
    }
    return 0;
}
#endif /* DSDNEO_B200_MBE_SHIM */
