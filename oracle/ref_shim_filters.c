/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Opens the UNMODIFIED sample-side filter unit of the reference (src/dsp/dsd_filters.c) for the design test: the unit is
 * compiled in place by including it here (-I$(REF); nothing of it is copied into this repository), which makes its static
 * design_sps_fir() and its static coefficient tables reachable.
 *   ref_filt_descriptor: the sps_fir descriptor of one filter (table, base sps, design kind, roll-off), dsd_filters.c:325-345
 *   ref_filt_design:     design_sps_fir() itself on a fresh descriptor (:153-170): tap count and normalised taps
 * which: 0 p25, 1 dmr, 2 nxdn, 3 dpmr, 4 m17.
 */
#include "src/dsp/dsd_filters.c"

static sps_fir*
filt_of(int which) {
    switch (which) {
        case 0: return &g_fir_p25;
        case 1: return &g_fir_dmr;
        case 2: return &g_fir_nxdn;
        case 3: return &g_fir_dpmr;
        case 4: return &g_fir_m17;
        default: return 0;
    }
}

int
ref_filt_descriptor(int which, float* base_out, int cap, int* base_sps, int* design_kind, float* rrc_alpha) {
    const sps_fir* f = filt_of(which);
    if (!f || f->base_len > cap) {
        return -1;
    }
    for (int i = 0; i < f->base_len; i++) {
        base_out[i] = f->base[i];
    }
    *base_sps = f->base_sps;
    *design_kind = f->design_kind;
    *rrc_alpha = f->rrc_alpha;
    return f->base_len;
}

int
ref_filt_design(int which, int sps, float* taps_out, int cap) {
    sps_fir* f = filt_of(which);
    if (!f) {
        return -1;
    }
    reset_sps_fir(f);
    design_sps_fir(f, sps);
    if (!f->ready) {
        return 0;
    }
    if (f->taps_len > cap) {
        return -f->taps_len;
    }
    for (int i = 0; i < f->taps_len; i++) {
        taps_out[i] = f->taps[i];
    }
    return f->taps_len;
}
