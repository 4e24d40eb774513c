/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the sample side of the hot path for the configuration the
 * batched path covers: RTL FSK-discriminator input (output kind 1), 4-level C4FM family (rf_mod == 0), no SNR
 * hooks installed (so the SNR weight is the reference's sentinel path, src/core/frames/dsd_dibit.c:504-546).
 *
 *   getSymbol            src/dsp/dsd_symbol.c:1853-1880 with :197-225 (window), :301-337 (matched filter choice),
 *                        :347-358 (sync clip), :360-397 (jitter), :423-460 (accumulate), :489-516 (timing nudge),
 *                        :1306-1387 (slicer reset, fractional samples-per-symbol)
 *   apply_sps_fir        src/dsp/dsd_filters.c:172-201
 *   use_symbol           src/core/frames/dsd_dibit.c:243-299 (+ :195-241 extrema, core/state.h:1388-1454 window sums)
 *   digitize + soft      src/core/frames/dsd_dibit.c:455-721, 963-1041
 *   getDibitSoft         src/core/frames/dsd_dibit.c:1043-1089
 *
 * Pinned against the compiled reference by tests/test_oracle_symbol.py (bit-exact symbols, dibits, LLRs).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

void
oracle_sym_init(oracle_sym_chan* c, int output_rate_hz, int symbol_rate_hz, int use_filter, int window_l, int track_minmax,
                int negative, const float* taps, int taps_len, int ssize, int msize) {
    memset(c, 0, sizeof(*c));
    c->output_rate_hz = output_rate_hz;
    c->symbol_rate_hz = symbol_rate_hz;
    c->use_filter = use_filter && taps && taps_len > 0;
    c->window_l = window_l;
    c->track_minmax = track_minmax;
    c->negative = negative;
    c->ssize = ssize;
    c->msize = msize;
    c->taps_len = c->use_filter ? taps_len : 0;
    if (c->use_filter) {
        memcpy(c->taps, taps, (size_t)taps_len * sizeof(float));
    }
    c->fir_head = -1;
    /* initState() values the path reads before the first-call reset (src/core/util/dsd_init.c:519-592) */
    c->jitter = -1;
    c->min = -15000.0f;
    c->max = 15000.0f;
    c->minref = -12000.0f;
    c->maxref = 12000.0f;
    for (int i = 0; i < 1024; i++) {
        c->minbuf[i] = -15000.0f;
        c->maxbuf[i] = 15000.0f;
    }
    c->sps = 10;
    c->center_idx = 4;
}

static void
reset_slicer(oracle_sym_chan* c) { /* dsd_symbol.c:1306-1326 */
    c->center = 0.0f;
    c->min = -30000.0f;
    c->max = 30000.0f;
    c->lmid = -20000.0f;
    c->umid = 20000.0f;
    c->minref = -24000.0f;
    c->maxref = 24000.0f;
    for (int i = 0; i < 1024; i++) {
        c->minbuf[i] = c->min;
        c->maxbuf[i] = c->max;
    }
    c->midx = 0;
    c->sum_window = 0;
}

static int
next_sps(oracle_sym_chan* c) { /* dsd_symbol.c:1328-1387 */
    if (c->sps_num != c->output_rate_hz || c->sps_den != c->symbol_rate_hz) {
        c->sps_num = c->output_rate_hz;
        c->sps_den = c->symbol_rate_hz;
        c->sps_accum = 0;
        c->jitter = -1;
        reset_slicer(c);
    }
    int whole = c->output_rate_hz / c->symbol_rate_hz;
    int rem = c->output_rate_hz % c->symbol_rate_hz;
    if (whole < 2) {
        whole = 2, rem = 0;
    }
    if (whole > 64) {
        whole = 64, rem = 0;
    }
    if (rem > 0 && c->sps_den > 0) {
        int acc = c->sps_accum + rem;
        if (acc >= c->sps_den) {
            whole++;
            acc -= c->sps_den;
        }
        c->sps_accum = acc;
        if (whole > 64) {
            whole = 64;
        }
    }
    return whole;
}

static float
matched_fir(oracle_sym_chan* c, float x) { /* dsd_filters.c:172-201 */
    if (!c->use_filter) {
        return x;
    }
    int head = c->fir_head + 1;
    if (head >= c->taps_len) {
        head = 0;
    }
    c->fir_hist[head] = x;
    c->fir_head = head;
    float acc = 0.0f;
    int last = c->taps_len - 1;
    for (int i = 0; i <= last; i++) {
        int idx = head - (last - i);
        if (idx < 0) {
            idx += c->taps_len;
        }
        acc += c->taps[i] * c->fir_hist[idx];
    }
    return acc;
}

/* One getSymbol() call.  `next(ctx)` yields the next discriminator sample. */
float
oracle_sym_get_symbol(oracle_sym_chan* c, int have_sync, float (*next)(void*), void* ctx) {
    c->sps = next_sps(c);
    c->center_idx = (c->sps - 1) / 2;
    /* select_window_c4fm / _gfsk (dsd_symbol.c:197-225) */
    const int l_edge = c->rf_mod == 2 ? 1 : c->window_l, r_edge = c->rf_mod == 2 ? 1 : 2;
    const int span = c->sps < 1 ? 1 : c->sps;
    if (span <= 1) {
        c->jitter = -1;
    }
    float sum = 0.0f;
    int count = 0;
    for (int i = 0; i < span; i++) {
        /* timing nudge, only at the first sample of an unsynchronised symbol (dsd_symbol.c:498-516, C4FM rule :489-496;
         * samples-per-symbol 20 uses the NXDN rule :462-468) */
        if (span > 1 && i == 0 && have_sync == 0 && c->jitter >= 0) {
            if (c->sps == 20) {
                if (c->jitter >= 7 && c->jitter <= 10) {
                    i--;
                } else if (c->jitter >= 11 && c->jitter <= 14) {
                    i++;
                }
            } else if (c->rf_mod == 2) { /* symbol_adjust_timing_gfsk (dsd_symbol.c:480-487) */
                if (c->jitter >= c->center_idx - 1 && c->jitter <= c->center_idx) {
                    i--;
                } else if (c->jitter >= c->center_idx + 1 && c->jitter <= c->center_idx + 2) {
                    i++;
                }
            } else {
                if (c->jitter > 0 && c->jitter <= c->center_idx) {
                    i--;
                } else if (c->jitter > c->center_idx && c->jitter < c->sps) {
                    i++;
                }
            }
            c->jitter = -1;
        }
        float s = next(ctx);
        s = matched_fir(c, s);
        if (have_sync == 1 && c->rf_mod == 0) { /* symbol_apply_sync_clip: C4FM only (dsd_symbol.c:347-358) */
            if (s > c->max) {
                s = c->max;
            } else if (s < c->min) {
                s = c->min;
            }
        }
        /* first zero crossing of the symbol (dsd_symbol.c:360-397, rf_mod == 0 branches) */
        if (s > c->center) {
            if (!(s > c->maxref * 1.25f)) {
                if (c->jitter < 0 && c->lastsample < c->center) {
                    c->jitter = i;
                }
            }
        } else {
            if (!(s < c->minref * 1.25f)) {
                if (c->jitter < 0 && c->lastsample > c->center) {
                    c->jitter = i;
                }
            }
        }
        /* window accumulation (dsd_symbol.c:423-460) */
        if (c->sps == 20 && i >= 7 && i <= 13) {
            sum += s;
            count++;
        }
        if (c->sps == 5 && i == 2) {
            sum += s;
            count++;
        } else if (c->rf_mod == 0) {
            if (i >= c->center_idx - l_edge && i <= c->center_idx + r_edge) {
                sum += s;
                count++;
            }
        } else { /* symbol_accumulate_other_window (dsd_symbol.c:428-434): the two edge samples only */
            const int hit = (c->sps <= 4) ? (i == c->center_idx) : (i == c->center_idx - l_edge || i == c->center_idx + r_edge);
            if (hit) {
                sum += s;
                count++;
            }
        }
        c->lastsample = s;
    }
    float symbol = count > 0 ? sum / (float)count : 0.0f;
    c->symbolcnt++;
    return symbol;
}

static void
extrema_avg2(const float* v, int n, float* omin, float* omax) { /* dsd_dibit.c:195-241 */
    if (n < 2) {
        *omin = *omax = 0.0f;
        return;
    }
    float mn1 = v[0], mn2 = v[1];
    if (mn2 < mn1) {
        float t = mn1;
        mn1 = mn2;
        mn2 = t;
    }
    float mx1 = v[0], mx2 = v[1];
    if (mx2 > mx1) {
        float t = mx1;
        mx1 = mx2;
        mx2 = t;
    }
    for (int i = 2; i < n; i++) {
        float x = v[i];
        if (x < mn1) {
            mn2 = mn1;
            mn1 = x;
        } else if (x < mn2) {
            mn2 = x;
        }
        if (x > mx1) {
            mx2 = mx1;
            mx1 = x;
        } else if (x > mx2) {
            mx2 = x;
        }
    }
    *omin = (mn1 + mn2) * 0.5f;
    *omax = (mx1 + mx2) * 0.5f;
}

static void
use_symbol(oracle_sym_chan* c) { /* dsd_dibit.c:243-299 */
    int cap = c->ssize;
    if (cap < 0) {
        cap = 0;
    }
    if (cap > 128) {
        cap = 128;
    }
    if (c->track_minmax) {
        float lmin, lmax;
        extrema_avg2(c->sbuf, cap, &lmin, &lmax);
        int window = c->msize < 1 ? 1 : (c->msize > 1024 ? 1024 : c->msize);
        if (c->sum_window != window) {
            double a = 0.0, b = 0.0;
            for (int i = 0; i < window; i++) {
                a += (double)c->minbuf[i];
                b += (double)c->maxbuf[i];
            }
            c->minbuf_sum = a;
            c->maxbuf_sum = b;
            c->sum_window = window;
            if (c->midx < 0 || c->midx >= window) {
                c->midx = 0;
            }
        }
        int idx = c->midx;
        if (idx < 0 || idx >= window) {
            idx = 0;
        }
        c->minbuf_sum += (double)lmin - (double)c->minbuf[idx];
        c->maxbuf_sum += (double)lmax - (double)c->maxbuf[idx];
        c->minbuf[idx] = lmin;
        c->maxbuf[idx] = lmax;
        idx++;
        c->midx = idx >= window ? 0 : idx;
        c->min = (float)(c->minbuf_sum / (double)window);
        c->max = (float)(c->maxbuf_sum / (double)window);
        c->center = (c->max + c->min) / 2.0f;
        c->umid = ((c->max - c->center) * 5.0f / 8.0f) + c->center;
        c->lmid = ((c->min - c->center) * 5.0f / 8.0f) + c->center;
        c->maxref = c->max * 0.80f;
        c->minref = c->min * 0.80f;
    } else {
        c->maxref = c->max;
        c->minref = c->min;
    }
    if (cap > 0) {
        if (c->sidx >= cap - 1) {
            c->sidx = 0;
        } else {
            c->sidx++;
        }
    }
}

static int
clamp255(int v) {
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

static int
bit_metric(float sym, const float ideal[4], int bit_index) { /* dsd_dibit.c:609-642 */
    float best0 = 3.4028234663852886e38f, best1 = 3.4028234663852886e38f, min_spacing = 3.4028234663852886e38f;
    for (int i = 0; i < 4; i++) {
        float d = (sym - ideal[i]) * (sym - ideal[i]);
        if (((i >> (1 - bit_index)) & 1) != 0) {
            if (d < best1) {
                best1 = d;
            }
        } else if (d < best0) {
            best0 = d;
        }
        for (int j = i + 1; j < 4; j++) {
            float sp = fabsf(ideal[i] - ideal[j]);
            if (sp > 1e-6f && sp < min_spacing) {
                min_spacing = sp;
            }
        }
    }
    if (min_spacing == 3.4028234663852886e38f) {
        min_spacing = 2.0f;
    }
    float scale = 255.0f / (min_spacing * min_spacing);
    return clamp255((int)lrintf(fabsf(best0 - best1) * scale));
}

static int
reliability(const oracle_sym_chan* c, float sym) { /* dsd_dibit.c:455-502 then :504-546 with no SNR hooks */
    const float eps = 1e-6f;
    int rel;
    if (sym > c->umid) {
        float span = c->max - c->umid;
        if (span < eps) {
            span = eps;
        }
        rel = (int)lrintf(((sym - c->umid) * 255.0f) / span);
    } else if (sym > c->center) {
        float d1 = sym - c->center, d2 = c->umid - sym, span = c->umid - c->center;
        if (span < eps) {
            span = eps;
        }
        float m = d1 < d2 ? d1 : d2;
        rel = (int)lrintf((m * 510.0f) / span);
    } else if (sym >= c->lmid) {
        float d1 = c->center - sym, d2 = sym - c->lmid, span = c->center - c->lmid;
        if (span < eps) {
            span = eps;
        }
        float m = d1 < d2 ? d1 : d2;
        rel = (int)lrintf((m * 510.0f) / span);
    } else {
        float span = c->lmid - c->min;
        if (span < eps) {
            span = eps;
        }
        rel = (int)lrintf(((c->lmid - sym) * 255.0f) / span);
    }
    rel = clamp255(rel);
    /* apply_c4fm_snr_weight with the hook sentinel (-100 dB): w256 = 0 => scale 204/256 */
    int scaled = (rel * 204) >> 8;
    return clamp255(scaled);
}

/* One getDibitSoft() call: returns the dibit (pre-inversion value, as the reference returns), fills symbol/soft. */
int
oracle_sym_get_dibit(oracle_sym_chan* c, float (*next)(void*), void* ctx, float* symbol_out, uint8_t* rel_out, int16_t llr_out[2]) {
    float sym = oracle_sym_get_symbol(c, 1, next, ctx);
    c->sbuf[c->sidx] = sym;
    use_symbol(c);
    /* digitize, four-level (dsd_dibit.c:963-976,1018-1041) */
    int dibit;
    if (sym > c->center) {
        dibit = sym > c->umid ? (c->negative ? 3 : 1) : (c->negative ? 2 : 0);
    } else {
        dibit = sym < c->lmid ? (c->negative ? 1 : 3) : (c->negative ? 0 : 2);
    }
    /* compute_dibit_soft_metric with build_standard_dibit_ideals (dsd_dibit.c:644-721) */
    float plus_one = 0.5f * (c->center + c->umid), minus_one = 0.5f * (c->lmid + c->center);
    float ideal[4];
    if (c->negative) {
        ideal[0] = minus_one, ideal[1] = c->min, ideal[2] = plus_one, ideal[3] = c->max;
    } else {
        ideal[0] = plus_one, ideal[1] = c->max, ideal[2] = minus_one, ideal[3] = c->min;
    }
    int mag0 = bit_metric(sym, ideal, 0), mag1 = bit_metric(sym, ideal, 1);
    int rel = reliability(c, sym);
    int min_mag = mag0 < mag1 ? mag0 : mag1;
    if (min_mag > 0 && rel < min_mag) {
        mag0 = (mag0 * rel) / min_mag;
        mag1 = (mag1 * rel) / min_mag;
    }
    mag0 = clamp255(mag0);
    mag1 = clamp255(mag1);
    llr_out[0] = (int16_t)(((dibit >> 1) & 1) ? mag0 : -mag0);
    llr_out[1] = (int16_t)((dibit & 1) ? mag1 : -mag1);
    int a0 = llr_out[0] < 0 ? -llr_out[0] : llr_out[0], a1 = llr_out[1] < 0 ? -llr_out[1] : llr_out[1];
    *rel_out = (uint8_t)clamp255(a1 < a0 ? a1 : a0);
    *symbol_out = sym;
    return dibit;
}

/* array drivers for ctypes */
typedef struct {
    const float* p;
    long n, pos;
} arr_src;

static float
arr_next(void* v) {
    arr_src* s = (arr_src*)v;
    return s->pos < s->n ? s->p[s->pos++] : 0.0f;
}

long
oracle_sym_run_symbols(oracle_sym_chan* c, int have_sync, const float* samples, long n, long reserve, float* out, long max_out,
                       long* consumed) {
    arr_src s = {samples, n, 0};
    long k = 0;
    while (k < max_out && (s.n - s.pos) >= reserve) {
        out[k++] = oracle_sym_get_symbol(c, have_sync, arr_next, &s);
    }
    *consumed = s.pos;
    return k;
}

long
oracle_sym_run_dibits(oracle_sym_chan* c, const float* samples, long n, long reserve, uint8_t* dibits, uint8_t* rel, int16_t* llr2,
                      float* symbols, long max_out, long* consumed) {
    arr_src s = {samples, n, 0};
    long k = 0;
    while (k < max_out && (s.n - s.pos) >= reserve) {
        dibits[k] = (uint8_t)oracle_sym_get_dibit(c, arr_next, &s, &symbols[k], &rel[k], &llr2[2 * k]);
        k++;
    }
    *consumed = s.pos;
    return k;
}


/* ---- frame-sync hunt (src/dsp/dsd_frame_sync.c:3098-3148): hunt-time slice (:2110-2127) symbol > 0 -> '1' else '3',
 * rolling character history, strcmp of its last strlen(pattern) characters against each pattern in table order; a
 * pattern is only tried once that many characters have been pushed (frame_sync_history_materialize).  Reports every
 * position where some pattern completes (the reference returns at the first and resumes after the frame).
 * hist (32 chars) / hist_count carry across calls; returns the number of hits found (only max_hits are stored). */
int
oracle_frame_sync_search(const float* symbols, int n, const char* const* patterns, const int* sync_types, int n_patterns,
                         char* hist32, int* hist_count, int* hit_pos, int* hit_type, int max_hits) {
    int found = 0;
    for (int p = 0; p < n; p++) {
        memmove(hist32, hist32 + 1, 31);
        hist32[31] = symbols[p] > 0.0f ? '1' : '3';
        if (*hist_count < 32) {
            (*hist_count)++;
        }
        for (int k = 0; k < n_patterns; k++) {
            int L = (int)strlen(patterns[k]);
            if (*hist_count >= L && strncmp(hist32 + 32 - L, patterns[k], (size_t)L) == 0) {
                if (found < max_hits) {
                    hit_pos[found] = p;
                    hit_type[found] = sync_types[k];
                }
                found++;
                break;
            }
        }
    }
    return found;
}

/* ---- acquisition: getFrameSync() (src/dsp/dsd_frame_sync.c:3098-3148) from the never-synchronised state -------------------
 *
 * Per symbol, in the reference's order: getSymbol(have_sync = 0) (timing nudges on, no matched filter before the first sync:
 * symbol_apply_matched_filter selects by lastsynctype, dsd_symbol.c:301-337); the 24-symbol level ring and sbuf write
 * (frame_sync_update_symbol_ring, :1747-1764); the rolling DMR payload dibit + reliability with the thresholds in force
 * (frame_sync_store_dmr_payload_symbol, :2161-2190, dmr_compute_reliability dsd_dibit.c:548-568); the hunt-time slice
 * symbol > 0 -> '1' else '3' (:2110-2127); from the 8th symbol on frame_sync_eval_window (:2638-2676): sorted-window level
 * estimate (frame_sync_level.c), maxref / minref = max / min for FSK profiles (:2332-2335), then the pattern compare.
 * A call gives up after 1800 symbols without sync (:3039) and the next call starts with an empty window, which is restated
 * here as a restart of the hunt context (the no-carrier hook of that moment is the host's business).
 * On a match: frame_sync_set_basic_lock (:386-392), then for P25 Phase 1 dsd_sync_warm_start_thresholds_outer_only(24) when
 * rf_mod == 0 (:611-622, src/dsp/sync_calibration.c:155-226), for DMR dmr_resample_on_sync (src/dsp/dmr_sync.c:108-126):
 * the same warm start, then the 66 dibits in front of the sync re-sliced with the new thresholds.
 * Pinned against the UNMODIFIED getFrameSync() by tests/test_oracle_symbol.py (ref_sym_frame_sync). */
static int
cmp_float(const void* a, const void* b) {
    const float x = *(const float*)a, y = *(const float*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static void
window_levels(const float* sorted, int count, float* omin, float* omax) { /* dsd_frame_sync_estimate_sorted_window_levels */
    if (count <= 0) {
        *omin = *omax = 0.0f;
        return;
    }
    if (count < 3) {
        float sum = 0.0f;
        for (int i = 0; i < count; i++) {
            sum += sorted[i];
        }
        *omin = *omax = sum / (float)count;
        return;
    }
    int min_idx = 0, max_idx = count - 3;
    if (count >= 13) {
        min_idx = 2;
        max_idx = count - 5;
    }
    if (max_idx + 2 >= count) {
        max_idx = count - 3;
    }
    *omin = (sorted[min_idx] + sorted[min_idx + 1] + sorted[min_idx + 2]) / 3.0f;
    *omax = (sorted[max_idx] + sorted[max_idx + 1] + sorted[max_idx + 2]) / 3.0f;
}

/* dsd_sync_warm_start_thresholds_outer_only over the newest `len` symbols (newest first, as the reference walks them).
 * Returns 1 when the thresholds were replaced (DSD_WARM_START_OK). */
int
oracle_warm_start_thresholds(oracle_sym_chan* c, const float* newest_first, int len) {
    float sum_pos = 0.0f, sum_neg = 0.0f;
    int n_pos = 0, n_neg = 0;
    for (int i = 0; i < len; i++) {
        const float v = newest_first[i];
        if (v > 0.0f) {
            sum_pos += v;
            n_pos++;
        } else {
            sum_neg += v;
            n_neg++;
        }
    }
    if (n_pos == 0 || n_neg == 0) {
        return 0;
    }
    const float mean_pos = sum_pos / (float)n_pos, mean_neg = sum_neg / (float)n_neg;
    if (fabsf(mean_pos - mean_neg) < 1.0f) {
        return 0;
    }
    c->max = mean_pos;
    c->min = mean_neg;
    c->center = (c->max + c->min) / 2.0f;
    c->umid = c->center + (c->max - c->center) * 0.625f;
    c->lmid = c->center + (c->min - c->center) * 0.625f;
    c->maxref = c->max * 0.80f;
    c->minref = c->min * 0.80f;
    int fill = c->msize > 1024 ? 1024 : c->msize;
    for (int i = 0; i < fill; i++) {
        c->maxbuf[i] = c->max;
        c->minbuf[i] = c->min;
    }
    c->sum_window = 0; /* dsd_state_invalidate_minmax_sums */
    return 1;
}

static int
payload_dibit(const oracle_sym_chan* c, float symbol) {
    if (symbol > c->center) {
        return symbol > c->umid ? 1 : 0;
    }
    return symbol < c->lmid ? 3 : 2;
}

long
oracle_sym_acquire(oracle_sym_chan* c, const float* samples, long n, long reserve, const oracle_acq_pattern* pats, int n_pats,
                   float* sym_out, uint8_t* dib_out, uint8_t* rel_out, long max_out, oracle_acq_result* res) {
    arr_src src = {samples, n, 0};
    memset(res, 0, sizeof(*res));
    res->sync_type = -1;
    float lbuf[48], sorted[48];
    int lidx = 0, level_count = 0, count = 0, since_start = 0;
    float lmin = c->min, lmax = c->max;
    char win[25];
    memset(win, 0, sizeof(win));
    long k = 0;
    while (k < max_out && (src.n - src.pos) >= reserve) {
        const float symbol = oracle_sym_get_symbol(c, 0, arr_next, &src);
        /* frame_sync_update_symbol_ring */
        lbuf[lidx] = symbol;
        if (level_count < 24) {
            level_count++;
        }
        c->sbuf[c->sidx] = symbol;
        lidx = (lidx == 23) ? 0 : lidx + 1;
        c->sidx = (c->sidx == c->ssize - 1) ? 0 : c->sidx + 1;
        /* rolling DMR payload buffer */
        sym_out[k] = symbol;
        dib_out[k] = (uint8_t)payload_dibit(c, symbol);
        rel_out[k] = (uint8_t)reliability(c, symbol);
        k++;
        memmove(win, win + 1, 23);
        win[23] = symbol > 0.0f ? '1' : '3';
        if (count < 24) {
            count++;
        }
        since_start++;
        if (since_start >= 8) { /* history_count >= 8: frame_sync_eval_window */
            memcpy(sorted, lbuf, (size_t)level_count * sizeof(float));
            qsort(sorted, (size_t)level_count, sizeof(float), cmp_float);
            window_levels(sorted, level_count, &lmin, &lmax);
            c->maxref = c->max;
            c->minref = c->min;
            for (int p = 0; p < n_pats; p++) {
                const int L = (int)strlen(pats[p].symbols);
                if (count < L || strncmp(win + 24 - L, pats[p].symbols, (size_t)L) != 0) {
                    continue;
                }
                /* frame_sync_set_basic_lock */
                c->max = (c->max + lmax) / 2;
                c->min = (c->min + lmin) / 2;
                float newest_first[24];
                for (int i = 0; i < 24; i++) {
                    newest_first[i] = (k - 1 - i) >= 0 ? sym_out[k - 1 - i] : 0.0f;
                }
                if (pats[p].kind == 1 || c->rf_mod == 0) {
                    res->warm_start = oracle_warm_start_thresholds(c, newest_first, 24);
                }
                if (pats[p].kind == 1 && k >= 90) { /* dmr_resample_cach: symbols [-90 .. -25] */
                    for (int i = 0; i < 66; i++) {
                        const float sv = sym_out[k - 90 + i];
                        res->resampled[i] = (uint8_t)payload_dibit(c, sv);
                        dib_out[k - 90 + i] = res->resampled[i];
                    }
                    res->resample_ok = 1;
                }
                /* the decoder state the frame handlers see from here on */
                c->use_filter = pats[p].use_filter && pats[p].taps && pats[p].taps_len > 0;
                c->taps_len = c->use_filter ? pats[p].taps_len : 0;
                if (c->use_filter) {
                    memcpy(c->taps, pats[p].taps, (size_t)pats[p].taps_len * sizeof(float));
                }
                c->window_l = pats[p].window_l;
                c->track_minmax = pats[p].track_minmax;
                c->negative = pats[p].negative;
                res->sync_type = pats[p].sync_type;
                res->hunt_symbols = k;
                res->consumed = src.pos;
                res->lmin = lmin;
                res->lmax = lmax;
                return k;
            }
        }
        if (since_start >= 1800) { /* frame_sync_handle_no_sync_timeout: the next getFrameSync() call starts afresh */
            lidx = 0, level_count = 0, count = 0, since_start = 0;
            lmin = c->min, lmax = c->max;
            memset(win, 0, sizeof(win));
        }
    }
    res->hunt_symbols = k;
    res->consumed = src.pos;
    res->lmin = lmin;
    res->lmax = lmax;
    return k;
}

/* ---- symbol-rate CQPSK input (output kind 2): the sample side behind the CQPSK chain ------------------------------------
 *
 *   getSymbol fast path        src/dsp/dsd_symbol.c:1583-1625 (fixed +-2 / 0 thresholds before every take, :744-765; one
 *                              stream float = one symbol)
 *   use_symbol                 src/core/frames/dsd_dibit.c:243-299 with rf_mod == 1: the rolling min / max tracker always runs
 *   digitize                   :1018-1041 -> select_four_level_dibit :978-1003: cqpsk_slice(symbol - center) (:329-349) through
 *                              the OP25 dibit orientation map (include/dsd-neo/core/p25_cqpsk_dibit.h:28-52) when the CQPSK
 *                              chain is active and the sync is a P25 one, else the threshold regions (:963-976)
 *   soft metric                compute_dibit_soft_metric :685-721 with build_cqpsk_dibit_ideals (:660-683) or the standard
 *                              ideals; reliability = cqpsk_reliability_raw (:376-401) weighted by the CQPSK SNR hook (:404-427)
 * Not restated: the DSD_NEO_CQPSK_SYNC_INV / _NEG debug switches (cfg->cqpsk_sync_inv / neg, default off).
 * Pinned against the compiled reference by tests/test_oracle_symbol.py (dibits, reliabilities, LLRs, thresholds). */
void
oracle_cqpsk_slicer_init(oracle_cqpsk_slicer* s, int negative, int p25_slice, int map_idx, double snr_db, int ssize, int msize) {
    oracle_sym_init(&s->base, 4800, 4800, 0, 2, 1, negative, 0, 0, ssize, msize);
    s->p25_slice = p25_slice;
    s->map_idx = (map_idx >= 0 && map_idx < 5) ? map_idx : 0;
    s->snr_db = snr_db;
}

static int
cqpsk_map_correct(int map_idx, int dibit) {
    static const uint8_t maps[5][4] = {{0, 1, 2, 3}, {2, 3, 0, 1}, {3, 2, 1, 0}, {1, 3, 0, 2}, {2, 0, 3, 1}};
    return maps[map_idx][dibit & 3];
}

static int
dibit_invert(int d) { /* dsd_dibit.c:300-311 */
    return (d + 2) & 3;
}

int
oracle_cqpsk_slicer_dibit(oracle_cqpsk_slicer* s, float sample, uint8_t* rel_out, int16_t llr_out[2]) {
    oracle_sym_chan* c = &s->base;
    /* symbol_try_rtl_symbol_rate_fast_path */
    c->center = 0.0f;
    c->min = -3.0f;
    c->max = 3.0f;
    c->lmid = -2.0f;
    c->umid = 2.0f;
    c->minref = -2.4f;
    c->maxref = 2.4f;
    const float sym = sample;
    c->lastsample = sym;
    c->symbolcnt++;
    /* get_dibit_and_analog_signal */
    c->sbuf[c->sidx] = sym;
    use_symbol(c);
    int dibit;
    float ideal[4];
    if (s->p25_slice) {
        const float v = sym - c->center;
        int raw = v >= 2.0f ? 1 : (v >= 0.0f ? 0 : (v >= -2.0f ? 2 : 3));
        dibit = cqpsk_map_correct(s->map_idx, raw);
        if (c->negative) {
            dibit = dibit_invert(dibit);
        }
        static const float base_ideal[4] = {1.0f, 3.0f, -1.0f, -3.0f};
        for (int d = 0; d < 4; d++) {
            const int corrected = c->negative ? dibit_invert(d) : d;
            int mapped = corrected;
            for (int raw_d = 0; raw_d < 4; raw_d++) { /* dsd_p25_cqpsk_raw_dibit_for_corrected */
                if (cqpsk_map_correct(s->map_idx, raw_d) == corrected) {
                    mapped = raw_d;
                    break;
                }
            }
            ideal[d] = c->center + base_ideal[mapped];
        }
    } else {
        if (sym > c->center) {
            dibit = sym > c->umid ? (c->negative ? 3 : 1) : (c->negative ? 2 : 0);
        } else {
            dibit = sym < c->lmid ? (c->negative ? 1 : 3) : (c->negative ? 0 : 2);
        }
        const float plus_one = 0.5f * (c->center + c->umid), minus_one = 0.5f * (c->lmid + c->center);
        if (c->negative) {
            ideal[0] = minus_one, ideal[1] = c->min, ideal[2] = plus_one, ideal[3] = c->max;
        } else {
            ideal[0] = plus_one, ideal[1] = c->max, ideal[2] = minus_one, ideal[3] = c->min;
        }
    }
    int mag0 = bit_metric(sym, ideal, 0), mag1 = bit_metric(sym, ideal, 1);
    /* dmr_compute_reliability, rf_mod == 1 */
    int rel;
    {
        const float sc = sym - c->center;
        const float id = sc >= 2.0f ? 3.0f : (sc >= 0.0f ? 1.0f : (sc >= -2.0f ? -1.0f : -3.0f));
        float err = fabsf(sc - id);
        if (err > 1.0f) {
            err = 1.0f;
        }
        rel = clamp255((int)((1.0f - err) * 255.0f + 0.5f));
        if (!(s->snr_db <= -50.0)) {
            int w256 = 0;
            if (s->snr_db >= 25.0) {
                w256 = 255;
            } else if (s->snr_db > 0.0) {
                w256 = (int)((s->snr_db / 25.0) * 255.0 + 0.5);
            }
            rel = clamp255((rel * (204 + (w256 >> 2))) >> 8);
        }
    }
    const int min_mag = mag0 < mag1 ? mag0 : mag1;
    if (min_mag > 0 && rel < min_mag) {
        mag0 = (mag0 * rel) / min_mag;
        mag1 = (mag1 * rel) / min_mag;
    }
    mag0 = clamp255(mag0);
    mag1 = clamp255(mag1);
    llr_out[0] = (int16_t)(((dibit >> 1) & 1) ? mag0 : -mag0);
    llr_out[1] = (int16_t)((dibit & 1) ? mag1 : -mag1);
    const int a0 = llr_out[0] < 0 ? -llr_out[0] : llr_out[0], a1 = llr_out[1] < 0 ? -llr_out[1] : llr_out[1];
    *rel_out = (uint8_t)clamp255(a1 < a0 ? a1 : a0);
    return dibit;
}

long
oracle_cqpsk_slicer_run(oracle_cqpsk_slicer* s, const float* symbols, long n, uint8_t* dibits, uint8_t* rel, int16_t* llr2) {
    for (long i = 0; i < n; i++) {
        dibits[i] = (uint8_t)oracle_cqpsk_slicer_dibit(s, symbols[i], &rel[i], &llr2[2 * i]);
    }
    return n;
}
