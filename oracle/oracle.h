/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the reference algorithms on the accelerated hot path of
 * arancormonk/dsd-neo @ 4d06905.  Plain C11, one channel / one codeword at a time, written for
 * clarity.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this; the product library (libdsdneo_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py) against
 *   (a) the reference's own known-answer vectors for the path, and
 *   (b) the unmodified reference sources compiled in place into oracle/_ref/libdsdneo_ref.so
 *       (oracle/Makefile), on seeded random inputs,
 * except the polyphase channelizer, which has no reference implementation (its oracle is the
 * mathematical direct form in float64) -- see DESIGN.md.
 *
 * Build: -O2 -fno-fast-math -ffp-contract=off (float results are order- and contraction-exact).
 */
#ifndef DSDNEO_ORACLE_H_
#define DSDNEO_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------- block side ------------------------------------------ */

#define ORACLE_LPF_MAX_TAPS 144

typedef struct oracle_demod_chan {
    /* configuration */
    int rate_out_hz;
    int lpf_enable;
    int lpf_profile;
    int fir_fma; /* 1: AVX2-kernel arithmetic (fused), 0: scalar/SSE2 arithmetic */
    float squelch_level;
    /* channel LPF plan + history (demod_state.channel_lpf_*) */
    int taps_len;
    float taps[ORACLE_LPF_MAX_TAPS];
    float hist_i[ORACLE_LPF_MAX_TAPS];
    float hist_q[ORACLE_LPF_MAX_TAPS];
    /* dsd_fsk_modem_state */
    float prev_i, prev_q;
    int have_prev;
    float dc_est;
    float peak_est;
    /* squelch */
    float channel_pwr;
    int channel_squelched;
} oracle_demod_chan;

int oracle_channel_lpf_design(int rate_out_hz, int profile, float* taps_out, int max_taps);
void oracle_fir_complex(const float* in, int in_len, float* out, float* hist_i, float* hist_q, const float* taps,
                        int taps_len, int fma);
float oracle_mean_power(const float* samples, int len, int step);
extern const float oracle_hb15_taps[15];
extern const float oracle_hb31_taps[31];
int oracle_hb_decim2_complex(const float* in, int in_len, float* out, float* hist_i, float* hist_q, const float* taps,
                             int taps_len, int fma);
int oracle_hb_cascade(const float* in, int in_len, int passes, float* hist, float* work, float* out, int fma);
int oracle_demod_chan_init(oracle_demod_chan* c, int rate_out_hz, int profile, int lpf_enable, float squelch_level,
                           int fir_fma);
/* One full_demod() call: n_floats interleaved I/Q in, n_floats/2 discriminator samples out.
 * `scratch` must hold n_floats floats. Returns result_len. */
int oracle_full_demod_block(oracle_demod_chan* c, const float* iq, int n_floats, float* scratch, float* out);

/* Polyphase channelizer, direct form in float64 (definition in DESIGN.md):
 *   y_k[n] = sum_{m=0}^{L-1} h[m] x[t_n - m] exp(-j 2 pi k (t_n - m) / M),  t_n = n*D + M - 1 - (hist offset)
 * x is preceded by `hist_len` = L - 1 (at least) earlier samples at x_with_hist[0..hist_len). */
void oracle_pfb_direct(const float* x_with_hist, int hist_len, int n_in, const float* h, int L, int M, int D,
                       const int* channels, int n_sel, double* out_re_im /* [n_sel][n_out][2] */, int n_out);

/* ------------------------------- CQPSK block side (oracle_cqpsk.c) --------------------------- */

#define ORACLE_FLL_MAX_TAPS 48
#define ORACLE_CQPSK_RING 64

typedef struct oracle_cqpsk_chan {
    /* configuration */
    int rate_out_hz, sps;
    float ted_gain;
    int ted_gain_is_set;
    /* cqpsk_rms_agc */
    float agc_avg;
    /* dsd_fll_band_edge_state_t */
    int fll_ntaps;
    float fll_alpha, fll_beta, fll_phase, fll_freq;
    float fll_lower_r[ORACLE_FLL_MAX_TAPS], fll_lower_i[ORACLE_FLL_MAX_TAPS];
    float fll_upper_r[ORACLE_FLL_MAX_TAPS], fll_upper_i[ORACLE_FLL_MAX_TAPS];
    /* history of FLL outputs (serves as the FLL delay line and the Gardner delay line) */
    float ring_r[ORACLE_CQPSK_RING], ring_j[ORACLE_CQPSK_RING];
    long pushed, consumed;
    /* ted_state_t */
    float mu, omega, omega_mid, omega_rel, last_r, last_j, lock_accum, ted_effective_gain;
    int lock_count, ted_span;
    /* op25_diff_phasor_cc */
    float diff_prev_r, diff_prev_j;
    /* dsd_costas_loop_state_t + per-block metrics */
    float costas_alpha, costas_beta, costas_phase, costas_freq, costas_err_smooth, costas_error;
    float m_err_abs, m_err_raw_abs, m_conf_acc;
    int m_zero_conf;
    int costas_err_avg_q14, costas_err_raw_avg_q14, costas_conf_avg_q14, costas_zero_conf_pct;
} oracle_cqpsk_chan;

const float* oracle_cqpsk_mmse_table(void); /* [17][8] */
int oracle_fll_band_edge_design(int sps, float* lower_r, float* lower_i, float* upper_r, float* upper_i, int max_taps);
int oracle_cqpsk_chan_init(oracle_cqpsk_chan* q, int rate_out_hz, int sps, float ted_gain, int ted_gain_is_set);
/* one un-squelched block of channel-filtered interleaved I/Q (>= 4 pairs) -> symbols near {-3,-1,+1,+3}; returns count */
int oracle_cqpsk_block(oracle_cqpsk_chan* q, const float* lp, int n_floats, float* out);
/* full_demod() with output_kind == SYMBOL_CQPSK: channel LPF + squelch (state in c) then the chain above */
int oracle_full_demod_cqpsk_block(oracle_demod_chan* c, oracle_cqpsk_chan* q, const float* iq, int n_floats,
                                  float* scratch, float* out);

/* ------------------------------- FEC leaves (oracle_fec.c) ----------------------------------- */

enum { ORACLE_HAMMING_7_4 = 0, ORACLE_HAMMING_12_8, ORACLE_HAMMING_13_9, ORACLE_HAMMING_15_11, ORACLE_HAMMING_16_11_4 };

void oracle_fec_init(void);
int oracle_hamming_decode(int code, uint8_t* bits, uint8_t* decoded);
int oracle_golay_24_12_decode(uint8_t* bits24);
void oracle_golay_24_12_encode(const uint8_t* data12, uint8_t* out24);
int oracle_golay_20_8_decode(uint8_t* bits20);
int oracle_dmr_r34_decode(const uint8_t* dibits98, const uint8_t* reliab98, uint8_t* out18);
int oracle_rs_12_9_decode(uint8_t* cw12, uint8_t* syndrome3, uint8_t* errors_found);
int oracle_qr_16_7_6_decode(uint8_t* bits16);
void oracle_bptc_deinterleave(const uint8_t* in196, uint8_t* out196);
unsigned oracle_bptc_196x96_extract(const uint8_t* in196, uint8_t* out96, uint8_t* r3, int* undefined_out);
unsigned oracle_bptc_128x77_extract(const uint8_t* in128, uint8_t* out77, int* undefined_out);
unsigned oracle_bptc_16x2_extract(const uint8_t* in32, uint8_t* out32, unsigned parity_odd, int* undefined_out);
int oracle_p25_12_soft_llr(const int16_t* llr196, uint8_t out12[12]);
int oracle_p25_12_soft_llr_list(const int16_t* llr196, uint8_t* cand_bytes, uint32_t* cand_metric, int max_candidates);
int oracle_rs63_decode(int tt, const int* in63, int* out63);
void oracle_rs63_encode(int tt, const int* data, int* cw63);
int oracle_p25_rs_decode(int n_total, int n_data, uint8_t* data_bits, const uint8_t* parity_bits);
int oracle_rs63_decode_with_erasures(int tt, const int* in63, int* out63, const int* erasures, int n_erasures);
int oracle_p25_rs_ranked_erasures(const uint8_t* data_rel, int n_data, const uint8_t* par_rel, int n_par, int min_er, int threshold,
                                  int* erasures, int max_er);
int oracle_p25_rs_soft_reliability(int n_total, int n_data, uint8_t* data_bits, const uint8_t* parity_bits, const uint8_t* data_rel,
                                   const uint8_t* par_rel, int threshold);
int oracle_p25_golay24_decode(int length, uint8_t* word, const uint8_t* parity, int* fixed_errors);
int oracle_hamming_10_6_3_decode(uint8_t* data6, const uint8_t* parity4);
/* check_and_fix_golay_24_6_soft / _24_12_soft (src/protocol/p25/phase1/p25p1_soft.cpp:477-593); length = 6 or 12 */
int oracle_p25_golay24_soft(int length, uint8_t* data, const uint8_t* parity, const int* reliab, int hard_override_enabled,
                            int threshold, int* fixed);
/* hamming_10_6_3_soft (src/protocol/p25/phase1/p25p1_soft.cpp:444-475) */
int oracle_hamming_10_6_3_soft(const uint8_t* bits10, const int* reliab10, int hard_override_enabled, int threshold, uint8_t* out10);
int oracle_bch_63_16_decode(const uint8_t* in63, uint8_t* out16, int* error_count);
/* p25p1_nid_decode (src/protocol/p25/phase1/p25p1_check_nid.cpp:322-354); returns NidResult status, fills nac / duid / errs */
/* sequential DMR BS data burst cutter (CACH, 196 info bits + reliabilities, 20 slot-type bits); 1 = burst complete */
int oracle_dmr_burst_cut(const uint8_t* dibits, const uint8_t* reliab, int count, int pos_last_sync, int inverted, uint8_t* cach24,
                         uint8_t* info196, uint8_t* rel98, uint8_t* slot20);
/* sequential P25p1 frame cutter (NID fields + status-stripped payload); bit 0 NID complete, bit 1 payload complete */
int oracle_p25p1_frame_cut(const uint8_t* dibits, const int16_t* llr, int count, int pos_last_sync, int n_payload,
                           uint8_t* code63, uint8_t* reliab63, uint8_t* parity, uint8_t* parity_reliab, uint8_t* payload_dibits,
                           int16_t* payload_llr);
int oracle_p25p1_nid_decode(const uint8_t* code63, const uint8_t* reliab63, int observed_nac, int parity, int parity_reliab,
                            int threshold, int* nac, int* duid, int* errs);
uint32_t oracle_viterbi_k5_decode(uint8_t* out, const uint16_t* in, int len);
uint32_t oracle_viterbi_k5_decode_punctured(uint8_t* out, const uint16_t* in, const uint8_t* punct, int in_len, int p_len);
void oracle_nxdn_conv_decode(const uint8_t* sym, const uint8_t* rel, int n_steps, int n_bits_out, uint16_t* metrics_io, uint8_t* out);

/* ------------------------------- sample side (oracle_symbol.c) --------------------------------- */

#define ORACLE_SYM_MAX_TAPS 256
typedef struct oracle_sym_chan {
    /* configuration */
    int output_rate_hz, symbol_rate_hz;
    int use_filter;      /* matched filter selected by lastsynctype (dsd_symbol.c:301-337) and opts->use_cosine_filter */
    int window_l;        /* left window edge: 2 (P25, NXDN96...) or 1 (YSF / DMR), dsd_symbol.c:197-211 */
    int track_minmax;    /* use_symbol() threshold tracking: rf_mod 0 and lastsynctype is P25p1 (dsd_dibit.c:264) */
    int negative;        /* is_four_level_neg_synctype(synctype) */
    int rf_mod;          /* 0 C4FM, 2 GFSK (window / accumulation / nudge rules, no sync clip) */
    int ssize, msize;
    int taps_len;
    float taps[ORACLE_SYM_MAX_TAPS];
    /* carried state */
    float fir_hist[ORACLE_SYM_MAX_TAPS];
    int fir_head;
    int sps_num, sps_den, sps_accum;
    int sps, center_idx, jitter;
    float lastsample;
    float min, max, center, umid, lmid, minref, maxref;
    float sbuf[128];
    int sidx;
    float minbuf[1024], maxbuf[1024];
    int midx, sum_window;
    double minbuf_sum, maxbuf_sum;
    long symbolcnt;
} oracle_sym_chan;

void oracle_sym_init(oracle_sym_chan* c, int output_rate_hz, int symbol_rate_hz, int use_filter, int window_l, int track_minmax,
                     int negative, const float* taps, int taps_len, int ssize, int msize);
float oracle_sym_get_symbol(oracle_sym_chan* c, int have_sync, float (*next)(void*), void* ctx);
int oracle_sym_get_dibit(oracle_sym_chan* c, float (*next)(void*), void* ctx, float* symbol_out, uint8_t* rel_out, int16_t llr_out[2]);
long oracle_sym_run_symbols(oracle_sym_chan* c, int have_sync, const float* samples, long n, long reserve, float* out, long max_out,
                            long* consumed);
long oracle_sym_run_dibits(oracle_sym_chan* c, const float* samples, long n, long reserve, uint8_t* dibits, uint8_t* rel,
                           int16_t* llr2, float* symbols, long max_out, long* consumed);

/* symbol-rate CQPSK input (output kind 2): thresholds / tracker / CQPSK slicer / soft metric of the sample side */
typedef struct oracle_cqpsk_slicer {
    oracle_sym_chan base; /* tracker state (sbuf, minbuf / maxbuf, running sums), thresholds, negative flag */
    int p25_slice;        /* is_cqpsk_active() && rf_mod == 1 && P25 sync type (dsd_dibit.c:950-961) */
    int map_idx;          /* state->p25_cqpsk_dibit_map_idx */
    double snr_db;        /* dsd_rtl_stream_metrics_hook_snr_cqpsk_db(); <= -50 = no hook */
} oracle_cqpsk_slicer;
void oracle_cqpsk_slicer_init(oracle_cqpsk_slicer* s, int negative, int p25_slice, int map_idx, double snr_db, int ssize, int msize);
int oracle_cqpsk_slicer_dibit(oracle_cqpsk_slicer* s, float sample, uint8_t* rel_out, int16_t llr_out[2]);
long oracle_cqpsk_slicer_run(oracle_cqpsk_slicer* s, const float* symbols, long n, uint8_t* dibits, uint8_t* rel, int16_t* llr2);

/* acquisition = getFrameSync() from the never-synchronised state (oracle_symbol.c) */
typedef struct oracle_acq_pattern {
    const char* symbols; /* '1' / '3' string, oldest first */
    int sync_type;
    int kind;            /* 0 P25 Phase 1 (threshold warm start when rf_mod == 0), 1 DMR (warm start + resample-on-sync) */
    int use_filter;      /* the decoder class from the sync on: matched filter, window, tracker, polarity */
    const float* taps;
    int taps_len;
    int window_l, track_minmax, negative;
} oracle_acq_pattern;
typedef struct oracle_acq_result {
    int sync_type;       /* -1: the samples ran out first */
    int warm_start;      /* DSD_WARM_START_OK */
    int resample_ok;
    long hunt_symbols, consumed;
    float lmin, lmax;
    uint8_t resampled[66];
} oracle_acq_result;
int oracle_warm_start_thresholds(oracle_sym_chan* c, const float* newest_first, int len);
long oracle_sym_acquire(oracle_sym_chan* c, const float* samples, long n, long reserve, const oracle_acq_pattern* pats, int n_pats,
                        float* sym_out, uint8_t* dib_out, uint8_t* rel_out, long max_out, oracle_acq_result* res);
int oracle_frame_sync_search(const float* symbols, int n, const char* const* patterns, const int* sync_types, int n_patterns,
                             char* hist32, int* hist_count, int* hit_pos, int* hit_type, int max_hits);
/* ------------------------------- MBE synthesis stage (oracle_mbe.c) -- PARITY UNPINNED, see its header --------- */
typedef struct { /* struct mbe_parameters of mbelib 1.3.0 (mbelib.h) */
    float w0;
    int L;
    int K;
    int Vl[57];
    float Ml[57];
    float log2Ml[57];
    float PHIl[57];
    float PSIl[57];
    float gamma;
    int un;
    int repeat;
} oracle_mbe_parms;
float oracle_mbe_uniform(uint64_t key, int band, int sample, int index, int stream);
void oracle_mbe_spectral_amp_enhance(oracle_mbe_parms* cur);
void oracle_mbe_synthesize_speechf(float* aout, oracle_mbe_parms* cur, oracle_mbe_parms* prev, int uvquality, uint64_t key);
void oracle_mbe_floattoshort(const float* in, int16_t* out);
void oracle_mbe_synth_frame(float* aout, int16_t* pcm, oracle_mbe_parms* cur, oracle_mbe_parms* prev_enhanced, int uvquality,
                            uint64_t key);

void oracle_libm_atan2f_array(const float* y, const float* x, float* out, long n);

#ifdef __cplusplus
}
#endif


/* ---- P25 Phase 1 frame handlers over a dibit stream (oracle_p25p1_frame.c) ---- */
/* same layout as dsdneo_b200_p25p1_frame / dsdneo_b200_p25p1_voice of include/dsdneo_b200.h */
typedef struct oracle_p25p1_frame {
    int64_t position;
    int32_t channel;
    int32_t voice_index;
    int16_t nac;
    int16_t nid_errs;
    int8_t nid_status;
    uint8_t duid;
    uint8_t n_tsbk;
    uint8_t tsbk_crc_ok;
    uint8_t rs_kind;
    uint8_t rs_status;
    uint8_t lsd_ok;
    uint8_t n_word_soft;
    uint8_t lsd[2];
    uint8_t reserved[6];
    uint8_t tsbk[3][12];
    uint8_t rs_data[20];
    uint8_t rs_in_data[20];
    uint8_t rs_in_parity[16];
} oracle_p25p1_frame;
typedef struct oracle_p25p1_voice {
    uint32_t bits[9][8];
    uint8_t reliab[9][8][23];
} oracle_p25p1_voice;
int oracle_p25p1_decode_frame(const uint8_t* dibits, const int16_t* llr, int count, int pos_last_sync, int observed_nac, int threshold,
                              oracle_p25p1_frame* f, oracle_p25p1_voice* voice);


/* ---- vocoder frame ECC + DMR voice burst cutter (oracle_mbe.c) -- PARITY UNPINNED for the ECC, see its header ---- */
int oracle_mbe_golay2312(const uint8_t* in23, uint8_t* out23);
unsigned oracle_mbe_golay2312_encode(unsigned data12);
int oracle_mbe_hamming1511(const uint8_t* in15, uint8_t* out15);
unsigned oracle_mbe_hamming1511_encode(unsigned data11);
void oracle_ambe3600x2450_decode(const uint8_t* ambe_fr, uint8_t* ambe_d, int* c0_errs, int* total_errs);
void oracle_imbe7200x4400_decode(const uint8_t* imbe_fr, uint8_t* imbe_d, int* c0_errs, int* total_errs);
void oracle_dmr_voice_cut(const uint8_t* burst144, int invert_first90, uint8_t* cach24, uint8_t* ambe_fr3, uint8_t* sync48);

#endif
