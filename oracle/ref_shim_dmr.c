/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference DMR data-burst reader dmr_data_sync()
 * (src/protocol/dmr/dmr_data.c:320-343) over a recorded dibit stream, to pin the device burst cutter
 * (dsdneo_b200_dmr_burst_cut_batch) and the slot-type / colour-code path behind it:
 *   - the first 90 dibits of a burst (CACH, first info half, slot-type prefix, sync) come from the reference's rolling payload
 *     buffer state->dmr_payload_p / dmr_soft_p (what getFrameSync fills while it hunts): set up here from caller arrays;
 *   - the remaining 54 dibits come from getDibitSoft(), supplied here as a replay (the handler's only sample-side call);
 *   - dmr_data_burst_handler() and dmr_cach() (protocol message parsing) are replaced through the linker's --wrap by recorders of
 *     exactly what dmr_data_sync hands them; Hamming(7,4), Golay(20,8) and the colour-code confidence gate are the real ones;
 *   - everything else (trunking state machine, block reset, debug formatting) is an inert generated stub.
 * Built into oracle/_ref/libdsdneo_ref_dmr.so by oracle/Makefile.  No reference source is copied.
 */
#include <dsd-neo/core/dibit.h>
#include <dsd-neo/core/opts.h>
#include <dsd-neo/core/state.h>
#include <dsd-neo/fec/block_codes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

void dmr_data_sync(dsd_opts* opts, dsd_state* state);
void dmr_confidence_reset(dsd_state* state);

typedef struct ref_dmr_burst {
    int32_t handler_called;  /* dmr_data_burst_handler was reached (slot type ok, confidence gate passed) */
    int32_t cach_called;
    int32_t burst;           /* data type handed to the handler */
    int32_t color_code;      /* state->color_code after the call (slot-type CC) */
    int32_t color_code_ok;
    int32_t dmr_color_code;  /* state->dmr_color_code after the call (16 = unknown; set by the confidence gate) */
    int32_t currentslot;
    int32_t live_dibits;     /* getDibitSoft calls + skipped dibits */
    uint8_t info[196];
    uint8_t rel98[98];
    uint8_t cach[25];
    uint8_t stereo_payload[144];
} ref_dmr_burst;

static struct {
    const uint8_t* dibits;
    const uint8_t* reliab;
    long n, pos;
    ref_dmr_burst* rec;
} g;

int
getDibitSoft(dsd_opts* opts, dsd_state* state, dsd_dibit_soft_t* out_soft) {
    (void)opts;
    (void)state;
    int d = 0;
    if (out_soft) {
        memset(out_soft, 0, sizeof(*out_soft));
    }
    if (g.pos < g.n) {
        d = g.dibits[g.pos] & 3;
        if (out_soft) {
            out_soft->reliability = g.reliab[g.pos];
        }
    }
    g.pos++;
    return d;
}

void
skipDibit(dsd_opts* opts, dsd_state* state, int count) {
    (void)opts;
    (void)state;
    g.pos += count;
}

void
__wrap_dmr_data_burst_handler(dsd_opts* opts, dsd_state* state, uint8_t info[196], uint8_t databurst, const uint8_t* reliab98) {
    (void)opts;
    (void)state;
    if (g.rec) {
        g.rec->handler_called = 1;
        g.rec->burst = databurst;
        memcpy(g.rec->info, info, 196);
        if (reliab98) {
            memcpy(g.rec->rel98, reliab98, 98);
        }
    }
}

uint8_t
__wrap_dmr_cach(dsd_opts* opts, dsd_state* state, uint8_t cach_bits[25]) {
    (void)opts;
    (void)state;
    if (g.rec) {
        g.rec->cach_called = 1;
        memcpy(g.rec->cach, cach_bits, 25);
    }
    return 0;
}

static dsd_opts* s_opts;
static dsd_state* s_state;
static int* s_payload;
static dsd_dibit_soft_t* s_soft;

/* Fresh decoder state (colour-code confidence gate unlocked). */
void
ref_dmr_reset(int inverted_dmr) {
    if (!s_opts) {
        s_opts = (dsd_opts*)calloc(1, sizeof(dsd_opts));
        s_state = (dsd_state*)calloc(1, sizeof(dsd_state));
        s_payload = (int*)calloc(4096, sizeof(int));
        s_soft = (dsd_dibit_soft_t*)calloc(4096, sizeof(dsd_dibit_soft_t));
        InitAllFecFunction();
    }
    memset(s_opts, 0, sizeof(*s_opts));
    memset(s_state, 0, sizeof(*s_state));
    s_opts->inverted_dmr = inverted_dmr ? 1 : 0;
    s_state->dmr_color_code = 16;
    dmr_confidence_reset(s_state);
}

/*
 * One burst: dibits[sync_end - 89 .. sync_end] are the 90 buffered dibits (raw, as the hunt stored them: NOT polarity corrected),
 * dibits[sync_end + 1 ..] what getDibitSoft returns afterwards.  Decoder state (confidence gate) persists across calls until
 * ref_dmr_reset.  Returns the number of live dibits consumed (incl. the skip of the next burst's first part).
 */
long
ref_dmr_data_sync(const uint8_t* dibits, const uint8_t* reliab, long n, long sync_end, ref_dmr_burst* out) {
    if (!s_opts || sync_end < 89 || sync_end >= n) {
        return -1;
    }
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 90; i++) {
        s_payload[1000 + i] = dibits[sync_end - 89 + i] & 3;
        s_soft[1000 + i].reliability = reliab[sync_end - 89 + i];
    }
    s_state->dmr_payload_buf = s_payload;
    s_state->dmr_payload_p = s_payload + 1090;
    s_state->dmr_soft_buf = s_soft;
    s_state->dmr_soft_p = s_soft + 1090;
    s_state->dmr_stereo = 0;
    g.dibits = dibits, g.reliab = reliab, g.n = n, g.pos = sync_end + 1;
    g.rec = out;
    fflush(stderr);
    const int saved = dup(2);
    FILE* devnull = fopen("/dev/null", "w");
    if (devnull) {
        dup2(fileno(devnull), 2);
    }
    dmr_data_sync(s_opts, s_state);
    fflush(stderr);
    if (devnull) {
        dup2(saved, 2);
        fclose(devnull);
    }
    close(saved);
    out->color_code = s_state->color_code;
    out->color_code_ok = s_state->color_code_ok;
    out->dmr_color_code = s_state->dmr_color_code;
    out->currentslot = s_state->currentslot;
    out->live_dibits = (int32_t)(g.pos - (sync_end + 1));
    for (int i = 0; i < 144; i++) {
        out->stereo_payload[i] = (uint8_t)s_state->dmr_stereo_payload[i];
    }
    g.rec = NULL;
    return g.pos - (sync_end + 1);
}

/* ---- voice: the UNMODIFIED dmrBSBootstrap() / dmrBS() (src/protocol/dmr/dmr_bs.c:898-990) on a replayed stream.  processMbeFrame
 * (the vocoder entry, src/core/vocoder/dsd_mbe.c) is replaced through --wrap by a recorder of the ambe_fr[4][24] it is handed;
 * every burst the loop reads is logged through its own dmr_stereo_payload copy. ---- */
void dmrBSBootstrap(dsd_opts* opts, dsd_state* state);
volatile uint8_t exitflag = 0; /* include/dsd-neo/runtime/exitflag.h:28 (read by the loop of dmrBS) */

static struct {
    uint8_t* frames; /* [max][96] */
    long* at;        /* [max]: live dibits consumed when the frame was handed over (names the burst) */
    int max, n;
} v;

void
__wrap_processMbeFrame(dsd_opts* opts, dsd_state* state, char imbe_fr[8][23], char ambe_fr[4][24], char imbe7100_fr[7][24]) {
    (void)opts;
    (void)state;
    (void)imbe_fr;
    (void)imbe7100_fr;
    if (ambe_fr && v.frames && v.n < v.max) {
        memcpy(v.frames + (size_t)v.n * 96, ambe_fr, 96);
        v.at[v.n] = g.pos;
    }
    v.n++;
}

/*
 * dibits[sync_end - 89 .. sync_end]: the 90 buffered dibits of the first voice burst (raw), the rest comes through getDibitSoft.
 * Records up to max_frames ambe_fr arrays in call order (frame_pos[i] = stream position after the burst that carried frame i); returns the number of processMbeFrame calls; *consumed = live dibits read.
 */
int
ref_dmr_voice_run(const uint8_t* dibits, const uint8_t* reliab, long n, long sync_end, uint8_t* frames96, long* frame_pos, int max_frames,
                  long* consumed) {
    if (!s_opts || sync_end < 89 || sync_end >= n) {
        return -1;
    }
    for (int i = 0; i < 90; i++) {
        s_payload[1000 + i] = dibits[sync_end - 89 + i] & 3;
        s_soft[1000 + i].reliability = reliab[sync_end - 89 + i];
    }
    s_state->dmr_payload_buf = s_payload;
    s_state->dmr_payload_p = s_payload + 1090;
    s_state->dmr_soft_buf = s_soft;
    s_state->dmr_soft_p = s_soft + 1090;
    g.dibits = dibits, g.reliab = reliab, g.n = n, g.pos = sync_end + 1;
    g.rec = NULL;
    v.frames = frames96, v.at = frame_pos, v.max = max_frames, v.n = 0;
    fflush(stderr);
    const int saved = dup(2);
    FILE* devnull = fopen("/dev/null", "w");
    if (devnull) {
        dup2(fileno(devnull), 2);
    }
    dmrBSBootstrap(s_opts, s_state);
    fflush(stderr);
    if (devnull) {
        dup2(saved, 2);
        fclose(devnull);
    }
    close(saved);
    if (consumed) {
        *consumed = g.pos - (sync_end + 1);
    }
    v.frames = NULL;
    return v.n;
}

int
ref_dmr_burst_size(void) {
    return (int)sizeof(ref_dmr_burst);
}
