/* SPDX-License-Identifier: GPL-3.0-or-later */
/*
 * TEST INFRASTRUCTURE ONLY.  Link-time stand-ins for reference symbols that the sample-side TUs
 * (src/dsp/dsd_symbol.c, src/core/frames/dsd_dibit.c) reference but that are unreachable on the path the oracle
 * drives (audio_in_type == AUDIO_IN_RTL, FSK-discriminator output kind): audio devices, sockets, WAV files, PCM
 * staging, analog filters.  Same technique as the reference's own tests (tests/dsp/test_rtl_symbol_cache_generation.c
 * defines the same set).  Every stub aborts, so a silently wrong oracle is impossible.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define UNREACHABLE(name)                                                                                               \
    do {                                                                                                                \
        fprintf(stderr, "oracle/ref_stubs.c: %s reached -- the oracle harness left the RTL discriminator path\n", name); \
        abort();                                                                                                        \
    } while (0)

#define STUB_INT(name)                                                                                                 \
    int name() {                                                                                                       \
        UNREACHABLE(#name);                                                                                            \
        return -1;                                                                                                     \
    }
#define STUB_VOID(name)                                                                                                \
    void name() { UNREACHABLE(#name); }

STUB_INT(Connect)
STUB_INT(dsd_audio_read)
STUB_INT(dsd_audio_reconfigure_output_for_input_policy)
STUB_VOID(dsd_audio_rescale_symbol_timing)
STUB_INT(dsd_audio_write)
STUB_INT(dsd_call_state_get)
STUB_INT(dsd_net_audio_input_hook_tcp_close)
STUB_INT(dsd_net_audio_input_hook_tcp_open)
STUB_INT(dsd_net_audio_input_hook_tcp_read_sample)
STUB_INT(dsd_net_audio_input_hook_udp_read_sample)
STUB_INT(dsd_socket_close)
STUB_VOID(dsd_udp_audio_hook_blast_analog)
STUB_INT(openAudioInput)
STUB_INT(sf_close)
STUB_INT(sf_read_short)
STUB_INT(sf_write_short)
STUB_VOID(sf_write_sync)

/* Analog-monitor audio filters: getSymbol() runs them on a COPY of the samples kept for listening to unsynchronised
 * audio (src/dsp/dsd_symbol.c:1041-1057); they never touch the symbol path.  No-ops, exactly like the reference's own
 * test (tests/dsp/test_rtl_symbol_cache_generation.c). */
void agsm_f() {}
void analog_gain_f() {}
/* hpf_f / lpf_f / pbf_f come from the reference's own src/core/util/dsd_misc.c (compiled for its Viterbi decoder) */

void* dsd_fopen_existing_regular_file() { return NULL; }
int dsd_frame_sync_active_nxdn_variant() { return 0; }
void dsd_sleep_ms(unsigned ms) { (void)ms; }
void dsd_sleep_ns(uint64_t ns) { (void)ns; }
uint64_t dsd_time_monotonic_ns(void) { return 0; }

/* the harness checks this flag after every call: set when the reference gave up reading samples */
int g_oracle_shutdown_requested = 0;
void dsd_request_shutdown() { g_oracle_shutdown_requested = 1; }

/* ---- getFrameSync (src/dsp/dsd_frame_sync.c) is compiled in for the acquisition harness (ref_sym_acquire): its UI, event,
 * telemetry and trunking side calls are inert here; the frame-sync hook table (src/runtime/frame_sync_hooks.c) is the
 * reference's own and stays empty, so its hooks are no-ops by the reference's design. ---- */
void dsd_event_sync_slot() {}
const char* dsd_format_local_datetime() { return ""; }
int dsd_telemetry_is_active() { return 0; }
void dsd_telemetry_publish_both_and_redraw() {}
void printFrameInfo() {}
void dsd_mark_cc_sync() {}
uint64_t dsd_time_monotonic_ms(void) { return 0; }
int dsd_rtl_channel_profile_for() { return 0; }
