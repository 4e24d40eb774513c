/* TEST INFRASTRUCTURE ONLY: minimal stand-in for <sndfile.h> (libsndfile is not installed here) so the reference's
 * src/dsp/dsd_symbol.c and src/core/frames/dsd_dibit.c compile unmodified for the oracle build.  None of these
 * functions is reachable on the RTL discriminator path the oracle drives; ref_stubs.c aborts if one is called. */
#ifndef ORACLE_STUB_SNDFILE_H
#define ORACLE_STUB_SNDFILE_H
#include <stdint.h>
typedef struct SNDFILE_tag SNDFILE;
typedef int64_t sf_count_t;
typedef struct {
    sf_count_t frames;
    int samplerate, channels, format, sections, seekable;
} SF_INFO;
int sf_close(SNDFILE* f);
sf_count_t sf_read_short(SNDFILE* f, short* p, sf_count_t n);
sf_count_t sf_write_short(SNDFILE* f, const short* p, sf_count_t n);
void sf_write_sync(SNDFILE* f);
#endif
