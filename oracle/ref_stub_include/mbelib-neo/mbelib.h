/* TEST INFRASTRUCTURE ONLY.  Minimal stand-in for <mbelib-neo/mbelib.h> (mbelib-neo 2.x is an un-vendored dependency of the
 * reference and is not on this machine): only the declarations src/engine/dispatch/dispatch_p25p1.c needs to compile.
 * The reference's own API-contract probe (CMakeLists.txt:622-655) names the same types and functions. */
#ifndef ORACLE_STUB_MBELIB_H
#define ORACLE_STUB_MBELIB_H
struct mbe_parameters;
typedef struct mbe_parameters mbe_parms;
void mbe_initMbeParms(mbe_parms* cur_mp, mbe_parms* prev_mp, mbe_parms* prev_mp_enhanced);
#endif
