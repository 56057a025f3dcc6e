// ljmd_debug.h -- tuning and debug entry points of libljmd.so.  NOT part of the product ABI
// (include/ljmd.h): exported for tools/ and the single-process rank emulation of the tests only;
// they may change or disappear without notice.
#pragma once
#include "ljmd.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Publishes the flag of the last completed step now (normally done by the first CTA of the next
 * step kernel); needed only when ranks are emulated by alternately stepped contexts of ONE process. */
int ljmd_ctx_slab_publish(ljmd_ctx *ctx);
/* Debug aid (neighbour-list path): order-independent XOR checksums of {next positions, v, a,
 * packed next positions} over all slots; equal on every rank right after a gather + rebuild. */
int ljmd_ctx_debug_checksum(ljmd_ctx *ctx, uint64_t *out4);

/* Debug aid (LJMD_SLAB_TIMELINE=1 at context creation): device globaltimer (ns) of {first CTA start, last CTA
 * past the flag wait, last CTA end, flag published} for the sharded steps seq = 0 .. max_steps-1 recorded since
 * the previous call (which re-arms the buffer; call it once before the run of interest).  Returns steps copied. */
int ljmd_ctx_debug_slab_timeline(ljmd_ctx *ctx, uint64_t *out, int32_t max_steps);

/* Tuning aid: per-warp {start, end, smid, items} clock records of the last independent-warp
 * all-pairs launch (context created with LJMD_SYM_DEBUG=1); returns the number of warps. */
int ljmd_ctx_debug_warp_times(ljmd_ctx *ctx, long long *out, int max_warps);

/* Self-benchmark of the symmetric all-pairs inner routine: cycles per (32 x 32*ri)-pair tile
 * divided by 32, for ri i-atoms per lane and rj rotation offsets per step, seen by one warp when
 * warps_per_cta warps share each SM.  Tuning aid (tools/chunk_cycles.py). */
int ljmd_bench_chunk_cycles(int32_t ri, int32_t rj, int32_t dense, int32_t warps_per_cta,
                            double *cycles_per_step);

#ifdef __cplusplus
}
#endif
