/* Development-only entry points of libni_b200_dev.so (NI_BUILD_TAG=dev python neural_imaging_b200/build.py): hardware probes and the in-kernel
 * role profiler behind DESIGN.md's "measured facts". They are NOT part of the shipping C-ABI (include/ni_b200.h / libni_b200.so). */
#ifndef NI_B200_DEV_H
#define NI_B200_DEV_H
#include "ni_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int ni_tc_selftest(const float* a /*128x32*/, const float* b /*64x32*/, float* d /*128x64*/, int mn_major, ni_stream_t stream);
/* measurement probe: TMA box streaming rate vs channel pitch / boxes in flight; returns the grid size (> 0) or an error (< 0) */
int ni_tma_probe(const float* x, int n, int h, int w, int c, int stages, int boxes_per_cta, long long* cycles_out, int max_grid, ni_stream_t stream);
/* measurement probes (tools/): tcgen05.mma issue / execution rate; 1-D bulk-copy (weight stream) ingest rate per SM */
int ni_mma_probe(int n, int ts, int rounds, int nacc, long long* cycles_out, int grid, ni_stream_t stream);
int ni_bulk_probe(const void* src, long long src_bytes, int bytes, int depth, int copies, int same, int warps, long long* cycles_out, int grid,
                  ni_stream_t stream);
/* debug: per-role clock64 spans of the persistent tcgen05 gemm [0,32) and of the wgrad kernel [32,64) (zeros unless built with -DNI_TC_PROFILE) */
int ni_tc_prof_read(long long* out64, int reset);
#ifdef __cplusplus
}
#endif
#endif
