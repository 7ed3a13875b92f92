/* semidetr_b200_debug.h -- instrumentation and microbenchmarks, NOT part of the reference-facing surface.
 *
 * These entry points live in a separate shared library, lib/libsemidetr_b200_debug.so (built by
 * `python -m semi_detr_b200.build --debug`, used only by tools/): the product library libsemidetr_b200.so exports
 * nothing declared here, and its GEMM kernels are compiled without the trace stamps.
 * The debug library carries its own copy of sdb_gemm_tf32 (same source, -DSDB_GEMM_TRACE=1) so a traced launch is
 * the same kernel plus the stamps.
 */
#ifndef SEMIDETR_B200_DEBUG_H_
#define SEMIDETR_B200_DEBUG_H_
#include "semidetr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Debug hook: device buffer of 64 x gridDim.x uint64 that subsequent sdb_gemm_tf32 launches fill with %globaltimer
 * stamps per warp role (tools/trace_gemm.py); NULL switches tracing off. */
int sdb_gemm_tf32_set_trace(unsigned long long* device_buffer);

/* Debug microbenchmark (csrc/umma_rate.cu): SM cycles for `iters` back-to-back tcgen05.mma kind::tf32 128 x n x 8 on
 * fixed operands; mode bit 0: A operand in tensor memory, bit 1: alternate between two accumulators. */
int sdb_debug_umma_rate(sdb_stream_t stream, int n, int mode, int iters, int grid, long long* cycles_out);

/* Debug microbenchmark (csrc/umma_rate.cu): streams a (rows, k) fp32 row-major matrix through every SM's shared memory
 * with TMA boxes of box_rows x 128 bytes x kblocks_per_box k-blocks, boxes_per_stage boxes per stage, a ring of `stages` stages and no consumer
 * work -- the feed rate the GEMM main loop can count on.  cycles_out: SM cycles of CTA 0. */
int sdb_debug_tma_rate(sdb_stream_t stream, const float* x, int rows, int k, int box_rows, int boxes_per_stage,
                       int kblocks_per_box, int stages, int grid, long long* cycles_out);

#ifdef __cplusplus
}
#endif
#endif /* SEMIDETR_B200_DEBUG_H_ */
