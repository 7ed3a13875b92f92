// Debug microbenchmark: issue rate of tcgen05.mma kind::tf32 on fixed shared-memory / tensor-memory operands (no
// loads, garbage data), one CTA per SM.  Answers "how long does one 128xNx8 TF32 MMA take back to back?" for the
// GEMM kernel's design (csrc/gemm_tf32.cu).  Not part of the reference-facing surface.
#include <cuda.h>

#include "common.cuh"

namespace sdb {
namespace {

__device__ __forceinline__ uint32_t smem_u32_(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_k128(uint32_t addr) {   // K-major, 128B swizzle (see gemm_tf32.cu)
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// mode bit 0: A from TMEM (TS) instead of shared memory; bit 1: alternate between two accumulators; bit 2: walk the
// B operand through a ring of 8 x 16 KB stages (a new stage every 4 MMAs) instead of re-reading one tile; bit 3:
// commit to an mbarrier after every 4 MMAs (as the GEMM main loop does)
template <int kN>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, int mode, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32_(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32_(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32_(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32_(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand area so the products are finite
  for (int i = threadIdx.x; i < (int)(((mode & 4) ? 9 * 16384 : (128 + kN) * 128) / 16); i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + 16u * i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = desc_k128(base), bdesc = desc_k128(base + 128 * 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (((mode & 2) && (i & 1)) ? (uint32_t)kN : 0u);
      const uint64_t ko = (uint64_t)((((i & 3) * 32) + ((mode & 4) ? ((i >> 2) & 7) * 16384 : 0)) >> 4);
      if (mode & 1) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(d), "r"(tmem + 2u * kN + (uint32_t)((i & 3) * 8)), "l"(bdesc + ko), "r"(idesc), "r"(1u), "r"(0u) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(adesc + (uint64_t)(((i & 3) * 32) >> 4)), "l"(bdesc + ko), "r"(idesc), "r"(1u) : "memory");
      }
      if ((mode & 8) && (i & 3) == 3)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32_(&bar2)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32_(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32_(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles_out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// cta_group::2: the leader of a CTA pair issues 256 x kN x 8 MMAs (each SM: 128 rows x kN columns, half of B from
// the peer's shared memory).  cycles_out[0] = leader cycles for `iters` instructions.
template <int kN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_rate_pair_kernel(int iters, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32_(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32_(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32_(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (128 + kN / 2) * 128 / 16; i += blockDim.x)      // A: 128 rows, B half: kN / 2 rows
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + 16u * i), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t adesc = desc_k128(base), bdesc = desc_k128(base + 128 * 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t ko = (uint64_t)(((i & 3) * 32) >> 4);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc + ko), "l"(bdesc + ko), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32_(&bar)), "h"((uint16_t)1) : "memory");
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32_(&bar)), "r"(0u) : "memory");
      if (!done && spin > (1u << 26)) __trap();
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles_out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- TMA feed-rate microbenchmark -------------------------------------------------------------------------
// What the GEMM traces point at: how fast can ONE SM pull K-major fp32 tiles (boxes of
// `box_rows` rows x 128 bytes, 128B swizzle) through a ring of `stages` stages when nothing consumes them?  A producer
// thread issues `boxes_per_stage` boxes per stage, a consumer thread frees every stage as soon as it has landed.
// Every CTA walks its own row tiles over the whole K extent, like the GEMM's A operand.
__device__ __forceinline__ void spin_wait_(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}

__global__ void __launch_bounds__(64, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int rows_total, int k_total, int box_rows, int boxes_per_stage,
                int kblocks_per_box, int stages, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 16];
  const uint32_t base = (smem_u32_(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32_(bars);
  const int box_bytes = box_rows * 128 * kblocks_per_box;          // smem image of one box: [k-block][row][32 floats]
  const int stage_bytes = boxes_per_stage * box_bytes;
  const int tile_rows = boxes_per_stage * box_rows;
  const int n_tiles = (rows_total + tile_rows - 1) / tile_rows;
  const int kbs = (k_total / 32 + kblocks_per_box - 1) / kblocks_per_box;   // stage loads per row tile
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * (16 + s)));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {                       // producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int kb = 0; kb < kbs; ++kb) {
        spin_wait_(bar0 + 8u * (16 + stage), phase ^ 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * stage), "r"((uint32_t)stage_bytes) : "memory");
        for (int b = 0; b < boxes_per_stage; ++b)      // 3-D map {32 floats, rows, k-blocks}: one box = several k-blocks
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                       ::"r"(base + (uint32_t)(stage * stage_bytes + b * box_bytes)), "l"(&tm), "r"(bar0 + 8u * stage),
                         "r"(0), "r"(tile * tile_rows + b * box_rows), "r"(kb * kblocks_per_box) : "memory");
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
  } else if (threadIdx.x == 32) {               // consumer: free the stage the moment it lands
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int kb = 0; kb < kbs; ++kb) {
        spin_wait_(bar0 + 8u * stage, phase);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8u * (16 + stage)) : "memory");
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles_out[0] = clock64() - t0;
}

}  // namespace
}  // namespace sdb

// Streams a (rows x k) fp32 row-major matrix once through every SM's shared memory with TMA and reports the SM cycles
// CTA 0 needed (the host times the launch with events for GB/s).  box_rows in {8..256}, boxes_per_stage >= 1,
// kblocks_per_box >= 1 (one TMA instruction fetches that many 128-byte k-blocks of every row of the box), stages <= 16,
// ring <= 200 KB, k % 32 == 0.
extern "C" int sdb_debug_tma_rate(sdb_stream_t stream, const float* x, int rows, int k, int box_rows, int boxes_per_stage,
                                  int kblocks_per_box, int stages, int grid, long long* cycles_out) {
  using namespace sdb;
  SDB_REQUIRE(x && cycles_out && rows > 0 && k > 0 && k % 32 == 0 && box_rows >= 8 && box_rows <= 256 &&
              boxes_per_stage >= 1 && kblocks_per_box >= 1 && kblocks_per_box <= 64 && stages >= 1 && stages <= 16 &&
              grid > 0, "debug_tma_rate: bad arguments");
  const size_t ring = (size_t)boxes_per_stage * box_rows * 128 * kblocks_per_box * stages;
  SDB_REQUIRE(ring <= 200 * 1024, "debug_tma_rate: ring larger than 200 KB");
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    set_error("debug_tma_rate: cuTensorMapEncodeTiled is not available");
    return SDB_ERR_CUDA;
  }
  CUtensorMap tm;
  // the matrix seen as {32 floats of one k-block, rows, k-blocks}: a box of kblocks_per_box k-blocks lands as that many
  // consecutive K-major 128B-swizzled tiles
  const cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(k / 32)};
  const cuuint64_t strides[2] = {(cuuint64_t)k * 4, 128};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)kblocks_per_box};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("debug_tma_rate: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return SDB_ERR_CUDA;
  }
  const size_t smem = ring + 1024;
  SDB_CUDA(cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tma_rate_kernel<<<grid, 64, smem, (cudaStream_t)stream>>>(tm, rows, k, box_rows, boxes_per_stage, kblocks_per_box, stages,
                                                             cycles_out);
  SDB_LAUNCH_CHECK("tma_rate_kernel");
  return SDB_OK;
}

// n: 128 or 256 (MMA N); mode bit 0: A from tensor memory, bit 1: two accumulators alternate; grid: CTAs (<= SMs).
// cycles_out (device, int64): SM cycles CTA 0 needed for `iters` back-to-back 128 x n x 8 MMAs.
extern "C" int sdb_debug_umma_rate(sdb_stream_t stream, int n, int mode, int iters, int grid, long long* cycles_out) {
  using namespace sdb;
  SDB_REQUIRE((n == 128 || n == 256) && iters > 0 && grid > 0 && cycles_out, "debug_umma_rate: bad arguments");
  if (mode & 16) {          // bit 4: cta_group::2 (grid = number of CTAs, rounded down to pairs)
    const size_t smem2 = (size_t)(128 + n / 2) * 128 + 1024;
    const unsigned ctas = (unsigned)(grid < 2 ? 2 : grid - grid % 2);
    if (n == 128) {
      SDB_CUDA(cudaFuncSetAttribute(umma_rate_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      umma_rate_pair_kernel<128><<<ctas, 128, smem2, (cudaStream_t)stream>>>(iters, cycles_out);
    } else {
      SDB_CUDA(cudaFuncSetAttribute(umma_rate_pair_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      umma_rate_pair_kernel<256><<<ctas, 128, smem2, (cudaStream_t)stream>>>(iters, cycles_out);
    }
    SDB_LAUNCH_CHECK("umma_rate_pair_kernel");
    return SDB_OK;
  }
  SDB_REQUIRE(!(n == 256 && (mode & 3)), "debug_umma_rate: n = 256 runs with one accumulator and a shared-memory A");
  const size_t smem = ((mode & 4) ? (size_t)9 * 16384 : (size_t)(128 + n) * 128) + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 128) {
    SDB_CUDA(cudaFuncSetAttribute(umma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_rate_kernel<128><<<grid, 128, smem, st>>>(iters, mode, cycles_out);
  } else {
    SDB_CUDA(cudaFuncSetAttribute(umma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_rate_kernel<256><<<grid, 128, smem, st>>>(iters, mode, cycles_out);
  }
  SDB_LAUNCH_CHECK("umma_rate_kernel");
  return SDB_OK;
}
