// Debug microbenchmark: issue rate of tcgen05.mma kind::tf32 on fixed shared-memory / tensor-memory operands (no
// loads, garbage data), one CTA per SM.  Answers "how long does one 128xNx8 TF32 MMA take back to back?" for the
// GEMM kernel's design (csrc/gemm_tf32.cu).  Not part of the reference-facing surface.
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

}  // namespace
}  // namespace sdb

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
