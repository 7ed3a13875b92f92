// Token-wise linears of the deformable transformer as tcgen05 tensor-core GEMMs (sm_100a).
//
//   Y[M,N] (row-major) = op(A)[M,K] . op(B)[K,N]  (+ bias[N]) (ReLU) (rows with row_mask != 0 -> 0)
//
// fp32 storage, TF32 tensor-core arithmetic (`tcgen05.mma.cta_group::1.kind::tf32`, fp32 accumulation in TMEM)
// -- the arithmetic the reference's `nn.Linear`s get from cuBLAS with TF32 allowed
// (detr_od/models/utils/ops/modules/ms_deform_attn.py:61-65 value_proj / sampling_offsets / attention_weights /
// output_proj, detr_od/models/utils/transformer.py:626-630 and :878-882 FFN).  One kernel serves the three
// products of a linear layer, selected by operand "majorness":
//   forward   y  = x . W^T     A = x  (M,K) K-major         B = W  (N,K) K-major
//   grad x    dx = dy . W      A = dy (M,K) K-major         B = W  (K,N) MN-major
//   grad W    dW = dy^T . x    A = dy (K,M) MN-major        B = x  (K,N) MN-major     (split along K, reduce-add)
//
// Structure (one persistent CTA per SM, 320 threads -> 10 warps):
//   warp 0   TMA producer: `cp.async.bulk.tensor.2d` boxes with the 128-byte swizzle into a 5-stage ring,
//            completion on `full[s]` mbarriers (transaction bytes)
//   warp 1   MMA issuer: one elected thread, 4 x `tcgen05.mma` (128x128x8) per 32-wide k-block, descriptors built
//            from the stage address; `tcgen05.commit` releases the stage (`empty[s]`) and, after the last k-block,
//            publishes the accumulator (`tmem_full[a]`).  It also owns the TMEM allocation (2 x 128 columns: the
//            epilogue of tile i overlaps the main loop of tile i+1).
//   warps 2-5  epilogue: `tcgen05.ld.32x32b.x32` (each thread owns one accumulator row), bias / ReLU / row mask in
//            registers; every warp drains its own 32 rows through a private, double-buffered, 128B-swizzled 4 KB
//            staging tile and its own `cp.async.bulk.tensor` stores (`cp.reduce.async.bulk.tensor ... .add` when the
//            product is split along K) -- no CTA-wide barrier in the epilogue.  (Measured alternatives: registers ->
//            `st.global.v4` is 1.2-1.4x slower -- 32 distinct lines per store instruction; a CTA-wide 128-row staging
//            tile with two `bar.sync` per chunk is on par.)
//   warps 6-9  operand rounding: the tensor core TRUNCATES fp32 words to TF32 (13 low mantissa bits ignored), a
//            systematic -7e-4 relative shrink of every product.  These warps round each landed stage to nearest
//            (`cvt.rna.tf32.f32`, in place, layout-agnostic) and hand it to the MMA warp through `ready[s]`, so the
//            result is the unbiased round-to-nearest TF32 product.  For an MN-major A (the grad-weight product) the
//            same warps also accumulate its column sums -- the bias gradient -- in registers (`a_column_sums`).
// K-major operands use the 128-byte swizzle; MN-major fp32 operands must use the "128B swizzle with 32-byte atoms"
// (UMMA layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 4 k-rows x 128 B per swizzle atom.
//
// Three variants share the roles: the streaming kernel above; `kBRes` (K <= 256: the CTA's 128 x K block of B stays
// in shared memory for the life of the CTA, only A streams through a 4-stage ring); and `gemm_tf32_wres_kernel`
// (opt-in: the weight block lives in TENSOR memory as the MMA's A operand, see its header).
// In HBM terms the projections are bandwidth-bound (K = 256: 91 MB per 44 446 x 256 x 256 product = 14 us at the
// measured peak vs 5 us of tensor time); measured, the main loop sustains one 128x128x8 MMA per ~185 cycles against
// 97-127 for the bare instruction stream (tools/umma_rate.py), see DESIGN.md section 4.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sdb {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;       // 32 fp32 = 128 B = one swizzle row
constexpr int kStages = 5;
constexpr int kResStages = 4;                       // A-only ring of the B-resident variant
constexpr int kResKB = 8;                           // resident B: up to 8 k-blocks (K <= 256)
constexpr int kTileBytes = kBM * kBK * 4;           // 16 KB per operand per stage
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kOutChunk = 32;                       // output columns per TMA store (128 B rows)
constexpr int kOutBufBytes = kBM * kOutChunk * 4;   // 16 KB
constexpr int kThreads = 320;
constexpr int kXformThreads = 128;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 2 * kBN;
constexpr size_t kDynSmem = (size_t)kStages * kStageBytes + 2 * kOutBufBytes + 1024;   // + alignment slack
constexpr size_t kDynSmemRes = (size_t)(kResKB + kResStages) * kTileBytes + 2 * kOutBufBytes + 1024;

// Per-role %globaltimer stamps exist only in the DEBUG build of this file (-DSDB_GEMM_TRACE=1 ->
// lib/libsemidetr_b200_debug.so, include/semidetr_b200_debug.h); the product library compiles them out.
#ifndef SDB_GEMM_TRACE
#define SDB_GEMM_TRACE 0
#endif
__device__ __forceinline__ void stamp(unsigned long long* trace, int slot) {
#if SDB_GEMM_TRACE
  if (trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    trace[(size_t)blockIdx.x * 64 + slot] = t;
    asm volatile("" ::: "memory");
  }
#else
  (void)trace; (void)slot;
#endif
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug becomes a trap (an error the host sees), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(src), "r"(x), "r"(y)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.  Fields in 16-byte units:
//   [0,14) start address   [16,30) leading byte offset   [32,46) stride byte offset   [46,48) version = 1
//   [61,64) layout type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
// K-major operand (rows of 32 fp32 along K), SWIZZLE_128B: 8-row groups are 1024 B apart (SBO = 64); LBO unused (1).
// MN-major operand (rows of 32 fp32 along M/N, one row per k), SWIZZLE_128B_BASE32B -- the only layout tcgen05
// accepts for MN-major 32-bit operands: the 128 M/N elements of a tile are four boxes of (kBK rows x 128 B) = 4096 B
// apart (LBO = 256); the swizzle atom holds 4 k-rows, so the two atoms of one MMA (k = 8) are 512 B apart (SBO = 32).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, bool mn_major) {
  const uint64_t lbo = mn_major ? (uint64_t)(kBK * 128 / 16) : 1ull;
  const uint64_t sbo = mn_major ? (uint64_t)(512 / 16) : (uint64_t)(1024 / 16);
  const uint64_t type = mn_major ? 1ull : 2ull;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}

// Instruction descriptor, kind::tf32: c_format F32 (1 @ bit 4), a/b format TF32 (2 @ bits 7 / 10), a_major bit 15,
// b_major bit 16, N >> 3 @ bit 17, M >> 4 @ bit 24.
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

struct GemmArgs {
  int M, N, K;
  int k_splits;           // > 1: every split reduce-adds its partial product into Y (Y must hold the initial value)
  int k_blocks_per_split;
  int relu;
  int round_a, round_b;   // round the operand to nearest TF32 in shared memory before the MMA reads it
  unsigned long long* trace;  // debug: 64 globaltimer stamps per CTA (sdb_gemm_tf32_set_trace), or null
  float* a_colsum;          // (M,) or null: column sums of an MN-major A are ADDED here (bias gradient)
  const float* bias;        // (N,) or null
  const uint8_t* row_mask;  // (M,) or null; nonzero -> the output row is zero
  const float* relu_src;    // (M, N) or null: y is zeroed where relu_src <= 0 (ReLU backward in the grad-input epilogue)
  float* y_colsum;          // (N,) or null: column sums of the finished y are ADDED here (the bias gradient behind it)
};

// ------------------------------------------------------------------------------------------------------------
// Epilogue of the 1-CTA kernels (warps 2..5; warp w drains TMEM lanes 32 (w % 4) .. + 31 = 32 output rows).
// Per 32-column chunk: tcgen05.ld (thread = row) -> bias / ReLU / row mask -> the warp's private 4 KB staging tile
// (128-byte swizzle, two tiles alternate) -> one TMA store.  kEpHmask adds a second pass over the staged tile in the
// ACTIVATION's memory layout (lane -> row (lane >> 3) + 4 i, 16-byte chunk lane & 7: four full 128-byte rows per
// request, where the accumulator's thread-per-row layout would touch 32 lines per request): zero where relu_src <= 0,
// column sums of what is left (two shuffle steps, one red.v4 per 4 columns).  The accumulator is handed back to the
// MMA warp as soon as its last chunk is in registers.
enum { kEpBias = 1, kEpRelu = 2, kEpMask = 4, kEpHmask = 8, kEpGeneric = 16 };

template <int F>
__device__ __forceinline__ void epilogue_tiles(const GemmArgs& g, const CUtensorMap* tmYw, uint32_t tmem_base,
                                               uint32_t out_base, uint32_t tfull0, uint32_t tempty0,
                                               long long num_items, int n_tiles, int warp, int lane) {
  constexpr bool kGen = (F & kEpGeneric) != 0;
  constexpr bool kHm = (F & kEpHmask) != 0;
  constexpr int kChunks = kBN / kOutChunk;
  const int quarter = warp & 3;
  const int row = quarter * 32 + lane;
  const uint32_t wbuf = out_base + (uint32_t)(quarter * 2) * 4096u;     // two 4 KB staging tiles, 1024-byte aligned
  const uint32_t l7 = (uint32_t)lane & 7u, r0 = (uint32_t)lane >> 3;
  const uint32_t srow = wbuf + (uint32_t)lane * 128u + (l7 << 4);       // ^ (q << 4): chunk q of this thread's row
  // second pass: row r0 + 4 i lies at r0 * 128 + i * 512, its chunk l7 at position l7 ^ (r & 7), r & 7 = r0 + 4 (i & 1)
  const uint32_t off_even = r0 * 128u + ((l7 ^ r0) << 4), off_odd = r0 * 128u + ((l7 ^ (r0 + 4u)) << 4);
  const bool relu = kGen ? g.relu != 0 : (F & kEpRelu) != 0;
  int acc = 0, epi_tile = 0;
  uint32_t acc_phase = 0;
  bool prev_odd = false;
  for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
    const int ks = (int)(item % g.k_splits);
    const long long tt = item / g.k_splits;
    const int n0 = (int)(tt % n_tiles) * kBN, m0 = (int)(tt / n_tiles) * kBM;
    const int nchunks = min(kChunks, (g.N - n0 + kOutChunk - 1) / kOutChunk);
    const bool use_mask = kGen ? g.row_mask != nullptr : (F & kEpMask) != 0;
    const bool masked = use_mask && (m0 + row) < g.M && g.row_mask[m0 + row] != 0;
    const bool add_bias = kGen ? (g.bias != nullptr && ks == 0) : (F & kEpBias) != 0;
    // activation block of the ReLU-backward epilogue
    const int rows_live = g.M - (m0 + quarter * 32);
    const bool h_full = rows_live >= 32 && n0 + kBN <= g.N;
    const float* hp0 = nullptr;
    size_t hstep = 0;
    if (kHm) {
      hp0 = g.relu_src + (size_t)(m0 + quarter * 32 + (int)r0) * g.N + n0 + 4 * (int)l7;
      hstep = (size_t)4 * g.N;
    }
    float4 h[2][8];
    auto fetch_h = [&](int c, float4 (&d)[8]) {
      if (h_full) {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = __ldg(reinterpret_cast<const float4*>(hp0 + i * hstep + c * kOutChunk));
      } else {                                   // ragged edge tile: rows past M / columns past N read as h = 0
        const int col = n0 + c * kOutChunk + 4 * (int)l7;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          d[i] = ((int)r0 + 4 * i < rows_live && col < g.N)
                     ? __ldg(reinterpret_cast<const float4*>(hp0 + i * hstep + c * kOutChunk))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (kHm) {
      fetch_h(0, h[0]);                          // does not depend on the product: requested before the wait
      // chunks are consumed ~1 us apart, less than an HBM round trip under load: ask L2 for the NEXT item's block now
      // (this thread's row, four 128-byte lines)
      const long long nxt = item + gridDim.x;
      if (nxt < num_items) {
        const long long ntt = nxt / g.k_splits;
        const int nn0 = (int)(ntt % n_tiles) * kBN, nm = (int)(ntt / n_tiles) * kBM + row;
        if (nm < g.M) {
          const float* pr = g.relu_src + (size_t)nm * g.N + nn0;
#pragma unroll
          for (int j = 0; j < kBN / 32; ++j)
            if (nn0 + 32 * j < g.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + 32 * j));
        }
      }
    }
    mbar_wait(tfull0 + 8u * acc, acc_phase);
    if (warp == 2 && lane == 0 && epi_tile < 6) stamp(g.trace, 32 + 2 * epi_tile);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kBN);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      if (c < nchunks) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(c * kOutChunk), v);
        if (kHm && c + 1 < nchunks) fetch_h(c + 1, h[(c + 1) & 1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == nchunks - 1) {                  // the whole accumulator has been read: back to the MMA warp
          tc_fence_before();
          mbar_arrive(tempty0 + 8u * acc);
        }
        // this chunk's staging tile was last read by the store issued two chunks ago (after a tile with an odd
        // number of chunks: by the most recent one)
        if (lane == 0) {
          if (c == 0 && prev_odd) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncwarp();
        const uint32_t tile0 = wbuf + (uint32_t)(c & 1) * 4096u;
        const uint32_t sb = srow + (uint32_t)(c & 1) * 4096u;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (add_bias)                          // warp-uniform address; columns past N are clipped by the store
            o = __ldg(reinterpret_cast<const float4*>(g.bias + min(n0 + c * kOutChunk + 4 * q, g.N - 4)));
          o.x += __uint_as_float(v[4 * q + 0]);
          o.y += __uint_as_float(v[4 * q + 1]);
          o.z += __uint_as_float(v[4 * q + 2]);
          o.w += __uint_as_float(v[4 * q + 3]);
          if (relu) {
            o.x = fmaxf(o.x, 0.f);
            o.y = fmaxf(o.y, 0.f);
            o.z = fmaxf(o.z, 0.f);
            o.w = fmaxf(o.w, 0.f);
          }
          if (masked) o = make_float4(0.f, 0.f, 0.f, 0.f);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb ^ (uint32_t)(q << 4)), "f"(o.x), "f"(o.y),
                       "f"(o.z), "f"(o.w)
                       : "memory");
        }
        if (kHm) {
          __syncwarp();
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                         : "r"(tile0 + ((i & 1) ? off_odd : off_even) + (uint32_t)(i * 512))
                         : "memory");
          float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4(&hc)[8] = h[c & 1];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[i].x = hc[i].x > 0.f ? x[i].x : 0.f;
            x[i].y = hc[i].y > 0.f ? x[i].y : 0.f;
            x[i].z = hc[i].z > 0.f ? x[i].z : 0.f;
            x[i].w = hc[i].w > 0.f ? x[i].w : 0.f;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile0 + ((i & 1) ? off_odd : off_even) +
                                                                        (uint32_t)(i * 512)),
                         "f"(x[i].x), "f"(x[i].y), "f"(x[i].z), "f"(x[i].w)
                         : "memory");
            cs.x += x[i].x;
            cs.y += x[i].y;
            cs.z += x[i].z;
            cs.w += x[i].w;
          }
          if (g.y_colsum) {                      // rows past M and columns past N staged zeros (zero-filled operands)
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
              cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
              cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
            }
            const int col = n0 + c * kOutChunk + 4 * lane;
            if (lane < 8 && col < g.N && rows_live > 0) red_add_f4(g.y_colsum + col, cs);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (kGen && g.k_splits > 1) tma_reduce_add_2d(tmYw, tile0, n0 + c * kOutChunk, m0 + quarter * 32);
          else tma_store_2d(tmYw, tile0, n0 + c * kOutChunk, m0 + quarter * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    prev_odd = (nchunks & 1) != 0;
    if (warp == 2 && lane == 0 && epi_tile < 6) stamp(g.trace, 33 + 2 * epi_tile);
    ++epi_tile;
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1u;
  }
}

// kBRes ("B resident"): for K <= 256 the whole (128 x K) B tile of one n-block -- the weight matrix of the
// value / offset / attention / output projections -- stays in shared memory for the life of the CTA (128 KB), is
// rounded once, and only A streams through a 4-stage ring.  L2 -> SM traffic drops from (A + B) per tile to A per tile.
template <bool kAMn, bool kBMn, bool kBRes>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmYw, const GemmArgs g) {
  constexpr int kNS = kBRes ? kResStages : kStages;                  // ring depth
  constexpr int kSB = kBRes ? kTileBytes : kStageBytes;              // bytes per ring stage
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * kStages + 6];
  __shared__ uint32_t tmem_slot;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bres = base;                                        // resident B: kResKB k-blocks of 16 KB
  const uint32_t ring = kBRes ? base + kResKB * kTileBytes : base;
  const uint32_t out_base = ring + kNS * kSB;
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto ready = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (3 * kStages + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (3 * kStages + 2 + a); };
  const uint32_t bfull = bar0 + 8u * (3 * kStages + 4), bready = bar0 + 8u * (3 * kStages + 5);
  // which operands the rounding warps touch per ring stage
  const bool round_ring_a = g.round_a != 0, round_ring_b = !kBRes && g.round_b != 0;
  const bool xform = round_ring_a || round_ring_b || (kAMn && g.a_colsum != nullptr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + kBM - 1) / kBM, n_tiles = (g.N + kBN - 1) / kBN;
  const long long num_items = (long long)m_tiles * n_tiles * g.k_splits;
  const int total_kb = (g.K + kBK - 1) / kBK;
  // kBRes: the host sizes the grid as a multiple of n_tiles, so every item of this CTA has the same n-block
  const int res_n0 = (int)(blockIdx.x % n_tiles) * kBN;

  if (threadIdx.x == 0) {
    stamp(g.trace, 0);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
      mbar_init(ready(s), kXformThreads);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), kEpiThreads);
    }
    mbar_init(bfull, 1);
    mbar_init(bready, kXformThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYw) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) stamp(g.trace, 1);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      auto load_b = [&](uint32_t sb, int n0, int kb, uint32_t bar) {
        if (kBMn) {
#pragma unroll
          for (int j = 0; j < kBN / 32; ++j) tma_load_2d(sb + j * (kBK * 128), &tmB, n0 + 32 * j, kb * kBK, bar);
        } else {
          tma_load_2d(sb, &tmB, kb * kBK, n0, bar);
        }
      };
      if (kBRes && blockIdx.x < num_items) {
        mbar_expect_tx(bfull, (uint32_t)(total_kb * kTileBytes));
        for (int kb = 0; kb < total_kb; ++kb) load_b(bres + kb * kTileBytes, res_n0, kb, bfull);
        stamp(g.trace, 2);
      }
      int tile_no = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int ks = (int)(item % g.k_splits);
        const long long t = item / g.k_splits;
        const int n0 = (int)(t % n_tiles) * kBN, m0 = (int)(t / n_tiles) * kBM;
        const int kb0 = ks * g.k_blocks_per_split;
        const int kb1 = min(total_kb, kb0 + g.k_blocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          if (kb == kb0 && tile_no < 6) stamp(g.trace, 3 + 2 * tile_no);
          mbar_expect_tx(full(stage), kSB);
          const uint32_t sa = ring + stage * kSB;
          if (kAMn) {
#pragma unroll
            for (int j = 0; j < kBM / 32; ++j) tma_load_2d(sa + j * (kBK * 128), &tmA, m0 + 32 * j, kb * kBK, full(stage));
          } else {
            tma_load_2d(sa, &tmA, kb * kBK, m0, full(stage));
          }
          if (!kBRes) load_b(sa + kTileBytes, n0, kb, full(stage));
          if (++stage == kNS) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (tile_no < 6) stamp(g.trace, 4 + 2 * tile_no);
        ++tile_no;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = make_idesc(kAMn, kBMn);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      if (kBRes && blockIdx.x < num_items) mbar_wait(g.round_b ? bready : bfull, 0);
      stamp(g.trace, 16);
      int tile_no = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int ks = (int)(item % g.k_splits);
        const int kb0 = ks * g.k_blocks_per_split;
        const int kb1 = min(total_kb, kb0 + g.k_blocks_per_split);
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(xform ? ready(stage) : full(stage), phase);
          if (kb == kb0 && tile_no < 6) stamp(g.trace, 17 + 2 * tile_no);
          tc_fence_after();
          const uint32_t sa = ring + stage * kSB;
          const uint32_t sb = kBRes ? bres + kb * kTileBytes : sa + kTileBytes;
          const uint64_t adesc = make_desc(sa, kAMn), bdesc = make_desc(sb, kBMn);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            // one MMA consumes 8 k: 32 B along a K-major row, or 8 rows (1024 B) of an MN-major box
            const uint64_t ao = (uint64_t)((kAMn ? k * 1024 : k * 32) >> 4);
            const uint64_t bo = (uint64_t)((kBMn ? k * 1024 : k * 32) >> 4);
            tc_mma_tf32(d_tmem, adesc + ao, bdesc + bo, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(empty(stage));
          if (++stage == kNS) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull(acc));
        if (tile_no < 6) stamp(g.trace, 18 + 2 * tile_no);
        ++tile_no;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 6) {
    // ---------------- operand rounding (warps 6..9) ----------------
    const int t = threadIdx.x - 192;            // 0..127
    auto round_region = [&](uint32_t p0, int n_vec) {   // n_vec x (128 threads x 16 B), in place
#pragma unroll 4
      for (int j = 0; j < n_vec; ++j) {
        const uint32_t p = p0 + 16u * t + (uint32_t)(j * 16 * kXformThreads);
        uint32_t a, b, c, d;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(p) : "memory");
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(a)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(__uint_as_float(b)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(__uint_as_float(c)));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(d)));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    };
    constexpr int kVecPerTile = kTileBytes / (16 * kXformThreads);   // 8
    if (kBRes && g.round_b && blockIdx.x < num_items) {
      mbar_wait(bfull, 0);
      if (t == 0) stamp(g.trace, 48);
      round_region(bres, total_kb * kVecPerTile);
      mbar_arrive(bready);
      if (t == 0) stamp(g.trace, 49);
    }
    if (xform) {
      // Column sums of an MN-major A (the bias gradient dy^T.1 of the grad-weight product): these warps read every
      // element of A anyway.  Thread t always lands on the same physical 16-byte slot of a 128-byte k-row and on
      // rows with the same (row % 4), so under the 32-byte-atom swizzle (32 B chunk index ^= row % 4) it always sees
      // the same four m-columns of each of the tile's four 32-column boxes: 16 register accumulators per thread,
      // folded over the warp with two shuffles each once per work item and added to global memory (one 16-byte
      // reduction per m-group and warp) by the n-block-0 items.
      const bool do_colsum = kAMn && g.a_colsum != nullptr;
      const int m_in_box = ((((t & 7) >> 1) ^ ((t >> 3) & 3)) << 3) + ((t & 1) << 2);
      float cs[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) cs[i][e] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int ks = (int)(item % g.k_splits);
        const long long tt = item / g.k_splits;
        const int n0 = (int)(tt % n_tiles) * kBN, m0 = (int)(tt / n_tiles) * kBM;
        const int kb0 = ks * g.k_blocks_per_split;
        const int kb1 = min(total_kb, kb0 + g.k_blocks_per_split);
        const bool sum_item = do_colsum && n0 == 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full(stage), phase);
          const uint32_t sa = ring + stage * kSB;
          if (round_ring_a || sum_item) {
#pragma unroll
            for (int j = 0; j < kVecPerTile; ++j) {
              const uint32_t p = sa + 16u * t + (uint32_t)(j * 16 * kXformThreads);
              uint32_t a, b, c, d;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(p) : "memory");
              if (sum_item) {            // row = t / 8 + 16 j -> box j / 2
                cs[j >> 1][0] += __uint_as_float(a);
                cs[j >> 1][1] += __uint_as_float(b);
                cs[j >> 1][2] += __uint_as_float(c);
                cs[j >> 1][3] += __uint_as_float(d);
              }
              if (round_ring_a) {
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(a)));
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(__uint_as_float(b)));
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(__uint_as_float(c)));
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(d)));
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
              }
            }
          }
          if (round_ring_b) round_region(sa + kTileBytes, kVecPerTile);
          else if (round_ring_a) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(ready(stage));
          if (++stage == kNS) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (sum_item) {
          // The four lanes of a warp that saw the same m-columns differ in (row % 4) = lane >> 3 and, through the
          // swizzle, in their physical chunk: lane ^ 10 and lane ^ 20 walk exactly that set.
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = cs[i][e];
              v += __shfl_xor_sync(0xffffffffu, v, 10);
              v += __shfl_xor_sync(0xffffffffu, v, 20);
              cs[i][e] = v;
            }
            const int m = m0 + 32 * i + m_in_box;
            if ((t & 31) < 8 && m < g.M)          // one writer per m-group and warp; M % 4 == 0
              red_add_f4(g.a_colsum + m, make_float4(cs[i][0], cs[i][1], cs[i][2], cs[i][3]));
#pragma unroll
            for (int e = 0; e < 4; ++e) cs[i][e] = 0.f;
          }
        }
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    // One instantiation per epilogue the step uses, chosen once per kernel: a single warp drains 32 rows, so the
    // epilogue's rate is its instruction count (ncu source view, docs/ROUND2_NOTES.md section 2) and every runtime
    // flag tested per element costs all of them.
    const uint32_t tf0 = tfull(0), te0 = tempty(0);
    const bool one = g.k_splits == 1;
    if (one && g.relu_src)
      epilogue_tiles<kEpHmask>(g, &tmYw, tmem_base, out_base, tf0, te0, num_items, n_tiles, warp, lane);
    else if (one && !g.bias && !g.relu && !g.row_mask)
      epilogue_tiles<0>(g, &tmYw, tmem_base, out_base, tf0, te0, num_items, n_tiles, warp, lane);
    else if (one && g.bias && g.relu && !g.row_mask)
      epilogue_tiles<kEpBias | kEpRelu>(g, &tmYw, tmem_base, out_base, tf0, te0, num_items, n_tiles, warp, lane);
    else if (one && g.bias && !g.relu && !g.row_mask)
      epilogue_tiles<kEpBias>(g, &tmYw, tmem_base, out_base, tf0, te0, num_items, n_tiles, warp, lane);
    else if (one && g.bias && !g.relu && g.row_mask)
      epilogue_tiles<kEpBias | kEpMask>(g, &tmYw, tmem_base, out_base, tf0, te0, num_items, n_tiles, warp, lane);
    else
      epilogue_tiles<kEpGeneric>(g, &tmYw, tmem_base, out_base, tf0, te0, num_items, n_tiles, warp, lane);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (warp == 2 && lane == 0) stamp(g.trace, 62);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(g.trace, 63);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// Weights-in-TMEM variant ("wres"): y[M,N] = a[M,K] . op(b)[K,N] for K <= 256 and a K-major a -- the forward and the
// grad-input products of every 256-wide projection (value / offsets / weights / output, enc_output), FFN linear1
// forward and FFN linear2 grad-input.
//
// Measured on the shared-memory-resident variant above (%globaltimer stamps per warp role, tools/trace_gemm.py): the
// main loop is LATENCY-bound -- 128 KB of resident weights leave room for four 16 KB stages, and four stages per
// ~1.6 us memory round trip are ~420 ns per k-block, twice the tensor time.  So the roles are swapped: the product is
// computed transposed, y^T[N,M] = op(b)^T . a^T, with the WEIGHTS as the MMA's A operand held in tensor memory
// (`tcgen05.mma ... [d_tmem], [a_tmem], b_desc`): the CTA's 128 x K weight block is loaded once through registers
// (rounded to TF32 on the way, any source layout), parked in 256 TMEM columns next to the two 128-column
// accumulators -- all 512 columns in use -- and shared memory is left to a 10-deep ring of activation tiles
// (160 KB in flight per SM) plus the epilogue's staging tiles.
// The accumulator is y^T: TMEM lane = output feature, column = token.  Each epilogue thread owns one output feature
// (one bias value), transposes through its warp's private staging tile and the warp stores 32x32 sub-tiles of y by TMA.
constexpr int kWStages = 10;
constexpr int kWEpiBufs = 4;                        // staging tiles per epilogue warp: one per 32-token chunk
constexpr size_t kDynSmemW = (size_t)kWStages * kTileBytes + 4 * kWEpiBufs * 4096 + 1024;

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

struct WresArgs {
  int M, N, K;            // y is (M tokens, N features); contraction K <= 256
  int relu, round_a, round_b;
  int b_mn;               // weights stored (K, N) row-major instead of (N, K)
  const float* b;         // weights
  const float* bias;      // (N,) or null
  const uint8_t* row_mask;  // (M,) or null
  const float* relu_src;    // (M, N) or null: y is zeroed where relu_src <= 0
  float* y_colsum;          // (N,) or null: column sums of the finished y are ADDED here
  unsigned long long* trace;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_wres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmYw,
                      const WresArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * kWStages + 5];
  __shared__ uint32_t tmem_slot;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = base + kWStages * kTileBytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kWStages + s); };
  auto ready = [&](int s) { return bar0 + 8u * (2 * kWStages + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (3 * kWStages + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (3 * kWStages + 2 + a); };
  const uint32_t wready = bar0 + 8u * (3 * kWStages + 4);
  const bool xform = g.round_a != 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + kBM - 1) / kBM, n_tiles = (g.N + kBN - 1) / kBN;
  const long long num_items = (long long)m_tiles * n_tiles;
  const int total_kb = (g.K + kBK - 1) / kBK;
  const int n0 = (int)(blockIdx.x % n_tiles) * kBN;     // grid is a multiple of n_tiles: one n-block per CTA

  if (threadIdx.x == 0) {
    stamp(g.trace, 0);
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
      mbar_init(ready(s), kXformThreads);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), kEpiThreads);
    }
    mbar_init(wready, kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYw) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t w_tmem = tmem_base + 256u;            // columns [256, 512): the weight block, column = k
  if (threadIdx.x == 0) stamp(g.trace, 1);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: activation tiles only ----------------
      int stage = 0, tile_no = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int m0 = (int)(item / n_tiles) * kBM;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          if (kb == 0 && tile_no < 6) stamp(g.trace, 3 + 2 * tile_no);
          mbar_expect_tx(full(stage), kTileBytes);
          tma_load_2d(base + stage * kTileBytes, &tmA, kb * kBK, m0, full(stage));
          if (++stage == kWStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (tile_no < 6) stamp(g.trace, 4 + 2 * tile_no);
        ++tile_no;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer: D[feature, token] += W_tmem[feature, k] . a_smem[token, k]^T ----------------
      const uint32_t idesc = make_idesc(false, false);
      int stage = 0, acc = 0, tile_no = 0;
      uint32_t phase = 0, acc_phase = 0;
      mbar_wait(wready, 0);
      tc_fence_after();
      stamp(g.trace, 16);
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kBN);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(xform ? ready(stage) : full(stage), phase);
          if (kb == 0 && tile_no < 6) stamp(g.trace, 17 + 2 * tile_no);
          tc_fence_after();
          const uint64_t bdesc = make_desc(base + stage * kTileBytes, false);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k)
            tc_mma_tf32_ts(d_tmem, w_tmem + (uint32_t)(kb * kBK + k * 8), bdesc + (uint64_t)((k * 32) >> 4), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(empty(stage));
          if (++stage == kWStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull(acc));
        if (tile_no < 6) stamp(g.trace, 18 + 2 * tile_no);
        ++tile_no;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 6) {
    // ---------------- activation rounding (warps 6..9) ----------------
    if (xform) {
      const int t = threadIdx.x - 192;
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(full(stage), phase);
          const uint32_t sa = base + stage * kTileBytes;
#pragma unroll
          for (int j = 0; j < kTileBytes / (16 * kXformThreads); ++j) {
            const uint32_t p = sa + 16u * t + (uint32_t)(j * 16 * kXformThreads);
            uint32_t a, b, c, d;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(p) : "memory");
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(a)));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(__uint_as_float(b)));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(__uint_as_float(c)));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(d)));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(ready(stage));
          if (++stage == kWStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ---------------- weight load, then epilogue (warps 2..5) ----------------
    const int quarter = warp & 3;
    const int feat = n0 + quarter * 32 + lane;            // output feature owned by this thread (TMEM lane)
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    {
      // This thread's weight row (TMEM lane = feature): element k is b[feat, k] (K-major source) or b[k, feat]
      // (MN-major source), zero past N / K.  The MN-major source is read coalesced as it is; the K-major source is
      // read coalesced row by row into the warp's staging tile and transposed there (33-float pitch).
      const bool live = feat < g.N;
      const uint32_t scratch = out_base + (uint32_t)(quarter * kWEpiBufs) * 4096u;
      for (int c = 0; c < kResKB; ++c) {                  // 32 k per tcgen05.st
        uint32_t v[32];
        if (g.b_mn) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int k = c * 32 + e;
            v[e] = (live && k < g.K) ? __float_as_uint(__ldg(g.b + (size_t)k * g.N + feat)) : 0u;
          }
        } else {
          const int k = c * 32 + lane;
          float w[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {                  // 32 independent coalesced loads in flight
            const int row = n0 + quarter * 32 + i;
            w[i] = (row < g.N && k < g.K) ? __ldg(g.b + (size_t)row * g.K + k) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(scratch + (uint32_t)((i * 33 + lane) * 4)), "f"(w[i]) : "memory");
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v[e]) : "r"(scratch + (uint32_t)((lane * 33 + e) * 4)) : "memory");
          __syncwarp();
        }
        if (g.round_b) {
#pragma unroll
          for (int e = 0; e < 32; ++e) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v[e]) : "f"(__uint_as_float(v[e])));
        }
        tmem_st32(w_tmem + lane_addr + (uint32_t)(c * 32), v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(wready);
      if (warp == 2 && lane == 0) stamp(g.trace, 49);
    }
    const float bias = (g.bias != nullptr && feat < g.N) ? __ldg(g.bias + feat) : 0.f;
    const uint32_t wbuf = out_base + (uint32_t)(quarter * kWEpiBufs) * 4096u;
    // ReLU-backward epilogue: in this kernel's layout (thread = output feature, 32 tokens per chunk) the activation
    // block is read as it lies in memory -- one full 128-byte row per request -- and the column sum of a feature is a
    // sum inside its own thread, carried over every tile of the CTA (one n-block per CTA) and added once at the end.
    const float* hsrc = g.relu_src;
    const int featc = min(feat, g.N - 1);                 // features past N: zero weights -> y = 0, clipped by the store
    float colsum = 0.f;
    int acc = 0, epi_tile = 0;
    uint32_t acc_phase = 0;
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int m0 = (int)(item / n_tiles) * kBM;
      const bool m_full = m0 + kBM <= g.M;
      float hh[2][32];
      auto fetch_h = [&](int c, float (&d)[32]) {
        if (m_full) {
          const float* p = hsrc + (size_t)(m0 + c * 32) * g.N + featc;
#pragma unroll
          for (int e = 0; e < 32; ++e) d[e] = __ldg(p + (size_t)e * g.N);
        } else {                                          // tokens past M: zero activations rows of A -> y = 0
#pragma unroll
          for (int e = 0; e < 32; ++e) d[e] = __ldg(hsrc + (size_t)min(m0 + c * 32 + e, g.M - 1) * g.N + featc);
        }
      };
      if (hsrc) fetch_h(0, hh[0]);
      mbar_wait(tfull(acc), acc_phase);
      if (warp == 2 && lane == 0 && epi_tile < 6) stamp(g.trace, 32 + 2 * epi_tile);
      tc_fence_after();
      // the staging tiles of the previous work item must have been read by their TMA stores
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      const uint32_t taddr = tmem_base + lane_addr + (uint32_t)(acc * kBN);
#pragma unroll
      for (int c = 0; c < kBM / 32; ++c) {                // 32 tokens per chunk
        if (m0 + c * 32 >= g.M) continue;
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        if (hsrc && c + 1 < kBM / 32 && m0 + (c + 1) * 32 < g.M) fetch_h(c + 1, hh[(c + 1) & 1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t mask_bits = 0;
        if (g.row_mask != nullptr) {                      // one byte per token, the same for every lane
          const int tok = m0 + c * 32 + lane;
          mask_bits = __ballot_sync(0xffffffffu, tok < g.M && g.row_mask[tok] != 0);
        }
        const uint32_t sbuf = wbuf + (uint32_t)c * 4096u;  // [32 tokens][32 features] fp32, 128-byte rows
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float x = __uint_as_float(v[e]) + bias;
          if (g.relu) x = fmaxf(x, 0.f);
          if ((mask_bits >> e) & 1u) x = 0.f;
          if (hsrc) {
            x = hh[c & 1][e] > 0.f ? x : 0.f;
            colsum += x;
          }
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbuf + (uint32_t)(e * 128 + lane * 4)), "f"(x) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && n0 + quarter * 32 < g.N) {
          tma_store_2d(&tmYw, sbuf, n0 + quarter * 32, m0 + c * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(tempty(acc));
      if (warp == 2 && lane == 0 && epi_tile < 6) stamp(g.trace, 33 + 2 * epi_tile);
      ++epi_tile;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (g.y_colsum != nullptr && feat < g.N) atomicAdd(g.y_colsum + feat, colsum);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (warp == 2 && lane == 0) stamp(g.trace, 62);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) stamp(g.trace, 63);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair variant (`cta_group::2`): the two CTAs of a cluster compute ONE 256 x 256 output tile.  Each CTA stages
// its own 128 rows of A and its own 128-column half of B (the same 32 KB per stage as the 1-CTA kernel), and one
// thread of the leader CTA issues `tcgen05.mma.cta_group::2` 256 x 256 x 8: each SM's tensor core multiplies its 128
// rows by all 256 columns, reading the other half of B from the peer's shared memory -- twice the flops per byte
// staged and per shared-memory byte read, which is what the measurements on the 1-CTA kernels ask for (DESIGN.md
// section 4).  Each CTA's accumulator (128 lanes x 256 columns, double-buffered = all 512 TMEM columns) is drained
// by its own epilogue warps.
// Barrier protocol: `full[s]` / `empty[s]` / `tfull[a]` are local to each CTA (TMA completion; `tcgen05.commit`
// multicast to both CTAs); `ready[s]` and `tempty[a]` live in the LEADER and count the rounding / epilogue threads of
// BOTH CTAs (remote `mbarrier.arrive` through `mapa`), because only the leader issues MMAs.
constexpr int kPairBN = 256;
constexpr size_t kDynSmem2 = (size_t)kStages * kStageBytes + 16 * 4096 + 1024;   // ring + 4 x 4 staging tiles

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 r;\n\tmapa.shared::cluster.u32 r, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool kAMn, bool kBMn>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmYw, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[3 * kStages + 4];
  __shared__ uint32_t tmem_slot;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = base + kStages * kStageBytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto ready = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (3 * kStages + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (3 * kStages + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = (g.M + 2 * kBM - 1) / (2 * kBM), n_tiles = (g.N + kPairBN - 1) / kPairBN;
  const long long num_items = (long long)m_tiles * n_tiles;
  const int total_kb = (g.K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    stamp(g.trace, 0);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
      mbar_init(ready(s), 2 * kXformThreads / 32); // used in the leader only: one arrival per rounding warp of both CTAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), 2 * kEpiThreads / 32);  // used in the leader only: one arrival per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYw) : "memory");
  }
  if (warp == 1) {      // the same warp of both CTAs allocates collectively
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) stamp(g.trace, 1);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: this CTA's 128 rows of A and its 128-column half of B ----------------
      int stage = 0, tile_no = 0;
      uint32_t phase = 0;
      for (long long item = pair; item < num_items; item += n_pairs) {
        const int n0 = (int)(item % n_tiles) * kPairBN + (int)rank * kBN;
        const int m0 = (int)(item / n_tiles) * 2 * kBM + (int)rank * kBM;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          if (tile_no == 0 && kb < 8) stamp(g.trace, 3 + kb);          // first tile: issue time of k-blocks 0..7
          mbar_expect_tx(full(stage), kStageBytes);
          const uint32_t sa = base + stage * kStageBytes, sb = sa + kTileBytes;
          if (kAMn) {
#pragma unroll
            for (int j = 0; j < kBM / 32; ++j) tma_load_2d(sa + j * (kBK * 128), &tmA, m0 + 32 * j, kb * kBK, full(stage));
          } else {
            tma_load_2d(sa, &tmA, kb * kBK, m0, full(stage));
          }
          if (kBMn) {
#pragma unroll
            for (int j = 0; j < kBN / 32; ++j) tma_load_2d(sb + j * (kBK * 128), &tmB, n0 + 32 * j, kb * kBK, full(stage));
          } else {
            tma_load_2d(sb, &tmB, kb * kBK, n0, full(stage));
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ++tile_no;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ---------------- MMA issuer (leader CTA only): 256 x 256 x 8 across the pair ----------------
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)kAMn << 15) | ((uint32_t)kBMn << 16) |
                             ((uint32_t)(kPairBN >> 3) << 17) | ((uint32_t)((2 * kBM) >> 4) << 24);
      int stage = 0, acc = 0, tile_no = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long item = pair; item < num_items; item += n_pairs) {
        mbar_wait_cluster(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kPairBN);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait_cluster(ready(stage), phase);
          if (tile_no == 0 && kb < 8) stamp(g.trace, 17 + kb);         // first tile: k-block kb ready for the MMA
          tc_fence_after();
          const uint32_t sa = base + stage * kStageBytes, sb = sa + kTileBytes;
          const uint64_t adesc = make_desc(sa, kAMn), bdesc = make_desc(sb, kBMn);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t ao = (uint64_t)((kAMn ? k * 1024 : k * 32) >> 4);
            const uint64_t bo = (uint64_t)((kBMn ? k * 1024 : k * 32) >> 4);
            tc_mma_tf32_pair(d_tmem, adesc + ao, bdesc + bo, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(empty(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_pair(tfull(acc));
        if (tile_no < 4) stamp(g.trace, 26 + tile_no);                   // tile committed
        ++tile_no;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 6) {
    // ---------------- operand rounding (warps 6..9): round the local stage, report to the leader ----------------
    const int t = threadIdx.x - 192;
    constexpr int kVecPerTile = kTileBytes / (16 * kXformThreads);
    const uint32_t first = g.round_a ? 0u : (uint32_t)kTileBytes;
    const int n_vec = ((g.round_a ? 1 : 0) + (g.round_b ? 1 : 0)) * kVecPerTile;
    int stage = 0, xf_tile = 0;
    uint32_t phase = 0;
    for (long long item = pair; item < num_items; item += n_pairs, ++xf_tile) {
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(full(stage), phase);
        if (t == 0 && xf_tile == 0 && kb < 8) stamp(g.trace, 40 + kb);   // first tile: k-block kb landed locally
        const uint32_t p0 = base + stage * kStageBytes + first + 16u * t;
#pragma unroll 4
        for (int j = 0; j < n_vec; ++j) {
          const uint32_t p = p0 + (uint32_t)(j * 16 * kXformThreads);
          uint32_t a, b, c, d;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(p) : "memory");
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(a)));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(__uint_as_float(b)));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(__uint_as_float(c)));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(d)));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready(stage), 0);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5): this CTA's 128 rows x 256 columns ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long item = pair; item < num_items; item += n_pairs) {
      const int n0 = (int)(item % n_tiles) * kPairBN;
      const int m0 = (int)(item / n_tiles) * 2 * kBM + (int)rank * kBM;
      const bool masked = g.row_mask != nullptr && (m0 + row) < g.M && g.row_mask[m0 + row] != 0;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kPairBN);
      const bool add_bias = g.bias != nullptr;
      const uint32_t wbuf = out_base + (uint32_t)(quarter * 4) * 4096u;   // four 4 KB staging tiles per warp
      constexpr int kChunks = kPairBN / kOutChunk;
      uint32_t v[2][32];
      tmem_ld32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 1 < kChunks) tmem_ld32(taddr + (uint32_t)((c + 1) * kOutChunk), v[(c + 1) & 1]);   // next chunk in flight
        if (n0 + c * kOutChunk < g.N) {
          // the staging tile written four chunks ago must have been read by its TMA store
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
          __syncwarp();
          const uint32_t sbuf = wbuf + (uint32_t)(c & 3) * 4096u;
          const uint32_t srow = sbuf + (uint32_t)lane * 128u;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            const int col = n0 + c * kOutChunk + 4 * q;
            if (add_bias && col < g.N) o = __ldg(reinterpret_cast<const float4*>(g.bias + col));
            float* op = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(v[c & 1][4 * q + e]) + op[e];
              if (g.relu) x = fmaxf(x, 0.f);
              op[e] = masked ? 0.f : x;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)(((q ^ (lane & 7)) << 4))),
                         "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmYw, sbuf, n0 + c * kOutChunk, m0 + quarter * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty(acc), 0);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  // nobody leaves (or frees tensor memory) while the peer may still read this CTA's shared memory or signal its barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 row-major matrix (rows x cols), box (box_rows x 32 columns), 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap* m, const float* ptr, long long rows, long long cols, int box_rows, bool atom32,
             const char* what, bool no_swizzle = false) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("gemm_tf32: cuTensorMapEncodeTiled is not available from this driver");
    return SDB_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  static int promo = -1;       // SDB_GEMM_L2PROMO=0..3: none / 64 B / 128 B / 256 B (experiments; default 256 B)
  if (promo < 0) {
    const char* e = getenv("SDB_GEMM_L2PROMO");
    promo = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
  }
  const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         no_swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                                    : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                         l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tf32: cuTensorMapEncodeTiled failed for %s (CUresult %d, rows=%lld cols=%lld)", what, (int)r, rows,
              cols);
    return SDB_ERR_CUDA;
  }
  return SDB_OK;
}

template <bool kAMn, bool kBMn, bool kBRes>
int launch(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& yw, const GemmArgs& g,
           long long items, int n_tiles) {
  static bool configured = false;
  auto k = gemm_tf32_kernel<kAMn, kBMn, kBRes>;
  const size_t smem = kBRes ? kDynSmemRes : kDynSmem;
  if (!configured) {
    SDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  long long grid = sm_count();
  if (grid > items) grid = items;
  if (kBRes) grid -= grid % n_tiles;   // every CTA keeps one n-block: items of CTA c are c, c + grid, ...
  k<<<(unsigned)grid, kThreads, smem, st>>>(a, b, yw, g);
  SDB_LAUNCH_CHECK("gemm_tf32_kernel");
  return SDB_OK;
}

}  // namespace
}  // namespace sdb

static unsigned long long* g_gemm_trace = nullptr;

// Debug hook (not part of the reference-facing surface): a device buffer of 64 x gridDim.x uint64 that the next launches
// fill with %globaltimer stamps per warp role (tools/trace_gemm.py prints the timeline); NULL switches it off.
#if SDB_GEMM_TRACE
extern "C" int sdb_gemm_tf32_set_trace(unsigned long long* device_buffer) {
  g_gemm_trace = device_buffer;
  return SDB_OK;
}
#endif

static int gemm_tf32_impl(sdb_stream_t stream, const float* a, int a_mn_major, const float* b, int b_mn_major,
                          float* y, int m, int n, int k, const float* bias, const uint8_t* row_mask, int relu,
                          int k_splits, int round_mode, float* a_column_sums, const float* relu_src,
                          float* y_column_sums) {
  using namespace sdb;
  SDB_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm_tf32: negative size");
  if (m == 0 || n == 0) return SDB_OK;
  SDB_REQUIRE(a && b && y, "gemm_tf32: null pointer");
  SDB_REQUIRE(k > 0, "gemm_tf32: k must be positive");
  SDB_REQUIRE(m % 4 == 0 || !a_mn_major, "gemm_tf32: an MN-major A needs m %% 4 == 0 (16-byte row pitch)");
  SDB_REQUIRE(n % 4 == 0, "gemm_tf32: n must be a multiple of 4 (16-byte row pitch of Y)");
  SDB_REQUIRE(k % 4 == 0 || (a_mn_major && b_mn_major), "gemm_tf32: K-major operands need k %% 4 == 0");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
              "gemm_tf32: operands must be 16-byte aligned");
  SDB_REQUIRE(k_splits >= 1, "gemm_tf32: k_splits must be >= 1");
  SDB_REQUIRE(round_mode >= 0 && round_mode <= 3, "gemm_tf32: round_mode is a bit mask (1 = a, 2 = b)");
  SDB_REQUIRE(!a_column_sums || a_mn_major, "gemm_tf32: column sums are taken of an MN-major a (the grad-weight product)");
  SDB_REQUIRE(k_splits == 1 || (!bias && !relu && !row_mask) , "gemm_tf32: a split product cannot carry an epilogue");
  CUtensorMap ta, tb, tyw;
  int rc;
  if (a_mn_major) rc = make_map(&ta, a, k, m, kBK, true, "A (k,m)");
  else rc = make_map(&ta, a, m, k, kBM, false, "A (m,k)");
  if (rc) return rc;
  if (b_mn_major) rc = make_map(&tb, b, k, n, kBK, true, "B (k,n)");
  else rc = make_map(&tb, b, n, k, kBN, false, "B (n,k)");
  if (rc) return rc;
  rc = make_map(&tyw, y, m, n, 32, false, "Y (m,n), 32-row boxes");
  if (rc) return rc;
  GemmArgs g;
  g.M = m; g.N = n; g.K = k;
  const int total_kb = (k + kBK - 1) / kBK;
  if (k_splits > total_kb) k_splits = total_kb;
  g.k_blocks_per_split = (total_kb + k_splits - 1) / k_splits;
  g.k_splits = (total_kb + g.k_blocks_per_split - 1) / g.k_blocks_per_split;   // no empty split
  g.relu = relu;
  g.round_a = (round_mode & 1) != 0;
  g.round_b = (round_mode & 2) != 0;
  g.bias = bias;
  g.row_mask = row_mask;
  g.a_colsum = a_column_sums;
  g.relu_src = relu_src;
  g.y_colsum = y_column_sums;
  g.trace = g_gemm_trace;
  const long long items = (long long)((m + kBM - 1) / kBM) * ((n + kBN - 1) / kBN) * g.k_splits;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_tiles = (n + kBN - 1) / kBN;
  // CTA-pair variant (cta_group::2, 256 x 256 tiles): opt-in with SDB_GEMM_2CTA=1 until it has been through the full
  // parity suite on hardware; no split-K, no fused column sums.
  static int pair_enabled = -1;
  if (pair_enabled < 0) {
    const char* e = getenv("SDB_GEMM_2CTA");
    pair_enabled = (e && strcmp(e, "1") == 0) ? 1 : 0;
  }
  const bool plain_epilogue = relu_src == nullptr && y_column_sums == nullptr;   // the other variants know no more
  if (pair_enabled && plain_epilogue && g.k_splits == 1 && !a_column_sums && sm_count() >= 2) {
    const long long pair_items = (long long)((m + 2 * kBM - 1) / (2 * kBM)) * ((n + kPairBN - 1) / kPairBN);
    long long pairs = sm_count() / 2;
    if (pairs > pair_items) pairs = pair_items;
    auto launch_pair = [&](auto kern) -> int {
      SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmem2));
      kern<<<(unsigned)(2 * pairs), kThreads, kDynSmem2, st>>>(ta, tb, tyw, g);
      SDB_LAUNCH_CHECK("gemm_tf32_pair_kernel");
      return SDB_OK;
    };
    if (a_mn_major) return b_mn_major ? launch_pair(gemm_tf32_pair_kernel<true, true>) : launch_pair(gemm_tf32_pair_kernel<true, false>);
    return b_mn_major ? launch_pair(gemm_tf32_pair_kernel<false, true>) : launch_pair(gemm_tf32_pair_kernel<false, false>);
  }
  // Weights-in-TMEM variant: K-major streamed operand, K <= 256, no split, enough m-tiles per CTA to amortise the
  // weight load.  Opt-in (SDB_GEMM_WRES=1): parity-green, but measured on par with the shared-memory-resident
  // variant (32 vs 30 us on the projection shape) -- the main loop of both runs at ~185 cycles per 128x128x8 MMA
  // against ~100-125 for the bare instruction stream (tools/umma_rate.py), so the ring depth it buys is not what
  // limits them; see DESIGN.md section 4.
  static int wres_enabled = -1;
  if (wres_enabled < 0) {
    const char* e = getenv("SDB_GEMM_WRES");
    wres_enabled = (e && strcmp(e, "1") == 0) ? 1 : 0;
  }
  // (its ReLU-backward epilogue reads the activation as 32 single-row requests per chunk: 258 us at the encoder FFN
  // shape against 201 us on the shared-memory-resident variant, so that product does not come here by default)
  if (wres_enabled && !a_mn_major && g.k_splits == 1 && total_kb <= kResKB && n_tiles <= sm_count() &&
      items >= 2ll * sm_count() && !a_column_sums) {
    CUtensorMap tyn;
    rc = make_map(&tyn, y, m, n, 32, false, "Y (m,n), 32x32 boxes", true);
    if (rc) return rc;
    WresArgs w;
    w.M = m; w.N = n; w.K = k;
    w.relu = relu; w.round_a = g.round_a; w.round_b = g.round_b;
    w.b_mn = b_mn_major ? 1 : 0;
    w.b = b; w.bias = bias; w.row_mask = row_mask; w.trace = g_gemm_trace;
    w.relu_src = relu_src; w.y_colsum = y_column_sums;
    static bool configured = false;
    if (!configured) {
      SDB_CUDA(cudaFuncSetAttribute(gemm_tf32_wres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemW));
      configured = true;
    }
    long long grid = sm_count();
    if (grid > items) grid = items;
    grid -= grid % n_tiles;
    gemm_tf32_wres_kernel<<<(unsigned)grid, kThreads, kDynSmemW, st>>>(ta, tyn, w);
    SDB_LAUNCH_CHECK("gemm_tf32_wres_kernel");
    return SDB_OK;
  }
  // B-resident variant: the whole K extent of one n-block fits the 128 KB resident region, and there is enough work
  // for every CTA to amortise loading it (at least two m-tiles per CTA on a full grid)
  const bool res = g.k_splits == 1 && total_kb <= kResKB && n_tiles <= sm_count() &&
                   items >= 2ll * sm_count();
  if (res) {
    if (a_mn_major) return b_mn_major ? launch<true, true, true>(st, ta, tb, tyw, g, items, n_tiles) : launch<true, false, true>(st, ta, tb, tyw, g, items, n_tiles);
    return b_mn_major ? launch<false, true, true>(st, ta, tb, tyw, g, items, n_tiles) : launch<false, false, true>(st, ta, tb, tyw, g, items, n_tiles);
  }
  if (a_mn_major) return b_mn_major ? launch<true, true, false>(st, ta, tb, tyw, g, items, n_tiles) : launch<true, false, false>(st, ta, tb, tyw, g, items, n_tiles);
  return b_mn_major ? launch<false, true, false>(st, ta, tb, tyw, g, items, n_tiles) : launch<false, false, false>(st, ta, tb, tyw, g, items, n_tiles);
}

extern "C" int sdb_gemm_tf32(sdb_stream_t stream, const float* a, int a_mn_major, const float* b, int b_mn_major,
                             float* y, int m, int n, int k, const float* bias, const uint8_t* row_mask, int relu,
                             int k_splits, int round_mode, float* a_column_sums) {
  return gemm_tf32_impl(stream, a, a_mn_major, b, b_mn_major, y, m, n, k, bias, row_mask, relu, k_splits, round_mode,
                        a_column_sums, nullptr, nullptr);
}

// Grad-input product of the layer BEHIND a ReLU, with that ReLU's backward and the bias gradient of the layer IN FRONT
// of it in the epilogue: y = (a . op(b)) * (relu_src > 0), y_column_sums[j] = sum_i y[i, j].  In the FFN
// (transformer.py:626-630, 878-882: linear2(dropout(relu(linear1(x))))) a = d(linear2 output), b = linear2.weight,
// relu_src = the hidden activation, y = d(linear1 pre-activation), y_column_sums = d(linear1.bias).
extern "C" int sdb_gemm_tf32_relu_grad(sdb_stream_t stream, const float* a, int a_mn_major, const float* b,
                                       int b_mn_major, float* y, int m, int n, int k, const float* relu_src,
                                       float* y_column_sums, int round_mode) {
  SDB_REQUIRE(relu_src != nullptr, "gemm_tf32_relu_grad: relu_src is null");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(relu_src) | reinterpret_cast<uintptr_t>(y_column_sums)) & 15) == 0,
              "gemm_tf32_relu_grad: relu_src and y_column_sums must be 16-byte aligned");
  if (y_column_sums && n > 0)
    SDB_CUDA(cudaMemsetAsync(y_column_sums, 0, sizeof(float) * (size_t)n, (cudaStream_t)stream));
  return gemm_tf32_impl(stream, a, a_mn_major, b, b_mn_major, y, m, n, k, nullptr, nullptr, 0, 1, round_mode, nullptr,
                        relu_src, y_column_sums);
}
