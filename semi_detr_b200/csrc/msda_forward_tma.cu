// Multi-scale deformable attention forward, TMA-staged variant for encoder self-attention.  sm_100a.
//
// Same operator and same lane ownership as msda_forward.cu (8 lanes per (query, head), float4 of channels each,
// 5 width-8 shuffles per point), but the value lines are not fetched through L1 on demand: for each work item
// (image, head, 8x8-pixel query tile) one elected thread issues ONE `cp.async.bulk.tensor.5d` per feature level
// -- a [B_l x B_l pixels x 32 channels] box of that head's value plane, centred on the tile's footprint at that
// level -- into shared memory, double-buffered behind an mbarrier, while the CTA gathers the previous item out of
// shared memory.  TMA's out-of-bounds zero fill IS the operator's zero padding, so in-box corners need no border
// logic.  A sample whose 2x2 footprint leaves the box (large learned offsets, or a coarse-level query sampling a
// much finer level) takes the same predicated global loads as the L1 kernel; the choice is a pointer select
// (generic loads), not a branch.
//
// Box sides {18, 14, 12, 11} cover offsets of +-4 px (+ bilinear footprint) around an 8x8 level-0 tile at the four
// pyramid levels of every shipped config: 785 pixels x 128 B = 100 KB per stage, 200 KB double-buffered, one
// 512-thread CTA per SM.  Why: the L1 kernel is bound by long-scoreboard stalls on the ~26% of corner reads that
// miss L1 (ncu: l1tex hit 74%, data pipe 63% busy); staging turns every in-box read into a fixed-latency shared
// memory access and moves the L2 traffic into bulk copies that overlap the gather.
#include <cuda.h>

#include <cstring>

#include "msda_common.cuh"

namespace sdb {

constexpr int kTmaThreads = 512;
constexpr int kTmaLevels = 4;
constexpr int kTmaTile = 8;                                   // 8 x 8 queries per item
__host__ __device__ constexpr int box_side(int l) { return l == 0 ? 18 : l == 1 ? 14 : l == 2 ? 12 : 11; }
constexpr int kStageFloats = (18 * 18 + 14 * 14 + 12 * 12 + 11 * 11) * 32;   // 25 120 floats = 100 480 B
constexpr int kStageBytes = kStageFloats * 4;
__host__ __device__ constexpr int box_base(int l) {                     // float offset of level l inside a stage
  return l == 0 ? 0 : l == 1 ? 18 * 18 * 32 : l == 2 ? (18 * 18 + 14 * 14) * 32 : (18 * 18 + 14 * 14 + 12 * 12) * 32;
}

struct TmaMaps {
  CUtensorMap lvl[kTmaLevels];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// Generic-address 16-byte load: the operand may point into the staged box (shared window) or into the image
// (global); spelled in PTX so the compiler cannot specialise the address space from the __restrict__ parameters.
__device__ __forceinline__ float4 ld_generic_f4(const float* p) {
  float4 r;
  asm volatile("ld.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

struct BoxOrigin {
  int x0[kTmaLevels], y0[kTmaLevels];
};

// top-left pixel of level l's box for the tile [tx0, tx0+8) x [ty0, ty0+8) of level lt
__device__ __forceinline__ void box_origin(const LevelTable& lt, int L, int tlvl, int tx0, int ty0, BoxOrigin& o) {
  const float cxn = ((float)tx0 + 0.5f * kTmaTile) / (float)lt.W[tlvl];
  const float cyn = ((float)ty0 + 0.5f * kTmaTile) / (float)lt.H[tlvl];
#pragma unroll
  for (int l = 0; l < kTmaLevels; ++l) {
    if (l < L) {
      const float half = 0.5f * (float)(box_side(l) - 1);
      o.x0[l] = (int)floorf(cxn * (float)lt.W[l] - 0.5f - half + 0.5f);
      o.y0[l] = (int)floorf(cyn * (float)lt.H[l] - 0.5f - half + 0.5f);
    } else {
      o.x0[l] = o.y0[l] = 0;
    }
  }
}

struct TmaPrep {
  int code;                  // < -2^30: INT_MIN + float offset inside the stage (in-box); else float offset in the image
  float w00, w01, w10, w11;  // corner weights * attention weight, 0 where the corner does not contribute
};

__device__ __forceinline__ TmaPrep tma_prep(const LevelTable& lt, const BoxOrigin& bo, int lvl, float x, float y,
                                            float a, int px_stride) {
  const int H = lt.H[lvl], W = lt.W[lvl];
  const Tap<float> t = make_tap<float>(x, y, H, W);
  TmaPrep r;
  const float hh = 1.f - t.lh, hw = 1.f - t.lw;
  r.w00 = t.c00 ? hh * hw * a : 0.f;
  r.w01 = t.c01 ? hh * t.lw * a : 0.f;
  r.w10 = t.c10 ? t.lh * hw * a : 0.f;
  r.w11 = t.c11 ? t.lh * t.lw * a : 0.f;
  const int B = box_side(lvl);
  const int bx = t.w0 - bo.x0[lvl], by = t.h0 - bo.y0[lvl];
  if (t.ok && bx >= 0 && by >= 0 && bx + 1 < B && by + 1 < B)
    r.code = (int)(0x80000000u | (unsigned)(box_base(lvl) + (by * B + bx) * 32));
  else
    r.code = (lt.start[lvl] + t.h0 * W + t.w0) * px_stride;
  return r;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}

// Warp roles: warps 0..15 gather (4 queries of the tile each), warp 16 is the TMA producer.  Stages are handed
// over with two mbarriers each -- full[s] (transaction count: the boxes have landed) and empty[s] (one arrival per
// gather warp: the stage may be overwritten) -- so the gather warps are never synchronised with one another: a warp
// that finishes an item early starts the next one as soon as its boxes are in.
constexpr int kGatherWarps = 16;

struct ItemOperands {      // this lane's two points of one item, loaded one item ahead
  float4 l4;
  float2 a2;
};

template <bool kFused>
__global__ void __launch_bounds__(kTmaThreads + 32, 1)
msda_fwd_tma_kernel(const __grid_constant__ TmaMaps maps, const float* __restrict__ value,
                    const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                    const float* __restrict__ loc, const float* __restrict__ attn, int batch, int S, int L, int Lq,
                    float* __restrict__ out, const float* __restrict__ ref, int ref_dim) {
  constexpr int M = 8, P = 4, px_stride = M * 32;
  extern __shared__ __align__(128) float stage[];           // 2 x kStageFloats
  __shared__ LevelTable lt;
  __shared__ __align__(8) uint64_t full[2], empty[2];
  load_levels<kTmaTile, kTmaTile>(lt, shapes, lsi, L, px_stride);
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty[0], kGatherWarps);
    mbar_init(&empty[1], kGatherWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_tiles = lt.tile_begin[L];
  const long long total = (long long)batch * n_tiles * M;
  const int LP = L * P;
  const int warp = threadIdx.x >> 5;

  auto decode = [&](long long item, int& m, int& n, TileCursor<kTmaTile, kTmaTile>& cur) {
    m = (int)(item % M);
    const long long t2 = item / M;
    n = (int)(t2 / n_tiles);
    cur.seek(lt, L, (int)(t2 % n_tiles), true, Lq);
  };

  if (warp == kGatherWarps) {
    // ===== producer: one lane walks the CTA's items and keeps two stages in flight =====
    if ((threadIdx.x & 31) == 0) {
      int it = 0;
      for (long long item = blockIdx.x; item < total; item += gridDim.x, ++it) {
        const int sidx = it & 1;
        if (it >= 2) mbar_wait(&empty[sidx], (uint32_t)(((it >> 1) - 1) & 1));   // all gather warps released it
        int m, n;
        TileCursor<kTmaTile, kTmaTile> cur;
        decode(item, m, n, cur);
        BoxOrigin bo;
        box_origin(lt, L, cur.lvl, cur.x0, cur.y0, bo);
        float* dst = stage + sidx * kStageFloats;
        uint32_t bytes = 0;
#pragma unroll
        for (int l = 0; l < kTmaLevels; ++l)
          if (l < L) bytes += box_side(l) * box_side(l) * 128;
        mbar_expect_tx(&full[sidx], bytes);
        // constant indices: the descriptors must be addressed in kernel-parameter space, never through a local copy
        tma_load_5d(dst + box_base(0), &maps.lvl[0], &full[sidx], 0, m, bo.x0[0], bo.y0[0], n);
        if (L > 1) tma_load_5d(dst + box_base(1), &maps.lvl[1], &full[sidx], 0, m, bo.x0[1], bo.y0[1], n);
        if (L > 2) tma_load_5d(dst + box_base(2), &maps.lvl[2], &full[sidx], 0, m, bo.x0[2], bo.y0[2], n);
        if (L > 3) tma_load_5d(dst + box_base(3), &maps.lvl[3], &full[sidx], 0, m, bo.x0[3], bo.y0[3], n);
      }
    }
    return;
  }

  // ===== gather warps =====
  const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;     // 64 groups = the 64 queries of a tile
  const int pt = 2 * j;
  const int lvj = min(pt / P, L - 1);

  auto load_operands = [&](long long item, ItemOperands& o) {
    o.l4 = make_float4(0.f, 0.f, 0.f, 0.f);
    o.a2 = make_float2(0.f, 0.f);
    if (item >= total) return;
    int m, n;
    TileCursor<kTmaTile, kTmaTile> cur;
    decode(item, m, n, cur);
    const int q = cur.query(grp, Lq);
    const bool live = q >= 0;
    const long long pair = ((long long)n * Lq + (live ? q : 0)) * M + m;
    if (kFused) {
      const FusedPoints fp = fused_prologue(lt, ref, ref_dim, loc, attn, (long long)n * Lq + (live ? q : 0), pair, L, P,
                                            LP, pt, live);
      o.l4 = fp.loc;
      o.a2 = fp.a;
    } else if (live && pt < LP) {
      o.l4 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 2 * pt));
      o.a2 = ld_stream_f2(reinterpret_cast<const float2*>(attn + pair * LP + pt));
    }
  };

  ItemOperands cur_op, next_op;
  load_operands(blockIdx.x, cur_op);
  int it = 0;
  for (long long item = blockIdx.x; item < total; item += gridDim.x, ++it) {
    const int sidx = it & 1;
    load_operands(item + gridDim.x, next_op);               // next item's locations / weights: latency hidden
    int m, n;
    TileCursor<kTmaTile, kTmaTile> cur;
    decode(item, m, n, cur);
    BoxOrigin bo;
    box_origin(lt, L, cur.lvl, cur.x0, cur.y0, bo);
    const float* vhead = value + (long long)n * S * px_stride + m * 32 + 4 * j;
    const float* sbase = stage + sidx * kStageFloats + 4 * j;
    const uint32_t sbase32 = smem_u32(sbase);
    const int q = cur.query(grp, Lq);
    const bool live = q >= 0;
    const long long pair = ((long long)n * Lq + (live ? q : 0)) * M + m;
    const TmaPrep p0 = tma_prep(lt, bo, lvj, cur_op.l4.x, cur_op.l4.y, cur_op.a2.x, px_stride);
    const TmaPrep p1 = tma_prep(lt, bo, lvj, cur_op.l4.z, cur_op.l4.w, cur_op.a2.y, px_stride);

    mbar_wait(&full[sidx], (uint32_t)((it >> 1) & 1));      // this item's boxes have landed

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 16; s += 2) {
      if (s >= LP) break;                                   // warp-uniform
      const int sl = s >> 1;
      const int lvl = min(s / P, L - 1);                    // both points of the step sit on one level
      const int Bs = box_side(lvl) * 32, ws = lt.wstr[lvl];
      int code[2];
      float w[2][4];
      code[0] = __shfl_sync(0xffffffffu, p0.code, sl, 8);
      code[1] = __shfl_sync(0xffffffffu, p1.code, sl, 8);
      w[0][0] = __shfl_sync(0xffffffffu, p0.w00, sl, 8); w[0][1] = __shfl_sync(0xffffffffu, p0.w01, sl, 8);
      w[0][2] = __shfl_sync(0xffffffffu, p0.w10, sl, 8); w[0][3] = __shfl_sync(0xffffffffu, p0.w11, sl, 8);
      w[1][0] = __shfl_sync(0xffffffffu, p1.w00, sl, 8); w[1][1] = __shfl_sync(0xffffffffu, p1.w01, sl, 8);
      w[1][2] = __shfl_sync(0xffffffffu, p1.w10, sl, 8); w[1][3] = __shfl_sync(0xffffffffu, p1.w11, sl, 8);
      const bool in0 = code[0] < -(1 << 30), in1 = code[1] < -(1 << 30);   // image offsets can be slightly negative
      float4 v[2][4];
      if (__all_sync(0xffffffffu, in0 && in1)) {
        // whole warp inside the staged boxes: plain shared-memory loads with 32-bit addresses
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t p = sbase32 + 4u * (uint32_t)(code[u] & 0x7fffffff);
          v[u][0] = lds_f4(w[u][0] != 0.f ? p : sbase32);
          v[u][1] = lds_f4(w[u][1] != 0.f ? p + 128u : sbase32);
          v[u][2] = lds_f4(w[u][2] != 0.f ? p + 4u * Bs : sbase32);
          v[u][3] = lds_f4(w[u][3] != 0.f ? p + 4u * Bs + 128u : sbase32);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool inbox = u ? in1 : in0;
          // staged box (shared window) or the image (global): one generic pointer per corner, no divergence.  A
          // corner that contributes nothing reads a harmless staged word; the accumulation below skips it.
          const float* p = inbox ? sbase + (code[u] & 0x7fffffff) : vhead + code[u];
          const int dx = 32 * (inbox ? 1 : M);              // next pixel: 128 B in the box, M*128 B in the image
          const int dy = inbox ? Bs : ws;
          v[u][0] = ld_generic_f4(w[u][0] != 0.f ? p : sbase);
          v[u][1] = ld_generic_f4(w[u][1] != 0.f ? p + dx : sbase);
          v[u][2] = ld_generic_f4(w[u][2] != 0.f ? p + dy : sbase);
          v[u][3] = ld_generic_f4(w[u][3] != 0.f ? p + dy + dx : sbase);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (w[u][k] != 0.f) {
            acc.x = fmaf(w[u][k], v[u][k].x, acc.x); acc.y = fmaf(w[u][k], v[u][k].y, acc.y);
            acc.z = fmaf(w[u][k], v[u][k].z, acc.z); acc.w = fmaf(w[u][k], v[u][k].w, acc.w);
          }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[sidx]);  // this warp no longer reads the stage
    if (live) st_stream_f4(reinterpret_cast<float4*>(out + pair * 32 + 4 * j), acc);
    cur_op = next_op;
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

int msda_forward_tma(cudaStream_t st, bool fused, const float* value, const int64_t* shapes_dev,
                     const int64_t* lsi_dev, const int64_t* shapes_host, const int64_t* lsi_host, const float* loc,
                     const float* attn, int batch, int S, int L, int Lq, float* out, const float* ref, int ref_dim) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("msda_forward_tma: cuTensorMapEncodeTiled is not available from this driver");
    return SDB_ERR_CUDA;
  }
  TmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int l = 0; l < kTmaLevels; ++l) {
    const int ll = l < L ? l : L - 1;                       // unused slots repeat the last level (never read)
    const cuuint64_t H = (cuuint64_t)shapes_host[2 * ll], W = (cuuint64_t)shapes_host[2 * ll + 1];
    void* base = (void*)(value + (size_t)lsi_host[ll] * 256);
    const cuuint64_t dims[5] = {32, 8, W, H, (cuuint64_t)batch};
    const cuuint64_t strides[4] = {128, 1024, W * 1024, (cuuint64_t)S * 1024};          // bytes, dims 1..4
    const cuuint32_t box[5] = {32, 1, (cuuint32_t)box_side(l), (cuuint32_t)box_side(l), 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(&maps.lvl[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("msda_forward_tma: cuTensorMapEncodeTiled failed for level %d (CUresult %d, H=%llu W=%llu)", l, (int)r,
                (unsigned long long)H, (unsigned long long)W);
      return SDB_ERR_CUDA;
    }
  }
  const size_t smem = 2 * (size_t)kStageBytes;
  static bool configured[2] = {false, false};
  auto k0 = msda_fwd_tma_kernel<false>;
  auto k1 = msda_fwd_tma_kernel<true>;
  if (!configured[fused]) {
    SDB_CUDA(cudaFuncSetAttribute(fused ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[fused] = true;
  }
  long long n_tiles = 0;
  for (int l = 0; l < L; ++l)
    n_tiles += ((shapes_host[2 * l] + kTmaTile - 1) / kTmaTile) * ((shapes_host[2 * l + 1] + kTmaTile - 1) / kTmaTile);
  long long grid = sm_count();
  const long long items = (long long)batch * n_tiles * 8;
  if (grid > items) grid = items;
  if (grid < 1) grid = 1;
  if (fused)
    k1<<<(unsigned)grid, kTmaThreads + 32, smem, st>>>(maps, value, shapes_dev, lsi_dev, loc, attn, batch, S, L, Lq, out,
                                                 ref, ref_dim);
  else
    k0<<<(unsigned)grid, kTmaThreads + 32, smem, st>>>(maps, value, shapes_dev, lsi_dev, loc, attn, batch, S, L, Lq, out,
                                                 ref, ref_dim);
  SDB_LAUNCH_CHECK("msda_fwd_tma_kernel");
  return SDB_OK;
}

}  // namespace sdb

// value (batch, S, 8, 32) fp32; encoder self-attention only (num_query == spatial_size), levels <= 4, points 4.
// `spatial_shapes_host` / `level_start_host` are HOST copies of the two index tensors (needed to encode the TMA
// descriptors); the device copies are still read by the kernel.  fused != 0: `sampling_loc` / `attn_weight` are the
// raw offsets / logits and (reference_points, ref_dim) the reference points (see sdb_msda_fused_forward_f32).
extern "C" int sdb_msda_forward_tma_f32(sdb_stream_t stream, int fused, const float* value,
                                        const int64_t* spatial_shapes, const int64_t* level_start_index,
                                        const int64_t* spatial_shapes_host, const int64_t* level_start_host,
                                        const float* reference_points, int ref_dim, const float* sampling_loc,
                                        const float* attn_weight, int batch, int spatial_size, int num_heads,
                                        int channels, int num_levels, int num_query, int num_point, float* out) {
  using namespace sdb;
  SDB_REQUIRE(batch >= 0 && spatial_size >= 0 && num_query >= 0, "msda_forward_tma: bad sizes");
  if (!(channels == 32 && num_heads == 8 && num_point == 4 && num_levels >= 1 && num_levels <= kTmaLevels &&
        num_query == spatial_size && (!fused || ref_dim == 2 || ref_dim == 4) &&
        (long long)spatial_size * 256 < (1ll << 30))) {
    set_error("msda_forward_tma: built for encoder self-attention with channels=32, heads=8, points=4, levels<=4 "
              "(got C=%d M=%d P=%d L=%d Lq=%d S=%d)", channels, num_heads, num_point, num_levels, num_query,
              spatial_size);
    return SDB_ERR_UNSUPPORTED;
  }
  if ((long long)batch * num_query == 0) return SDB_OK;
  SDB_REQUIRE(value && spatial_shapes && level_start_index && spatial_shapes_host && level_start_host && sampling_loc &&
              attn_weight && out && (!fused || reference_points), "msda_forward_tma: null pointer");
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(value) & 127) == 0, "msda_forward_tma: value must be 128-byte aligned");
  return msda_forward_tma((cudaStream_t)stream, fused != 0, value, spatial_shapes, level_start_index,
                          spatial_shapes_host, level_start_host, sampling_loc, attn_weight, batch, spatial_size,
                          num_levels, num_query, out, reference_points, ref_dim);
}
