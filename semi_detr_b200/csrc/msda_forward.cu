// Multi-scale deformable attention, forward.  sm_100a.
//
// Replaces ms_deformable_im2col_cuda / ms_deformable_im2col_gpu_kernel
// (/root/reference/detr_od/models/utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:923-954, 237-299).
//
// Tuned path (fp32, head dim 32 -- every shipped config):
//  * 8 lanes own one (query, head) pair; lane j holds channels 4j..4j+3, so every bilinear corner is ONE
//    128-byte line read as 8 x LDG.128 (the reference reads it as 32 scalar loads per corner and re-reads
//    the sampling location / weight in all 32 lanes).  A warp works on 4 pairs at a time.
//  * the 8 lanes of a pair load its sampling locations / weights cooperatively (one coalesced float4 +
//    float2 per lane = 2 points), each lane turns its 2 points into (pixel offset, row stride, 4 corner
//    weights pre-multiplied by the attention weight) and the group broadcasts them with width-8 shuffles;
//  * work is a persistent loop over (image, head, query-tile) items.  When num_query == spatial_size
//    (encoder self-attention) tiles are TH x TW pixel blocks of one level, so the value lines a CTA gathers
//    for ONE head (128 B per pixel) stay L1-resident: 16x16 queries touch ~1200 distinct lines (154 KB) for
//    16384 corner reads.  Streamed operands (loc, attn, out) bypass L1 (L1::no_allocate);
//  * `out` is written exactly once with 16-byte stores -- no zero-fill pass.
// Roofline: the gather moves N*Lq*M*L*P*4*128 B through L1 (36 TB/s aggregate) against 4*(value+loc+attn+out)
// bytes of HBM traffic, so the kernel is L1-wavefront bound at ~30% of the HBM roofline.
#include "msda_common.cuh"

namespace sdb {

static int g_fwd_variant = 0;
int g_bwd_variant = 0;

// ------------------------------------------------------------------------------------------------
// generic kernel: any channel count, float or double.  One thread per output element.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const T* __restrict__ loc,
                        const T* __restrict__ attn, long long total, int S, int M, int D, int L, int Lq,
                        int P, T* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % D);
    const long long pair = e / D;
    const int m = (int)(pair % M);
    const long long n = pair / ((long long)M * Lq);
    const T* lp = loc + pair * L * P * 2;
    const T* ap = attn + pair * L * P;
    const T* vb = value + (n * S * M + m) * (long long)D + c;
    const long long px = (long long)M * D;  // elements per pixel
    T acc = 0;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const T* vl = vb + lsi[l] * px;
      for (int p = 0; p < P; ++p) {
        const Tap<T> t = make_tap<T>(lp[0], lp[1], H, W);
        const T a = ap[0];
        lp += 2;
        ap += 1;
        if (!t.ok) continue;
        const T hh = 1 - t.lh, hw = 1 - t.lw;
        const T* v00 = vl + ((long long)t.h0 * W + t.w0) * px;
        T s = 0;
        if (t.c00) s += hh * hw * v00[0];
        if (t.c01) s += hh * t.lw * v00[px];
        if (t.c10) s += t.lh * hw * v00[(long long)W * px];
        if (t.c11) s += t.lh * t.lw * v00[(long long)(W + 1) * px];
        acc += s * a;
      }
    }
    out[e] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// tuned kernel: fp32, D == 32, L*P even.  kHeads / kPoints > 0 bake num_heads / num_point in (8 / 4 in every
// shipped config) so corner strides are immediates and a point's level is a shift; 0 = runtime values.
// ------------------------------------------------------------------------------------------------
struct PointPrep {
  int off;    // float offset of pixel (h0,w0) relative to the image's value base (head / channel offset excluded)
  int wstr;   // float stride between rows of the point's level
  float w00, w01, w10, w11;  // corner weights * attention weight, 0 where the corner does not contribute
};

__device__ __forceinline__ PointPrep prep_point(const LevelTable& lt, int lvl, float x, float y, float a,
                                                int px_stride) {
  const int H = lt.H[lvl], W = lt.W[lvl];
  const Tap<float> t = make_tap<float>(x, y, H, W);
  PointPrep r;
  const float hh = 1.f - t.lh, hw = 1.f - t.lw;
  r.w00 = t.c00 ? hh * hw * a : 0.f;
  r.w01 = t.c01 ? hh * t.lw * a : 0.f;
  r.w10 = t.c10 ? t.lh * hw * a : 0.f;
  r.w11 = t.c11 ? t.lh * t.lw * a : 0.f;
  r.off = (lt.start[lvl] + t.h0 * W + t.w0) * px_stride;
  r.wstr = lt.wstr[lvl];
  return r;
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}

constexpr unsigned kFull = 0xffffffffu;

// kFused: `loc` / `attn` hold the RAW sampling offsets / attention logits and (ref, ref_dim) the reference points;
// locations and softmax weights are formed in registers (fused_prologue).
// V: storage type of `value` and `out` (float, or __nv_bfloat16 for the bf16 configuration; arithmetic is fp32).
template <int kThreads, int TH, int TW, int kMinBlocks, int kHeads, int kPoints, int kStep, bool kFused = false,
          typename V = float>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
msda_fwd_d32_kernel(const V* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                    const float* __restrict__ attn, int batch, int S, int M, int L, int Lq, int P,
                    int tiled, V* __restrict__ out, const float* __restrict__ ref = nullptr, int ref_dim = 0) {
  __shared__ LevelTable lt;
  const int px_stride = kHeads > 0 ? kHeads * 32 : M * 32;
  load_levels<TH, TW>(lt, shapes, lsi, L, px_stride);
  constexpr int TQ = TH * TW;
  constexpr int kGroups = kThreads / 8;
  static_assert(TQ % 4 == 0 && kGroups % 4 == 0, "warp-uniform trip counts need whole warps of groups");
  const int n_tiles = tiled ? lt.tile_begin[L] : (Lq + TQ - 1) / TQ;
  const long long total = (long long)batch * n_tiles * M;
  const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
  const int LP = L * P;

  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long t2 = item / M;
    const int tile = (int)(t2 % n_tiles);
    const int n = (int)(t2 / n_tiles);
    TileCursor<TH, TW> cur;
    cur.seek(lt, L, tile, tiled != 0, Lq);
    const V* vhead = value + (long long)n * S * px_stride + m * 32 + 4 * j;

    // every group of the warp runs the same trip counts: shuffles stay full-mask and convergent
    static_assert(TQ % kGroups == 0, "tile must be a whole number of CTA passes");
#pragma unroll 1
    for (int it = 0; it < TQ / kGroups; ++it) {   // compile-time trip count: provably convergent shuffles
      const int i = grp + it * kGroups;
      const int q = cur.query(i, Lq);
      const bool live = q >= 0;
      const long long pair = ((long long)n * Lq + (live ? q : 0)) * M + m;
      const float* lp = loc + pair * LP * 2;
      const float* ap = attn + pair * LP;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

      for (int c0 = 0; c0 < LP; c0 += 16) {
        const int pt = c0 + 2 * j;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 a2 = make_float2(0.f, 0.f);   // weight 0 -> a dead group / padded point gathers nothing
        if (kFused) {
          const FusedPoints fp = fused_prologue(lt, ref, ref_dim, loc, attn, (long long)n * Lq + (live ? q : 0), pair,
                                                L, P, LP, pt, live);
          l4 = fp.loc;
          a2 = fp.a;
        } else if (live && pt < LP) {
          l4 = ld_stream_f4(reinterpret_cast<const float4*>(lp + 2 * pt));
          a2 = ld_stream_f2(reinterpret_cast<const float2*>(ap + pt));
        }
        int lv0, lv1;
        if constexpr (kPoints > 0) { lv0 = pt / kPoints; lv1 = (pt + 1) / kPoints; }
        else                       { lv0 = pt / P;       lv1 = (pt + 1) / P; }
        lv0 = min(lv0, L - 1);
        lv1 = min(lv1, L - 1);
        const PointPrep p0 = prep_point(lt, lv0, l4.x, l4.y, a2.x, px_stride);
        const PointPrep p1 = prep_point(lt, lv1, l4.z, l4.w, a2.y, px_stride);
        const int npt = min(16, LP - c0);
#pragma unroll
        for (int s = 0; s < 16; s += kStep) {
          if (s >= npt) break;  // warp-uniform (L*P is a multiple of kStep on this path)
          // kStep points per step: all their corner loads are issued before the first use
          int off[kStep], ws[kStep];
          float w[kStep][4];
#pragma unroll
          for (int u = 0; u < kStep; ++u) {
            const int sl = (s + u) >> 1;  // lane of the group that prepared point s+u
            const PointPrep& src = ((s + u) & 1) ? p1 : p0;
            off[u] = __shfl_sync(kFull, src.off, sl, 8);
            w[u][0] = __shfl_sync(kFull, src.w00, sl, 8);
            w[u][1] = __shfl_sync(kFull, src.w01, sl, 8);
            w[u][2] = __shfl_sync(kFull, src.w10, sl, 8);
            w[u][3] = __shfl_sync(kFull, src.w11, sl, 8);
            if constexpr (kPoints > 0 && kPoints % kStep == 0) ws[u] = lt.wstr[min((c0 + s) / kPoints, L - 1)];
            else ws[u] = __shfl_sync(kFull, src.wstr, sl, 8);
          }
          float4 v[kStep][4];
#pragma unroll
          for (int u = 0; u < kStep; ++u) {
            const V* pa = vhead + off[u];
            const V* pa2 = pa + ws[u];
            // a corner with zero weight (outside the map, or a dead / padded point) is neither loaded nor used
            if (w[u][0] != 0.f) v[u][0] = Chan4<V>::gather(pa);
            if (w[u][1] != 0.f) v[u][1] = Chan4<V>::gather(pa + px_stride);
            if (w[u][2] != 0.f) v[u][2] = Chan4<V>::gather(pa2);
            if (w[u][3] != 0.f) v[u][3] = Chan4<V>::gather(pa2 + px_stride);
          }
#pragma unroll
          for (int u = 0; u < kStep; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (w[u][k] != 0.f) fma4(acc, w[u][k], v[u][k]);
        }
      }
      if (live) Chan4<V>::stream_out(out + pair * 32 + 4 * j, acc);
    }
  }
}

template <int kThreads, int TH, int TW, int kMinBlocks, int kHeads, int kPoints, int kStep, bool kFused = false,
          typename V = float>
static int launch_fwd_d32(cudaStream_t st, const V* value, const int64_t* shapes, const int64_t* lsi,
                          const float* loc, const float* attn, int batch, int S, int M, int L, int Lq, int P,
                          V* out, const float* ref = nullptr, int ref_dim = 0) {
  auto kern = msda_fwd_d32_kernel<kThreads, TH, TW, kMinBlocks, kHeads, kPoints, kStep, kFused, V>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    // all of the unified L1/shared array as cache: the kernel's reuse lives in L1
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    int b = 0;
    SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kThreads, 0));
    blocks_per_sm = b > 0 ? b : 1;
  }
  const int tiled = (Lq == S) ? 1 : 0;
  // upper bound on items so tiny problems do not launch idle CTAs
  const long long approx_items = (long long)batch * M * ((Lq + TH * TW - 1) / (TH * TW) + (tiled ? 4 * L : 0));
  long long grid = (long long)sm_count() * blocks_per_sm;
  if (grid > approx_items) grid = approx_items;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, tiled, out, ref,
                                           ref_dim);
  SDB_LAUNCH_CHECK("msda_fwd_d32_kernel");
  return SDB_OK;
}

#define SDB_FWD_ARGS st, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, out

template <typename T>
static int msda_forward(cudaStream_t st, const T* value, const int64_t* shapes, const int64_t* lsi,
                        const T* loc, const T* attn, int batch, int S, int M, int D, int L, int Lq, int P,
                        T* out) {
  SDB_REQUIRE(batch >= 0 && S >= 0 && M > 0 && D > 0 && L > 0 && Lq >= 0 && P > 0,
              "msda_forward: bad sizes batch=%d spatial=%d heads=%d channels=%d levels=%d query=%d point=%d",
              batch, S, M, D, L, Lq, P);
  const long long total = (long long)batch * Lq * M * D;
  if (total == 0) return SDB_OK;
  SDB_REQUIRE(value && shapes && lsi && loc && attn && out, "msda_forward: null pointer");
  if constexpr (sizeof(T) == 4) {
    const bool fits32 = (long long)S * M * D < (1ll << 31);
    const bool fast_ok = D == 32 && ((L * P) % 2 == 0) && L <= kMaxLevels && fits32 &&
                         ((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(loc) |
                           reinterpret_cast<uintptr_t>(attn) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    int v = g_fwd_variant;
    if (v != 9 && fast_ok) {
      if (v == 0) v = 1;  // measured best on B200 (profiles/)
      if (M == 8 && P == 4) {
        switch (v) {
          case 1:
            if (Lq == S) return launch_fwd_d32<256, 8, 8, 3, 8, 4, 2>(SDB_FWD_ARGS);
            return launch_fwd_d32<256, 4, 8, 3, 8, 4, 2>(SDB_FWD_ARGS);
          case 2: return launch_fwd_d32<256, 4, 8, 4, 8, 4, 2>(SDB_FWD_ARGS);
          case 3: return launch_fwd_d32<256, 4, 8, 2, 8, 4, 4>(SDB_FWD_ARGS);
          case 4: return launch_fwd_d32<256, 4, 8, 3, 8, 4, 4>(SDB_FWD_ARGS);
          case 5: return launch_fwd_d32<512, 8, 8, 1, 8, 4, 4>(SDB_FWD_ARGS);
          case 6: return launch_fwd_d32<128, 4, 8, 4, 8, 4, 4>(SDB_FWD_ARGS);
          case 7: return launch_fwd_d32<256, 16, 16, 3, 8, 4, 2>(SDB_FWD_ARGS);
          default: return launch_fwd_d32<256, 4, 8, 3, 0, 0, 2>(SDB_FWD_ARGS);
        }
      }
      return launch_fwd_d32<256, 4, 8, 3, 0, 0, 2>(SDB_FWD_ARGS);
    }
  }
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  msda_fwd_generic_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(value, shapes, lsi, loc, attn, total, S, M, D, L,
                                                               Lq, P, out);
  SDB_LAUNCH_CHECK("msda_fwd_generic_kernel");
  return SDB_OK;
}

}  // namespace sdb

extern "C" int sdb_msda_forward_f32(sdb_stream_t stream, const float* value, const int64_t* spatial_shapes,
                                    const int64_t* level_start_index, const float* sampling_loc,
                                    const float* attn_weight, int batch, int spatial_size, int num_heads,
                                    int channels, int num_levels, int num_query, int num_point, float* out) {
  return sdb::msda_forward<float>((cudaStream_t)stream, value, spatial_shapes, level_start_index, sampling_loc,
                                  attn_weight, batch, spatial_size, num_heads, channels, num_levels, num_query,
                                  num_point, out);
}

extern "C" int sdb_msda_forward_f64(sdb_stream_t stream, const double* value, const int64_t* spatial_shapes,
                                    const int64_t* level_start_index, const double* sampling_loc,
                                    const double* attn_weight, int batch, int spatial_size, int num_heads,
                                    int channels, int num_levels, int num_query, int num_point, double* out) {
  return sdb::msda_forward<double>((cudaStream_t)stream, value, spatial_shapes, level_start_index, sampling_loc,
                                   attn_weight, batch, spatial_size, num_heads, channels, num_levels, num_query,
                                   num_point, out);
}

// bf16 storage for `value` and `out` (BASELINE.json configs[3]: bf16 value / output, fp32 sampling math).  Tuned
// shape only -- the reference op has no bf16 path at all, so there is nothing generic to mirror.
extern "C" int sdb_msda_forward_bf16(sdb_stream_t stream, const uint16_t* value, const int64_t* spatial_shapes,
                                     const int64_t* level_start_index, const float* sampling_loc,
                                     const float* attn_weight, int batch, int spatial_size, int num_heads,
                                     int channels, int num_levels, int num_query, int num_point, uint16_t* out) {
  using namespace sdb;
  const int S = spatial_size, M = num_heads, L = num_levels, Lq = num_query, P = num_point;
  SDB_REQUIRE(batch >= 0 && S >= 0 && M > 0 && channels > 0 && L > 0 && Lq >= 0 && P > 0,
              "msda_forward_bf16: bad sizes batch=%d spatial=%d heads=%d channels=%d levels=%d query=%d point=%d",
              batch, S, M, channels, L, Lq, P);
  if (!(channels == 32 && M == 8 && P == 4 && L <= kMaxLevels && (long long)S * M * channels < (1ll << 31))) {
    set_error("msda_forward_bf16: built for channels=32, heads=8, points=4, levels<=%d (got C=%d M=%d P=%d L=%d)",
              kMaxLevels, channels, M, P, L);
    return SDB_ERR_UNSUPPORTED;
  }
  if ((long long)batch * Lq == 0) return SDB_OK;
  SDB_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
              "msda_forward_bf16: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(out)) & 7) == 0 &&
              ((reinterpret_cast<uintptr_t>(sampling_loc) | reinterpret_cast<uintptr_t>(attn_weight)) & 15) == 0,
              "msda_forward_bf16: value/out must be 8-byte aligned, sampling_loc/attn_weight 16-byte aligned");
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(value);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  cudaStream_t st = (cudaStream_t)stream;
  if (Lq == S)
    return launch_fwd_d32<256, 8, 8, 3, 8, 4, 2, false, __nv_bfloat16>(st, v, spatial_shapes, level_start_index,
                                                                      sampling_loc, attn_weight, batch, S, M, L, Lq,
                                                                      P, o);
  return launch_fwd_d32<256, 4, 8, 3, 8, 4, 2, false, __nv_bfloat16>(st, v, spatial_shapes, level_start_index,
                                                                    sampling_loc, attn_weight, batch, S, M, L, Lq, P,
                                                                    o);
}

extern "C" int sdb_msda_fused_forward_f32(sdb_stream_t stream, const float* value, const int64_t* spatial_shapes,
                                          const int64_t* level_start_index, const float* reference_points,
                                          int ref_dim, const float* sampling_offsets, const float* attn_logits,
                                          int batch, int spatial_size, int num_heads, int channels, int num_levels,
                                          int num_query, int num_point, float* out) {
  using namespace sdb;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = spatial_size, M = num_heads, L = num_levels, Lq = num_query, P = num_point;
  SDB_REQUIRE(batch >= 0 && S >= 0 && Lq >= 0, "msda_fused_forward: bad sizes");
  if (!(channels == 32 && M == 8 && P == 4 && L * P <= 16 && L <= kMaxLevels && (ref_dim == 2 || ref_dim == 4) &&
        (long long)S * M * channels < (1ll << 31))) {
    set_error("msda_fused_forward: built for channels=32, heads=8, points=4, levels*points<=16, ref_dim 2|4 "
              "(got C=%d M=%d P=%d L=%d ref_dim=%d); use the unfused entry point", channels, M, P, L, ref_dim);
    return SDB_ERR_UNSUPPORTED;
  }
  if ((long long)batch * Lq == 0) return SDB_OK;
  SDB_REQUIRE(value && spatial_shapes && level_start_index && reference_points && sampling_offsets && attn_logits &&
              out, "msda_fused_forward: null pointer");
#define SDB_FFWD_ARGS st, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, batch, S, M, L, Lq, P, \
                      out, reference_points, ref_dim
  if (Lq == S) {
    switch (g_fwd_variant) {   // sdb_msda_set_variant: register budget / points in flight (tools/bwd_variants.py)
      case 3: return launch_fwd_d32<256, 8, 8, 2, 8, 4, 4, true>(SDB_FFWD_ARGS);
      case 5: return launch_fwd_d32<512, 8, 8, 1, 8, 4, 4, true>(SDB_FFWD_ARGS);
      case 6: return launch_fwd_d32<128, 8, 8, 4, 8, 4, 4, true>(SDB_FFWD_ARGS);
      case 8: return launch_fwd_d32<256, 8, 8, 2, 8, 4, 2, true>(SDB_FFWD_ARGS);
      default: return launch_fwd_d32<256, 8, 8, 3, 8, 4, 2, true>(SDB_FFWD_ARGS);
    }
  }
  switch (g_fwd_variant) {
    case 3: return launch_fwd_d32<256, 4, 8, 2, 8, 4, 4, true>(SDB_FFWD_ARGS);
    case 6: return launch_fwd_d32<128, 4, 8, 4, 8, 4, 4, true>(SDB_FFWD_ARGS);
    default: return launch_fwd_d32<256, 4, 8, 3, 8, 4, 2, true>(SDB_FFWD_ARGS);
  }
#undef SDB_FFWD_ARGS
}

extern "C" int sdb_msda_set_variant(int forward_variant, int backward_variant) {
  sdb::g_fwd_variant = forward_variant;
  sdb::g_bwd_variant = backward_variant;
  return SDB_OK;
}
