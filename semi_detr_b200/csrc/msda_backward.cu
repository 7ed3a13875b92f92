// Multi-scale deformable attention, backward.  sm_100a.
//
// Replaces ms_deformable_col2im_cuda and its seven kernel variants
// (/root/reference/detr_od/models/utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:956-1327, 301-920); the
// production one there is ..._shm_blocksize_aware_reduce_v1<float,32> (:301-403): a 32-thread block per
// (query, head), two __syncthreads and a serial 32-element sum by thread 0 for each of the 16 points, and
// 64 scalar atomicAdd per thread.
//
// Tuned path (fp32, head dim 32):
//  * same 8-lanes-per-(query, head) ownership and persistent tiled work list as the forward kernel;
//  * grad_value: one `red.global.add.v4.f32` per lane per corner -- a whole 128-byte line per 8 lanes and
//    4x fewer L2 atomic transactions than scalar atomicAdd;
//  * grad_sampling_loc / grad_attn_weight: per point each lane reduces its 4 channels to four corner dot
//    products; a width-8 reduce-scatter (4 shuffles per point, no shared memory, no barrier) sums them over the
//    head and one lane pair finishes the point, so these two tensors are written exactly once (no zero-fill
//    pass; the reference memsets all three gradients);
//  * grad_value is zero-filled with one cudaMemsetAsync on the same stream.
// Summation order differs from the reference's (fp32 atomics are unordered there too); parity is to 1e-3 rel.
#include "msda_common.cuh"

namespace sdb {

extern int g_bwd_variant;

// encoder self-attention (num_query == spatial_size): tile-combined grad_value (msda_backward_tile.cu), the default
int msda_backward_tile_bf16(cudaStream_t st, const __nv_bfloat16* grad_out, const __nv_bfloat16* value,
                            const int64_t* shapes, const int64_t* lsi, const float* loc, const float* attn, int batch,
                            int S, int L, float* grad_value, float* grad_loc, float* grad_attn);
int msda_backward_tile(cudaStream_t st, const float* grad_out, const float* value, const int64_t* shapes,
                       const int64_t* lsi, const float* loc, const float* attn, int batch, int S, int L,
                       float* grad_value, float* grad_loc, float* grad_attn, const float* ref);

// ------------------------------------------------------------------------------------------------
// generic kernel: any channel count, float or double.  One warp per (n, q, m).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
msda_bwd_generic_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                        const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                        const T* __restrict__ loc, const T* __restrict__ attn, long long pairs, int S, int M,
                        int D, int L, int Lq, int P, T* __restrict__ grad_value, T* __restrict__ grad_loc,
                        T* __restrict__ grad_attn) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long px = (long long)M * D;
  for (long long pair = warp0; pair < pairs; pair += nwarps) {
    const int m = (int)(pair % M);
    const long long n = pair / ((long long)M * Lq);
    const T* lp = loc + pair * L * P * 2;
    const T* ap = attn + pair * L * P;
    T* glp = grad_loc + pair * L * P * 2;
    T* gap = grad_attn + pair * L * P;
    const T* go = grad_out + pair * D;
    const long long img = (n * S * M + m) * (long long)D;
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
      const long long lbase = img + lsi[l] * px;
      for (int p = 0; p < P; ++p) {
        const Tap<T> t = make_tap<T>(lp[0], lp[1], H, W);
        const T a = ap[0];
        T gw = 0, gh = 0, ga = 0;
        if (t.ok) {
          const T hh = 1 - t.lh, hw = 1 - t.lw;
          const long long o00 = lbase + ((long long)t.h0 * W + t.w0) * px;
          const long long o01 = o00 + px, o10 = o00 + (long long)W * px, o11 = o10 + px;
          for (int c = lane; c < D; c += 32) {
            const T g = go[c];
            const T tgv = g * a;
            T v00 = 0, v01 = 0, v10 = 0, v11 = 0;
            if (t.c00) { v00 = value[o00 + c]; atomicAdd(grad_value + o00 + c, hh * hw * tgv); }
            if (t.c01) { v01 = value[o01 + c]; atomicAdd(grad_value + o01 + c, hh * t.lw * tgv); }
            if (t.c10) { v10 = value[o10 + c]; atomicAdd(grad_value + o10 + c, t.lh * hw * tgv); }
            if (t.c11) { v11 = value[o11 + c]; atomicAdd(grad_value + o11 + c, t.lh * t.lw * tgv); }
            const T val = hh * hw * v00 + hh * t.lw * v01 + t.lh * hw * v10 + t.lh * t.lw * v11;
            ga += g * val;
            gh += tgv * (hw * (v10 - v00) + t.lw * (v11 - v01));
            gw += tgv * (hh * (v01 - v00) + t.lh * (v11 - v10));
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gw += __shfl_xor_sync(0xffffffffu, gw, o);
          gh += __shfl_xor_sync(0xffffffffu, gh, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        if (lane == 0) {
          glp[0] = (T)W * gw;
          glp[1] = (T)H * gh;
          gap[0] = ga;
        }
        lp += 2; ap += 1; glp += 2; gap += 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tuned kernel: fp32, D == 32, num_heads == 8, num_point == 4 (every shipped config).
//
// Per point every lane needs only FOUR scalars from its 4 channels: d_k = <grad_out, value_corner_k>.  With
// tgv = a * grad_out the reference's three sums over the 32 channels collapse to
//   grad_attn = sum_k w_k d_k,  grad_x = W a (hh (d01-d00) + lh (d11-d10)),  grad_y = H a (hw (d10-d00) + lw (d11-d01))
// so the cross-lane work is a reduce-scatter of 4 values per point (4 shuffles per point, in batches of 4
// points) instead of the reference's shared-memory staging + serial sum.
// ------------------------------------------------------------------------------------------------
struct BwdPrep {
  int offm;        // float offset of pixel (h0,w0) in the image (multiple of 256) | corner-validity bits 0..3
  float lh, lw, a; // fractional offsets, attention weight (all 0 when the sample is out of range / dead)
};

__device__ __forceinline__ BwdPrep bwd_prep(const LevelTable& lt, int lvl, float x, float y, float a,
                                            int px_stride) {
  const int H = lt.H[lvl], W = lt.W[lvl];
  const Tap<float> t = make_tap<float>(x, y, H, W);
  BwdPrep r;
  r.lh = t.lh;
  r.lw = t.lw;
  r.a = a;   // not masked: an out-of-range sample has no valid corner, so every term it feeds is already zero
  const int mask = (t.c00 ? 1 : 0) | (t.c01 ? 2 : 0) | (t.c10 ? 4 : 0) | (t.c11 ? 8 : 0);
  r.offm = ((lt.start[lvl] + t.h0 * W + t.w0) * px_stride) | mask;
  return r;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

constexpr unsigned kFull = 0xffffffffu;

// kFused: `loc` / `attn` are the RAW sampling offsets / attention logits, (ref, ref_dim) the reference points, and
// the outputs are the gradients w.r.t. those raw tensors (location arithmetic and softmax differentiated here).
// V: storage type of `grad_out` and `value` (float or __nv_bfloat16); every gradient is accumulated and written in fp32.
template <int kThreads, int TH, int TW, int kMinBlocks, bool kFused = false, typename V = float>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
msda_bwd_d32_kernel(const V* __restrict__ grad_out, const V* __restrict__ value,
                    const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                    const float* __restrict__ loc, const float* __restrict__ attn, int batch, int S, int L,
                    int Lq, int tiled, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_attn, const float* __restrict__ ref = nullptr, int ref_dim = 0) {
  constexpr int M = 8, P = 4;
  constexpr int px_stride = M * 32;
  __shared__ LevelTable lt;
  load_levels<TH, TW>(lt, shapes, lsi, L, px_stride);
  constexpr int TQ = TH * TW;
  constexpr int kGroups = kThreads / 8;
  static_assert(TQ % kGroups == 0 && kGroups % 4 == 0, "tile must be a whole number of CTA passes");
  const int n_tiles = tiled ? lt.tile_begin[L] : (Lq + TQ - 1) / TQ;
  const long long total = (long long)batch * n_tiles * M;
  const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
  const bool hi4 = (j & 4) != 0, hi2 = (j & 2) != 0;
  const int LP = L * P;

  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long t2 = item / M;
    const int tile = (int)(t2 % n_tiles);
    const int n = (int)(t2 / n_tiles);
    TileCursor<TH, TW> cur;
    cur.seek(lt, L, tile, tiled != 0, Lq);
    const long long img = (long long)n * S * px_stride + m * 32 + 4 * j;
    const V* vhead = value + img;
    float* gvhead = grad_value + img;

#pragma unroll 1
    for (int it = 0; it < TQ / kGroups; ++it) {   // compile-time trip count: shuffles stay convergent
      const int q = cur.query(grp + it * kGroups, Lq);
      const bool live = q >= 0;
      const long long pair = ((long long)n * Lq + (live ? q : 0)) * M + m;
      const float* lp = loc + pair * LP * 2;
      const float* ap = attn + pair * LP;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) g = Chan4<V>::stream_in(grad_out + pair * 32 + 4 * j);

#pragma unroll 1
      for (int c0 = 0; c0 < LP; c0 += 16) {
        const int pt = c0 + 2 * j;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 a2 = make_float2(0.f, 0.f);
        const long long nq = (long long)n * Lq + (live ? q : 0);
        if (kFused) {
          const FusedPoints fp = fused_prologue(lt, ref, ref_dim, loc, attn, nq, pair, L, P, LP, pt, live);
          l4 = fp.loc;
          a2 = fp.a;
        } else if (live && pt < LP) {
          l4 = ld_stream_f4(reinterpret_cast<const float4*>(lp + 2 * pt));
          a2 = ld_stream_f2(reinterpret_cast<const float2*>(ap + pt));
        }
        float sm_a[4] = {0.f, 0.f, 0.f, 0.f}, sm_g[4] = {0.f, 0.f, 0.f, 0.f};   // fused: softmax backward state
        const int lvj = min(pt / P, L - 1);   // both of this lane's points sit on one level (P = 4)
        const BwdPrep p0 = bwd_prep(lt, lvj, l4.x, l4.y, a2.x, px_stride);
        const BwdPrep p1 = bwd_prep(lt, lvj, l4.z, l4.w, a2.y, px_stride);
        const int nb = min(4, (LP - c0) / 4);
#pragma unroll
        for (int b = 0; b < 4; ++b) {   // batch of 4 points = one level
          if (b >= nb) break;           // warp-uniform
          const int lvl = min((c0 + 4 * b) / P, L - 1);
          const int ws = lt.wstr[lvl];
          float d[4][4];
          float klh = 0.f, klw = 0.f, ka = 0.f;   // the point this lane finalises: batch index j >> 1
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int sl = 2 * b + (r >> 1);
            const BwdPrep& src = (r & 1) ? p1 : p0;
            const int offm = __shfl_sync(kFull, src.offm, sl, 8);
            const float lh = __shfl_sync(kFull, src.lh, sl, 8);
            const float lw = __shfl_sync(kFull, src.lw, sl, 8);
            const float a = __shfl_sync(kFull, src.a, sl, 8);
            if ((j >> 1) == r) { klh = lh; klw = lw; ka = a; }
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float4 tg = make_float4(g.x * a, g.y * a, g.z * a, g.w * a);
            const int off = offm & ~15;
            const V* pv = vhead + off;
            float* pg = gvhead + off;
            float4 v00, v01, v10, v11;
            const bool q00 = offm & 1, q01 = offm & 2, q10 = offm & 4, q11 = offm & 8;
            if (q00) v00 = Chan4<V>::gather(pv);
            if (q01) v01 = Chan4<V>::gather(pv + px_stride);
            if (q10) v10 = Chan4<V>::gather(pv + ws);
            if (q11) v11 = Chan4<V>::gather(pv + ws + px_stride);
            {
              // the reductions need no loaded data: they fill the wait for the four corner loads
              const float w00 = hh * hw, w01 = hh * lw, w10 = lh * hw, w11 = lh * lw;
              red_add_f4_if(q00, pg, w00 * tg.x, w00 * tg.y, w00 * tg.z, w00 * tg.w);
              red_add_f4_if(q01, pg + px_stride, w01 * tg.x, w01 * tg.y, w01 * tg.z, w01 * tg.w);
              red_add_f4_if(q10, pg + ws, w10 * tg.x, w10 * tg.y, w10 * tg.z, w10 * tg.w);
              red_add_f4_if(q11, pg + ws + px_stride, w11 * tg.x, w11 * tg.y, w11 * tg.z, w11 * tg.w);
            }
            d[r][0] = q00 ? dot4(g, v00) : 0.f;
            d[r][1] = q01 ? dot4(g, v01) : 0.f;
            d[r][2] = q10 ? dot4(g, v10) : 0.f;
            d[r][3] = q11 ? dot4(g, v11) : 0.f;
          }
          // reduce-scatter over the 8 lanes: afterwards lanes (j, j^1) hold the sums of point j >> 1
          float e[2][4], f[4];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float keep = hi4 ? d[r + 2][k] : d[r][k];
              const float send = hi4 ? d[r][k] : d[r + 2][k];
              e[r][k] = keep + __shfl_xor_sync(kFull, send, 4, 8);
            }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float keep = hi2 ? e[1][k] : e[0][k];
            const float send = hi2 ? e[0][k] : e[1][k];
            f[k] = keep + __shfl_xor_sync(kFull, send, 2, 8);
            f[k] += __shfl_xor_sync(kFull, f[k], 1, 8);
          }
          const float hh = 1.f - klh, hw = 1.f - klw;
          const int point = c0 + 4 * b + (j >> 1);
          if (live) {
            if ((j & 1) == 0) {
              float gx = (float)lt.W[lvl] * ka * (hh * (f[1] - f[0]) + klh * (f[3] - f[2]));
              float gy = (float)lt.H[lvl] * ka * (hw * (f[2] - f[0]) + klw * (f[3] - f[1]));
              if (kFused) {   // chain through loc = ref + off * (sx, sy)
                if (ref_dim == 2) {
                  gx *= 1.f / (float)lt.W[lvl];
                  gy *= 1.f / (float)lt.H[lvl];
                } else {
                  const float* rp = ref + (nq * L + lvl) * 4;
                  gx *= rp[2] * 0.5f / (float)P;
                  gy *= rp[3] * 0.5f / (float)P;
                }
              }
              st_stream_f2(reinterpret_cast<float2*>(grad_loc + (pair * LP + point) * 2), make_float2(gx, gy));
            } else {
              const float ga = hh * (hw * f[0] + klw * f[1]) + klh * (hw * f[2] + klw * f[3]);
              if (kFused) { sm_a[b] = ka; sm_g[b] = ga; }
              else grad_attn[pair * LP + point] = ga;
            }
          }
        }
        if (kFused) {
          // softmax backward over the pair's L*P points: dlogit_i = a_i * (ga_i - sum_j a_j ga_j)
          const float dotp = group8_sum(sm_a[0] * sm_g[0] + sm_a[1] * sm_g[1] + sm_a[2] * sm_g[2] + sm_a[3] * sm_g[3]);
          if (live && (j & 1)) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (b < nb) grad_attn[pair * LP + c0 + 4 * b + (j >> 1)] = sm_a[b] * (sm_g[b] - dotp);
          }
        }
      }
    }
  }
}

template <int kThreads, int TH, int TW, int kMinBlocks, bool kFused = false, typename V = float>
static int launch_bwd_d32(cudaStream_t st, const V* grad_out, const V* value, const int64_t* shapes,
                          const int64_t* lsi, const float* loc, const float* attn, int batch, int S, int M,
                          int L, int Lq, int P, float* grad_value, float* grad_loc, float* grad_attn,
                          const float* ref = nullptr, int ref_dim = 0) {
  auto kern = msda_bwd_d32_kernel<kThreads, TH, TW, kMinBlocks, kFused, V>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    int b = 0;
    SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kThreads, 0));
    blocks_per_sm = b > 0 ? b : 1;
  }
  const int tiled = (Lq == S) ? 1 : 0;
  const long long approx_items = (long long)batch * M * ((Lq + TH * TW - 1) / (TH * TW) + (tiled ? 4 * L : 0));
  long long grid = (long long)sm_count() * blocks_per_sm;
  if (grid > approx_items) grid = approx_items;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(grad_out, value, shapes, lsi, loc, attn, batch, S, L, Lq, tiled,
                                           grad_value, grad_loc, grad_attn, ref, ref_dim);
  SDB_LAUNCH_CHECK("msda_bwd_d32_kernel");
  return SDB_OK;
}

#define SDB_BWD_ARGS st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn

template <typename T>
static int msda_backward(cudaStream_t st, const T* grad_out, const T* value, const int64_t* shapes,
                         const int64_t* lsi, const T* loc, const T* attn, int batch, int S, int M, int D, int L,
                         int Lq, int P, T* grad_value, T* grad_loc, T* grad_attn) {
  SDB_REQUIRE(batch >= 0 && S >= 0 && M > 0 && D > 0 && L > 0 && Lq >= 0 && P > 0,
              "msda_backward: bad sizes batch=%d spatial=%d heads=%d channels=%d levels=%d query=%d point=%d",
              batch, S, M, D, L, Lq, P);
  const long long nv = (long long)batch * S * M * D;
  const long long pairs = (long long)batch * Lq * M;
  if (nv > 0) {
    SDB_REQUIRE(grad_value, "msda_backward: null grad_value");
    SDB_CUDA(cudaMemsetAsync(grad_value, 0, sizeof(T) * (size_t)nv, st));
  }
  if (pairs == 0) return SDB_OK;
  SDB_REQUIRE(grad_out && value && shapes && lsi && loc && attn && grad_loc && grad_attn,
              "msda_backward: null pointer");
  if constexpr (sizeof(T) == 4) {
    const bool fits32 = (long long)S * M * D < (1ll << 31);
    const uintptr_t al = reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(loc) |
                         reinterpret_cast<uintptr_t>(attn) | reinterpret_cast<uintptr_t>(grad_out) |
                         reinterpret_cast<uintptr_t>(grad_value) | reinterpret_cast<uintptr_t>(grad_loc) |
                         reinterpret_cast<uintptr_t>(grad_attn);
    const bool fast_ok = D == 32 && M == 8 && P == 4 && L <= kMaxLevels && fits32 && (al & 15) == 0;
    int v = g_bwd_variant;
    if (v != 9 && fast_ok) {
      // default: the tile-combining kernel for encoder self-attention (one reduction line per touched pixel and tile
      // instead of one per sampled corner), the 8-lane kernel for everything else (profiles/msda_bwd_variants_r2.txt)
      if ((v == 0 || (v >= 20 && v <= 26)) && Lq == S && L <= 8 && Lq > 0)
        return msda_backward_tile(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                  grad_attn, nullptr);
      if (v == 0 || (v >= 20 && v <= 26)) v = 5;   // spill-free 115-register build
      switch (v) {
        case 1: return launch_bwd_d32<128, 4, 8, 6>(SDB_BWD_ARGS);
        case 2: return launch_bwd_d32<256, 4, 8, 3>(SDB_BWD_ARGS);
        case 3: return launch_bwd_d32<256, 4, 8, 2>(SDB_BWD_ARGS);
        case 4: return launch_bwd_d32<128, 4, 8, 8>(SDB_BWD_ARGS);
        case 5: return launch_bwd_d32<128, 4, 8, 4>(SDB_BWD_ARGS);
        default: return launch_bwd_d32<512, 8, 8, 1>(SDB_BWD_ARGS);
      }
    }
  }
  long long blocks = (pairs + 7) / 8;  // 8 warps per block
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  msda_bwd_generic_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(grad_out, value, shapes, lsi, loc, attn, pairs, S,
                                                               M, D, L, Lq, P, grad_value, grad_loc, grad_attn);
  SDB_LAUNCH_CHECK("msda_bwd_generic_kernel");
  return SDB_OK;
}

}  // namespace sdb

extern "C" int sdb_msda_backward_f32(sdb_stream_t stream, const float* grad_out, const float* value,
                                     const int64_t* spatial_shapes, const int64_t* level_start_index,
                                     const float* sampling_loc, const float* attn_weight, int batch,
                                     int spatial_size, int num_heads, int channels, int num_levels,
                                     int num_query, int num_point, float* grad_value,
                                     float* grad_sampling_loc, float* grad_attn_weight) {
  return sdb::msda_backward<float>((cudaStream_t)stream, grad_out, value, spatial_shapes, level_start_index,
                                   sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                                   num_levels, num_query, num_point, grad_value, grad_sampling_loc,
                                   grad_attn_weight);
}

extern "C" int sdb_msda_fused_backward_f32(sdb_stream_t stream, const float* grad_out, const float* value,
                                           const int64_t* spatial_shapes, const int64_t* level_start_index,
                                           const float* reference_points, int ref_dim,
                                           const float* sampling_offsets, const float* attn_logits, int batch,
                                           int spatial_size, int num_heads, int channels, int num_levels,
                                           int num_query, int num_point, float* grad_value, float* grad_offsets,
                                           float* grad_attn_logits) {
  using namespace sdb;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = spatial_size, M = num_heads, L = num_levels, Lq = num_query, P = num_point;
  SDB_REQUIRE(batch >= 0 && S >= 0 && Lq >= 0, "msda_fused_backward: bad sizes");
  if (!(channels == 32 && M == 8 && P == 4 && L * P <= 16 && L <= kMaxLevels && (ref_dim == 2 || ref_dim == 4) &&
        (long long)S * M * channels < (1ll << 31))) {
    set_error("msda_fused_backward: built for channels=32, heads=8, points=4, levels*points<=16, ref_dim 2|4");
    return SDB_ERR_UNSUPPORTED;
  }
  const long long nv = (long long)batch * S * M * channels;
  if (nv > 0) {
    SDB_REQUIRE(grad_value, "msda_fused_backward: null grad_value");
    SDB_CUDA(cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)nv, st));
  }
  if ((long long)batch * Lq == 0) return SDB_OK;
  SDB_REQUIRE(grad_out && value && spatial_shapes && level_start_index && reference_points && sampling_offsets &&
              attn_logits && grad_offsets && grad_attn_logits, "msda_fused_backward: null pointer");
#define SDB_FBWD_ARGS st, grad_out, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, batch, S, M, L, \
                      Lq, P, grad_value, grad_offsets, grad_attn_logits, reference_points, ref_dim
  if ((g_bwd_variant == 0 || (g_bwd_variant >= 20 && g_bwd_variant <= 26)) && Lq == S && ref_dim == 2)
    return msda_backward_tile(st, grad_out, value, spatial_shapes, level_start_index, sampling_offsets, attn_logits,
                              batch, S, L, grad_value, grad_offsets, grad_attn_logits, reference_points);
  switch (g_bwd_variant) {     // sdb_msda_set_variant: register budget / occupancy trade-off (tools/microbench.py)
    // measured at the train-step shapes (profiles/msda_bwd_variants_r1.txt): 582 / 536 / 527 / 534 / 546 us (encoder)
    // for 6 / 5 / 4 / 3 CTAs of 128 threads and 2 of 256 per SM -- the spill-free 115-register build wins
    case 2: return launch_bwd_d32<256, 4, 8, 2, true>(SDB_FBWD_ARGS);
    case 3: return launch_bwd_d32<128, 4, 8, 3, true>(SDB_FBWD_ARGS);   // 143 registers, 12 warps / SM
    case 5: return launch_bwd_d32<128, 4, 8, 5, true>(SDB_FBWD_ARGS);   // 96 registers (spills 20 B), 20 warps / SM
    case 6: return launch_bwd_d32<128, 4, 8, 6, true>(SDB_FBWD_ARGS);   // 80 registers (spills 108 B), 24 warps / SM
    default: return launch_bwd_d32<128, 4, 8, 4, true>(SDB_FBWD_ARGS);  // 115 registers, no spills, 16 warps / SM
  }
#undef SDB_FBWD_ARGS
}

// bf16 storage for `grad_out` and `value`; the three gradients are fp32 (grad_value is accumulated with fp32
// reductions -- the caller narrows it if it needs a bf16 gradient).
extern "C" int sdb_msda_backward_bf16(sdb_stream_t stream, const uint16_t* grad_out, const uint16_t* value,
                                      const int64_t* spatial_shapes, const int64_t* level_start_index,
                                      const float* sampling_loc, const float* attn_weight, int batch,
                                      int spatial_size, int num_heads, int channels, int num_levels,
                                      int num_query, int num_point, float* grad_value, float* grad_sampling_loc,
                                      float* grad_attn_weight) {
  using namespace sdb;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = spatial_size, M = num_heads, L = num_levels, Lq = num_query, P = num_point;
  SDB_REQUIRE(batch >= 0 && S >= 0 && M > 0 && channels > 0 && L > 0 && Lq >= 0 && P > 0,
              "msda_backward_bf16: bad sizes batch=%d spatial=%d heads=%d channels=%d levels=%d query=%d point=%d",
              batch, S, M, channels, L, Lq, P);
  if (!(channels == 32 && M == 8 && P == 4 && L <= kMaxLevels && (long long)S * M * channels < (1ll << 31))) {
    set_error("msda_backward_bf16: built for channels=32, heads=8, points=4, levels<=%d (got C=%d M=%d P=%d L=%d)",
              kMaxLevels, channels, M, P, L);
    return SDB_ERR_UNSUPPORTED;
  }
  const long long nv = (long long)batch * S * M * channels;
  if (nv > 0) {
    SDB_REQUIRE(grad_value, "msda_backward_bf16: null grad_value");
    SDB_CUDA(cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)nv, st));
  }
  if ((long long)batch * Lq == 0) return SDB_OK;
  SDB_REQUIRE(grad_out && value && spatial_shapes && level_start_index && sampling_loc && attn_weight &&
              grad_sampling_loc && grad_attn_weight, "msda_backward_bf16: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(grad_out)) & 7) == 0 &&
              ((reinterpret_cast<uintptr_t>(sampling_loc) | reinterpret_cast<uintptr_t>(attn_weight) |
                reinterpret_cast<uintptr_t>(grad_value) | reinterpret_cast<uintptr_t>(grad_sampling_loc) |
                reinterpret_cast<uintptr_t>(grad_attn_weight)) & 15) == 0,
              "msda_backward_bf16: value/grad_out must be 8-byte aligned, the fp32 tensors 16-byte aligned");
  // encoder self-attention: the tile-combining kernel (msda_backward_tile.cu), as for fp32
  if ((g_bwd_variant == 0 || (g_bwd_variant >= 20 && g_bwd_variant <= 26)) && Lq == S && L <= 8)
    return msda_backward_tile_bf16(st, reinterpret_cast<const __nv_bfloat16*>(grad_out),
                                   reinterpret_cast<const __nv_bfloat16*>(value), spatial_shapes, level_start_index,
                                   sampling_loc, attn_weight, batch, S, L, grad_value, grad_sampling_loc,
                                   grad_attn_weight);
  return launch_bwd_d32<128, 4, 8, 4, false, __nv_bfloat16>(
      st, reinterpret_cast<const __nv_bfloat16*>(grad_out), reinterpret_cast<const __nv_bfloat16*>(value),
      spatial_shapes, level_start_index, sampling_loc, attn_weight, batch, S, M, L, Lq, P, grad_value,
      grad_sampling_loc, grad_attn_weight);
}

extern "C" int sdb_msda_backward_f64(sdb_stream_t stream, const double* grad_out, const double* value,
                                     const int64_t* spatial_shapes, const int64_t* level_start_index,
                                     const double* sampling_loc, const double* attn_weight, int batch,
                                     int spatial_size, int num_heads, int channels, int num_levels,
                                     int num_query, int num_point, double* grad_value,
                                     double* grad_sampling_loc, double* grad_attn_weight) {
  return sdb::msda_backward<double>((cudaStream_t)stream, grad_out, value, spatial_shapes, level_start_index,
                                    sampling_loc, attn_weight, batch, spatial_size, num_heads, channels,
                                    num_levels, num_query, num_point, grad_value, grad_sampling_loc,
                                    grad_attn_weight);
}
