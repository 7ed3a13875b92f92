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
//  * grad_sampling_loc / grad_attn_weight: the sum over the 32 channels is 4 in-register adds + a 3-step
//    width-8 shuffle tree, no shared memory, no barrier; the lane that loaded a point keeps its result and the
//    group writes them back as one coalesced float4 + float2 per lane, so these two tensors are written
//    exactly once (no zero-fill pass; the reference memsets all three gradients);
//  * grad_value is zero-filled with one cudaMemsetAsync on the same stream.
// Summation order differs from the reference's (fp32 atomics are unordered there too); parity is to 1e-3 rel.
#include "msda_common.cuh"

namespace sdb {

extern int g_bwd_variant;

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
// tuned kernel: fp32, D == 32, L*P even.
// ------------------------------------------------------------------------------------------------
struct BwdPrep {
  int off, wstr;     // as in the forward kernel
  float lh, lw, a;   // fractional offsets and attention weight (0,0,0 when the sample is out of range)
  int info;          // bits 0..3 corner validity, bits 4.. level index
};

__device__ __forceinline__ BwdPrep bwd_prep(const LevelTable& lt, int lvl, float x, float y, float a,
                                            int px_stride, int head_off) {
  const int H = lt.H[lvl], W = lt.W[lvl];
  const Tap<float> t = make_tap<float>(x, y, H, W);
  BwdPrep r;
  r.lh = t.lh;
  r.lw = t.lw;
  r.a = t.ok ? a : 0.f;
  r.info = (t.c00 ? 1 : 0) | (t.c01 ? 2 : 0) | (t.c10 ? 4 : 0) | (t.c11 ? 8 : 0) | (lvl << 4);
  r.off = (lt.start[lvl] + t.h0 * W + t.w0) * px_stride + head_off;
  r.wstr = W * px_stride;
  return r;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

template <int kThreads, int TH, int TW, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
msda_bwd_d32_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                    const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                    const float* __restrict__ loc, const float* __restrict__ attn, int batch, int S, int M,
                    int L, int Lq, int P, int tiled, float* __restrict__ grad_value,
                    float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  __shared__ LevelTable lt;
  load_levels<TH, TW>(lt, shapes, lsi, L);
  constexpr int TQ = TH * TW;
  constexpr int kGroups = kThreads / 8;
  const int n_tiles = tiled ? lt.tile_begin[L] : (Lq + TQ - 1) / TQ;
  const long long total = (long long)batch * n_tiles * M;
  const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  const int LP = L * P;
  const int px_stride = M * 32;
  const int ps4 = px_stride >> 2;

  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long t2 = item / M;
    const int tile = (int)(t2 % n_tiles);
    const int n = (int)(t2 / n_tiles);
    TileCursor<TH, TW> cur;
    cur.seek(lt, L, tile, tiled != 0, Lq);
    const long long img = (long long)n * S * px_stride;
    const float* vimg = value + img;
    float* gvimg = grad_value + img;
    const int head_off = m * 32;

    for (int i = grp; i < TQ; i += kGroups) {
      const int q = cur.query(i, Lq);
      if (q < 0) continue;  // group-uniform
      const long long pair = ((long long)n * Lq + q) * M + m;
      const float* lp = loc + pair * LP * 2;
      const float* ap = attn + pair * LP;
      const float4 g = ld_stream_f4(reinterpret_cast<const float4*>(grad_out + pair * 32 + 4 * j));

      for (int c0 = 0; c0 < LP; c0 += 16) {
        const int pt = c0 + 2 * j;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 a2 = make_float2(0.f, 0.f);
        if (pt < LP) {
          l4 = ld_stream_f4(reinterpret_cast<const float4*>(lp + 2 * pt));
          a2 = ld_stream_f2(reinterpret_cast<const float2*>(ap + pt));
        }
        const int lv0 = min(pt / P, L - 1), lv1 = min((pt + 1) / P, L - 1);
        const BwdPrep p0 = bwd_prep(lt, lv0, l4.x, l4.y, a2.x, px_stride, head_off);
        const BwdPrep p1 = bwd_prep(lt, lv1, l4.z, l4.w, a2.y, px_stride, head_off);
        float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);  // (d/dx, d/dy) of this lane's two points
        float2 gatt = make_float2(0.f, 0.f);
        const int npt = min(16, LP - c0);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          if (s >= npt) break;  // uniform
          const BwdPrep& src = (s & 1) ? p1 : p0;
          const int sl = s >> 1;
          const int off = __shfl_sync(gmask, src.off, sl, 8) + 4 * j;
          const int ws4 = __shfl_sync(gmask, src.wstr, sl, 8) >> 2;
          const float lh = __shfl_sync(gmask, src.lh, sl, 8);
          const float lw = __shfl_sync(gmask, src.lw, sl, 8);
          const float a = __shfl_sync(gmask, src.a, sl, 8);
          const int info = __shfl_sync(gmask, src.info, sl, 8);
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float4 tg = make_float4(g.x * a, g.y * a, g.z * a, g.w * a);
          const float4* b = reinterpret_cast<const float4*>(vimg + off);
          float* gb = gvimg + off;
          float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
          if (info & 1) {
            v00 = __ldg(b);
            const float w = hh * hw;
            red_add_f4(gb, make_float4(w * tg.x, w * tg.y, w * tg.z, w * tg.w));
          }
          if (info & 2) {
            v01 = __ldg(b + ps4);
            const float w = hh * lw;
            red_add_f4(gb + px_stride, make_float4(w * tg.x, w * tg.y, w * tg.z, w * tg.w));
          }
          if (info & 4) {
            v10 = __ldg(b + ws4);
            const float w = lh * hw;
            red_add_f4(gb + 4 * ws4, make_float4(w * tg.x, w * tg.y, w * tg.z, w * tg.w));
          }
          if (info & 8) {
            v11 = __ldg(b + ws4 + ps4);
            const float w = lh * lw;
            red_add_f4(gb + 4 * ws4 + px_stride, make_float4(w * tg.x, w * tg.y, w * tg.z, w * tg.w));
          }
          // d(out)/d(attn) = g . bilinear(value);  d/dh, d/dw via the corner differences
          const float4 top = make_float4(v01.x - v00.x, v01.y - v00.y, v01.z - v00.z, v01.w - v00.w);
          const float4 bot = make_float4(v11.x - v10.x, v11.y - v10.y, v11.z - v10.z, v11.w - v10.w);
          const float4 lef = make_float4(v10.x - v00.x, v10.y - v00.y, v10.z - v00.z, v10.w - v00.w);
          const float4 rig = make_float4(v11.x - v01.x, v11.y - v01.y, v11.z - v01.z, v11.w - v01.w);
          const float4 val = make_float4(
              hh * (hw * v00.x + lw * v01.x) + lh * (hw * v10.x + lw * v11.x),
              hh * (hw * v00.y + lw * v01.y) + lh * (hw * v10.y + lw * v11.y),
              hh * (hw * v00.z + lw * v01.z) + lh * (hw * v10.z + lw * v11.z),
              hh * (hw * v00.w + lw * v01.w) + lh * (hw * v10.w + lw * v11.w));
          float ga = dot4(g, val);
          float gw = hh * dot4(tg, top) + lh * dot4(tg, bot);
          float gh = hw * dot4(tg, lef) + lw * dot4(tg, rig);
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            ga += __shfl_xor_sync(gmask, ga, o, 8);
            gw += __shfl_xor_sync(gmask, gw, o, 8);
            gh += __shfl_xor_sync(gmask, gh, o, 8);
          }
          if (j == sl) {
            const int lvl = info >> 4;
            const float Wf = (float)lt.W[lvl], Hf = (float)lt.H[lvl];
            if (s & 1) { gl.z = Wf * gw; gl.w = Hf * gh; gatt.y = ga; }
            else       { gl.x = Wf * gw; gl.y = Hf * gh; gatt.x = ga; }
          }
        }
        if (pt < LP) {
          st_stream_f4(reinterpret_cast<float4*>(grad_loc + pair * LP * 2 + 2 * pt), gl);
          st_stream_f2(reinterpret_cast<float2*>(grad_attn + pair * LP + pt), gatt);
        }
      }
    }
  }
}

template <int kThreads, int TH, int TW, int kMinBlocks>
static int launch_bwd_d32(cudaStream_t st, const float* grad_out, const float* value, const int64_t* shapes,
                          const int64_t* lsi, const float* loc, const float* attn, int batch, int S, int M,
                          int L, int Lq, int P, float* grad_value, float* grad_loc, float* grad_attn) {
  auto kern = msda_bwd_d32_kernel<kThreads, TH, TW, kMinBlocks>;
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
  kern<<<(unsigned)grid, kThreads, 0, st>>>(grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P,
                                           tiled, grad_value, grad_loc, grad_attn);
  SDB_LAUNCH_CHECK("msda_bwd_d32_kernel");
  return SDB_OK;
}

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
    const bool fast_ok = D == 32 && ((L * P) % 2 == 0) && L <= kMaxLevels && fits32 && (al & 15) == 0;
    int v = g_bwd_variant;
    if (v != 9 && fast_ok) {
      if (v == 0) v = (Lq == S) ? 2 : 4;
      switch (v) {
        case 1: return launch_bwd_d32<1024, 16, 16, 1>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn);
        case 2: return launch_bwd_d32<512, 8, 16, 2>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn);
        case 3: return launch_bwd_d32<256, 8, 8, 4>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn);
        case 5: return launch_bwd_d32<512, 16, 16, 2>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn);
        case 6: return launch_bwd_d32<256, 16, 16, 4>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn);
        default: return launch_bwd_d32<256, 4, 8, 4>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, M, L, Lq, P, grad_value, grad_loc, grad_attn);
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
