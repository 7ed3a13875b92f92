// Bilinear-tap arithmetic shared by the MSDA forward / backward kernels.
//
// Semantics follow the reference op: pixel coordinate = loc * size - 0.5,
// a sample participates iff -1 < h < H and -1 < w < W (ms_deform_im2col_cuda.cuh:285-288), corners
// outside [0,H-1]x[0,W-1] read as zero (:55-78), weights (1-lh)(1-lw), (1-lh)lw, lh(1-lw), lh*lw (:80).
#pragma once
#include "common.cuh"

namespace sdb {

constexpr int kMaxLevels = 16;  // tuned kernels keep per-level geometry in shared memory

template <typename T>
struct Tap {
  int h0, w0;   // top-left corner (may be -1)
  T lh, lw;     // fractional offsets
  bool ok;      // sample inside the (-1, size) window
  bool c00, c01, c10, c11;  // corner validity: (h0,w0) (h0,w0+1) (h0+1,w0) (h0+1,w0+1)
};

template <typename T>
__device__ __forceinline__ Tap<T> make_tap(T x, T y, int H, int W) {
  Tap<T> t;
  const T h_im = y * (T)H - (T)0.5;
  const T w_im = x * (T)W - (T)0.5;
  t.ok = (h_im > (T)-1) && (w_im > (T)-1) && (h_im < (T)H) && (w_im < (T)W);
  const T hf = floor(h_im), wf = floor(w_im);
  t.lh = h_im - hf;
  t.lw = w_im - wf;
  t.h0 = t.ok ? (int)hf : 0;
  t.w0 = t.ok ? (int)wf : 0;
  if (!t.ok) { t.lh = 0; t.lw = 0; }
  const bool hlo = t.h0 >= 0, wlo = t.w0 >= 0, hhi = t.h0 + 1 <= H - 1, whi = t.w0 + 1 <= W - 1;
  t.c00 = t.ok && hlo && wlo;
  t.c01 = t.ok && hlo && whi;
  t.c10 = t.ok && hhi && wlo;
  t.c11 = t.ok && hhi && whi;
  return t;
}

// Per-level geometry staged once per CTA.  In "tiled" mode (num_query == spatial_size, i.e. encoder
// self-attention where query i sits on pixel i of the level pyramid) the work list is cut into
// TH x TW pixel tiles per level so that the value lines a CTA gathers stay L1-resident; otherwise
// tiles are TH*TW consecutive queries.
struct LevelTable {
  int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels], tiles_x[kMaxLevels];
  int wstr[kMaxLevels];  // floats between two rows of the level: W * num_heads * 32
  int tile_begin[kMaxLevels + 1];
};

template <int TH, int TW>
__device__ __forceinline__ void load_levels(LevelTable& t, const int64_t* __restrict__ shapes,
                                            const int64_t* __restrict__ lsi, int L, int px_stride = 0) {
  if (threadIdx.x < L) {
    const int l = threadIdx.x;
    t.H[l] = (int)shapes[2 * l];
    t.W[l] = (int)shapes[2 * l + 1];
    t.start[l] = (int)lsi[l];
    t.tiles_x[l] = (t.W[l] + TW - 1) / TW;
    t.wstr[l] = t.W[l] * px_stride;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < L; ++l) {
      t.tile_begin[l] = acc;
      acc += ((t.H[l] + TH - 1) / TH) * t.tiles_x[l];
    }
    t.tile_begin[L] = acc;
  }
  __syncthreads();
}

// query index of local slot i of a tile, or -1
template <int TH, int TW>
struct TileCursor {
  int lvl, y0, x0, Hl, Wl, start, first_q;
  bool tiled;
  __device__ __forceinline__ void seek(const LevelTable& t, int L, int tile, bool tiled_, int) {
    tiled = tiled_;
    if (tiled) {
      lvl = 0;
      while (lvl + 1 < L && tile >= t.tile_begin[lvl + 1]) ++lvl;
      const int tt = tile - t.tile_begin[lvl];
      y0 = (tt / t.tiles_x[lvl]) * TH;
      x0 = (tt % t.tiles_x[lvl]) * TW;
      Hl = t.H[lvl];
      Wl = t.W[lvl];
      start = t.start[lvl];
    } else {
      first_q = tile * (TH * TW);
    }
  }
  __device__ __forceinline__ int query(int i, int Lq) const {
    if (tiled) {
      const int y = y0 + i / TW, x = x0 + i % TW;
      return (y < Hl && x < Wl) ? start + y * Wl + x : -1;
    }
    const int q = first_q + i;
    return q < Lq ? q : -1;
  }
};

// ------------------------------------------------------------------------------------------------
// Fused prologue: the module's softmax over the L*P attention logits and its
// sampling-location arithmetic (modules/ms_deform_attn.py:98-112) evaluated by the lane that owns the two points,
// so the (N, Lq, M, L, P, 2) locations and (N, Lq, M, L, P) weights are never materialised in HBM.
//   ref_dim == 2 (encoder):  loc = ref + off / (W_l, H_l)
//   ref_dim == 4 (decoder):  loc = ref_xy + off / P * ref_wh * 0.5
// `sx`, `sy` return d(loc)/d(off) for the backward.  One chunk only: L*P <= 16.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float group8_max(float v) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, 8));
  return v;
}
__device__ __forceinline__ float group8_sum(float v) {
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 8);
  return v;
}

struct FusedPoints {
  float4 loc;   // x0 y0 x1 y1
  float2 a;     // softmax weights of the two points
  float sx, sy; // d loc / d offset (same for both points: they sit on one level)
};

__device__ __forceinline__ FusedPoints fused_prologue(const LevelTable& lt, const float* __restrict__ ref,
                                                      int ref_dim, const float* __restrict__ offsets,
                                                      const float* __restrict__ logits, long long nq, long long pair,
                                                      int L, int P, int LP, int pt, bool live) {
  FusedPoints r;
  const bool on = live && pt < LP;
  float4 off = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 lg = make_float2(live ? -INFINITY : 0.f, live ? -INFINITY : 0.f);
  if (on) {
    off = ld_stream_f4(reinterpret_cast<const float4*>(offsets + pair * LP * 2 + 2 * pt));
    lg = ld_stream_f2(reinterpret_cast<const float2*>(logits + pair * LP + pt));
  }
  const float mx = group8_max(fmaxf(lg.x, lg.y));
  const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx);
  const float inv = 1.f / group8_sum(e0 + e1);
  r.a = make_float2(on ? e0 * inv : 0.f, on ? e1 * inv : 0.f);
  const int lvl = min(pt / P, L - 1);
  const float* rp = ref + (nq * L + lvl) * ref_dim;
  float rx = 0.f, ry = 0.f;
  r.sx = r.sy = 0.f;
  if (on) {
    rx = rp[0];
    ry = rp[1];
    if (ref_dim == 2) {
      r.sx = 1.f / (float)lt.W[lvl];
      r.sy = 1.f / (float)lt.H[lvl];
      r.loc = make_float4(rx + off.x / (float)lt.W[lvl], ry + off.y / (float)lt.H[lvl],
                          rx + off.z / (float)lt.W[lvl], ry + off.w / (float)lt.H[lvl]);
    } else {
      const float rw = rp[2], rh = rp[3];
      r.sx = rw * 0.5f / (float)P;
      r.sy = rh * 0.5f / (float)P;
      r.loc = make_float4(rx + off.x / (float)P * rw * 0.5f, ry + off.y / (float)P * rh * 0.5f,
                          rx + off.z / (float)P * rw * 0.5f, ry + off.w / (float)P * rh * 0.5f);
    }
  } else {
    r.loc = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return r;
}

}  // namespace sdb
