// Multi-scale deformable attention, backward -- "x8" lane mapping (EXPERIMENTAL, opt-in: backward variant 7).
//
// Same arithmetic, same reference semantics (ms_deform_im2col_cuda.cuh:87-159, 301-403) and same persistent tiled
// work list as msda_bwd_d32_kernel (msda_backward.cu); what changes is who owns what:
//
//   d32 kernel : 8 lanes per (query, head), 4 channels per lane, a warp covers 4 pairs
//   x8 kernel  : 4 lanes per (query, head), 8 channels per lane, a warp covers 8 pairs
//
// Why: the SASS of the d32 kernel (tools/sass_by_line.py, profiles/sass_msda_bwd_by_line_r1.txt) spends ~48 of its
// ~205 instructions per warp and point on work that scales with the channels a lane holds (corner loads, weighted
// products, reductions, dot products) and the other ~155 on per-point bookkeeping every lane repeats whatever it
// holds: broadcast shuffles, corner-validity bits, 64-bit addresses, the predication around each reduction, the
// reduce-scatter.  Halving the lanes per pair halves that bookkeeping per pair; the per-channel work is unchanged.
// Cost: each 128-byte value line is now fetched / reduced as two 64-byte halves by two instructions (twice the L1
// wavefronts and L2 reduction requests, same sectors), and ~40 more registers per thread.
//
// Ownership inside a group of 4 lanes (lane j, channels 8j..8j+7):
//   * lane j loads and prepares the 4 points of level (c0/4 + j) of the pair (one 32-byte + one 16-byte load,
//     the group reads 128 + 64 contiguous bytes), so a batch of 4 points = the points one lane prepared;
//   * after the 4 points of a batch a two-stage reduce-scatter (xor 2, xor 1) leaves lane j with the four corner dot
//     products of point j, which it turns into grad_x, grad_y and grad_attn of that point.
//
// NOT YET RUN ON HARDWARE (written after round 1's GPU budget was spent).  tests/test_msda_x8_gpu.py is gated by
// SDB_RUN_UNVALIDATED=1; tools/bwd_variants.py times it next to the d32 variants.
#include "msda_common.cuh"

namespace sdb {

namespace {

struct X8Prep {
  int offm;         // float offset of pixel (h0,w0) in the image (multiple of 256) | corner-validity bits 0..3
  float lh, lw, a;  // fractional offsets, attention weight
};

__device__ __forceinline__ X8Prep x8_prep(const LevelTable& lt, int lvl, float x, float y, float a) {
  const int H = lt.H[lvl], W = lt.W[lvl];
  const Tap<float> t = make_tap<float>(x, y, H, W);
  X8Prep r;
  r.lh = t.lh;
  r.lw = t.lw;
  r.a = a;   // not masked: a sample without a valid corner feeds only zero terms
  const int mask = (t.c00 ? 1 : 0) | (t.c01 ? 2 : 0) | (t.c10 ? 4 : 0) | (t.c11 ? 8 : 0);
  r.offm = ((lt.start[lvl] + t.h0 * W + t.w0) * 256) | mask;
  return r;
}

__device__ __forceinline__ float dot8(const float4& a0, const float4& a1, const float4& b0, const float4& b1) {
  float s = a0.x * b0.x;
  s = fmaf(a0.y, b0.y, s);
  s = fmaf(a0.z, b0.z, s);
  s = fmaf(a0.w, b0.w, s);
  s = fmaf(a1.x, b1.x, s);
  s = fmaf(a1.y, b1.y, s);
  s = fmaf(a1.z, b1.z, s);
  return fmaf(a1.w, b1.w, s);
}

constexpr unsigned kAll = 0xffffffffu;

// both 16-byte halves of a lane's 8 channels under ONE branch (ptxas turns a predicated vector `red` into a branch
// region of its own anyway: see profiles/sass_msda_bwd_by_line_r1.txt)
__device__ __forceinline__ void red_add_f8_if(bool pred, float* p, float w, const float4& t0, const float4& t1) {
  if (pred)
    asm volatile(
        "red.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n\t"
        "red.global.add.v4.f32 [%0+16], {%5,%6,%7,%8};"
        ::"l"(p), "f"(w * t0.x), "f"(w * t0.y), "f"(w * t0.z), "f"(w * t0.w), "f"(w * t1.x), "f"(w * t1.y),
          "f"(w * t1.z), "f"(w * t1.w));   // no "memory" clobber: grad_value is write-only here
}

__device__ __forceinline__ float group4_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(kAll, v, 2, 4));
  return fmaxf(v, __shfl_xor_sync(kAll, v, 1, 4));
}
__device__ __forceinline__ float group4_sum(float v) {
  v += __shfl_xor_sync(kAll, v, 2, 4);
  return v + __shfl_xor_sync(kAll, v, 1, 4);
}

// kFused: `loc` / `attn` are the RAW sampling offsets / attention logits and (ref, ref_dim) the reference points; the
// outputs are the gradients of those raw tensors (modules/ms_deform_attn.py:98-112 differentiated here).  L*P <= 16.
// kRolled: the loop over the 4 batches of a chunk stays a loop (4x less code: the instruction cache hit rate of the
// fully unrolled 8-lane kernel is 85.7 %, profiles/ncu_msda_stalls_r1.txt); the batch index becomes a run-time value.
template <int kThreads, int TH, int TW, int kMinBlocks, bool kFused, bool kRolled>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
msda_bwd_x8_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                   const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                   const float* __restrict__ loc, const float* __restrict__ attn, int batch, int S, int L, int Lq,
                   int tiled, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                   float* __restrict__ grad_attn, const float* __restrict__ ref, int ref_dim) {
  constexpr int M = 8, P = 4;
  constexpr int px_stride = M * 32;
  __shared__ LevelTable lt;
  load_levels<TH, TW>(lt, shapes, lsi, L, px_stride);
  constexpr int TQ = TH * TW;
  constexpr int kGroups = kThreads / 4;
  static_assert(TQ % kGroups == 0 && kGroups % 8 == 0, "tile must be a whole number of CTA passes of whole warps");
  const int n_tiles = tiled ? lt.tile_begin[L] : (Lq + TQ - 1) / TQ;
  const long long total = (long long)batch * n_tiles * M;
  const int grp = threadIdx.x >> 2, j = threadIdx.x & 3;
  const bool hi2 = (j & 2) != 0, hi1 = (j & 1) != 0;
  const int LP = L * P;

  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long t2 = item / M;
    const int tile = (int)(t2 % n_tiles);
    const int n = (int)(t2 / n_tiles);
    TileCursor<TH, TW> cur;
    cur.seek(lt, L, tile, tiled != 0, Lq);
    const long long img = (long long)n * S * px_stride + m * 32 + 8 * j;
    const float* vhead = value + img;
    float* gvhead = grad_value + img;

#pragma unroll 1
    for (int it = 0; it < TQ / kGroups; ++it) {   // compile-time trip count: every shuffle below is convergent
      const int q = cur.query(grp + it * kGroups, Lq);
      const bool live = q >= 0;
      const long long nq = (long long)n * Lq + (live ? q : 0);
      const long long pair = nq * M + m;
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      if (live) {
        g0 = ld_stream_f4(reinterpret_cast<const float4*>(grad_out + pair * 32 + 8 * j));
        g1 = ld_stream_f4(reinterpret_cast<const float4*>(grad_out + pair * 32 + 8 * j + 4));
      }

#pragma unroll 1
      for (int c0 = 0; c0 < LP; c0 += 16) {
        // ---- this lane's level: 4 points -----------------------------------------------------------------
        const int lv = c0 / P + j;
        const bool on = live && lv < L;
        const int lvc = min(lv, L - 1);
        float4 xy0 = make_float4(0.f, 0.f, 0.f, 0.f), xy1 = xy0;   // x0 y0 x1 y1 | x2 y2 x3 y3
        float4 aw = make_float4(0.f, 0.f, 0.f, 0.f);
        float sx = 0.f, sy = 0.f;                                   // fused: d loc / d offset of this level
        if (kFused) {
          float4 o0 = xy0, o1 = xy0;
          const float dead = live ? -INFINITY : 0.f;                // a dead group must not produce inf - inf
          float4 lg = make_float4(dead, dead, dead, dead);
          if (on) {
            o0 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv));
            o1 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv + 4));
            lg = ld_stream_f4(reinterpret_cast<const float4*>(attn + pair * LP + 4 * lv));
          }
          const float mx = group4_max(fmaxf(fmaxf(lg.x, lg.y), fmaxf(lg.z, lg.w)));
          const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx), e2 = expf(lg.z - mx), e3 = expf(lg.w - mx);
          const float inv = 1.f / group4_sum((e0 + e1) + (e2 + e3));
          if (on) {
            aw = make_float4(e0 * inv, e1 * inv, e2 * inv, e3 * inv);
            const float* rp = ref + (nq * L + lvc) * ref_dim;
            const float rx = rp[0], ry = rp[1];
            if (ref_dim == 2) {
              const float Wf = (float)lt.W[lvc], Hf = (float)lt.H[lvc];
              sx = 1.f / Wf;
              sy = 1.f / Hf;
              xy0 = make_float4(rx + o0.x / Wf, ry + o0.y / Hf, rx + o0.z / Wf, ry + o0.w / Hf);
              xy1 = make_float4(rx + o1.x / Wf, ry + o1.y / Hf, rx + o1.z / Wf, ry + o1.w / Hf);
            } else {
              const float rw = rp[2], rh = rp[3];
              sx = rw * 0.5f / (float)P;
              sy = rh * 0.5f / (float)P;
              xy0 = make_float4(rx + o0.x / (float)P * rw * 0.5f, ry + o0.y / (float)P * rh * 0.5f,
                                rx + o0.z / (float)P * rw * 0.5f, ry + o0.w / (float)P * rh * 0.5f);
              xy1 = make_float4(rx + o1.x / (float)P * rw * 0.5f, ry + o1.y / (float)P * rh * 0.5f,
                                rx + o1.z / (float)P * rw * 0.5f, ry + o1.w / (float)P * rh * 0.5f);
            }
          }
        } else if (on) {
          xy0 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv));
          xy1 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv + 4));
          aw = ld_stream_f4(reinterpret_cast<const float4*>(attn + pair * LP + 4 * lv));
        }
        X8Prep p[4];
        p[0] = x8_prep(lt, lvc, xy0.x, xy0.y, aw.x);
        p[1] = x8_prep(lt, lvc, xy0.z, xy0.w, aw.y);
        p[2] = x8_prep(lt, lvc, xy1.x, xy1.y, aw.z);
        p[3] = x8_prep(lt, lvc, xy1.z, xy1.w, aw.w);

        float sm_a[4] = {0.f, 0.f, 0.f, 0.f}, sm_g[4] = {0.f, 0.f, 0.f, 0.f};   // fused: softmax backward state
        const int nb = min(4, (LP - c0) / 4);
        // batch b = the 4 points of level c0/4 + b, prepared by lane b
        auto batch = [&](const int b) {
          const int lvl = min(c0 / P + b, L - 1);
          const int ws = lt.wstr[lvl];
          float d[4][4];
          float klh = 0.f, klw = 0.f, ka = 0.f;   // the point this lane finalises: point j of the batch
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int offm = __shfl_sync(kAll, p[r].offm, b, 4);
            const float lh = __shfl_sync(kAll, p[r].lh, b, 4);
            const float lw = __shfl_sync(kAll, p[r].lw, b, 4);
            const float a = __shfl_sync(kAll, p[r].a, b, 4);
            if (j == r) { klh = lh; klw = lw; ka = a; }
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float4 t0 = make_float4(g0.x * a, g0.y * a, g0.z * a, g0.w * a);
            const float4 t1 = make_float4(g1.x * a, g1.y * a, g1.z * a, g1.w * a);
            const int off = offm & ~15;
            const float* pv = vhead + off;
            float* pg = gvhead + off;
            const bool q00 = offm & 1, q01 = offm & 2, q10 = offm & 4, q11 = offm & 8;
            float4 a00, b00, a01, b01, a10, b10, a11, b11;   // corner k: channels 8j..8j+3 | 8j+4..8j+7
            if (q00) { a00 = __ldg(reinterpret_cast<const float4*>(pv)); b00 = __ldg(reinterpret_cast<const float4*>(pv + 4)); }
            if (q01) { a01 = __ldg(reinterpret_cast<const float4*>(pv + px_stride)); b01 = __ldg(reinterpret_cast<const float4*>(pv + px_stride + 4)); }
            if (q10) { a10 = __ldg(reinterpret_cast<const float4*>(pv + ws)); b10 = __ldg(reinterpret_cast<const float4*>(pv + ws + 4)); }
            if (q11) { a11 = __ldg(reinterpret_cast<const float4*>(pv + ws + px_stride)); b11 = __ldg(reinterpret_cast<const float4*>(pv + ws + px_stride + 4)); }
            {
              // the reductions need no loaded data: they fill the wait for the corner loads
              const float w00 = hh * hw, w01 = hh * lw, w10 = lh * hw, w11 = lh * lw;
              red_add_f8_if(q00, pg, w00, t0, t1);
              red_add_f8_if(q01, pg + px_stride, w01, t0, t1);
              red_add_f8_if(q10, pg + ws, w10, t0, t1);
              red_add_f8_if(q11, pg + ws + px_stride, w11, t0, t1);
            }
            d[r][0] = q00 ? dot8(g0, g1, a00, b00) : 0.f;
            d[r][1] = q01 ? dot8(g0, g1, a01, b01) : 0.f;
            d[r][2] = q10 ? dot8(g0, g1, a10, b10) : 0.f;
            d[r][3] = q11 ? dot8(g0, g1, a11, b11) : 0.f;
          }
          // reduce-scatter over the 4 lanes: afterwards lane j holds the four corner sums of point j of the batch
          float e[2][4], f[4];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float keep = hi2 ? d[r + 2][k] : d[r][k];
              const float send = hi2 ? d[r][k] : d[r + 2][k];
              e[r][k] = keep + __shfl_xor_sync(kAll, send, 2, 4);
            }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float keep = hi1 ? e[1][k] : e[0][k];
            const float send = hi1 ? e[0][k] : e[1][k];
            f[k] = keep + __shfl_xor_sync(kAll, send, 1, 4);
          }
          float sxb = 0.f, syb = 0.f;
          if (kFused) {   // chain factors of the batch's level live in lane b
            sxb = __shfl_sync(kAll, sx, b, 4);
            syb = __shfl_sync(kAll, sy, b, 4);
          }
          const float hh = 1.f - klh, hw = 1.f - klw;
          const int point = c0 + 4 * b + j;
          float gx = (float)lt.W[lvl] * ka * (hh * (f[1] - f[0]) + klh * (f[3] - f[2]));
          float gy = (float)lt.H[lvl] * ka * (hw * (f[2] - f[0]) + klw * (f[3] - f[1]));
          const float ga = hh * (hw * f[0] + klw * f[1]) + klh * (hw * f[2] + klw * f[3]);
          if (kFused) {   // loc = ref + off * (sx, sy)
            gx *= sxb;
            gy *= syb;
            if constexpr (kRolled) {   // run-time b: select chain keeps the state in registers
#pragma unroll
              for (int bb = 0; bb < 4; ++bb)
                if (bb == b) { sm_a[bb] = ka; sm_g[bb] = ga; }
            } else {
              sm_a[b] = ka;
              sm_g[b] = ga;
            }
          }
          if (live) {
            st_stream_f2(reinterpret_cast<float2*>(grad_loc + (pair * LP + point) * 2), make_float2(gx, gy));
            if (!kFused) grad_attn[pair * LP + point] = ga;
          }
        };
        if constexpr (kRolled) {
#pragma unroll 1
          for (int b = 0; b < nb; ++b) batch(b);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b >= nb) break;           // warp-uniform
            batch(b);
          }
        }
        if (kFused) {
          // softmax backward over the pair's L*P points: dlogit_i = a_i * (ga_i - sum_k a_k ga_k)
          const float dotp =
              group4_sum((sm_a[0] * sm_g[0] + sm_a[1] * sm_g[1]) + (sm_a[2] * sm_g[2] + sm_a[3] * sm_g[3]));
          if (live) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
              if (b < nb) grad_attn[pair * LP + c0 + 4 * b + j] = sm_a[b] * (sm_g[b] - dotp);
          }
        }
      }
    }
  }
}

template <int kThreads, int TH, int TW, int kMinBlocks, bool kFused, bool kRolled>
int launch_x8(cudaStream_t st, const float* grad_out, const float* value, const int64_t* shapes, const int64_t* lsi,
              const float* loc, const float* attn, int batch, int S, int L, int Lq, float* grad_value,
              float* grad_loc, float* grad_attn, const float* ref, int ref_dim) {
  auto kern = msda_bwd_x8_kernel<kThreads, TH, TW, kMinBlocks, kFused, kRolled>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    int b = 0;
    SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kThreads, 0));
    blocks_per_sm = b > 0 ? b : 1;
  }
  const int tiled = (Lq == S) ? 1 : 0;
  const long long approx_items = (long long)batch * 8 * ((Lq + TH * TW - 1) / (TH * TW) + (tiled ? 4 * L : 0));
  long long grid = (long long)sm_count() * blocks_per_sm;
  if (grid > approx_items) grid = approx_items;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, 0, st>>>(grad_out, value, shapes, lsi, loc, attn, batch, S, L, Lq, tiled,
                                           grad_value, grad_loc, grad_attn, ref, ref_dim);
  SDB_LAUNCH_CHECK("msda_bwd_x8_kernel");
  return SDB_OK;
}

}  // namespace

// Called by msda_backward.cu for backward variants 7 (unrolled) and 8 (rolled batch loop) (num_heads == 8, channels == 32, num_point == 4, 16-byte aligned
// tensors, 32-bit image offsets -- the caller has checked; grad_value is already zero-filled on `st`).
// `ref == nullptr`: loc / attn are sampling locations / attention weights; otherwise the fused form (L*P <= 16).
int msda_backward_x8(cudaStream_t st, const float* grad_out, const float* value, const int64_t* shapes,
                     const int64_t* lsi, const float* loc, const float* attn, int batch, int S, int L, int Lq,
                     float* grad_value, float* grad_loc, float* grad_attn, const float* ref, int ref_dim, bool rolled) {
  // a 4 x 8 pixel tile = 32 pairs of one head = one pass of a 128-thread CTA
#define SDB_X8_ARGS st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, Lq, grad_value, grad_loc, grad_attn
  if (ref != nullptr) {
    if (rolled) return launch_x8<128, 4, 8, 3, true, true>(SDB_X8_ARGS, ref, ref_dim);
    return launch_x8<128, 4, 8, 3, true, false>(SDB_X8_ARGS, ref, ref_dim);
  }
  if (rolled) return launch_x8<128, 4, 8, 3, false, true>(SDB_X8_ARGS, nullptr, 0);
  return launch_x8<128, 4, 8, 3, false, false>(SDB_X8_ARGS, nullptr, 0);
#undef SDB_X8_ARGS
}

}  // namespace sdb
