// Multi-scale deformable attention, backward, for encoder self-attention (num_query == spatial_size).  sm_100a.
//
// Reference semantics: ms_deformable_col2im_cuda / ..._shm_blocksize_aware_reduce_v1
// (/root/reference/detr_od/models/utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159, 301-403, 956-1327).
//
// Why a second backward kernel.  msda_bwd_d32_kernel issues one 128-byte `red.global.add.v4.f32` line per sampled
// corner: 64 per (query, head), 81.7 M L2 reduction sectors per encoder launch at the train-step shape -- and the
// fp32 reduction rate of the L2 slices is what it runs at (every variant that kept the reductions measured 520-600 us
// whatever its occupancy, loads in flight or instruction count; the one that split each line into half-sector
// requests took exactly twice as long: profiles/msda_bwd_variants_r2.txt).  In encoder self-attention query i sits on
// pixel i, so the 64 queries of an 8 x 8 pixel tile sample the SAME few hundred value pixels of each level over and
// over (a level-0 tile: 4096 corner updates onto ~250 distinct pixels).  This kernel combines them on chip:
//
//   A  one thread per sampling point: bilinear tap -> record {pixel offset | corner bits, lh, lw, attention weight} in
//      shared memory; the point is counted into the bucket of its top-left pixel inside a per-level WINDOW around the
//      tile's footprint (integer shared-memory atomics, which are native; fp32 shared atomics are CAS loops);
//   B  prefix sum over the buckets + index scatter = the tile's points sorted by pixel;
//   C  gather: 4 lanes x 8 channels per (query, head) read the records (broadcast LDS), fetch the four corner lines,
//      reduce to the four corner dot products and finish grad_sampling_loc / grad_attn_weight exactly like the d32
//      kernel (closed forms of the corner dot products, two-stage reduce-scatter) -- but issue NO reductions;
//   D  scatter turned into a gather: each window pixel is owned by 4 lanes x 8 channels, which walk the (at most
//      four) buckets whose points touch it, accumulate weight x grad_out from shared memory in registers in a fixed
//      order, and issue ONE 128-byte reduction line per touched pixel.
//
// Points whose top-left pixel falls outside the window (sampling offsets beyond +-kHalo pixels, or a level whose
// footprint would not fit the bucket table) take the direct `red` path inside step C, so any input is handled; only
// the speed depends on locality.  grad_value is still accumulated with reductions because neighbouring tiles overlap.
#include "msda_common.cuh"

namespace sdb {

namespace {

constexpr int kTT = 256;          // threads per CTA
constexpr int kTH = 8, kTW = 8;   // query tile (pixels of the query's own level)
constexpr int kTQ = kTH * kTW;
constexpr int kHalo = 5;          // sampling offsets up to +-kHalo pixels stay inside the window
constexpr int kMaxBuckets = 2040; // bucket table (ints); + guard entries = 2048
constexpr int kTableInts = 2048;
constexpr int kMaxWinLevels = 8;
constexpr unsigned kAllLanes = 0xffffffffu;

struct Windows {
  int y0[kMaxWinLevels], x0[kMaxWinLevels], h[kMaxWinLevels], w[kMaxWinLevels], base[kMaxWinLevels];
};

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

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}

template <int kWidth>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
  for (int o = kWidth / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kAllLanes, v, o, kWidth));
  return v;
}
template <int kWidth>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int o = kWidth / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kAllLanes, v, o, kWidth);
  return v;
}

// Shared-memory layout (dynamic): records | grad_out tile | bucket table | sorted point ids
template <int kSlots>
struct TileSmem {
  static constexpr int kRecStride = kSlots + 1;                 // records per query row (+1: bank spread)
  static constexpr int kRecBytes = kTQ * kRecStride * 16;
  static constexpr int kGBytes = kTQ * 32 * 4;
  static constexpr int kTableBytes = kTableInts * 4;
  static constexpr int kIdxBytes = kTQ * kSlots * 2;
  static constexpr int kTotal = kRecBytes + kGBytes + kTableBytes + kIdxBytes;
};

// kSlots: point slots per (query, head) -- 16 (L <= 4) or 32 (L <= 8); P == 4, 8 heads x 32 channels.
// kFused: `loc` / `attn` are the RAW sampling offsets / attention logits, `ref` the (N, Lq, L, 2) reference points and
// the outputs are the gradients of the raw tensors (ms_deform_attn.py:98-105 differentiated here); kSlots == 16 only.
template <int kSlots, bool kFused>
__global__ void __launch_bounds__(kTT, 2)
msda_bwd_tile_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                     const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const float* __restrict__ loc, const float* __restrict__ attn, int batch, int S, int L,
                     float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn,
                     const float* __restrict__ ref) {
  constexpr int M = 8, P = 4;
  constexpr int px_stride = M * 32;
  using SM = TileSmem<kSlots>;
  constexpr int RS = SM::kRecStride;
  constexpr int kPtsPerThread = kTQ * kSlots / kTT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* rec = reinterpret_cast<float4*>(smem_raw);
  float* gtile = reinterpret_cast<float*>(smem_raw + SM::kRecBytes);
  int* table = reinterpret_cast<int*>(smem_raw + SM::kRecBytes + SM::kGBytes);
  unsigned short* sidx = reinterpret_cast<unsigned short*>(smem_raw + SM::kRecBytes + SM::kGBytes + SM::kTableBytes);
  __shared__ LevelTable lt;
  __shared__ Windows win;
  __shared__ int warp_tot[kTT / 32];

  load_levels<kTH, kTW>(lt, shapes, lsi, L, px_stride);
  const int Lq = S;
  const int n_tiles = lt.tile_begin[L];
  const long long total = (long long)batch * n_tiles * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int LP = L * P;

  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long t2 = item / M;
    const int tile = (int)(t2 % n_tiles);
    const int n = (int)(t2 / n_tiles);
    TileCursor<kTH, kTW> cur;
    cur.seek(lt, L, tile, true, Lq);

    // ---- window geometry (thread 0) + cleared bucket table ------------------------------------------------------
    __syncthreads();   // previous item's phase D has finished with the table / records
    {
      int4* t4 = reinterpret_cast<int4*>(table);
      for (int i = tid; i < kTableInts / 4; i += kTT) t4[i] = make_int4(0, 0, 0, 0);
    }
    if (tid == 0) {
      int acc = 1;   // bucket 0 is a guard (always empty)
      for (int l = L - 1; l >= 0; --l) {   // coarse levels first: smallest windows, most reuse
        const float sy = (float)lt.H[l] / (float)cur.Hl, sx = (float)lt.W[l] / (float)cur.Wl;
        const int wy0 = (int)floorf((float)cur.y0 * sy) - kHalo - 1;
        const int wx0 = (int)floorf((float)cur.x0 * sx) - kHalo - 1;
        const int wy1 = (int)ceilf((float)(cur.y0 + kTH) * sy) + kHalo + 1;
        const int wx1 = (int)ceilf((float)(cur.x0 + kTW) * sx) + kHalo + 1;
        const int wh = wy1 - wy0 + 1, ww = wx1 - wx0 + 1;
        const bool use = l < kMaxWinLevels && acc + wh * ww <= kMaxBuckets;
        win.y0[l] = wy0;
        win.x0[l] = wx0;
        win.h[l] = use ? wh : 0;
        win.w[l] = use ? ww : 0;
        win.base[l] = acc;
        if (use) acc += wh * ww;
      }
    }
    __syncthreads();

    // ---- A: records + bucket counts ------------------------------------------------------------------------------
    int key[kPtsPerThread], rank[kPtsPerThread];
#pragma unroll
    for (int i = 0; i < kPtsPerThread; ++i) {
      const int e = tid + i * kTT;
      const int ql = e / kSlots, pt = e % kSlots;
      const int q = cur.query(ql, Lq);
      const bool live = q >= 0;
      const bool on = live && pt < LP;
      const int lvl = min(pt / P, L - 1);
      const long long nq = (long long)n * Lq + (live ? q : 0);
      const long long pair = nq * M + m;
      float x = 0.f, y = 0.f, a = 0.f;
      if (kFused) {
        float2 off = make_float2(0.f, 0.f);
        float lg = live ? -INFINITY : 0.f;
        if (on) {
          off = ld_stream_f2(reinterpret_cast<const float2*>(loc + pair * LP * 2 + 2 * pt));
          lg = __ldg(attn + pair * LP + pt);
        }
        const float mx = seg_max<kSlots>(lg);
        const float ex = expf(lg - mx);
        const float inv = 1.f / seg_sum<kSlots>(ex);
        if (on) {
          a = ex * inv;
          const float2 rp = __ldg(reinterpret_cast<const float2*>(ref + (nq * L + lvl) * 2));
          x = rp.x + off.x / (float)lt.W[lvl];
          y = rp.y + off.y / (float)lt.H[lvl];
        }
      } else if (on) {
        const float2 xy = ld_stream_f2(reinterpret_cast<const float2*>(loc + pair * LP * 2 + 2 * pt));
        x = xy.x;
        y = xy.y;
        a = __ldg(attn + pair * LP + pt);
      }
      const int H = lt.H[lvl], W = lt.W[lvl];
      const Tap<float> t = make_tap<float>(x, y, H, W);
      int mask = on ? ((t.c00 ? 1 : 0) | (t.c01 ? 2 : 0) | (t.c10 ? 4 : 0) | (t.c11 ? 8 : 0)) : 0;
      key[i] = -1;
      rank[i] = 0;
      if (mask) {
        const int ww = win.w[lvl];
        const int wy = t.h0 - win.y0[lvl], wx = t.w0 - win.x0[lvl];
        if (ww > 0 && wy >= 0 && wy < win.h[lvl] - 1 && wx >= 0 && wx < ww - 1) {
          key[i] = win.base[lvl] + wy * ww + wx;
          rank[i] = atomicAdd(&table[key[i] + 1], 1);
        } else {
          mask |= 16;   // outside the window: step C reduces this point directly
        }
      }
      const int offm = ((lt.start[lvl] + t.h0 * W + t.w0) * px_stride) | mask;
      rec[ql * RS + pt] = make_float4(__int_as_float(offm), t.lh, t.lw, a);
    }
    __syncthreads();

    // ---- B: inclusive scan of table[1..] in place: bucket k = [table[k], table[k+1]) ------------------------------
    {
      int4* t4 = reinterpret_cast<int4*>(table);
      int4 v0 = t4[2 * tid], v1 = t4[2 * tid + 1];
      v0.y += v0.x; v0.z += v0.y; v0.w += v0.z;
      v1.x += v0.w; v1.y += v1.x; v1.z += v1.y; v1.w += v1.z;
      int incl = v1.w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kAllLanes, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      int before = incl - v1.w;
#pragma unroll
      for (int w = 0; w < kTT / 32; ++w)
        if (w < warp) before += warp_tot[w];
      v0.x += before; v0.y += before; v0.z += before; v0.w += before;
      v1.x += before; v1.y += before; v1.z += before; v1.w += before;
      t4[2 * tid] = v0;
      t4[2 * tid + 1] = v1;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPtsPerThread; ++i)
      if (key[i] >= 0) sidx[table[key[i]] + rank[i]] = (unsigned short)(tid + i * kTT);

    // ---- C: gather; grad_sampling_loc, grad_attn_weight -----------------------------------------------------------
    {
      const int grp = tid >> 2, j = tid & 3;   // query slot of the tile, channel octet
      const bool hi2 = (j & 2) != 0, hi1 = (j & 1) != 0;
      const int q = cur.query(grp, Lq);
      const bool live = q >= 0;
      const long long nq = (long long)n * Lq + (live ? q : 0);
      const long long pair = nq * M + m;
      const long long img = (long long)n * S * px_stride + m * 32 + 8 * j;
      const float* vhead = value + img;
      float* gvhead = grad_value + img;
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      if (live) {
        g0 = ld_stream_f4(reinterpret_cast<const float4*>(grad_out + pair * 32 + 8 * j));
        g1 = ld_stream_f4(reinterpret_cast<const float4*>(grad_out + pair * 32 + 8 * j + 4));
      }
      *reinterpret_cast<float4*>(gtile + grp * 32 + 8 * j) = g0;
      *reinterpret_cast<float4*>(gtile + grp * 32 + 8 * j + 4) = g1;
      const float4* myrec = rec + grp * RS;
      float sm_a[kSlots / 4], sm_g[kSlots / 4];   // fused: softmax backward state (point j of every batch)
#pragma unroll
      for (int b = 0; b < kSlots / 4; ++b) { sm_a[b] = 0.f; sm_g[b] = 0.f; }
#pragma unroll
      for (int b = 0; b < kSlots / 4; ++b) {   // batch b = the 4 points of level b
        if (b >= L) break;                     // warp-uniform
        const int ws = lt.wstr[b];
        float d[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 R = myrec[4 * b + r];
          const int offm = __float_as_int(R.x);
          const int off = offm & ~31;
          const float* pv = vhead + off;
          const bool q00 = offm & 1, q01 = offm & 2, q10 = offm & 4, q11 = offm & 8;
          float4 a00, b00, a01, b01, a10, b10, a11, b11;   // corner k: channels 8j..8j+3 | 8j+4..8j+7
          if (q00) { a00 = __ldg(reinterpret_cast<const float4*>(pv)); b00 = __ldg(reinterpret_cast<const float4*>(pv + 4)); }
          if (q01) { a01 = __ldg(reinterpret_cast<const float4*>(pv + px_stride)); b01 = __ldg(reinterpret_cast<const float4*>(pv + px_stride + 4)); }
          if (q10) { a10 = __ldg(reinterpret_cast<const float4*>(pv + ws)); b10 = __ldg(reinterpret_cast<const float4*>(pv + ws + 4)); }
          if (q11) { a11 = __ldg(reinterpret_cast<const float4*>(pv + ws + px_stride)); b11 = __ldg(reinterpret_cast<const float4*>(pv + ws + px_stride + 4)); }
          if (offm & 16) {   // outside the window (rare): reduce directly, like the d32 kernel
            const float lh = R.y, lw = R.z, a = R.w;
            const float hh = 1.f - lh, hw = 1.f - lw;
            float* pg = gvhead + off;
            const float w00 = hh * hw * a, w01 = hh * lw * a, w10 = lh * hw * a, w11 = lh * lw * a;
            if (q00) { red_add_f4(pg, make_float4(w00 * g0.x, w00 * g0.y, w00 * g0.z, w00 * g0.w)); red_add_f4(pg + 4, make_float4(w00 * g1.x, w00 * g1.y, w00 * g1.z, w00 * g1.w)); }
            if (q01) { red_add_f4(pg + px_stride, make_float4(w01 * g0.x, w01 * g0.y, w01 * g0.z, w01 * g0.w)); red_add_f4(pg + px_stride + 4, make_float4(w01 * g1.x, w01 * g1.y, w01 * g1.z, w01 * g1.w)); }
            if (q10) { red_add_f4(pg + ws, make_float4(w10 * g0.x, w10 * g0.y, w10 * g0.z, w10 * g0.w)); red_add_f4(pg + ws + 4, make_float4(w10 * g1.x, w10 * g1.y, w10 * g1.z, w10 * g1.w)); }
            if (q11) { red_add_f4(pg + ws + px_stride, make_float4(w11 * g0.x, w11 * g0.y, w11 * g0.z, w11 * g0.w)); red_add_f4(pg + ws + px_stride + 4, make_float4(w11 * g1.x, w11 * g1.y, w11 * g1.z, w11 * g1.w)); }
          }
          d[r][0] = q00 ? dot8(g0, g1, a00, b00) : 0.f;
          d[r][1] = q01 ? dot8(g0, g1, a01, b01) : 0.f;
          d[r][2] = q10 ? dot8(g0, g1, a10, b10) : 0.f;
          d[r][3] = q11 ? dot8(g0, g1, a11, b11) : 0.f;
        }
        // reduce-scatter over the 4 lanes: afterwards lane j holds the four corner sums of point j of the batch
        float e2[2][4], f[4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float keep = hi2 ? d[r + 2][k] : d[r][k];
            const float send = hi2 ? d[r][k] : d[r + 2][k];
            e2[r][k] = keep + __shfl_xor_sync(kAllLanes, send, 2, 4);
          }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float keep = hi1 ? e2[1][k] : e2[0][k];
          const float send = hi1 ? e2[0][k] : e2[1][k];
          f[k] = keep + __shfl_xor_sync(kAllLanes, send, 1, 4);
        }
        const float4 K = myrec[4 * b + j];   // the point this lane finalises
        const float klh = K.y, klw = K.z, ka = K.w;
        const float hh = 1.f - klh, hw = 1.f - klw;
        const int point = 4 * b + j;
        float gx = (float)lt.W[b] * ka * (hh * (f[1] - f[0]) + klh * (f[3] - f[2]));
        float gy = (float)lt.H[b] * ka * (hw * (f[2] - f[0]) + klw * (f[3] - f[1]));
        const float ga = hh * (hw * f[0] + klw * f[1]) + klh * (hw * f[2] + klw * f[3]);
        if (kFused) {   // loc = ref + off / (W, H)
          gx *= 1.f / (float)lt.W[b];
          gy *= 1.f / (float)lt.H[b];
          sm_a[b] = ka;
          sm_g[b] = ga;
        }
        if (live) {
          st_stream_f2(reinterpret_cast<float2*>(grad_loc + (pair * LP + point) * 2), make_float2(gx, gy));
          if (!kFused) grad_attn[pair * LP + point] = ga;
        }
      }
      if (kFused) {
        // softmax backward over the pair's L*P points: dlogit_i = a_i * (ga_i - sum_k a_k ga_k)
        float part = 0.f;
#pragma unroll
        for (int b = 0; b < kSlots / 4; ++b) part = fmaf(sm_a[b], sm_g[b], part);
        const float dotp = seg_sum<4>(part);
        if (live) {
#pragma unroll
          for (int b = 0; b < kSlots / 4; ++b)
            if (b < L) grad_attn[pair * LP + 4 * b + j] = sm_a[b] * (sm_g[b] - dotp);
        }
      }
    }
    __syncthreads();

    // ---- C': records -> the four corner coefficients a * w_k (0 for corners that do not contribute) ---------------
#pragma unroll
    for (int i = 0; i < kPtsPerThread; ++i) {
      const int e = tid + i * kTT;
      const int ri = (e / kSlots) * RS + (e % kSlots);
      const float4 R = rec[ri];
      const int offm = __float_as_int(R.x);
      const float lh = R.y, lw = R.z, a = R.w;
      const float hh = 1.f - lh, hw = 1.f - lw;
      rec[ri] = make_float4((offm & 1) ? hh * hw * a : 0.f, (offm & 2) ? hh * lw * a : 0.f,
                            (offm & 4) ? lh * hw * a : 0.f, (offm & 8) ? lh * lw * a : 0.f);
    }
    __syncthreads();

    // ---- D: every window pixel gathers its contributions; one reduction line per touched pixel --------------------
    {
      const int grp = tid >> 2, j = tid & 3;   // pixel owner group; lane j owns channels 4j..4j+3 and 16+4j..16+4j+3
      const float* recf = reinterpret_cast<const float*>(rec);
      float* gvimg = grad_value + (long long)n * S * px_stride + m * 32 + 4 * j;
      for (int l = 0; l < L; ++l) {
        const int ww = win.w[l];
        if (ww == 0) continue;
        const int npx = win.h[l] * ww, wb = win.base[l];
        const int H = lt.H[l], W = lt.W[l];
        int wy = grp / ww, wx = grp - wy * ww;
        for (int p = grp; p < npx; p += kTT / 4) {
          const int k = wb + p;
          // same-row buckets (wx-1 -> this pixel is their corner 01, wx -> corner 00)
          const int b0 = table[k - 1], b1 = table[k], b2 = table[k + 1];
          // previous-row buckets (wx-1 -> corner 11, wx -> corner 10); row 0 has none
          int a0 = 0, a1 = 0, a2 = 0;
          if (wy > 0) { a0 = table[k - ww - 1]; a1 = table[k - ww]; a2 = table[k - ww + 1]; }
          if (b2 > b0 || a2 > a0) {
            float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
            for (int i = a0; i < a2; ++i) {
              const int e = sidx[i];
              const int ql = e / kSlots;
              const float c = recf[(ql * RS + (e % kSlots)) * 4 + (i < a1 ? 3 : 2)];
              fma4(acc0, c, *reinterpret_cast<const float4*>(gtile + ql * 32 + 4 * j));
              fma4(acc1, c, *reinterpret_cast<const float4*>(gtile + ql * 32 + 16 + 4 * j));
            }
            for (int i = b0; i < b2; ++i) {
              const int e = sidx[i];
              const int ql = e / kSlots;
              const float c = recf[(ql * RS + (e % kSlots)) * 4 + (i < b1 ? 1 : 0)];
              fma4(acc0, c, *reinterpret_cast<const float4*>(gtile + ql * 32 + 4 * j));
              fma4(acc1, c, *reinterpret_cast<const float4*>(gtile + ql * 32 + 16 + 4 * j));
            }
            const int y = win.y0[l] + wy, x = win.x0[l] + wx;
            if (y >= 0 && y < H && x >= 0 && x < W) {   // an out-of-level pixel only ever collects zero coefficients
              float* pg = gvimg + (long long)(lt.start[l] + y * W + x) * px_stride;
              red_add_f4(pg, acc0);
              red_add_f4(pg + 16, acc1);
            }
          }
          wx += kTT / 4;
          while (wx >= ww) { wx -= ww; ++wy; }
        }
      }
    }
  }
}

template <int kSlots, bool kFused>
int launch_tile(cudaStream_t st, const float* grad_out, const float* value, const int64_t* shapes, const int64_t* lsi,
                const float* loc, const float* attn, int batch, int S, int L, float* grad_value, float* grad_loc,
                float* grad_attn, const float* ref) {
  auto kern = msda_bwd_tile_kernel<kSlots, kFused>;
  constexpr int smem = TileSmem<kSlots>::kTotal;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // shared memory for two CTAs, the rest of the 228 KB stays L1 for the corner gathers
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (2 * (smem + 2048) * 100 + 233471) / 233472));
    int b = 0;
    SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kTT, smem));
    blocks_per_sm = b > 0 ? b : 1;
  }
  const long long approx_items = (long long)batch * 8 * ((S + kTQ - 1) / kTQ + 4 * L);
  long long grid = (long long)sm_count() * blocks_per_sm;
  if (grid > approx_items) grid = approx_items;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kTT, smem, st>>>(grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                          grad_attn, ref);
  SDB_LAUNCH_CHECK("msda_bwd_tile_kernel");
  return SDB_OK;
}

}  // namespace

// Encoder self-attention backward (num_query == spatial_size), 8 heads x 32 channels x 4 points, L <= 8 levels,
// 16-byte aligned tensors, 32-bit image offsets -- the caller has checked; grad_value is already zero-filled on `st`.
// `ref == nullptr`: loc / attn are sampling locations / attention weights; otherwise the fused form (L <= 4, (N, Lq, L, 2)
// reference points).
int msda_backward_tile(cudaStream_t st, const float* grad_out, const float* value, const int64_t* shapes,
                       const int64_t* lsi, const float* loc, const float* attn, int batch, int S, int L,
                       float* grad_value, float* grad_loc, float* grad_attn, const float* ref) {
  if (ref != nullptr)
    return launch_tile<16, true>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                 grad_attn, ref);
  if (L <= 4)
    return launch_tile<16, false>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                  grad_attn, nullptr);
  return launch_tile<32, false>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                grad_attn, nullptr);
}

}  // namespace sdb
