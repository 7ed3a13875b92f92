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
// over (a level-0 tile: 4096 corner updates onto 200-550 distinct pixels).  This kernel turns the scatter into an
// on-chip gather: the VALUE PIXEL owns the work, not the query.
//
//   A  one thread per (query, level): bilinear taps of its 4 points -> record {pixel offset | corner bits, lh, lw, a};
//      the four corner coefficients a*w_k go to the point's result slots; every valid corner ("visit") is counted into
//      the bucket of the pixel it touches inside a per-level WINDOW around the tile's footprint (integer shared-memory
//      atomics, which are native -- fp32 shared atomics are CAS loops);
//   B  prefix sum over the buckets, each padded to whole TASKS of kT visits, and scatter of the visit codes: the tile's
//      visits sorted by pixel, cut into equal-sized tasks (a warp's 8 tasks run in lock step, no divergence);
//   D  one task per 4 lanes x 8 channels: load the task's value pixel ONCE, then for each visit read grad_out of the
//      visiting query from shared memory, accumulate coefficient x grad_out (grad_value) and reduce <grad_out, value>
//      (the corner dot product, which replaces the coefficient in the point's result slot); one reduction line per task;
//   X  points whose top-left pixel falls outside the window (offsets beyond +-kHalo pixels, levels finer than the
//      tile's own, or a tile whose visits overflow the tables) are handled the direct way: corner loads, dot
//      products, one reduction line per corner -- any input is handled, only the speed depends on locality;
//   F  one thread per point: grad_sampling_loc / grad_attn_weight from the four corner dot products (closed forms,
//      like the d32 kernel), in the fused form followed by the location chain rule and the softmax backward.
//
// grad_value is still accumulated with reductions (neighbouring tiles overlap), but ~4-5x fewer of them; no query ever
// gathers a corner line from global memory in the common case.
#include <algorithm>

#include "msda_common.cuh"

namespace sdb {

extern int g_bwd_variant;   // msda_backward.cu (sdb_msda_set_variant): 20 = no look-ahead prefetch, 21 / 22 / 23 = 2 / 3 / 4 tasks, 24 / 25 / 26 = windows for 1 / 2 / 0 finer levels

namespace {

constexpr int kTT = 256;          // threads per CTA
constexpr int kTH = 8, kTW = 8;   // query tile (pixels of the query's own level)
constexpr int kTQ = kTH * kTW;
constexpr int kHalo = 5;          // sampling offsets up to +-kHalo pixels stay inside the window
constexpr int kT = 4;             // visits per task
constexpr int kTableInts = 2048;  // bucket table incl. guard entries
constexpr int kMaxBuckets = kTableInts - 8;
constexpr int kVisitCap = 8192;   // visit slots (padded)
constexpr int kMaxWinLevels = 8;
constexpr unsigned kAllLanes = 0xffffffffu;
constexpr int kDefaultFinerLevels = 1;     // a tile also windows the next finer level when it fits (435 -> 419 us); variant 26: none
constexpr int kDefaultPrefetchAhead = 2;   // tasks of look-ahead for the L1 prefetch of value lines (447 -> 435 us; 0 = off)

struct Windows {
  int y0[kMaxWinLevels], x0[kMaxWinLevels], h[kMaxWinLevels], w[kMaxWinLevels], base[kMaxWinLevels];
  int magic[kMaxWinLevels];   // floor(p / w) == (p * magic) >> 16 for p < 2048, w <= 32
};

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float dot4acc(const float4& a, const float4& b, float s) {
  s = fmaf(a.x, b.x, s);
  s = fmaf(a.y, b.y, s);
  s = fmaf(a.z, b.z, s);
  return fmaf(a.w, b.w, s);
}
// shared-memory accesses through 32-bit shared-space addresses (no generic-to-shared conversion in the loop)
__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ float lds_f(unsigned a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ int lds_i(unsigned a) {
  int r;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts_f(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
// two fp32 lanes per instruction (FFMA2)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ void ffma2(f32x2& d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ float group4_sum(float v) {
  v += __shfl_xor_sync(kAllLanes, v, 2, 4);
  return v + __shfl_xor_sync(kAllLanes, v, 1, 4);
}
// Shared-memory layout (dynamic).  kSlots = point slots per (query, head): 16 (L <= 4), 20 (L = 5: two CTAs still fit
// an SM, which the 32-slot layout does not allow) or 32 (L <= 8).
template <int kSlots>
struct TileSmem {
  // point e lives at slot e + (e >> 3): a thread's 4 records are 64 contiguous bytes and the extra 16 bytes per 8
  // points spread a quarter-warp's 16-byte stores over all 32 banks
  static constexpr int kPointSlots = kTQ * kSlots + kTQ * kSlots / 8;
  static constexpr int kRecBytes = kPointSlots * 16;             // {offm, lh, lw, a} per point
  static constexpr int kResFloats = kPointSlots * 4 + 4;         // 4 result slots per point (+ 1 dummy, padded)
  static constexpr int kResBytes = kResFloats * 4;
  static constexpr int kGBytes = (kTQ + 1) * 32 * 4;             // grad_out rows of the tile (+ 1 zero row)
  static constexpr int kTableBytes = kTableInts * 4;
  static constexpr int kVisBytes = kVisitCap * 4;                // visit word: result slot byte offset | grad_out row byte offset << 16
  static constexpr int kKeyBytes = (kVisitCap / kT) * 4;         // value-pixel index of every task
  static constexpr int kDirectBytes = kTQ * kSlots * 2;
  static constexpr int oRes = kRecBytes, oG = oRes + kResBytes, oTable = oG + kGBytes, oVis = oTable + kTableBytes,
                       oKey = oVis + kVisBytes, oDirect = oKey + kKeyBytes;
  static constexpr int kTotal = oDirect + kDirectBytes;
};

// kFused: `loc` / `attn` are the RAW sampling offsets / attention logits, `ref` the (N, Lq, L, 2) reference points and
// the outputs are the gradients of the raw tensors (ms_deform_attn.py:98-105 differentiated here).
// V: storage type of `grad_out` and `value` (float, or __nv_bfloat16 widened to fp32 in registers: BASELINE.json
// configs[3]); locations, weights, every accumulation and all three gradients are fp32 either way.
template <int kSlots, bool kFused, typename V = float>
__global__ void __launch_bounds__(kTT, 2)
msda_bwd_tile_kernel(const V* __restrict__ grad_out, const V* __restrict__ value,
                     const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const float* __restrict__ loc, const float* __restrict__ attn, int batch, int S, int L,
                     float* __restrict__ grad_value, float* __restrict__ grad_loc, float* __restrict__ grad_attn,
                     const float* __restrict__ ref, int prefetch_ahead, int finer_levels) {
  constexpr int M = 8, P = 4;
  constexpr int px_stride = M * 32;
  constexpr int kLv = kSlots / P;                        // level slots per query
  constexpr int kQLPasses = (kTQ * kLv + kTT - 1) / kTT;   // (query, level) pairs per thread: 1 or 2
  constexpr bool kExact = kTQ * kLv % kTT == 0;          // otherwise the last pass has threads without a pair
  static_assert(!kFused || (kLv & (kLv - 1)) == 0, "the fused prologue reduces over kLv adjacent lanes");
  using SM = TileSmem<kSlots>;
  constexpr int kNullRes = SM::kPointSlots * 4;          // the null visit: a zero coefficient slot ...
  constexpr unsigned kNullVisit = (unsigned)(kNullRes * 4) | ((unsigned)(kTQ * 128) << 16);   // ... and the zero grad_out row
  static_assert(kNullRes * 4 + 4 <= 65536 && kTQ * 128 < 65536, "visit words hold 16-bit byte offsets");
  auto slot = [](int e) { return e + (e >> 3); };        // padded point index
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* rec = reinterpret_cast<float4*>(smem_raw);
  float* res = reinterpret_cast<float*>(smem_raw + SM::oRes);
  float* gtile = reinterpret_cast<float*>(smem_raw + SM::oG);
  int* table = reinterpret_cast<int*>(smem_raw + SM::oTable);
  unsigned* vis = reinterpret_cast<unsigned*>(smem_raw + SM::oVis);
  int* tpix = reinterpret_cast<int*>(smem_raw + SM::oKey);
  unsigned short* direct = reinterpret_cast<unsigned short*>(smem_raw + SM::oDirect);
  __shared__ LevelTable lt;
  __shared__ Windows win;
  __shared__ int warp_tot[kTT / 32];
  __shared__ int n_direct, n_slots, overflow;

  load_levels<kTH, kTW>(lt, shapes, lsi, L, px_stride);
  const int Lq = S;
  const int n_tiles = lt.tile_begin[L];
  const long long total = (long long)batch * n_tiles * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int LP = L * P;
  if (tid < 32) gtile[kTQ * 32 + tid] = 0.f;   // the null visit's grad_out row

  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int m = (int)(item % M);
    const long long t2 = item / M;
    const int tile = (int)(t2 % n_tiles);
    const int n = (int)(t2 / n_tiles);
    TileCursor<kTH, kTW> cur;
    cur.seek(lt, L, tile, true, Lq);
    const long long img = (long long)n * S * px_stride + m * 32;

    // ---- window geometry (thread 0), cleared tables, grad_out tile ------------------------------------------------
    __syncthreads();   // the previous item is finished with shared memory
    {
      int4* t4 = reinterpret_cast<int4*>(table);
      for (int i = tid; i < kTableInts / 4; i += kTT) t4[i] = make_int4(0, 0, 0, 0);
      // every visit slot starts as the null visit (padding of the last task of a pixel)
      uint4* v4 = reinterpret_cast<uint4*>(vis);
      for (int i = tid; i < kVisitCap / 4; i += kTT) v4[i] = make_uint4(kNullVisit, kNullVisit, kNullVisit, kNullVisit);
      const int ql = tid >> 2, j = tid & 3;    // grad_out rows: 4 lanes x 8 channels per query
      const int q = cur.query(ql, Lq);
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      if (q >= 0) {
        const V* gp = grad_out + (((long long)n * Lq + q) * M + m) * 32 + 8 * j;
        g0 = Chan4<V>::stream_in(gp);
        g1 = Chan4<V>::stream_in(gp + 4);
      }
      *reinterpret_cast<float4*>(gtile + ql * 32 + 8 * j) = g0;
      *reinterpret_cast<float4*>(gtile + ql * 32 + 8 * j + 4) = g1;
    }
    if (tid == 0) {
      int acc = 1;   // bucket 0 is a guard
      n_direct = 0;
      res[kNullRes] = 0.f;
      for (int l = L - 1; l >= 0; --l) {   // coarse levels first: smallest windows, most reuse
        const float sy = (float)lt.H[l] / (float)cur.Hl, sx = (float)lt.W[l] / (float)cur.Wl;
        const int wy0 = (int)floorf((float)cur.y0 * sy) - kHalo - 1;
        const int wx0 = (int)floorf((float)cur.x0 * sx) - kHalo - 1;
        const int wy1 = (int)ceilf((float)(cur.y0 + kTH) * sy) + kHalo + 1;
        const int wx1 = (int)ceilf((float)(cur.x0 + kTW) * sx) + kHalo + 1;
        const int wh = wy1 - wy0 + 1, ww = wx1 - wx0 + 1;
        // windows only where the tile's footprint is at most its own size (own level and coarser): finer levels
        // spread 64 queries over >= 256 pixels -- little to combine -- and take the direct path
        const bool use = l >= cur.lvl - finer_levels && l < kMaxWinLevels && ww <= 32 && acc + wh * ww <= kMaxBuckets;
        win.y0[l] = wy0;
        win.x0[l] = wx0;
        win.h[l] = use ? wh : 0;
        win.w[l] = use ? ww : 0;
        win.base[l] = acc;
        win.magic[l] = use ? 65536 / ww + 1 : 0;
        if (use) acc += wh * ww;
      }
    }
    __syncthreads();

    // ---- A: records, coefficients, visit counts --------------------------------------------------------------------
    unsigned short vkey[kQLPasses][P][4], vrank[kQLPasses][P][4];   // bucket / rank of every visit (0xffff: none)
#pragma unroll
    for (int ps = 0; ps < kQLPasses; ++ps) {
      const int u = tid + ps * kTT;
      const bool live = kExact || u < kTQ * kLv;        // a thread without a pair aliases pair 0 and writes nothing
      const int ql = live ? u / kLv : 0, lv = live ? u % kLv : 0;
      const int q = cur.query(ql, Lq);
      const bool on = live && q >= 0 && lv < L;
      const int lvl = min(lv, L - 1);
      const long long nq = (long long)n * Lq + (q >= 0 ? q : 0);
      const long long pair = nq * M + m;
      float xs[P], ys[P], as[P];
#pragma unroll
      for (int p = 0; p < P; ++p) { xs[p] = 0.f; ys[p] = 0.f; as[p] = 0.f; }
      const int H = lt.H[lvl], W = lt.W[lvl];
      if (kFused) {
        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
        const float dead = q >= 0 ? -INFINITY : 0.f;
        float4 lg = make_float4(dead, dead, dead, dead);
        if (on) {
          o0 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv));
          o1 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv + 4));
          lg = ld_stream_f4(reinterpret_cast<const float4*>(attn + pair * LP + 4 * lv));
        }
        // softmax over the query's L*P logits: the kLv threads of a query are adjacent lanes
        float mx = fmaxf(fmaxf(lg.x, lg.y), fmaxf(lg.z, lg.w));
#pragma unroll
        for (int o = kLv / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kAllLanes, mx, o, kLv));
        const float e0 = expf(lg.x - mx), e1 = expf(lg.y - mx), e2 = expf(lg.z - mx), e3 = expf(lg.w - mx);
        float sum = (e0 + e1) + (e2 + e3);
#pragma unroll
        for (int o = kLv / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(kAllLanes, sum, o, kLv);
        const float inv = 1.f / sum;
        if (on) {
          as[0] = e0 * inv; as[1] = e1 * inv; as[2] = e2 * inv; as[3] = e3 * inv;
          const float2 rp = __ldg(reinterpret_cast<const float2*>(ref + (nq * L + lvl) * 2));
          const float Wf = (float)W, Hf = (float)H;
          xs[0] = rp.x + o0.x / Wf; ys[0] = rp.y + o0.y / Hf;
          xs[1] = rp.x + o0.z / Wf; ys[1] = rp.y + o0.w / Hf;
          xs[2] = rp.x + o1.x / Wf; ys[2] = rp.y + o1.y / Hf;
          xs[3] = rp.x + o1.z / Wf; ys[3] = rp.y + o1.w / Hf;
        }
      } else if (on) {
        const float4 o0 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv));
        const float4 o1 = ld_stream_f4(reinterpret_cast<const float4*>(loc + pair * LP * 2 + 8 * lv + 4));
        const float4 aw = ld_stream_f4(reinterpret_cast<const float4*>(attn + pair * LP + 4 * lv));
        xs[0] = o0.x; ys[0] = o0.y; xs[1] = o0.z; ys[1] = o0.w;
        xs[2] = o1.x; ys[2] = o1.y; xs[3] = o1.z; ys[3] = o1.w;
        as[0] = aw.x; as[1] = aw.y; as[2] = aw.z; as[3] = aw.w;
      }
      const int ww = win.w[lvl], wh = win.h[lvl], wy0 = win.y0[lvl], wx0 = win.x0[lvl], wbase = win.base[lvl];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const Tap<float> t = make_tap<float>(xs[p], ys[p], H, W);
        int mask = on ? ((t.c00 ? 1 : 0) | (t.c01 ? 2 : 0) | (t.c10 ? 4 : 0) | (t.c11 ? 8 : 0)) : 0;
        const float a = as[p];
        const float hh = 1.f - t.lh, hw = 1.f - t.lw;
        const int e = ql * kSlots + lv * P + p;
#pragma unroll
        for (int r = 0; r < 4; ++r) { vkey[ps][p][r] = 0xffff; vrank[ps][p][r] = 0; }
        if (mask) {
          const int wy = t.h0 - wy0, wx = t.w0 - wx0;
          if (ww > 0 && wy >= 0 && wy < wh - 1 && wx >= 0 && wx < ww - 1) {
            const int k00 = wbase + wy * ww + wx;
#pragma unroll
            for (int r = 0; r < 4; ++r)
              if (mask & (1 << r)) {
                const int k = k00 + (r & 1) + (r >> 1) * ww;
                vkey[ps][p][r] = (unsigned short)k;
                vrank[ps][p][r] = (unsigned short)atomicAdd(&table[k + 1], 1);
              }
          } else {
            mask |= 16;   // outside the window: step X handles this point
            direct[atomicAdd(&n_direct, 1)] = (unsigned short)e;
          }
        }
        const int offm = ((lt.start[lvl] + t.h0 * W + t.w0) * px_stride) | mask;
        if (live) {
          rec[slot(e)] = make_float4(__int_as_float(offm), t.lh, t.lw, a);
          *reinterpret_cast<float4*>(res + 4 * slot(e)) =
              make_float4((mask & 1) ? hh * hw * a : 0.f, (mask & 2) ? hh * t.lw * a : 0.f,
                          (mask & 4) ? t.lh * hw * a : 0.f, (mask & 8) ? t.lh * t.lw * a : 0.f);
        }
      }
    }
    __syncthreads();

    // ---- B: counts -> padded to whole tasks -> inclusive scan in place: bucket k = [table[k], table[k+1]) ---------
    {
      int4* t4 = reinterpret_cast<int4*>(table);
      int4 v0 = t4[2 * tid], v1 = t4[2 * tid + 1];
      auto pad = [](int c) { return (c + kT - 1) & ~(kT - 1); };
      v0.x = pad(v0.x); v0.y = pad(v0.y) + v0.x; v0.z = pad(v0.z) + v0.y; v0.w = pad(v0.w) + v0.z;
      v1.x = pad(v1.x) + v0.w; v1.y = pad(v1.y) + v1.x; v1.z = pad(v1.z) + v1.y; v1.w = pad(v1.w) + v1.z;
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
      if (tid == kTT - 1) {
        n_slots = v1.w;
        overflow = v1.w > kVisitCap ? 1 : 0;
      }
    }
    __syncthreads();
    const bool all_direct = overflow != 0;
    if (!all_direct) {
#pragma unroll
      for (int ps = 0; ps < kQLPasses; ++ps) {
        const int u = tid + ps * kTT;
        const bool live = kExact || u < kTQ * kLv;      // (dead threads hold no visits: every vkey is 0xffff)
        const int ql = live ? u / kLv : 0, lv = live ? u % kLv : 0;
        const int lvl = min(lv, L - 1);
        const int ww = win.w[lvl], wbase = win.base[lvl], magic = win.magic[lvl];
        const int pix0 = lt.start[lvl] + win.y0[lvl] * lt.W[lvl] + win.x0[lvl], Wl = lt.W[lvl];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int k = vkey[ps][p][r];
            if (k != 0xffff) {
              const int pos = table[k] + vrank[ps][p][r];
              vis[pos] = (unsigned)(((slot(ql * kSlots + lv * P + p) << 2) | r) << 2) | ((unsigned)(ql * 128) << 16);
              const int pidx = k - wbase;
              const int wy = (pidx * magic) >> 16;
              tpix[pos / kT] = pix0 + wy * Wl + (pidx - wy * ww);   // same value from every visit of the task
            }
          }
      }
    }
    __syncthreads();

    // 4 lanes x 8 channels per task.  Lane j owns channels 4j..4j+3 and 16+4j..16+4j+3; the two 16-byte halves are
    // touched in opposite order by even and odd groups, so the two groups of a quarter-warp always read different
    // halves (banks 0-15 / 16-31) of their grad_out rows: no shared-memory bank conflicts whatever the rows.
    const int j = tid & 3;
    const int c0 = 4 * j + 16 * ((lane >> 2) & 1), c1 = 4 * j + 16 * (1 - ((lane >> 2) & 1));
    const V* vimg = value + img;
    float* gvimg = grad_value + img;

    // ---- D: tasks (pixel, <= kT visits) in contiguous runs, one run per 4 lanes ------------------------------------
    // Group G (4 lanes) owns tasks [G * per, G * per + per): consecutive tasks of one pixel -- every pixel with more than
    // kT visits -- meet in the same group, which keeps the value line and the accumulated grad_value line in
    // registers across them: one line load and one reduction line per PIXEL RUN instead of per task.  Loop bounds are
    // warp-uniform (every lane takes part in the width-4 shuffles); a group past its range runs null visits.  The next
    // distinct pixel's value line is fetched while the current task is processed.
    if (!all_direct) {
      const int n_tasks = n_slots / kT;
      // tasks per group, rounded up to an odd number: the 8 groups of a warp read their 16-byte visit words at a stride
      // of 16 * per bytes, which touches 8 different bank groups exactly when per is odd (an even per put the
      // visit-word loads at 2.9 wavefronts per instruction instead of 1)
      const int per = ((n_tasks + kTT / 4 - 1) / (kTT / 4)) | 1;
      int i = (tid >> 2) * per;
      const int i_end = min(i + per, n_tasks);
      const unsigned res_s = (unsigned)__cvta_generic_to_shared(res);
      const unsigned g_s = (unsigned)__cvta_generic_to_shared(gtile) + 4u * (unsigned)c0;
      const int g_d = 4 * (c1 - c0);                       // second half of the row: +-64 bytes
      const unsigned vis_s = (unsigned)__cvta_generic_to_shared(vis);
      const unsigned pix_s = (unsigned)__cvta_generic_to_shared(tpix);
      const bool hi2 = (j & 2) != 0, hi1 = (j & 1) != 0;
      int cur_pix = -1;
      int nxt_pix = i < i_end ? lds_i(pix_s + 4u * (unsigned)i) : -1;
      float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0;
      if (nxt_pix >= 0) {
        const V* pv = vimg + (long long)nxt_pix * px_stride;
        n0 = Chan4<V>::gather(pv + c0);
        n1 = Chan4<V>::gather(pv + c1);
      }
      f32x2 w[4] = {0ull, 0ull, 0ull, 0ull};               // value line of cur_pix: channels c0..c0+3, c1..c1+3
      f32x2 acc[4] = {0ull, 0ull, 0ull, 0ull};             // grad_value line of cur_pix
      auto flush = [&]() {
        float* pg = gvimg + (long long)cur_pix * px_stride;
        const float2 a0 = unpack2(acc[0]), a1 = unpack2(acc[1]), a2 = unpack2(acc[2]), a3 = unpack2(acc[3]);
        red_add_f4(pg + c0, make_float4(a0.x, a0.y, a1.x, a1.y));
        red_add_f4(pg + c1, make_float4(a2.x, a2.y, a3.x, a3.y));
      };
      for (int it = 0; it < per; ++it, ++i) {
        const bool valid = i < i_end;
        const int pix = nxt_pix;
        uint4 codes = make_uint4(kNullVisit, kNullVisit, kNullVisit, kNullVisit);
        unsigned my_code = kNullVisit;
        if (valid) {
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(codes.x), "=r"(codes.y), "=r"(codes.z), "=r"(codes.w)
                       : "r"(vis_s + 16u * (unsigned)i));
          my_code = (unsigned)lds_i(vis_s + 16u * (unsigned)i + 4u * (unsigned)j);
        }
        nxt_pix = i + 1 < i_end ? lds_i(pix_s + 4u * (unsigned)i + 4u) : -1;
        if (prefetch_ahead > 0 && j == 0 && i + prefetch_ahead < i_end) {
          // the value line a few tasks ahead goes to L1 now: the register load one task ahead then finds it there
          // instead of paying an L2 round trip at every change of pixel
          const int pf = lds_i(pix_s + 4u * (unsigned)(i + prefetch_ahead));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(vimg + (long long)pf * px_stride));
        }
        if (valid && pix != cur_pix) {                     // a new pixel run: hand over the finished line
          if (cur_pix >= 0) flush();
          acc[0] = acc[1] = acc[2] = acc[3] = 0ull;
          w[0] = pack2(n0.x, n0.y); w[1] = pack2(n0.z, n0.w);
          w[2] = pack2(n1.x, n1.y); w[3] = pack2(n1.z, n1.w);
          cur_pix = pix;
        }
        if (nxt_pix >= 0 && nxt_pix != pix) {
          const V* pv = vimg + (long long)nxt_pix * px_stride;
          n0 = Chan4<V>::gather(pv + c0);
          n1 = Chan4<V>::gather(pv + c1);
        }
        const unsigned cur_codes[kT] = {codes.x, codes.y, codes.z, codes.w};
        float d[kT];
#pragma unroll
        for (int s = 0; s < kT; ++s) {
          const float c = lds_f(res_s + (cur_codes[s] & 0xffffu));
          const unsigned ga = g_s + (cur_codes[s] >> 16);
          const float4 g0 = lds_f4(ga);
          const float4 g1 = lds_f4(ga + g_d);
          const f32x2 cc = pack2(c, c);
          const f32x2 p0 = pack2(g0.x, g0.y), p1 = pack2(g0.z, g0.w), p2 = pack2(g1.x, g1.y), p3 = pack2(g1.z, g1.w);
          ffma2(acc[0], cc, p0);
          ffma2(acc[1], cc, p1);
          ffma2(acc[2], cc, p2);
          ffma2(acc[3], cc, p3);
          f32x2 da = 0ull, db = 0ull;
          ffma2(da, p0, w[0]);
          ffma2(db, p1, w[1]);
          ffma2(da, p2, w[2]);
          ffma2(db, p3, w[3]);
          const float2 xa = unpack2(da), xb = unpack2(db);
          d[s] = (xa.x + xa.y) + (xb.x + xb.y);
        }
        // reduce-scatter of the 4 partial dot products over the 4 lanes: lane j ends with the full sum of visit j
        // (3 shuffles instead of 8) and stores it; the coefficient in that slot has been consumed above
        static_assert(kT == 4, "the reduce-scatter below is written for 4 visits per task");
        const float k0 = hi2 ? d[2] : d[0], k1 = hi2 ? d[3] : d[1];
        const float s0 = hi2 ? d[0] : d[2], s1 = hi2 ? d[1] : d[3];
        const float e0 = k0 + __shfl_xor_sync(kAllLanes, s0, 2, 4);   // visit (hi2 ? 2 : 0), over lane pairs
        const float e1 = k1 + __shfl_xor_sync(kAllLanes, s1, 2, 4);   // visit (hi2 ? 3 : 1)
        const float keep = hi1 ? e1 : e0, send = hi1 ? e0 : e1;
        const float full = keep + __shfl_xor_sync(kAllLanes, send, 1, 4);   // visit j
        sts_f(res_s + (my_code & 0xffffu), full);   // <grad_out, value_corner> (the null visit's slot just gets its 0 back)
      }
      if (cur_pix >= 0) flush();
    }

    // ---- X: points outside the windows -- corner loads, dot products, one reduction line per corner ----------------
    {
      const int nd = all_direct ? kTQ * kSlots : n_direct;
      for (int i0 = warp * 8; i0 < nd; i0 += kTT / 4) {
        const bool valid = i0 + (lane >> 2) < nd;
        const int i = valid ? i0 + (lane >> 2) : nd - 1;
        const int e = all_direct ? i : direct[i];
        const float4 R = rec[slot(e)];
        const int offm = valid ? __float_as_int(R.x) : 0;
        const int ql = e / kSlots;
        const int lvl = (e % kSlots) / P;
        const int ws = lt.wstr[lvl];
        const int off = offm & ~31;
        const V* pv = vimg + off;
        float* pg = gvimg + off;
        const float4 g0 = *reinterpret_cast<const float4*>(gtile + ql * 32 + c0);
        const float4 g1 = *reinterpret_cast<const float4*>(gtile + ql * 32 + c1);
        const float4 C = *reinterpret_cast<const float4*>(res + 4 * slot(e));
        float d[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          d[r] = 0.f;
          if (offm & (1 << r)) {
            const int o = (r & 1) * px_stride + (r >> 1) * ws;
            const float4 x0 = Chan4<V>::gather(pv + o + c0);
            const float4 x1 = Chan4<V>::gather(pv + o + c1);
            const float c = r == 0 ? C.x : (r == 1 ? C.y : (r == 2 ? C.z : C.w));
            red_add_f4(pg + o + c0, make_float4(c * g0.x, c * g0.y, c * g0.z, c * g0.w));
            red_add_f4(pg + o + c1, make_float4(c * g1.x, c * g1.y, c * g1.z, c * g1.w));
            d[r] = dot4acc(g1, x1, dot4acc(g0, x0, 0.f));
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) d[r] = group4_sum(d[r]);
        if (valid && (offm & 15) && j == 0)
          *reinterpret_cast<float4*>(res + 4 * slot(e)) = make_float4(d[0], d[1], d[2], d[3]);
      }
    }
    __syncthreads();

    // ---- F: grad_sampling_loc, grad_attn_weight from the corner dot products ----------------------------------------
#pragma unroll
    for (int ps = 0; ps < kQLPasses; ++ps) {
      const int u = tid + ps * kTT;
      const bool live = kExact || u < kTQ * kLv;
      const int ql = live ? u / kLv : 0, lv = live ? u % kLv : 0;
      const int q = cur.query(ql, Lq);
      const bool on = live && q >= 0 && lv < L;
      const int lvl = min(lv, L - 1);
      const long long pair = ((long long)n * Lq + (q >= 0 ? q : 0)) * M + m;
      const float Wf = (float)lt.W[lvl], Hf = (float)lt.H[lvl];
      float gx[P], gy[P], ga[P], aa[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const int e = ql * kSlots + lv * P + p;
        const float4 R = rec[slot(e)];
        const float4 D = *reinterpret_cast<const float4*>(res + 4 * slot(e));
        const int offm = __float_as_int(R.x);
        const float f0 = (offm & 1) ? D.x : 0.f, f1 = (offm & 2) ? D.y : 0.f;
        const float f2 = (offm & 4) ? D.z : 0.f, f3 = (offm & 8) ? D.w : 0.f;
        const float lh = R.y, lw = R.z, a = R.w;
        const float hh = 1.f - lh, hw = 1.f - lw;
        gx[p] = Wf * a * (hh * (f1 - f0) + lh * (f3 - f2));
        gy[p] = Hf * a * (hw * (f2 - f0) + lw * (f3 - f1));
        ga[p] = hh * (hw * f0 + lw * f1) + lh * (hw * f2 + lw * f3);
        aa[p] = a;
      }
      if (kFused) {
        // loc = ref + off / (W, H); softmax backward: dlogit_i = a_i * (ga_i - sum_k a_k ga_k)
        float part = (aa[0] * ga[0] + aa[1] * ga[1]) + (aa[2] * ga[2] + aa[3] * ga[3]);
#pragma unroll
        for (int o = kLv / 2; o > 0; o >>= 1) part += __shfl_xor_sync(kAllLanes, part, o, kLv);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          gx[p] *= 1.f / Wf;
          gy[p] *= 1.f / Hf;
          ga[p] = aa[p] * (ga[p] - part);
        }
      }
      if (on) {
        float* gl = grad_loc + (pair * LP + lv * P) * 2;
        st_stream_f4(reinterpret_cast<float4*>(gl), make_float4(gx[0], gy[0], gx[1], gy[1]));
        st_stream_f4(reinterpret_cast<float4*>(gl + 4), make_float4(gx[2], gy[2], gx[3], gy[3]));
        st_stream_f4(reinterpret_cast<float4*>(grad_attn + pair * LP + lv * P), make_float4(ga[0], ga[1], ga[2], ga[3]));
      }
    }
  }
}

template <int kSlots, bool kFused, typename V = float>
int launch_tile(cudaStream_t st, const V* grad_out, const V* value, const int64_t* shapes, const int64_t* lsi,
                const float* loc, const float* attn, int batch, int S, int L, float* grad_value, float* grad_loc,
                float* grad_attn, const float* ref) {
  auto kern = msda_bwd_tile_kernel<kSlots, kFused, V>;
  constexpr int smem = TileSmem<kSlots>::kTotal;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // shared memory for two CTAs; what is left of the 228 KB stays L1 for the value-pixel loads
    SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  (int)std::min(100LL, (2LL * (smem + 1024) * 100 + 233471) / 233472)));
    int b = 0;
    SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kTT, smem));
    blocks_per_sm = b > 0 ? b : 1;
  }
  const long long approx_items = (long long)batch * 8 * ((S + kTQ - 1) / kTQ + 4 * L);
  long long grid = (long long)sm_count() * blocks_per_sm;
  if (grid > approx_items) grid = approx_items;
  if (grid < 1) grid = 1;
  const int v = g_bwd_variant;
  const int ahead = v == 20 ? 0 : (v == 21 ? 2 : (v == 22 ? 3 : (v == 23 ? 4 : kDefaultPrefetchAhead)));
  const int finer = v == 24 ? 1 : (v == 25 ? 2 : (v == 26 ? 0 : kDefaultFinerLevels));   // windows also for levels finer than the tile's own
  kern<<<(unsigned)grid, kTT, smem, st>>>(grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                          grad_attn, ref, ahead, finer);
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
  if (L == 5)
    return launch_tile<20, false>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                  grad_attn, nullptr);
  return launch_tile<32, false>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value, grad_loc,
                                grad_attn, nullptr);
}

// bf16 storage of grad_out / value (the unfused form: locations and weights given): BASELINE.json configs[3], where the
// 5-level encoder backward is a fifth of the step
int msda_backward_tile_bf16(cudaStream_t st, const __nv_bfloat16* grad_out, const __nv_bfloat16* value,
                            const int64_t* shapes, const int64_t* lsi, const float* loc, const float* attn, int batch,
                            int S, int L, float* grad_value, float* grad_loc, float* grad_attn) {
  if (L <= 4)
    return launch_tile<16, false, __nv_bfloat16>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value,
                                                 grad_loc, grad_attn, nullptr);
  if (L == 5)
    return launch_tile<20, false, __nv_bfloat16>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value,
                                                 grad_loc, grad_attn, nullptr);
  return launch_tile<32, false, __nv_bfloat16>(st, grad_out, value, shapes, lsi, loc, attn, batch, S, L, grad_value,
                                               grad_loc, grad_attn, nullptr);
}

}  // namespace sdb
