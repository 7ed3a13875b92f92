// Decoder self-attention core, forward and backward, as one kernel each way.  sm_100a.
//
// The DINO decoder layer runs nn.MultiheadAttention over the ~1 100 content + denoising queries with the (T, T)
// denoising mask (/root/reference/detr_od/models/utils/transformer.py:765, 795-812; mask built by
// dense_heads/dn_components.py:97-113): per (image, head) scores = (q d^-1/2) k^T + mask, softmax over the keys,
// out = probs v.  Written out with library calls that is a (N*H, T, T) score tensor formed, normalised and consumed in
// three passes forward and six backward (~0.9 GB of HBM traffic per layer at T = 1 100).  Here the scores never leave
// the SM: one CTA owns 64 query rows of one (image, head), streams the keys / values through shared memory in tiles of
// 64 and keeps the running maximum / sum / output in registers; the backward recomputes the probabilities from the
// saved log-sum-exp (one kernel per gradient side: dq per query block, dk + dv per key block).
//
// Head dimension 32, fp32 storage.  The four products (q k^T, p v, and their transposes in the backward) are small
// dense contractions: warp-level `mma.sync.m16n8k8` TF32 with fp32 accumulation -- the same operand rounding the
// library products of this step use when torch's TF32 switch is on, which is the only mode the host layer routes here
// (layers/attention.py); tcgen05 would need 128-row tiles for 16 (image, head) problems of 1 100 rows.  The
// probabilities feed the second product straight from the accumulator registers: the accumulator fragment of an
// m16n8 tile holds columns (2t, 2t+1) per thread, the A fragment wants (t, t+4), so the key index inside each group
// of 8 is permuted (slot t <-> key 2t, slot t+4 <-> key 2t+1) on both operands -- a sum over keys does not care.
#include "common.cuh"

namespace sdb {

namespace {

constexpr int kHd = 32;        // head dimension
constexpr int kTile = 64;      // rows per CTA and columns per streamed tile
constexpr int kLd = 36;        // shared-memory row stride in words: conflict-free for both fragment patterns
constexpr int kAttThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// element (t, b, h, c) of a (T, B, H*32)-shaped operand lives at p[t * tok + b * bat + h * 32 + c]
struct View {
  float* p;
  long long tok, bat;
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 64 rows x 32 channels of one (image, head), rows r0.. of the operand, go into shared memory as TF32 words (x scale;
// rows past T are zero) in two steps, so that the global loads of tile i + 1 are in flight while tile i is computed on:
// fetch = this thread's 4 x 16 bytes into registers, stash = scale, round to TF32, store to shared memory.
__device__ __forceinline__ void fetch_tile(float4 (&r)[4], const float* base, long long tok, int r0, int T) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + u * kAttThreads;
    const int row = i >> 3, c4 = i & 7;
    r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + row < T) r[u] = __ldg(reinterpret_cast<const float4*>(base + (long long)(r0 + row) * tok + 4 * c4));
  }
}
__device__ __forceinline__ void stash_tile(uint32_t* s, const float4 (&r)[4], float scale) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.x + u * kAttThreads;
    const int row = i >> 3, c4 = i & 7;
    *reinterpret_cast<uint4*>(s + row * kLd + 4 * c4) =
        make_uint4(to_tf32(r[u].x * scale), to_tf32(r[u].y * scale), to_tf32(r[u].z * scale), to_tf32(r[u].w * scale));
  }
}
// first tile at or after c0 that is not fully masked (T when there is none)
__device__ __forceinline__ int next_live_tile(const unsigned char* flags, int row_block, int nb, int c0, int T,
                                              bool transposed) {
  if (flags)
    while (c0 < T && (flags[transposed ? (c0 / kTile) * nb + row_block : row_block * nb + c0 / kTile] & 2)) c0 += kTile;
  return c0;
}

// A fragments (16 rows x 32 channels, 4 k-steps) of the warp's own rows, straight from global memory
__device__ __forceinline__ void load_afrag(uint32_t (&a)[4][4], const float* base, long long tok, int row_lo, int T,
                                           int t, float scale) {
  const int row_hi = row_lo + 8;
  const float* p0 = base + (long long)row_lo * tok;
  const float* p1 = base + (long long)row_hi * tok;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = to_tf32(row_lo < T ? __ldg(p0 + 8 * ks + t) * scale : 0.f);
    a[ks][1] = to_tf32(row_hi < T ? __ldg(p1 + 8 * ks + t) * scale : 0.f);
    a[ks][2] = to_tf32(row_lo < T ? __ldg(p0 + 8 * ks + t + 4) * scale : 0.f);
    a[ks][3] = to_tf32(row_hi < T ? __ldg(p1 + 8 * ks + t + 4) * scale : 0.f);
  }
}

// c (16 x 64) = a (16 x 32) . s^T, s = 64 rows x 32 channels in shared memory
__device__ __forceinline__ void mm_nt(float (&c)[8][4], const uint32_t (&a)[4][4], const uint32_t* s, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    const uint32_t* r = s + (8 * nt + g) * kLd + t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma8(c[nt], a[ks], r[8 * ks], r[8 * ks + 4]);
  }
}

// acc (16 x 32) += p (16 x 64, accumulator layout) . s, s = 64 rows x 32 channels in shared memory
__device__ __forceinline__ void mm_pn(float (&acc)[4][4], const float (&p)[8][4], const uint32_t* s, int g, int t) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t a[4] = {to_tf32(p[ks][0]), to_tf32(p[ks][2]), to_tf32(p[ks][1]), to_tf32(p[ks][3])};
    const uint32_t* r0 = s + (8 * ks + 2 * t) * kLd + g;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma8(acc[nt], a, r0[8 * nt], r0[kLd + 8 * nt]);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---- forward ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttThreads)
mha_fwd_kernel(View q, View k, View v, const float* __restrict__ mask, const unsigned char* __restrict__ flags, int T,
               int H, float scale, float* __restrict__ out, float* __restrict__ lse) {
  __shared__ __align__(16) uint32_t Ks[kTile * kLd];
  __shared__ __align__(16) uint32_t Vs[kTile * kLd];
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * kTile + 16 * warp + g, row1 = row0 + 8;
  const float* qb = q.p + b * q.bat + h * kHd;
  const float* kb = k.p + b * k.bat + h * kHd;
  const float* vb = v.p + b * v.bat + h * kHd;
  uint32_t qa[4][4];
  load_afrag(qa, qb, q.tok, row0, T, t, scale);
  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max (log2 units) and this thread's share of the sum
  const float* mrow0 = mask ? mask + (long long)min(row0, T - 1) * T : nullptr;
  const float* mrow1 = mask ? mask + (long long)min(row1, T - 1) * T : nullptr;

  const int nb = (T + kTile - 1) / kTile;
  // tile flags (bit 0: some element of the tile carries a mask value, bit 1: every element is -inf): a fully masked
  // tile contributes nothing and is skipped, an unmasked one does not touch the mask tensor
  float4 kreg[4], vreg[4];
  int c0 = next_live_tile(flags, blockIdx.x, nb, 0, T, false);
  if (c0 < T) {
    fetch_tile(kreg, kb, k.tok, c0, T);
    fetch_tile(vreg, vb, v.tok, c0, T);
  }
  for (; c0 < T;) {
    const int f = flags ? flags[blockIdx.x * nb + c0 / kTile] : 1;
    const float* mr0 = (f & 1) ? mrow0 : nullptr;
    const float* mr1 = (f & 1) ? mrow1 : nullptr;
    __syncthreads();
    stash_tile(Ks, kreg, 1.f);
    stash_tile(Vs, vreg, 1.f);
    __syncthreads();
    const int cn = next_live_tile(flags, blockIdx.x, nb, c0 + kTile, T, false);
    if (cn < T) {                                  // next tile's loads fly under this tile's products
      fetch_tile(kreg, kb, k.tok, cn, T);
      fetch_tile(vreg, vb, v.tok, cn, T);
    }
    float s[8][4];
    mm_nt(s, qa, Ks, g, t);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = c0 + 8 * nt + 2 * t + e;
        const bool in = col < T;
        const float a0 = in ? (mr0 ? mr0[col] : 0.f) : -INFINITY;
        const float a1 = in ? (mr1 ? mr1[col] : 0.f) : -INFINITY;
        s[nt][e] = (s[nt][e] + a0) * kLog2e;
        s[nt][2 + e] = (s[nt][2 + e] + a1) * kLog2e;
        mx0 = fmaxf(mx0, s[nt][e]);
        mx1 = fmaxf(mx1, s[nt][2 + e]);
      }
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
    const float u0 = n0 == -INFINITY ? 0.f : n0, u1 = n1 == -INFINITY ? 0.f : n1;   // a row masked so far: exp2(-inf - 0) = 0
    const float f0 = exp2f(m0 - u0), f1 = exp2f(m1 - u1);
    m0 = n0;
    m1 = n1;
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - u0);
      s[nt][1] = exp2f(s[nt][1] - u0);
      s[nt][2] = exp2f(s[nt][2] - u1);
      s[nt][3] = exp2f(s[nt][3] - u1);
      r0 += s[nt][0] + s[nt][1];
      r1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * f0 + r0;
    l1 = l1 * f1 + r1;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      o[nt][0] *= f0;
      o[nt][1] *= f0;
      o[nt][2] *= f1;
      o[nt][3] *= f1;
    }
    mm_pn(o, s, Vs, g, t);
    c0 = cn;
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int C = H * kHd;
  if (row0 < T) {
    float* p = out + ((long long)row0 * gridDim.y / H + b) * C + h * kHd + 2 * t;   // (T, B, C): B = gridDim.y / H
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<float2*>(p + 8 * nt) = make_float2(o[nt][0] * i0, o[nt][1] * i0);
    if (t == 0) lse[(long long)bh * T + row0] = (m0 + log2f(l0)) * kLn2;
  }
  if (row1 < T) {
    float* p = out + ((long long)row1 * gridDim.y / H + b) * C + h * kHd + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<float2*>(p + 8 * nt) = make_float2(o[nt][2] * i1, o[nt][3] * i1);
    if (t == 0) lse[(long long)bh * T + row1] = (m1 + log2f(l1)) * kLn2;
  }
}

// ---- backward, query side: delta = rowsum(dout * out), dq = scale * (p o (dp - delta)) k -----------------------------
__global__ void __launch_bounds__(kAttThreads)
mha_bwd_dq_kernel(View q, View k, View v, const float* __restrict__ mask, const unsigned char* __restrict__ flags,
                  const float* __restrict__ out,
                  const float* __restrict__ dout, const float* __restrict__ lse, int T, int H, float scale, View dq,
                  float* __restrict__ delta) {
  __shared__ __align__(16) uint32_t Ks[kTile * kLd];
  __shared__ __align__(16) uint32_t Vs[kTile * kLd];
  const int bh = blockIdx.y, b = bh / H, h = bh % H, B = gridDim.y / H, C = H * kHd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * kTile + 16 * warp + g, row1 = row0 + 8;
  const float* qb = q.p + b * q.bat + h * kHd;
  const float* kb = k.p + b * k.bat + h * kHd;
  const float* vb = v.p + b * v.bat + h * kHd;
  const float* ob = out + (long long)b * C + h * kHd;
  const float* gb = dout + (long long)b * C + h * kHd;
  const long long otok = (long long)B * C;
  uint32_t qa[4][4], ga[4][4];
  load_afrag(qa, qb, q.tok, row0, T, t, scale);
  load_afrag(ga, gb, otok, row0, T, t, 1.f);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 8 * ks + t + 4 * e;
      if (row0 < T) d0 = fmaf(__ldg(gb + (long long)row0 * otok + c), __ldg(ob + (long long)row0 * otok + c), d0);
      if (row1 < T) d1 = fmaf(__ldg(gb + (long long)row1 * otok + c), __ldg(ob + (long long)row1 * otok + c), d1);
    }
  d0 = quad_sum(d0);
  d1 = quad_sum(d1);
  float e0 = 0.f, e1 = 0.f;   // log-sum-exp in log2 units
  if (row0 < T) {
    e0 = lse[(long long)bh * T + row0] * kLog2e;
    if (t == 0) delta[(long long)bh * T + row0] = d0;
  }
  if (row1 < T) {
    e1 = lse[(long long)bh * T + row1] * kLog2e;
    if (t == 0) delta[(long long)bh * T + row1] = d1;
  }
  const float* mrow0 = mask ? mask + (long long)min(row0, T - 1) * T : nullptr;
  const float* mrow1 = mask ? mask + (long long)min(row1, T - 1) * T : nullptr;
  float acc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;

  const int nb = (T + kTile - 1) / kTile;
  float4 kreg[4], vreg[4];
  int c0 = next_live_tile(flags, blockIdx.x, nb, 0, T, false);   // fully masked tiles: p = 0, no contribution to dq
  if (c0 < T) {
    fetch_tile(kreg, kb, k.tok, c0, T);
    fetch_tile(vreg, vb, v.tok, c0, T);
  }
  for (; c0 < T;) {
    const int f = flags ? flags[blockIdx.x * nb + c0 / kTile] : 1;
    const float* mr0 = (f & 1) ? mrow0 : nullptr;
    const float* mr1 = (f & 1) ? mrow1 : nullptr;
    __syncthreads();
    stash_tile(Ks, kreg, 1.f);
    stash_tile(Vs, vreg, 1.f);
    __syncthreads();
    const int cn = next_live_tile(flags, blockIdx.x, nb, c0 + kTile, T, false);
    if (cn < T) {
      fetch_tile(kreg, kb, k.tok, cn, T);
      fetch_tile(vreg, vb, v.tok, cn, T);
    }
    float s[8][4], dp[8][4];
    mm_nt(s, qa, Ks, g, t);
    mm_nt(dp, ga, Vs, g, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = c0 + 8 * nt + 2 * t + e;
        const bool in = col < T;
        const float a0 = in ? (mr0 ? mr0[col] : 0.f) : -INFINITY;
        const float a1 = in ? (mr1 ? mr1[col] : 0.f) : -INFINITY;
        const float p0 = exp2f((s[nt][e] + a0) * kLog2e - e0);
        const float p1 = exp2f((s[nt][2 + e] + a1) * kLog2e - e1);
        s[nt][e] = p0 * (dp[nt][e] - d0);
        s[nt][2 + e] = p1 * (dp[nt][2 + e] - d1);
      }
    mm_pn(acc, s, Ks, g, t);
    c0 = cn;
  }
  if (row0 < T) {
    float* p = dq.p + (long long)row0 * dq.tok + b * dq.bat + h * kHd + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<float2*>(p + 8 * nt) = make_float2(acc[nt][0] * scale, acc[nt][1] * scale);
  }
  if (row1 < T) {
    float* p = dq.p + (long long)row1 * dq.tok + b * dq.bat + h * kHd + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<float2*>(p + 8 * nt) = make_float2(acc[nt][2] * scale, acc[nt][3] * scale);
  }
}

// ---- backward, key side: everything transposed (rows = keys, columns = queries): dv = p^T dout, dk = ds^T (scale q) ----
__global__ void __launch_bounds__(kAttThreads)
mha_bwd_dkv_kernel(View q, View k, View v, const float* __restrict__ mask_t, const unsigned char* __restrict__ flags,
                   const float* __restrict__ dout,
                   const float* __restrict__ lse, const float* __restrict__ delta, int T, int H, float scale, View dk,
                   View dv) {
  __shared__ __align__(16) uint32_t Qs[kTile * kLd];
  __shared__ __align__(16) uint32_t Gs[kTile * kLd];
  __shared__ float Ls[kTile], Ds[kTile];
  const int bh = blockIdx.y, b = bh / H, h = bh % H, B = gridDim.y / H, C = H * kHd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * kTile + 16 * warp + g, row1 = row0 + 8;   // keys
  const float* qb = q.p + b * q.bat + h * kHd;
  const float* kb = k.p + b * k.bat + h * kHd;
  const float* vb = v.p + b * v.bat + h * kHd;
  const float* gb = dout + (long long)b * C + h * kHd;
  const long long otok = (long long)B * C;
  uint32_t ka[4][4], va[4][4];
  load_afrag(ka, kb, k.tok, row0, T, t, 1.f);
  load_afrag(va, vb, v.tok, row0, T, t, 1.f);
  const float* mrow0 = mask_t ? mask_t + (long long)min(row0, T - 1) * T : nullptr;
  const float* mrow1 = mask_t ? mask_t + (long long)min(row1, T - 1) * T : nullptr;
  float ak[4][4], av[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    ak[nt][0] = ak[nt][1] = ak[nt][2] = ak[nt][3] = 0.f;
    av[nt][0] = av[nt][1] = av[nt][2] = av[nt][3] = 0.f;
  }

  const int nb = (T + kTile - 1) / kTile;
  float4 qreg[4], greg[4];
  int c0 = next_live_tile(flags, blockIdx.x, nb, 0, T, true);     // flags are (query block, key block)
  if (c0 < T) {
    fetch_tile(qreg, qb, q.tok, c0, T);
    fetch_tile(greg, gb, otok, c0, T);
  }
  for (; c0 < T;) {   // query tiles
    const int f = flags ? flags[(c0 / kTile) * nb + blockIdx.x] : 1;
    const float* mr0 = (f & 1) ? mrow0 : nullptr;
    const float* mr1 = (f & 1) ? mrow1 : nullptr;
    __syncthreads();
    stash_tile(Qs, qreg, scale);
    stash_tile(Gs, greg, 1.f);
    if (threadIdx.x < kTile) {
      const int c = c0 + threadIdx.x;
      Ls[threadIdx.x] = c < T ? lse[(long long)bh * T + c] * kLog2e : 0.f;
      Ds[threadIdx.x] = c < T ? delta[(long long)bh * T + c] : 0.f;
    }
    __syncthreads();
    const int cn = next_live_tile(flags, blockIdx.x, nb, c0 + kTile, T, true);
    if (cn < T) {
      fetch_tile(qreg, qb, q.tok, cn, T);
      fetch_tile(greg, gb, otok, cn, T);
    }
    float s[8][4], dp[8][4];
    mm_nt(s, ka, Qs, g, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cl = 8 * nt + 2 * t + e, col = c0 + cl;
        const bool in = col < T;
        const float a0 = in ? (mr0 ? mr0[col] : 0.f) : -INFINITY;
        const float a1 = in ? (mr1 ? mr1[col] : 0.f) : -INFINITY;
        s[nt][e] = exp2f((s[nt][e] + a0) * kLog2e - Ls[cl]);
        s[nt][2 + e] = exp2f((s[nt][2 + e] + a1) * kLog2e - Ls[cl]);
      }
    mm_pn(av, s, Gs, g, t);
    mm_nt(dp, va, Gs, g, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float d = Ds[8 * nt + 2 * t + e];
        s[nt][e] *= dp[nt][e] - d;
        s[nt][2 + e] *= dp[nt][2 + e] - d;
      }
    mm_pn(ak, s, Qs, g, t);
    c0 = cn;
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int row = half ? row1 : row0;
    if (row < T) {
      float* pk = dk.p + (long long)row * dk.tok + b * dk.bat + h * kHd + 2 * t;
      float* pv = dv.p + (long long)row * dv.tok + b * dv.bat + h * kHd + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<float2*>(pk + 8 * nt) = make_float2(ak[nt][2 * half], ak[nt][2 * half + 1]);
        *reinterpret_cast<float2*>(pv + 8 * nt) = make_float2(av[nt][2 * half], av[nt][2 * half + 1]);
      }
    }
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool view_ok(const float* p, long long tok, long long bat) { return aligned16(p) && tok % 4 == 0 && bat % 4 == 0; }

}  // namespace

}  // namespace sdb

extern "C" {

int sdb_mha_forward_f32(sdb_stream_t stream, const float* q, int64_t q_tok, int64_t q_bat, const float* k,
                        int64_t k_tok, int64_t k_bat, const float* v, int64_t v_tok, int64_t v_bat,
                        const float* mask_add, const unsigned char* tile_flags, int T, int B, int H, int D, float scale,
                        float* out, float* lse) {
  using namespace sdb;
  SDB_REQUIRE(q && k && v && out && lse, "sdb_mha_forward_f32: null pointer");
  SDB_REQUIRE(T > 0 && B > 0 && H > 0 && (long long)B * H <= 65535, "sdb_mha_forward_f32: bad sizes T=%d B=%d H=%d", T, B, H);
  if (D != kHd) {
    set_error("sdb_mha_forward_f32: head dimension %d (built for 32)", D);
    return SDB_ERR_UNSUPPORTED;
  }
  SDB_REQUIRE(view_ok(q, q_tok, q_bat) && view_ok(k, k_tok, k_bat) && view_ok(v, v_tok, v_bat) && aligned16(out),
              "sdb_mha_forward_f32: operands must be 16-byte aligned with strides that are multiples of 4 floats");
  const dim3 grid((T + kTile - 1) / kTile, B * H);
  mha_fwd_kernel<<<grid, kAttThreads, 0, (cudaStream_t)stream>>>(
      View{const_cast<float*>(q), q_tok, q_bat}, View{const_cast<float*>(k), k_tok, k_bat},
      View{const_cast<float*>(v), v_tok, v_bat}, mask_add, mask_add ? tile_flags : nullptr, T, H, scale, out, lse);
  SDB_LAUNCH_CHECK("mha_fwd_kernel");
  return SDB_OK;
}

int sdb_mha_backward_f32(sdb_stream_t stream, const float* q, int64_t q_tok, int64_t q_bat, const float* k,
                         int64_t k_tok, int64_t k_bat, const float* v, int64_t v_tok, int64_t v_bat,
                         const float* mask_add, const float* mask_add_t, const unsigned char* tile_flags,
                         const float* out, const float* dout,
                         const float* lse, int T, int B, int H, int D, float scale, float* dq, int64_t dq_tok,
                         int64_t dq_bat, float* dk, int64_t dk_tok, int64_t dk_bat, float* dv, int64_t dv_tok,
                         int64_t dv_bat, float* delta) {
  using namespace sdb;
  SDB_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && delta, "sdb_mha_backward_f32: null pointer");
  SDB_REQUIRE((mask_add == nullptr) == (mask_add_t == nullptr),
              "sdb_mha_backward_f32: the mask and its transpose come together");
  SDB_REQUIRE(T > 0 && B > 0 && H > 0 && (long long)B * H <= 65535, "sdb_mha_backward_f32: bad sizes T=%d B=%d H=%d", T, B, H);
  if (D != kHd) {
    set_error("sdb_mha_backward_f32: head dimension %d (built for 32)", D);
    return SDB_ERR_UNSUPPORTED;
  }
  SDB_REQUIRE(view_ok(q, q_tok, q_bat) && view_ok(k, k_tok, k_bat) && view_ok(v, v_tok, v_bat) &&
                  view_ok(dq, dq_tok, dq_bat) && view_ok(dk, dk_tok, dk_bat) && view_ok(dv, dv_tok, dv_bat) &&
                  aligned16(out) && aligned16(dout),
              "sdb_mha_backward_f32: operands must be 16-byte aligned with strides that are multiples of 4 floats");
  const dim3 grid((T + kTile - 1) / kTile, B * H);
  const View Q{const_cast<float*>(q), q_tok, q_bat}, K{const_cast<float*>(k), k_tok, k_bat},
      V{const_cast<float*>(v), v_tok, v_bat};
  const unsigned char* fl = mask_add ? tile_flags : nullptr;
  mha_bwd_dq_kernel<<<grid, kAttThreads, 0, (cudaStream_t)stream>>>(Q, K, V, mask_add, fl, out, dout, lse, T, H, scale,
                                                                    View{dq, dq_tok, dq_bat}, delta);
  SDB_LAUNCH_CHECK("mha_bwd_dq_kernel");
  mha_bwd_dkv_kernel<<<grid, kAttThreads, 0, (cudaStream_t)stream>>>(Q, K, V, mask_add_t, fl, dout, lse, delta, T, H, scale,
                                                                     View{dk, dk_tok, dk_bat}, View{dv, dv_tok, dv_bat});
  SDB_LAUNCH_CHECK("mha_bwd_dkv_kernel");
  return SDB_OK;
}

}  // extern "C"
