// MSDeformAttn's softmax + sampling-location arithmetic as ONE pass each way, for shapes the fused MSDA kernels do not
// take (levels x points > 16, e.g. the 5-level model of BASELINE.json configs[3]).  sm_100a.
//
// Reference: /root/reference/detr_od/models/utils/ops/modules/ms_deform_attn.py:98-112
//   attention_weights = softmax(logits.view(N, Lq, M, L*P), -1)
//   ref_dim 2: loc = ref[:, :, None, :, None, :] + offsets / (W_l, H_l)
//   ref_dim 4: loc = ref[..., :2] + offsets / P * ref[..., 2:] * 0.5
// Written with tensor ops at the 5-scale shape that is a broadcast division, a broadcast add, a softmax and two dtype
// conversions over 57 M offsets per encoder layer, and as many passes again in the backward: 1.5 ms per layer.  Here one
// thread owns one (image, query, head): it reads the head's L*P offsets and logits (fp32 or bf16 storage, contiguous
// per thread and across the 8 heads of a query), keeps them in registers, and writes the fp32 locations and weights
// the MSDA kernels consume.  The backward maps grad_loc / grad_attn (fp32) to the gradients of the raw tensors in the
// storage type.  The reference points get no gradient (DINO detaches them, transformer.py:1030-1036).
#include "common.cuh"

namespace sdb {

namespace {

constexpr int kProThreads = 256;
constexpr int kProMaxLP = 32;     // levels x points per head (8 levels x 4 points)
constexpr int kProMaxL = 8;

struct ProShapes {
  float w[kProMaxL], h[kProMaxL];
};

// All per-thread traffic moves in groups of 4 elements (Chan4<V>: 16 bytes of fp32 or 8 bytes of bf16 per access, fp32
// outputs as float4): L*P and 2*L*P are multiples of 4 because P == 4.
template <typename V, int kP>
__global__ void __launch_bounds__(kProThreads)
msda_prologue_fwd_kernel(const V* __restrict__ offsets, const V* __restrict__ logits, const float* __restrict__ ref,
                         int ref_dim, ProShapes sh, long long pairs, int M, int L, float* __restrict__ loc,
                         float* __restrict__ attn) {
  static_assert(kP == 4, "one level = one group of 4 points");
  const long long i = (long long)blockIdx.x * kProThreads + threadIdx.x;   // (n, q, m)
  if (i >= pairs) return;
  const long long nq = i / M;
  const int LP = L * kP;
  const V* lg = logits + i * LP;
  float4 e[kProMaxL];
  float mx = -INFINITY;
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L) {
      e[l] = Chan4<V>::stream_in(lg + 4 * l);
      mx = fmaxf(mx, fmaxf(fmaxf(e[l].x, e[l].y), fmaxf(e[l].z, e[l].w)));
    }
  float sum = 0.f;
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L) {
      e[l] = make_float4(expf(e[l].x - mx), expf(e[l].y - mx), expf(e[l].z - mx), expf(e[l].w - mx));
      sum += (e[l].x + e[l].y) + (e[l].z + e[l].w);
    }
  const float inv = 1.f / sum;
  float4* ao = reinterpret_cast<float4*>(attn + i * LP);
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L) st_stream_f4(ao + l, make_float4(e[l].x * inv, e[l].y * inv, e[l].z * inv, e[l].w * inv));
  const V* of = offsets + i * LP * 2;
  float4* lo = reinterpret_cast<float4*>(loc + i * LP * 2);
  const float* rp = ref + nq * L * ref_dim;
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L) {
      const float rx = rp[l * ref_dim], ry = rp[l * ref_dim + 1];
      const float4 a = Chan4<V>::stream_in(of + 8 * l), b = Chan4<V>::stream_in(of + 8 * l + 4);   // x0 y0 x1 y1 | x2 y2 x3 y3
      float4 pa, pb;
      if (ref_dim == 2) {
        const float sx = sh.w[l], sy = sh.h[l];
        pa = make_float4(rx + a.x / sx, ry + a.y / sy, rx + a.z / sx, ry + a.w / sy);
        pb = make_float4(rx + b.x / sx, ry + b.y / sy, rx + b.z / sx, ry + b.w / sy);
      } else {   // offsets / P * ref_wh * 0.5, in the reference's order of operations
        const float sx = rp[l * ref_dim + 2], sy = rp[l * ref_dim + 3];
        pa = make_float4(rx + a.x / (float)kP * sx * 0.5f, ry + a.y / (float)kP * sy * 0.5f,
                         rx + a.z / (float)kP * sx * 0.5f, ry + a.w / (float)kP * sy * 0.5f);
        pb = make_float4(rx + b.x / (float)kP * sx * 0.5f, ry + b.y / (float)kP * sy * 0.5f,
                         rx + b.z / (float)kP * sx * 0.5f, ry + b.w / (float)kP * sy * 0.5f);
      }
      st_stream_f4(lo + 2 * l, pa);
      st_stream_f4(lo + 2 * l + 1, pb);
    }
}

template <typename V, int kP>
__global__ void __launch_bounds__(kProThreads)
msda_prologue_bwd_kernel(const float* __restrict__ grad_loc, const float* __restrict__ grad_attn,
                         const float* __restrict__ attn, const float* __restrict__ ref, int ref_dim, ProShapes sh,
                         long long pairs, int M, int L, V* __restrict__ grad_offsets, V* __restrict__ grad_logits) {
  static_assert(kP == 4, "one level = one group of 4 points");
  const long long i = (long long)blockIdx.x * kProThreads + threadIdx.x;
  if (i >= pairs) return;
  const long long nq = i / M;
  const int LP = L * kP;
  const float4* a4 = reinterpret_cast<const float4*>(attn + i * LP);
  const float4* g4 = reinterpret_cast<const float4*>(grad_attn + i * LP);
  float4 av[kProMaxL], gv[kProMaxL];
  float dot = 0.f;
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L) {
      av[l] = ld_stream_f4(a4 + l);
      gv[l] = ld_stream_f4(g4 + l);
      dot = fmaf(av[l].x, gv[l].x, dot);
      dot = fmaf(av[l].y, gv[l].y, dot);
      dot = fmaf(av[l].z, gv[l].z, dot);
      dot = fmaf(av[l].w, gv[l].w, dot);
    }
  V* gl = grad_logits + i * LP;
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L)   // softmax backward
      Chan4<V>::stream_out(gl + 4 * l, make_float4(av[l].x * (gv[l].x - dot), av[l].y * (gv[l].y - dot),
                                                   av[l].z * (gv[l].z - dot), av[l].w * (gv[l].w - dot)));
  const float4* gloc = reinterpret_cast<const float4*>(grad_loc + i * LP * 2);
  V* go = grad_offsets + i * LP * 2;
  const float* rp = ref + nq * L * ref_dim;
#pragma unroll
  for (int l = 0; l < kProMaxL; ++l)
    if (l < L) {
      float fx, fy;
      if (ref_dim == 2) {
        fx = 1.f / sh.w[l];
        fy = 1.f / sh.h[l];
      } else {
        fx = rp[l * ref_dim + 2] * 0.5f / (float)kP;
        fy = rp[l * ref_dim + 3] * 0.5f / (float)kP;
      }
      const float4 a = ld_stream_f4(gloc + 2 * l), b = ld_stream_f4(gloc + 2 * l + 1);
      Chan4<V>::stream_out(go + 8 * l, make_float4(a.x * fx, a.y * fy, a.z * fx, a.w * fy));
      Chan4<V>::stream_out(go + 8 * l + 4, make_float4(b.x * fx, b.y * fy, b.z * fx, b.w * fy));
    }
}

int fill_shapes(const int64_t* shapes_host, int L, ProShapes& sh) {
  for (int l = 0; l < L; ++l) {
    sh.h[l] = (float)shapes_host[2 * l];
    sh.w[l] = (float)shapes_host[2 * l + 1];
  }
  return SDB_OK;
}

template <typename V>
int prologue_forward(cudaStream_t st, const V* offsets, const V* logits, const float* ref, int ref_dim,
                     const int64_t* shapes_host, int batch, int Lq, int M, int L, int P, float* loc, float* attn) {
  SDB_REQUIRE(batch >= 0 && Lq >= 0 && M > 0 && L > 0 && L <= kProMaxL && P == 4 && L * P <= kProMaxLP,
              "msda_prologue_forward: bad sizes batch=%d query=%d heads=%d levels=%d points=%d", batch, Lq, M, L, P);
  SDB_REQUIRE(ref_dim == 2 || ref_dim == 4, "msda_prologue_forward: ref_dim=%d (2 or 4)", ref_dim);
  const long long pairs = (long long)batch * Lq * M;
  if (pairs == 0) return SDB_OK;
  SDB_REQUIRE(offsets && logits && ref && shapes_host && loc && attn, "msda_prologue_forward: null pointer");
  ProShapes sh{};
  fill_shapes(shapes_host, L, sh);
  const long long grid = (pairs + kProThreads - 1) / kProThreads;
  msda_prologue_fwd_kernel<V, 4><<<(unsigned)grid, kProThreads, 0, st>>>(offsets, logits, ref, ref_dim, sh, pairs, M, L,
                                                                         loc, attn);
  SDB_LAUNCH_CHECK("msda_prologue_fwd_kernel");
  return SDB_OK;
}

template <typename V>
int prologue_backward(cudaStream_t st, const float* grad_loc, const float* grad_attn, const float* attn, const float* ref,
                      int ref_dim, const int64_t* shapes_host, int batch, int Lq, int M, int L, int P, V* grad_offsets,
                      V* grad_logits) {
  SDB_REQUIRE(batch >= 0 && Lq >= 0 && M > 0 && L > 0 && L <= kProMaxL && P == 4 && L * P <= kProMaxLP,
              "msda_prologue_backward: bad sizes batch=%d query=%d heads=%d levels=%d points=%d", batch, Lq, M, L, P);
  SDB_REQUIRE(ref_dim == 2 || ref_dim == 4, "msda_prologue_backward: ref_dim=%d (2 or 4)", ref_dim);
  const long long pairs = (long long)batch * Lq * M;
  if (pairs == 0) return SDB_OK;
  SDB_REQUIRE(grad_loc && grad_attn && attn && ref && shapes_host && grad_offsets && grad_logits,
              "msda_prologue_backward: null pointer");
  ProShapes sh{};
  fill_shapes(shapes_host, L, sh);
  const long long grid = (pairs + kProThreads - 1) / kProThreads;
  msda_prologue_bwd_kernel<V, 4><<<(unsigned)grid, kProThreads, 0, st>>>(grad_loc, grad_attn, attn, ref, ref_dim, sh, pairs,
                                                                         M, L, grad_offsets, grad_logits);
  SDB_LAUNCH_CHECK("msda_prologue_bwd_kernel");
  return SDB_OK;
}

}  // namespace

}  // namespace sdb

extern "C" {

int sdb_msda_prologue_forward_f32(sdb_stream_t stream, const float* offsets, const float* logits, const float* ref,
                                  int ref_dim, const int64_t* spatial_shapes_host, int batch, int num_query,
                                  int num_heads, int num_levels, int num_point, float* sampling_loc, float* attn_weight) {
  return sdb::prologue_forward<float>((cudaStream_t)stream, offsets, logits, ref, ref_dim, spatial_shapes_host, batch,
                                      num_query, num_heads, num_levels, num_point, sampling_loc, attn_weight);
}
int sdb_msda_prologue_forward_bf16(sdb_stream_t stream, const uint16_t* offsets, const uint16_t* logits, const float* ref,
                                   int ref_dim, const int64_t* spatial_shapes_host, int batch, int num_query,
                                   int num_heads, int num_levels, int num_point, float* sampling_loc,
                                   float* attn_weight) {
  return sdb::prologue_forward<__nv_bfloat16>((cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(offsets),
                                              reinterpret_cast<const __nv_bfloat16*>(logits), ref, ref_dim,
                                              spatial_shapes_host, batch, num_query, num_heads, num_levels, num_point,
                                              sampling_loc, attn_weight);
}
int sdb_msda_prologue_backward_f32(sdb_stream_t stream, const float* grad_loc, const float* grad_attn,
                                   const float* attn_weight, const float* ref, int ref_dim,
                                   const int64_t* spatial_shapes_host, int batch, int num_query, int num_heads,
                                   int num_levels, int num_point, float* grad_offsets, float* grad_logits) {
  return sdb::prologue_backward<float>((cudaStream_t)stream, grad_loc, grad_attn, attn_weight, ref, ref_dim,
                                       spatial_shapes_host, batch, num_query, num_heads, num_levels, num_point,
                                       grad_offsets, grad_logits);
}
int sdb_msda_prologue_backward_bf16(sdb_stream_t stream, const float* grad_loc, const float* grad_attn,
                                    const float* attn_weight, const float* ref, int ref_dim,
                                    const int64_t* spatial_shapes_host, int batch, int num_query, int num_heads,
                                    int num_levels, int num_point, uint16_t* grad_offsets, uint16_t* grad_logits) {
  return sdb::prologue_backward<__nv_bfloat16>((cudaStream_t)stream, grad_loc, grad_attn, attn_weight, ref, ref_dim,
                                               spatial_shapes_host, batch, num_query, num_heads, num_levels, num_point,
                                               reinterpret_cast<__nv_bfloat16*>(grad_offsets),
                                               reinterpret_cast<__nv_bfloat16*>(grad_logits));
}

}  // extern "C"
