// LayerNorm over the channel dimension of the transformer (d_model = 256), forward and backward.  sm_100a.
//
// The DINO encoder / decoder apply nn.LayerNorm after every attention and FFN block
// (/root/reference/detr_od/models/utils/transformer.py:606-642, 762-791, 1039): 37 calls per pass over up to
// 44 446 tokens x 256 channels.  torch's backward for the affine parameters (GammaBetaBackward) is the single
// largest non-GEMM kernel of the step on B200 (4.5 ms / step); this pair is HBM-bound instead:
//   forward : read x, write y (+ mean, rstd)                     8 B / element
//   backward: read dy, x, write dx; dgamma / dbeta are summed in registers over a persistent row loop, reduced
//             across the CTA in shared memory and finished by a tiny second kernel      12 B / element
// One warp owns one row: lane j holds channels 4j..4j+3 and 128+4j..128+4j+3 (two coalesced 16-byte accesses).
// Statistics follow torch (biased variance, eps inside the sqrt, fp32).
#include "common.cuh"

namespace sdb {

constexpr int kLnCols = 256;
constexpr int kLnThreads = 256;           // 8 warps = 8 rows in flight per CTA
constexpr int kLnMaxGrid = 148 * 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kLnThreads)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
                     const float* __restrict__ beta, long long rows, float eps, float* __restrict__ y,
                     float* __restrict__ mean, float* __restrict__ rstd, const float* __restrict__ pos,
                     float* __restrict__ y_plus_pos) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * kLnThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kLnThreads) >> 5;
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane], g1 = reinterpret_cast<const float4*>(gamma)[32 + lane];
  const float4 b0 = reinterpret_cast<const float4*>(beta)[lane], b1 = reinterpret_cast<const float4*>(beta)[32 + lane];
  for (long long r = warp; r < rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * kLnCols);
    float4 a = ld_stream_f4(xr + lane), b = ld_stream_f4(xr + 32 + lane);
    if (residual) {   // y = LN(x + residual): the post-norm residual add of the layer rides in this pass
      const float4* rr = reinterpret_cast<const float4*>(residual + r * kLnCols);
      const float4 ra = ld_stream_f4(rr + lane), rb = ld_stream_f4(rr + 32 + lane);
      a = make_float4(a.x + ra.x, a.y + ra.y, a.z + ra.z, a.w + ra.w);
      b = make_float4(b.x + rb.x, b.y + rb.y, b.z + rb.z, b.w + rb.w);
    }
    const float mu = warp_sum(a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w) * (1.f / kLnCols);
    const float4 da = make_float4(a.x - mu, a.y - mu, a.z - mu, a.w - mu);
    const float4 db = make_float4(b.x - mu, b.y - mu, b.z - mu, b.w - mu);
    const float var = warp_sum(da.x * da.x + da.y * da.y + da.z * da.z + da.w * da.w + db.x * db.x + db.y * db.y +
                               db.z * db.z + db.w * db.w) * (1.f / kLnCols);
    const float rs = rsqrtf(var + eps);
    float4* yr = reinterpret_cast<float4*>(y + r * kLnCols);
    const float4 ya = make_float4(da.x * rs * g0.x + b0.x, da.y * rs * g0.y + b0.y, da.z * rs * g0.z + b0.z,
                                  da.w * rs * g0.w + b0.w);
    const float4 yb = make_float4(db.x * rs * g1.x + b1.x, db.y * rs * g1.y + b1.y, db.z * rs * g1.z + b1.z,
                                  db.w * rs * g1.w + b1.w);
    st_stream_f4(yr + lane, ya);
    st_stream_f4(yr + 32 + lane, yb);
    if (y_plus_pos) {   // the next layer's query = output + positional embedding, written from registers
      const float4* pr = reinterpret_cast<const float4*>(pos + r * kLnCols);
      const float4 pa = ld_stream_f4(pr + lane), pb = ld_stream_f4(pr + 32 + lane);
      float4* qr = reinterpret_cast<float4*>(y_plus_pos + r * kLnCols);
      st_stream_f4(qr + lane, make_float4(ya.x + pa.x, ya.y + pa.y, ya.z + pa.z, ya.w + pa.w));
      st_stream_f4(qr + 32 + lane, make_float4(yb.x + pb.x, yb.y + pb.y, yb.z + pb.z, yb.w + pb.w));
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

__global__ void __launch_bounds__(kLnThreads)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2, const float* __restrict__ x,
                     const float* __restrict__ residual, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, long long rows,
                     float* __restrict__ dx, float* __restrict__ partial /* [grid][2][256] */) {
  __shared__ float s_acc[kLnThreads / 32][2][kLnCols];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warp = ((long long)blockIdx.x * kLnThreads + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * kLnThreads) >> 5;
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane], g1 = reinterpret_cast<const float4*>(gamma)[32 + lane];
  float dg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = warp; r < rows; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * kLnCols);
    const float4* dr = reinterpret_cast<const float4*>(dy + r * kLnCols);
    float4 xa = ld_stream_f4(xr + lane), xb = ld_stream_f4(xr + 32 + lane);
    float4 da = ld_stream_f4(dr + lane), db = ld_stream_f4(dr + 32 + lane);
    if (residual) {   // the normalised input was x + residual
      const float4* rr = reinterpret_cast<const float4*>(residual + r * kLnCols);
      const float4 ra = ld_stream_f4(rr + lane), rb = ld_stream_f4(rr + 32 + lane);
      xa = make_float4(xa.x + ra.x, xa.y + ra.y, xa.z + ra.z, xa.w + ra.w);
      xb = make_float4(xb.x + rb.x, xb.y + rb.y, xb.z + rb.z, xb.w + rb.w);
    }
    if (dy2) {        // gradient of the second output (y + pos) adds to the gradient of y
      const float4* d2 = reinterpret_cast<const float4*>(dy2 + r * kLnCols);
      const float4 ea = ld_stream_f4(d2 + lane), eb = ld_stream_f4(d2 + 32 + lane);
      da = make_float4(da.x + ea.x, da.y + ea.y, da.z + ea.z, da.w + ea.w);
      db = make_float4(db.x + eb.x, db.y + eb.y, db.z + eb.z, db.w + eb.w);
    }
    const float mu = mean[r], rs = rstd[r];
    const float xh[8] = {(xa.x - mu) * rs, (xa.y - mu) * rs, (xa.z - mu) * rs, (xa.w - mu) * rs,
                         (xb.x - mu) * rs, (xb.y - mu) * rs, (xb.z - mu) * rs, (xb.w - mu) * rs};
    const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float s1 = 0.f, s2 = 0.f, gd[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      gd[k] = d[k] * gm[k];
      s1 += gd[k];
      s2 += gd[k] * xh[k];
      dg[k] += d[k] * xh[k];
      dbt[k] += d[k];
    }
    s1 = warp_sum(s1) * (1.f / kLnCols);
    s2 = warp_sum(s2) * (1.f / kLnCols);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = rs * (gd[k] - s1 - xh[k] * s2);
    float4* ox = reinterpret_cast<float4*>(dx + r * kLnCols);
    st_stream_f4(ox + lane, make_float4(o[0], o[1], o[2], o[3]));
    st_stream_f4(ox + 32 + lane, make_float4(o[4], o[5], o[6], o[7]));
  }
  // CTA reduction of the affine-parameter partials, then one row of partials per CTA
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s_acc[wid][0][4 * lane + k] = dg[k];
    s_acc[wid][0][128 + 4 * lane + k] = dg[4 + k];
    s_acc[wid][1][4 * lane + k] = dbt[k];
    s_acc[wid][1][128 + 4 * lane + k] = dbt[4 + k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * kLnCols; c += kLnThreads) {
    const int which = c / kLnCols, col = c % kLnCols;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) v += s_acc[w][which][col];
    partial[((long long)blockIdx.x * 2 + which) * kLnCols + col] = v;
  }
}

__global__ void __launch_bounds__(256)
layernorm_bwd_finish_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ dgamma,
                            float* __restrict__ dbeta) {
  // grid = 2 * 256 / 8 blocks; each warp sums one column over the partial rows
  const int lane = threadIdx.x & 31;
  const int col2 = blockIdx.x * 8 + (threadIdx.x >> 5);       // 0 .. 511  (which * 256 + col)
  const int which = col2 / kLnCols, col = col2 % kLnCols;
  float v = 0.f;
  for (int b = lane; b < nblocks; b += 32) v += partial[((long long)b * 2 + which) * kLnCols + col];
  v = warp_sum(v);
  if (lane == 0) (which == 0 ? dgamma : dbeta)[col] = v;
}

static int ln_grid(long long rows) {
  long long g = (rows + (kLnThreads / 32) - 1) / (kLnThreads / 32);
  const long long cap = (long long)sm_count() * 8 < kLnMaxGrid ? (long long)sm_count() * 8 : kLnMaxGrid;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace sdb

extern "C" int sdb_layernorm_bwd_workspace_floats(void) { return sdb::kLnMaxGrid * 2 * sdb::kLnCols; }

extern "C" int sdb_layernorm_forward_f32(sdb_stream_t stream, const float* x, const float* gamma, const float* beta,
                                         int64_t rows, int cols, float eps, float* y, float* mean, float* rstd) {
  SDB_REQUIRE(rows >= 0 && cols > 0, "layernorm_forward: bad sizes rows=%lld cols=%d", (long long)rows, cols);
  if (cols != sdb::kLnCols) {
    sdb::set_error("layernorm_forward: cols=%d (only d_model=%d is built)", cols, sdb::kLnCols);
    return SDB_ERR_UNSUPPORTED;
  }
  if (rows == 0) return SDB_OK;
  SDB_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_forward: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "layernorm_forward: pointers must be 16-byte aligned");
  sdb::layernorm_fwd_kernel<<<sdb::ln_grid(rows), sdb::kLnThreads, 0, (cudaStream_t)stream>>>(
      x, nullptr, gamma, beta, rows, eps, y, mean, rstd, nullptr, nullptr);
  SDB_LAUNCH_CHECK("layernorm_fwd_kernel");
  return SDB_OK;
}

extern "C" int sdb_add_layernorm_forward_f32(sdb_stream_t stream, const float* x, const float* residual,
                                             const float* gamma, const float* beta, int64_t rows, int cols, float eps,
                                             float* y, float* mean, float* rstd, const float* pos, float* y_plus_pos) {
  SDB_REQUIRE(rows >= 0 && cols > 0, "add_layernorm_forward: bad sizes rows=%lld cols=%d", (long long)rows, cols);
  if (cols != sdb::kLnCols) {
    sdb::set_error("add_layernorm_forward: cols=%d (only d_model=%d is built)", cols, sdb::kLnCols);
    return SDB_ERR_UNSUPPORTED;
  }
  if (rows == 0) return SDB_OK;
  SDB_REQUIRE(x && gamma && beta && y && mean && rstd, "add_layernorm_forward: null pointer");
  SDB_REQUIRE((pos == nullptr) == (y_plus_pos == nullptr), "add_layernorm_forward: pos and y_plus_pos go together");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(residual) |
                reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(y_plus_pos)) & 15) == 0,
              "add_layernorm_forward: pointers must be 16-byte aligned");
  sdb::layernorm_fwd_kernel<<<sdb::ln_grid(rows), sdb::kLnThreads, 0, (cudaStream_t)stream>>>(
      x, residual, gamma, beta, rows, eps, y, mean, rstd, pos, y_plus_pos);
  SDB_LAUNCH_CHECK("layernorm_fwd_kernel");
  return SDB_OK;
}

static int layernorm_backward(sdb_stream_t stream, const float* dy, const float* dy2, const float* x,
                              const float* residual, const float* gamma, const float* mean, const float* rstd,
                              int64_t rows, int cols, float* dx, float* dgamma, float* dbeta, float* workspace) {
  SDB_REQUIRE(rows >= 0 && cols > 0, "layernorm_backward: bad sizes rows=%lld cols=%d", (long long)rows, cols);
  if (cols != sdb::kLnCols) {
    sdb::set_error("layernorm_backward: cols=%d (only d_model=%d is built)", cols, sdb::kLnCols);
    return SDB_ERR_UNSUPPORTED;
  }
  SDB_REQUIRE(dgamma && dbeta && workspace, "layernorm_backward: null pointer");
  SDB_REQUIRE(rows == 0 || (dy && x && gamma && mean && rstd && dx), "layernorm_backward: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dy2) |
                reinterpret_cast<uintptr_t>(residual)) & 15) == 0, "layernorm_backward: pointers must be 16-byte aligned");
  const int grid = sdb::ln_grid(rows);
  sdb::layernorm_bwd_kernel<<<grid, sdb::kLnThreads, 0, (cudaStream_t)stream>>>(dy, dy2, x, residual, gamma, mean, rstd,
                                                                                rows, dx, workspace);
  SDB_LAUNCH_CHECK("layernorm_bwd_kernel");
  sdb::layernorm_bwd_finish_kernel<<<2 * sdb::kLnCols / 8, 256, 0, (cudaStream_t)stream>>>(workspace, grid, dgamma,
                                                                                          dbeta);
  SDB_LAUNCH_CHECK("layernorm_bwd_finish_kernel");
  return SDB_OK;
}

extern "C" int sdb_layernorm_backward_f32(sdb_stream_t stream, const float* dy, const float* x, const float* gamma,
                                          const float* mean, const float* rstd, int64_t rows, int cols, float* dx,
                                          float* dgamma, float* dbeta, float* workspace) {
  return layernorm_backward(stream, dy, nullptr, x, nullptr, gamma, mean, rstd, rows, cols, dx, dgamma, dbeta, workspace);
}

extern "C" int sdb_add_layernorm_backward_f32(sdb_stream_t stream, const float* dy, const float* dy_plus_pos,
                                              const float* x, const float* residual, const float* gamma,
                                              const float* mean, const float* rstd, int64_t rows, int cols, float* dx,
                                              float* dgamma, float* dbeta, float* workspace) {
  return layernorm_backward(stream, dy, dy_plus_pos, x, residual, gamma, mean, rstd, rows, cols, dx, dgamma, dbeta,
                            workspace);
}
