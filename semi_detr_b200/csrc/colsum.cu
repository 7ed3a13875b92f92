// Column sums of a row-major (rows x cols) fp32 matrix: the bias gradient of every nn.Linear on the path
// (grad_bias = grad_output.sum(0)).  sm_100a.
//
// In the DINO step these reductions run over 44 446 tokens x {128, 256, 2048} channels for each of the encoder /
// decoder linears (/root/reference/detr_od/models/utils/transformer.py:596-630, 765-791;
// ops/modules/ms_deform_attn.py:52-55); torch's generic reduce kernel made them the largest non-GEMM item of the
// step on B200 (3.7 ms / step).  This kernel is a single streaming pass: each CTA owns a 128-column strip and a
// slice of the rows, every thread keeps a float4 of running sums (so a warp reads 512 contiguous bytes per row),
// 8 row-lanes per CTA are combined in shared memory and CTAs along the row direction are combined with one
// atomicAdd per column.  4 B read per element: HBM-bound.
#include "common.cuh"

namespace sdb {

constexpr int kCsThreads = 256;          // 32 column-lanes (x float4 = 128 columns) x 8 row-lanes
constexpr int kCsCols = 128;

// V: storage type of the matrix (float, or __nv_bfloat16 widened in registers: the bf16 autocast step of BASELINE.json
// configs[3]); sums and the output are fp32 either way.
template <typename V>
__global__ void __launch_bounds__(kCsThreads)
colsum_kernel(const V* __restrict__ x, long long rows, int cols, float* __restrict__ out) {
  __shared__ float4 s_part[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * kCsCols + 4 * cl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < cols) {
    const long long stride = (long long)gridDim.y * 8;
    for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += stride) {
      const float4 v = Chan4<V>::stream_in(x + r * cols + col);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  s_part[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && col < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = s_part[k][cl];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(out + col, acc.x);
    atomicAdd(out + col + 1, acc.y);
    atomicAdd(out + col + 2, acc.z);
    atomicAdd(out + col + 3, acc.w);
  }
}

// ReLU backward fused with the column sums: g[r, c] = (y[r, c] > 0 ? dy[r, c] : 0) is written once and its column
// sums (the bias gradient of the linear layer before the ReLU) accumulate in the same pass -- 12 B per element
// instead of 12 (threshold_backward) + 4 (a separate column-sum pass over g).
template <typename V>
__global__ void __launch_bounds__(kCsThreads)
relu_bwd_colsum_kernel(const V* __restrict__ dy, const V* __restrict__ y, long long rows, int cols,
                       V* __restrict__ g, float* __restrict__ out) {
  __shared__ float4 s_part[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * kCsCols + 4 * cl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < cols) {
    const long long stride = (long long)gridDim.y * 8;
    for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += stride) {
      const float4 d = Chan4<V>::stream_in(dy + r * cols + col);
      const float4 a = Chan4<V>::stream_in(y + r * cols + col);
      const float4 v = make_float4(a.x > 0.f ? d.x : 0.f, a.y > 0.f ? d.y : 0.f, a.z > 0.f ? d.z : 0.f,
                                   a.w > 0.f ? d.w : 0.f);
      Chan4<V>::stream_out(g + r * cols + col, v);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  s_part[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && col < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = s_part[k][cl];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(out + col, acc.x);
    atomicAdd(out + col + 1, acc.y);
    atomicAdd(out + col + 2, acc.z);
    atomicAdd(out + col + 3, acc.w);
  }
}

}  // namespace sdb

namespace sdb {

template <typename V>
int relu_backward_colsum(cudaStream_t st, const V* dy, const V* y, long long rows, int cols, V* g, float* colsum,
                         const char* what) {
  SDB_REQUIRE(rows >= 0 && cols > 0, "%s: bad sizes rows=%lld cols=%d", what, rows, cols);
  SDB_REQUIRE(colsum != nullptr, "%s: null output", what);
  SDB_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)cols, st));
  if (rows == 0) return SDB_OK;
  SDB_REQUIRE(dy && y && g, "%s: null pointer", what);
  if (cols % 4 != 0 || ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(g)) &
                        Chan4<V>::kAlignMask) != 0) {
    set_error("%s: cols=%d must be a multiple of 4 and the tensors aligned to 4 elements", what, cols);
    return SDB_ERR_UNSUPPORTED;
  }
  const int strips = (cols + kCsCols - 1) / kCsCols;
  long long row_ctas = (long long)sm_count() * 8 / strips;
  const long long max_useful = (rows + 63) / 64;
  if (row_ctas > max_useful) row_ctas = max_useful;
  if (row_ctas < 1) row_ctas = 1;
  if (row_ctas > 65535) row_ctas = 65535;
  relu_bwd_colsum_kernel<V><<<dim3(strips, (unsigned)row_ctas), kCsThreads, 0, st>>>(dy, y, rows, cols, g, colsum);
  SDB_LAUNCH_CHECK("relu_bwd_colsum_kernel");
  return SDB_OK;
}

template <typename V>
int colsum(cudaStream_t st, const V* x, long long rows, int cols, float* out, const char* what) {
  SDB_REQUIRE(rows >= 0 && cols > 0, "%s: bad sizes rows=%lld cols=%d", what, rows, cols);
  SDB_REQUIRE(out != nullptr, "%s: null output", what);
  SDB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, st));
  if (rows == 0) return SDB_OK;
  SDB_REQUIRE(x != nullptr, "%s: null input", what);
  if (cols % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & Chan4<V>::kAlignMask) != 0) {
    set_error("%s: cols=%d must be a multiple of 4 and x aligned to 4 elements", what, cols);
    return SDB_ERR_UNSUPPORTED;
  }
  const int strips = (cols + kCsCols - 1) / kCsCols;
  long long row_ctas = (long long)sm_count() * 8 / strips;           // ~8 CTAs per SM in total
  const long long max_useful = (rows + 63) / 64;                      // at least 8 rows per row-lane
  if (row_ctas > max_useful) row_ctas = max_useful;
  if (row_ctas < 1) row_ctas = 1;
  if (row_ctas > 65535) row_ctas = 65535;
  colsum_kernel<V><<<dim3(strips, (unsigned)row_ctas), kCsThreads, 0, st>>>(x, rows, cols, out);
  SDB_LAUNCH_CHECK("colsum_kernel");
  return SDB_OK;
}

}  // namespace sdb

extern "C" int sdb_relu_backward_colsum_f32(sdb_stream_t stream, const float* dy, const float* y, int64_t rows, int cols,
                                            float* g, float* colsum) {
  return sdb::relu_backward_colsum<float>((cudaStream_t)stream, dy, y, rows, cols, g, colsum, "relu_backward_colsum");
}

extern "C" int sdb_colsum_f32(sdb_stream_t stream, const float* x, int64_t rows, int cols, float* out) {
  return sdb::colsum<float>((cudaStream_t)stream, x, rows, cols, out, "colsum");
}

extern "C" int sdb_relu_backward_colsum_bf16(sdb_stream_t stream, const uint16_t* dy, const uint16_t* y, int64_t rows,
                                             int cols, uint16_t* g, float* colsum) {
  return sdb::relu_backward_colsum<__nv_bfloat16>((cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(dy),
                                                  reinterpret_cast<const __nv_bfloat16*>(y), rows, cols,
                                                  reinterpret_cast<__nv_bfloat16*>(g), colsum, "relu_backward_colsum_bf16");
}

extern "C" int sdb_colsum_bf16(sdb_stream_t stream, const uint16_t* x, int64_t rows, int cols, float* out) {
  return sdb::colsum<__nv_bfloat16>((cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(x), rows, cols, out,
                                    "colsum_bf16");
}
