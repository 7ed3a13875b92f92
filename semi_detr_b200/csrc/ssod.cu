// Pseudo-label side path of the teacher-student step on the device.  sm_100a.
//
//  sdb_pseudo_label_nms_f32   teacher detections -> pseudo boxes, one CTA per image, no host round trip:
//      class-wise greedy NMS (IoU > 0.6 suppresses, score > 0.01, first max_per_img survivors in score order) --
//      mmdet multiclass_nms / mmcv batched_nms as called by
//      /root/reference/detr_od/models/dense_heads/dino_detr_ssod_head.py:1371-1395 -- followed by the
//      `score >= mean + std` and `w > 0, h > 0` filter of /root/reference/detr_ssod/models/dino_detr_ssod.py:921-939.
//      The reference does this per image with nonzero() / boolean indexing / a per-class loop (hundreds of stream
//      synchronisations per step when an untrained teacher puts every (query, class) pair above 0.01).
//  sdb_gmm_threshold_f32      two-component 1-D Gaussian mixture on the pooled matched costs -> cost threshold
//      (dino_detr_ssod.py:832-890: sklearn GaussianMixture(2, diag, reg_covar 1e-5, means_init [min, max], weights
//      [.5, .5], precisions 1, tol 1e-3, max_iter 100), threshold = cost of the most likely sample of component 0,
//      falling back to component 1), EM in float64, one CTA; the result stays on the device.
#include "common.cuh"

namespace sdb {

namespace {

constexpr int kNmsThreads = 256;
constexpr int kNmsMaxKeep = 1024;   // capacity of the kept list in shared memory (max_per_img <= this)

__device__ __forceinline__ float iou_xyxy(const float4& a, const float4& b) {
  // torchvision / mmcv nms (offset 0): intersection over union of two xyxy boxes
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  const float inter = w * h;
  const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
  return inter / (sa + sb - inter);
}

// scores_sorted / index_sorted: (B, K) candidates of every image in descending score order, index = query * C + class.
// boxes: (B, Q, 4) xyxy in pixels.  Outputs (B, max_keep, ...) + counts; rows past the count are zero.
__global__ void __launch_bounds__(kNmsThreads)
pseudo_label_nms_kernel(const float* __restrict__ scores_sorted, const int64_t* __restrict__ index_sorted,
                        const float* __restrict__ boxes, int K, int Q, int C, float score_thr, float iou_thr,
                        int max_keep, int apply_filter, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                        int64_t* __restrict__ out_labels, int* __restrict__ out_count, int* __restrict__ nms_count) {
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ float4 kept_box[kNmsMaxKeep];
  __shared__ float kept_score[kNmsMaxKeep];
  __shared__ int kept_label[kNmsMaxKeep];
  __shared__ float4 c_box[kNmsThreads];
  __shared__ int c_label[kNmsThreads];
  __shared__ float c_score[kNmsThreads];
  __shared__ unsigned c_dead[kNmsThreads / 32];                 // suppressed by an earlier kept box
  __shared__ unsigned sup[kNmsThreads][kNmsThreads / 32];       // sup[i] bit j: candidate i suppresses candidate j > i
  __shared__ int n_kept, done;
  if (t == 0) { n_kept = 0; done = 0; }
  __syncthreads();
  const float* sc = scores_sorted + (long long)b * K;
  const int64_t* ix = index_sorted + (long long)b * K;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (long long)b * Q;

  for (int base = 0; base < K; base += kNmsThreads) {
    if (done) break;
    const int c = base + t;
    float s = c < K ? sc[c] : -1.f;
    const bool cand = s > score_thr;
    int label = -1;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cand) {
      const long long id = ix[c];
      label = (int)(id % C);
      box = bx[id / C];
    }
    c_box[t] = box; c_label[t] = label; c_score[t] = s;
    // (1) against the boxes kept so far
    bool dead = !cand;
    const int nk = n_kept;
    for (int k = 0; k < nk && !dead; ++k)
      if (kept_label[k] == label && iou_xyxy(kept_box[k], box) > iou_thr) dead = true;
    const unsigned dead_mask = __ballot_sync(0xffffffffu, dead);
    if ((t & 31) == 0) c_dead[t >> 5] = dead_mask;
    __syncthreads();
    // (2) pairwise inside the chunk: row t = which later candidates t would suppress
#pragma unroll
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      unsigned m = 0;
      if (!dead) {
        for (int j = 0; j < 32; ++j) {
          const int o = w * 32 + j;
          if (o > t && c_label[o] == label && iou_xyxy(box, c_box[o]) > iou_thr) m |= 1u << j;
        }
      }
      sup[t][w] = m;
    }
    __syncthreads();
    // (3) one thread walks the chunk in score order
    if (t == 0) {
      unsigned removed[kNmsThreads / 32];
      for (int w = 0; w < kNmsThreads / 32; ++w) removed[w] = c_dead[w];
      int n = n_kept;
      for (int i = 0; i < kNmsThreads && n < max_keep; ++i) {
        if (removed[i >> 5] & (1u << (i & 31))) continue;
        kept_box[n] = c_box[i]; kept_label[n] = c_label[i]; kept_score[n] = c_score[i];
        ++n;
        for (int w = 0; w < kNmsThreads / 32; ++w) removed[w] |= sup[i][w];
      }
      n_kept = n;
      // sorted input: once a chunk holds a score at or below the threshold nothing later can qualify
      if (n >= max_keep || !(c_score[kNmsThreads - 1] > score_thr)) done = 1;
    }
    __syncthreads();
  }

  // ---- score >= mean + std (unbiased) and non-degenerate boxes; compaction keeps the order -------------------------
  const int n = n_kept;
  __shared__ float red[kNmsThreads / 32];
  __shared__ float s_mean, s_thr;
  float part = 0.f;
  for (int i = t; i < n; i += kNmsThreads) part += kept_score[i];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((t & 31) == 0) red[t >> 5] = part;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int w = 0; w < kNmsThreads / 32; ++w) tot += red[w];
    s_mean = n > 0 ? tot / (float)n : 0.f;
  }
  __syncthreads();
  part = 0.f;
  for (int i = t; i < n; i += kNmsThreads) { const float d = kept_score[i] - s_mean; part += d * d; }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = part;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int w = 0; w < kNmsThreads / 32; ++w) tot += red[w];
    // torch.std of one element is NaN (0 / 0): the comparison below is then false for every box, like the reference
    s_thr = s_mean + sqrtf(tot / (float)(n - 1));
    if (n == 1) s_thr = __int_as_float(0x7fc00000);
  }
  __syncthreads();
  if (t == 0) {   // n <= max_keep (a few hundred): a serial stable compaction is a microsecond
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const float4 bb = kept_box[i];
      const bool ok = !apply_filter || (kept_score[i] >= s_thr && bb.z - bb.x > 0.f && bb.w - bb.y > 0.f);
      if (ok) { kept_box[m] = bb; kept_score[m] = kept_score[i]; kept_label[m] = kept_label[i]; ++m; }
    }
    n_kept = m;
    out_count[b] = m;
    if (nms_count) nms_count[b] = n;
  }
  __syncthreads();
  const int m = n_kept;
  for (int i = t; i < max_keep; i += kNmsThreads) {
    const bool on = i < m;
    reinterpret_cast<float4*>(out_boxes)[(long long)b * max_keep + i] = on ? kept_box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    out_scores[(long long)b * max_keep + i] = on ? kept_score[i] : 0.f;
    out_labels[(long long)b * max_keep + i] = on ? (int64_t)kept_label[i] : 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int kGmmThreads = 256;
constexpr int kGmmMax = 4096;

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < kGmmThreads / 32; ++w) tot += scratch[w];   // same order in every thread
  return tot;
}

// costs: `num_segs` segments of `seg_stride` floats, segment s holding seg_counts[s] values (the padded all-gather
// buffer of the ranks; one segment for a single rank).  threshold[0] <- the cost threshold, threshold[1] <- number of
// pooled values (as float).
__global__ void __launch_bounds__(kGmmThreads)
gmm_threshold_kernel(const float* __restrict__ costs, const int* __restrict__ seg_counts, int num_segs, int seg_stride,
                     float tol, int max_iter, double reg_covar, float* __restrict__ threshold) {
  __shared__ double x[kGmmMax];
  __shared__ double scratch[kGmmThreads / 32];
  __shared__ int seg_start[65];
  const int t = threadIdx.x;
  if (t == 0) {
    int acc = 0;
    for (int s = 0; s < num_segs; ++s) {
      seg_start[s] = acc;
      int c = seg_counts[s];
      c = c < 0 ? 0 : c;
      if (acc + c > kGmmMax) c = kGmmMax - acc;
      acc += c;
    }
    seg_start[num_segs] = acc;
  }
  __syncthreads();
  const int n = seg_start[num_segs];
  for (int s = 0; s < num_segs; ++s) {
    const int c = seg_start[s + 1] - seg_start[s];
    for (int i = t; i < c; i += kGmmThreads) x[seg_start[s] + i] = (double)costs[(long long)s * seg_stride + i];
  }
  __syncthreads();
  if (n == 0) { if (t == 0) { threshold[0] = 0.f; threshold[1] = 0.f; } return; }
  if (n == 1) { if (t == 0) { threshold[0] = (float)x[0]; threshold[1] = 1.f; } return; }
  // the reference sorts the costs first; the fit is order-independent up to float summation order, the threshold is a
  // sample value -- so only min / max are needed here
  double lo = 1e300, hi = -1e300;
  for (int i = t; i < n; i += kGmmThreads) { lo = fmin(lo, x[i]); hi = fmax(hi, x[i]); }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ double mm[2][kGmmThreads / 32];
  if ((t & 31) == 0) { mm[0][t >> 5] = lo; mm[1][t >> 5] = hi; }
  __syncthreads();
  for (int w = 0; w < kGmmThreads / 32; ++w) { lo = fmin(lo, mm[0][w]); hi = fmax(hi, mm[1][w]); }

  double w0 = 0.5, w1 = 0.5, mu0 = lo, mu1 = hi, var0 = 1.0, var1 = 1.0;
  const double log2pi = 1.8378770664093453;
  double lower = -1e300;
  for (int it = 0; it < max_iter; ++it) {
    const double prev = lower;
    const double lw0 = log(w0), lw1 = log(w1), lv0 = log(var0), lv1 = log(var1);
    double s_ll = 0, s_r0 = 0, s_r1 = 0, s_x0 = 0, s_x1 = 0, s_xx0 = 0, s_xx1 = 0;
    for (int i = t; i < n; i += kGmmThreads) {
      const double xi = x[i];
      const double a0 = -0.5 * (log2pi + lv0 + (xi - mu0) * (xi - mu0) / var0) + lw0;
      const double a1 = -0.5 * (log2pi + lv1 + (xi - mu1) * (xi - mu1) / var1) + lw1;
      const double m = fmax(a0, a1);
      const double ln = m + log(exp(a0 - m) + exp(a1 - m));
      const double r0 = exp(a0 - ln), r1 = exp(a1 - ln);
      s_ll += ln; s_r0 += r0; s_r1 += r1; s_x0 += r0 * xi; s_x1 += r1 * xi; s_xx0 += r0 * xi * xi; s_xx1 += r1 * xi * xi;
    }
    s_ll = block_sum(s_ll, scratch);
    s_r0 = block_sum(s_r0, scratch); s_r1 = block_sum(s_r1, scratch);
    s_x0 = block_sum(s_x0, scratch); s_x1 = block_sum(s_x1, scratch);
    s_xx0 = block_sum(s_xx0, scratch); s_xx1 = block_sum(s_xx1, scratch);
    lower = s_ll / (double)n;
    const double eps10 = 10.0 * 2.220446049250313e-16;
    const double n0 = s_r0 + eps10, n1 = s_r1 + eps10;
    mu0 = s_x0 / n0; mu1 = s_x1 / n1;
    var0 = s_xx0 / n0 - 2.0 * mu0 * s_x0 / n0 + mu0 * mu0 + reg_covar;
    var1 = s_xx1 / n1 - 2.0 * mu1 * s_x1 / n1 + mu1 * mu1 + reg_covar;
    w0 = n0 / (n0 + n1); w1 = n1 / (n0 + n1);
    if (fabs(lower - prev) < (double)tol) break;
  }
  // predict / score_samples: most likely sample of component 0, else of component 1 (ties: the smaller cost, which is
  // the first one in the reference's ascending order)
  const double lw0 = log(w0), lw1 = log(w1), lv0 = log(var0), lv1 = log(var1);
  double best[2] = {-1e300, -1e300}, bx[2] = {1e300, 1e300};
  for (int i = t; i < n; i += kGmmThreads) {
    const double xi = x[i];
    const double a0 = -0.5 * (log2pi + lv0 + (xi - mu0) * (xi - mu0) / var0) + lw0;
    const double a1 = -0.5 * (log2pi + lv1 + (xi - mu1) * (xi - mu1) / var1) + lw1;
    const double m = fmax(a0, a1);
    const double ln = m + log(exp(a0 - m) + exp(a1 - m));
    const int comp = a1 > a0 ? 1 : 0;          // argmax, first index on ties
    if (ln > best[comp] || (ln == best[comp] && xi < bx[comp])) { best[comp] = ln; bx[comp] = xi; }
  }
  __shared__ double cand[2][2][kGmmThreads];
  for (int c = 0; c < 2; ++c) { cand[c][0][t] = best[c]; cand[c][1][t] = bx[c]; }
  __syncthreads();
  if (t == 0) {
    float out = (float)x[0];
    for (int c = 0; c < 2; ++c) {
      double b = -1e300, bv = 1e300;
      for (int i = 0; i < kGmmThreads; ++i)
        if (cand[c][0][i] > b || (cand[c][0][i] == b && cand[c][1][i] < bv)) { b = cand[c][0][i]; bv = cand[c][1][i]; }
      if (b > -1e299) { out = (float)bv; break; }
    }
    threshold[0] = out;
    threshold[1] = (float)n;
  }
}

}  // namespace
}  // namespace sdb

extern "C" int sdb_pseudo_label_nms_f32(sdb_stream_t stream, const float* scores_sorted, const int64_t* index_sorted,
                                        const float* boxes_xyxy, int batch, int num_candidates, int num_query,
                                        int num_classes, float score_thr, float iou_thr, int max_per_img,
                                        int apply_mean_std_filter, float* out_boxes, float* out_scores,
                                        int64_t* out_labels, int32_t* out_count, int32_t* nms_count) {
  using namespace sdb;
  SDB_REQUIRE(batch >= 0 && num_candidates >= 0 && num_query > 0 && num_classes > 0 && max_per_img > 0 &&
              max_per_img <= kNmsMaxKeep,
              "pseudo_label_nms: bad sizes batch=%d candidates=%d query=%d classes=%d max_per_img=%d (<= %d)", batch,
              num_candidates, num_query, num_classes, max_per_img, kNmsMaxKeep);
  if (batch == 0) return SDB_OK;
  SDB_REQUIRE(scores_sorted && index_sorted && boxes_xyxy && out_boxes && out_scores && out_labels && out_count,
              "pseudo_label_nms: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(boxes_xyxy) | reinterpret_cast<uintptr_t>(out_boxes)) & 15) == 0,
              "pseudo_label_nms: box tensors must be 16-byte aligned");
  pseudo_label_nms_kernel<<<batch, kNmsThreads, 0, (cudaStream_t)stream>>>(
      scores_sorted, index_sorted, boxes_xyxy, num_candidates, num_query, num_classes, score_thr, iou_thr, max_per_img,
      apply_mean_std_filter, out_boxes, out_scores, out_labels, out_count, nms_count);
  SDB_LAUNCH_CHECK("pseudo_label_nms_kernel");
  return SDB_OK;
}

extern "C" int sdb_gmm_threshold_f32(sdb_stream_t stream, const float* costs, const int32_t* seg_counts, int num_segs,
                                     int seg_stride, float tol, int max_iter, double reg_covar, float* threshold) {
  using namespace sdb;
  SDB_REQUIRE(num_segs >= 1 && num_segs <= 64 && seg_stride >= 0 && max_iter >= 1,
              "gmm_threshold: bad sizes segs=%d stride=%d max_iter=%d", num_segs, seg_stride, max_iter);
  SDB_REQUIRE(costs && seg_counts && threshold, "gmm_threshold: null pointer");
  gmm_threshold_kernel<<<1, kGmmThreads, 0, (cudaStream_t)stream>>>(costs, seg_counts, num_segs, seg_stride, tol,
                                                                    max_iter, reg_covar, threshold);
  SDB_LAUNCH_CHECK("gmm_threshold_kernel");
  return SDB_OK;
}
