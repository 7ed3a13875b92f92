// DETR Hungarian matching on the device: cost build + exact rectangular LSAP.  sm_100a.
// This translation unit is compiled with -fmad=false so the fp32 cost follows torch's separately rounded
// elementwise ops (the reference builds the cost with eager torch kernels).
//
// Replaces HungarianAssigner.assign
// (/root/reference/thirdparty/mmdetection/mmdet/core/bbox/assigners/hungarian_assigner.py:96-148):
// three torch cost kernels chains, `cost.detach().cpu()` (a device->host sync per image per decoder layer),
// scipy.optimize.linear_sum_assignment on the host, and two host->device index copies.  Here all P = layers x
// images problems are built by one launch and solved by one launch (one CTA each), with no host round trip.
//
// The solver is the shortest-augmenting-path algorithm scipy implements (Crouse 2016), kept step-for-step
// identical to oracle/lsap.c so the indices are bit-identical to scipy's -- including ties: the candidate list
// is filled in reverse, removal swaps with the last candidate, and among equal minima an unassigned column met
// later wins.  The O(nc) scan of each Dijkstra step is spread over the CTA; the sequential tie rule is turned
// into an order-independent key so the parallel arg-min returns exactly what the sequential scan would.
#include <math.h>

#include "common.cuh"

namespace sdb {

// ------------------------------------------------------------------------------------------------
// cost build: one thread per (problem, query), loop over the problem's ground truths
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
match_cost_kernel(const float* __restrict__ cls_pred, const float* __restrict__ bbox_pred,
                  const float* __restrict__ gt_bboxes, const int64_t* __restrict__ gt_labels,
                  const int32_t* __restrict__ prob_seg, const int32_t* __restrict__ seg_offsets,
                  const float* __restrict__ seg_img_wh, const int64_t* __restrict__ cost_offsets, int Q, int C,
                  float w_cls, float w_l1, float w_iou, float* __restrict__ cost_qg,
                  float* __restrict__ cost_solver) {
  const int p = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int seg = prob_seg[p];
  const int g0 = seg_offsets[seg], G = seg_offsets[seg + 1] - g0;
  if (G <= 0) return;
  const float iw = seg_img_wh[2 * seg], ih = seg_img_wh[2 * seg + 1];
  const long long base = cost_offsets[p];
  const float4 pb = reinterpret_cast<const float4*>(bbox_pred)[(long long)p * Q + q];  // cx cy w h
  const float* cl = cls_pred + ((long long)p * Q + q) * C;
  // bbox_cxcywh_to_xyxy(bbox_pred) * factor        (transforms.py:222-234, hungarian_assigner.py:126)
  const float px1 = (pb.x - 0.5f * pb.z) * iw, py1 = (pb.y - 0.5f * pb.w) * ih;
  const float px2 = (pb.x + 0.5f * pb.z) * iw, py2 = (pb.y + 0.5f * pb.w) * ih;
  const float area1 = (px2 - px1) * (py2 - py1);
  const bool transposed = Q > G;
  for (int g = 0; g < G; ++g) {
    const float4 gb = reinterpret_cast<const float4*>(gt_bboxes)[g0 + g];  // x1 y1 x2 y2 (pixels)
    const int label = (int)gt_labels[g0 + g];
    // FocalLossCost, match_cost.py:93-99 (alpha .25, gamma 2, eps 1e-12)
    // a label outside [0, C) would read out of bounds: its cost becomes NaN, the solver reports status 1 for the
    // problem and HungarianAssigner.check_status() raises what scipy raises on invalid entries
    const float x = (label >= 0 && label < C) ? cl[label] : __int_as_float(0x7fc00000);
    const float s = 1.0f / (1.0f + expf(-x));
    const float neg = (-logf((1.0f - s) + 1e-12f)) * 0.75f * (s * s);
    const float pos = (-logf(s + 1e-12f)) * 0.25f * ((1.0f - s) * (1.0f - s));
    const float c_cls = (pos - neg) * w_cls;
    // BBoxL1Cost on normalised cxcywh, match_cost.py:45-50, hungarian_assigner.py:123-124
    const float nx1 = gb.x / iw, ny1 = gb.y / ih, nx2 = gb.z / iw, ny2 = gb.w / ih;
    const float gcx = (nx1 + nx2) / 2.0f, gcy = (ny1 + ny2) / 2.0f, gw = nx2 - nx1, gh = ny2 - ny1;
    const float l1 = fabsf(pb.x - gcx) + fabsf(pb.y - gcy) + fabsf(pb.z - gw) + fabsf(pb.w - gh);
    const float c_l1 = l1 * w_l1;
    // IoUCost(giou) -> bbox_overlaps, iou2d_calculator.py:218-260 (eps 1e-6)
    const float area2 = (gb.z - gb.x) * (gb.w - gb.y);
    const float ow = fmaxf(fminf(px2, gb.z) - fmaxf(px1, gb.x), 0.0f);
    const float oh = fmaxf(fminf(py2, gb.w) - fmaxf(py1, gb.y), 0.0f);
    const float overlap = ow * oh;
    const float uni = fmaxf(area1 + area2 - overlap, 1e-6f);
    const float iou = overlap / uni;
    const float ew = fmaxf(fmaxf(px2, gb.z) - fminf(px1, gb.x), 0.0f);
    const float eh = fmaxf(fmaxf(py2, gb.w) - fminf(py1, gb.y), 0.0f);
    const float earea = fmaxf(ew * eh, 1e-6f);
    const float giou = iou - (earea - uni) / earea;
    const float c_iou = (-giou) * w_iou;
    const float c = (c_cls + c_l1) + c_iou;
    if (cost_qg) cost_qg[base + (long long)q * G + g] = c;
    cost_solver[base + (transposed ? (long long)g * Q + q : (long long)q * G + g)] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// LSAP: one CTA per problem
// ------------------------------------------------------------------------------------------------
constexpr int kLsapThreads = 128;

struct Best {
  double d;
  int key;   // smaller wins among equal d; encodes "unassigned later beats everything, else earliest"
  int t;     // position in the candidate list
};

__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  return a.d < b.d || (a.d == b.d && a.key < b.key);
}

__global__ void __launch_bounds__(kLsapThreads)
lsap_kernel(const float* __restrict__ cost_solver, const int64_t* __restrict__ cost_offsets,
            const int32_t* __restrict__ prob_seg, const int32_t* __restrict__ seg_offsets,
            const int64_t* __restrict__ gt_labels, int Q, int NCmax, int NRmax, int64_t* __restrict__ gt_inds,
            int64_t* __restrict__ labels, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x;
  const int tid = threadIdx.x;
  const int seg = prob_seg[p];
  const int g0 = seg_offsets[seg], G = seg_offsets[seg + 1] - g0;
  int64_t* gi = gt_inds + (long long)p * Q;
  int64_t* lb = labels ? labels + (long long)p * Q : nullptr;
  if (G <= 0 || Q <= 0) {  // hungarian_assigner.py:108-114: no GT -> everything background
    for (int q = tid; q < Q; q += kLsapThreads) {
      gi[q] = 0;
      if (lb) lb[q] = -1;
    }
    if (tid == 0 && status) status[p] = 0;
    return;
  }
  const bool transposed = Q > G;
  const int nr = transposed ? G : Q, nc = transposed ? Q : G;
  const float* cost = cost_solver + cost_offsets[p];

  double* v = reinterpret_cast<double*>(smem_raw);
  double* dist = v + NCmax;
  double* u = dist + NCmax;
  int* path = reinterpret_cast<int*>(u + NRmax);
  int* row4col = path + NCmax;
  int* cand = row4col + NCmax;
  int* col4row = cand + NCmax;
  unsigned char* in_sc = reinterpret_cast<unsigned char*>(col4row + NRmax);
  unsigned char* in_sr = in_sc + NCmax;

  __shared__ Best s_best[kLsapThreads / 32];
  __shared__ int s_i, s_sink, s_left, s_bad;
  __shared__ double s_min;

  if (tid == 0) s_bad = 0;
  for (int j = tid; j < nc; j += kLsapThreads) { v[j] = 0.0; path[j] = -1; row4col[j] = -1; }
  for (int i = tid; i < nr; i += kLsapThreads) { u[i] = 0.0; col4row[i] = -1; }
  __syncthreads();
  // scipy rejects NaN and -inf up front
  {
    int bad = 0;
    for (long long k = tid; k < (long long)nr * nc; k += kLsapThreads) {
      const float c = cost[k];
      if (c != c || c == -INFINITY) bad = 1;
    }
    if (bad) s_bad = 1;
  }
  __syncthreads();
  int rc = s_bad ? 1 : 0;

  for (int cur = 0; cur < nr && rc == 0; ++cur) {
    for (int j = tid; j < nc; j += kLsapThreads) { cand[j] = nc - j - 1; dist[j] = INFINITY; in_sc[j] = 0; }
    for (int i = tid; i < nr; i += kLsapThreads) in_sr[i] = 0;
    if (tid == 0) { s_i = cur; s_sink = -1; s_left = nc; s_min = 0.0; }
    __syncthreads();

    while (true) {
      const int i = s_i, left = s_left;
      const double min_val = s_min;
      const double ui = u[i];
      const float* crow = cost + (long long)i * nc;
      Best best;
      best.d = INFINITY; best.key = 0x7fffffff; best.t = -1;
      for (int t = tid; t < left; t += kLsapThreads) {
        const int j = cand[t];
        const double r = ((min_val + (double)crow[j]) - ui) - v[j];
        double d = dist[j];
        if (r < d) { path[j] = i; dist[j] = r; d = r; }
        Best c;
        c.d = d;
        c.t = t;
        c.key = (row4col[j] < 0) ? (SDB_LSAP_MAX_DIM - t) : (SDB_LSAP_MAX_DIM + 1 + t);
        // the first candidate always replaces the INFINITY sentinel, like `index` in the sequential scan
        if (best.t < 0 || better(c, best)) best = c;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Best other;
        other.d = __shfl_xor_sync(0xffffffffu, best.d, o);
        other.key = __shfl_xor_sync(0xffffffffu, best.key, o);
        other.t = __shfl_xor_sync(0xffffffffu, best.t, o);
        if (other.t >= 0 && (best.t < 0 || better(other, best))) best = other;
      }
      if ((tid & 31) == 0) s_best[tid >> 5] = best;
      __syncthreads();
      if (tid == 0) {
        Best b = s_best[0];
#pragma unroll
        for (int w = 1; w < kLsapThreads / 32; ++w) {
          const Best o = s_best[w];
          if (o.t >= 0 && (b.t < 0 || better(o, b))) b = o;
        }
        in_sr[i] = 1;
        s_min = b.d;
        if (b.d == INFINITY) {
          s_sink = -2;  // infeasible
        } else {
          const int j = cand[b.t];
          if (row4col[j] < 0) s_sink = j; else s_i = row4col[j];
          in_sc[j] = 1;
          cand[b.t] = cand[left - 1];
          s_left = left - 1;
        }
      }
      __syncthreads();
      if (s_sink != -1) break;
    }
    if (s_sink == -2) { rc = 2; break; }

    const double min_val = s_min;
    const int sink = s_sink;
    // dual updates (all reads of dist/col4row happen before the augmentation below)
    for (int i = tid; i < nr; i += kLsapThreads) {
      if (i == cur) u[i] += min_val;
      else if (in_sr[i]) u[i] += min_val - dist[col4row[i]];
    }
    for (int j = tid; j < nc; j += kLsapThreads)
      if (in_sc[j]) v[j] -= min_val - dist[j];
    __syncthreads();
    if (tid == 0) {
      int j = sink;
      while (true) {
        const int i = path[j];
        row4col[j] = i;
        const int t = col4row[i];
        col4row[i] = j;
        j = t;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }

  // hungarian_assigner.py:142-148
  for (int q = tid; q < Q; q += kLsapThreads) {
    int g = -1;
    if (rc == 0) g = transposed ? row4col[q] : col4row[q];
    gi[q] = g + 1;
    if (lb) lb[q] = (g >= 0 && gt_labels) ? gt_labels[g0 + g] : -1;
  }
  if (tid == 0 && status) status[p] = rc;
}

static size_t lsap_smem_bytes(int NC, int NR) {
  return sizeof(double) * (2 * (size_t)NC + NR) + sizeof(int) * (3 * (size_t)NC + NR) + (size_t)NC + NR + 16;
}

}  // namespace sdb

extern "C" int sdb_match_cost_f32(sdb_stream_t stream, const float* cls_pred, const float* bbox_pred,
                                  const float* gt_bboxes, const int64_t* gt_labels, const int32_t* prob_seg,
                                  const int32_t* seg_offsets, const float* seg_img_wh,
                                  const int64_t* cost_offsets, int num_problems, int num_query,
                                  int num_classes, float w_cls, float w_l1, float w_iou, float* cost_qg,
                                  float* cost_solver) {
  SDB_REQUIRE(num_problems >= 0 && num_query >= 0 && num_classes > 0, "match_cost: bad sizes P=%d Q=%d C=%d",
              num_problems, num_query, num_classes);
  if (num_problems == 0 || num_query == 0) return SDB_OK;
  SDB_REQUIRE(cls_pred && bbox_pred && prob_seg && seg_offsets && seg_img_wh && cost_offsets && cost_solver,
              "match_cost: null pointer");
  SDB_REQUIRE(num_problems <= 65535, "match_cost: too many problems (%d)", num_problems);
  dim3 grid((num_query + 127) / 128, num_problems);
  sdb::match_cost_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
      cls_pred, bbox_pred, gt_bboxes, gt_labels, prob_seg, seg_offsets, seg_img_wh, cost_offsets, num_query,
      num_classes, w_cls, w_l1, w_iou, cost_qg, cost_solver);
  SDB_LAUNCH_CHECK("match_cost_kernel");
  return SDB_OK;
}

extern "C" int sdb_lsap_solve_f32(sdb_stream_t stream, const float* cost_solver, const int64_t* cost_offsets,
                                  const int32_t* prob_seg, const int32_t* seg_offsets,
                                  const int64_t* gt_labels, int num_problems, int num_query, int max_gt,
                                  int64_t* gt_inds, int64_t* labels, int32_t* status) {
  SDB_REQUIRE(num_problems >= 0 && num_query >= 0 && max_gt >= 0, "lsap_solve: bad sizes P=%d Q=%d maxG=%d",
              num_problems, num_query, max_gt);
  if (num_problems == 0) return SDB_OK;
  SDB_REQUIRE(cost_offsets && prob_seg && seg_offsets && gt_inds, "lsap_solve: null pointer");
  SDB_REQUIRE(cost_solver || max_gt == 0 || num_query == 0, "lsap_solve: null cost");
  const int NC = num_query > max_gt ? num_query : max_gt;
  const int NR = num_query > max_gt ? max_gt : num_query;
  if (NC > SDB_LSAP_MAX_DIM) {
    sdb::set_error("lsap_solve: max(Q=%d, max_gt=%d) exceeds SDB_LSAP_MAX_DIM=%d", num_query, max_gt,
                   SDB_LSAP_MAX_DIM);
    return SDB_ERR_UNSUPPORTED;
  }
  const size_t smem = sdb::lsap_smem_bytes(NC > 0 ? NC : 1, NR > 0 ? NR : 1);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SDB_CUDA(cudaFuncSetAttribute(sdb::lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  sdb::lsap_kernel<<<num_problems, sdb::kLsapThreads, smem, (cudaStream_t)stream>>>(
      cost_solver, cost_offsets, prob_seg, seg_offsets, gt_labels, num_query, NC > 0 ? NC : 1, NR > 0 ? NR : 1,
      gt_inds, labels, status);
  SDB_LAUNCH_CHECK("lsap_kernel");
  return SDB_OK;
}

extern "C" int sdb_hungarian_assign_f32(sdb_stream_t stream, const float* cls_pred, const float* bbox_pred,
                                        const float* gt_bboxes, const int64_t* gt_labels,
                                        const int32_t* prob_seg, const int32_t* seg_offsets,
                                        const float* seg_img_wh, const int64_t* cost_offsets, int num_problems,
                                        int num_query, int num_classes, int max_gt, float w_cls, float w_l1,
                                        float w_iou, float* workspace, float* cost_qg, int64_t* gt_inds,
                                        int64_t* labels, int32_t* status) {
  if (max_gt > 0 && num_query > 0) {
    int rc = sdb_match_cost_f32(stream, cls_pred, bbox_pred, gt_bboxes, gt_labels, prob_seg, seg_offsets,
                                seg_img_wh, cost_offsets, num_problems, num_query, num_classes, w_cls, w_l1,
                                w_iou, cost_qg, workspace);
    if (rc != SDB_OK) return rc;
  }
  return sdb_lsap_solve_f32(stream, workspace, cost_offsets, prob_seg, seg_offsets, gt_labels, num_problems,
                            num_query, max_gt, gt_inds, labels, status);
}
