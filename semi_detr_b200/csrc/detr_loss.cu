// Classification + box losses of the DINO head for ALL (decoder layer, image) problems in one launch, targets gathered
// from the assignment inside the kernel; and the matching backward.  sm_100a.
//
// Replaces, per train step, the 13 x `loss_single` of
// /root/reference/detr_od/models/dense_heads/dino_detr_head.py:634-736 (and the denoising twin :739-819, :895-980):
// `get_targets` scatter of labels / box targets from the assignment, FocalLoss (mmdet losses/focal_loss.py:12-57;
// mmcv's sigmoid_focal_loss op on CUDA -- same arithmetic), L1Loss on normalised cxcywh (smooth_l1_loss.py:34-46) with
// the xy / hw diagnostics (:726-734), GIoULoss on pixel xyxy (iou_loss.py:101-116; iou2d_calculator.py:204-260 aligned
// form) -- ~20 elementwise launches + 2-3 host syncs per call there.
//
// Forward: one CTA per problem p = (layer, image); sums[p] = {focal, l1, l1_xy, l1_hw, (1 - giou)} summed over the
// problem's queries in a FIXED order (warp shuffle tree, then the warps in index order) -- bitwise reproducible.
// Normalisers and loss weights stay outside (they involve a cross-rank mean, dino_detr_head.py:698-723).
// Backward: one thread per logit / per box, no reductions.
//
// Target of query q of problem p: gi = gt_inds[p, q] (0 = background, k + 1 = GT k of the problem's segment
// s = prob_seg[p]); label = gt_labels[seg_offsets[s] + k] (background: no positive class), box target = the GT box
// / (w, h, w, h) converted to cxcywh (dino_detr_head.py:969-976).
//
// Subgradient conventions follow torch autograd (the reference differentiates these expressions with it):
// clamp(min=c) passes the gradient where x >= c; min / max split it evenly on ties; |x|' = sign(x).
#include "common.cuh"

namespace sdb {

namespace {

constexpr int kLossThreads = 512;

struct BoxTarget {
  float cx, cy, w, h;   // normalised cxcywh
  float fw, fh;         // image width / height (pixel factor)
  long long label;
};

__device__ __forceinline__ BoxTarget load_target(const float* __restrict__ gt_bboxes,
                                                 const int64_t* __restrict__ gt_labels,
                                                 const float* __restrict__ img_wh, int seg, int g) {
  BoxTarget t;
  t.fw = img_wh[2 * seg];
  t.fh = img_wh[2 * seg + 1];
  const float x1 = gt_bboxes[4 * g] / t.fw, y1 = gt_bboxes[4 * g + 1] / t.fh;
  const float x2 = gt_bboxes[4 * g + 2] / t.fw, y2 = gt_bboxes[4 * g + 3] / t.fh;
  t.cx = (x1 + x2) / 2.f;
  t.cy = (y1 + y2) / 2.f;
  t.w = x2 - x1;
  t.h = y2 - y1;
  t.label = gt_labels[g];
  return t;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// focal term of one logit (py_sigmoid_focal_loss): bce_with_logits(x, t) * (alpha t + (1 - alpha)(1 - t)) * pt^gamma
__device__ __forceinline__ float focal_term(float x, bool pos, float alpha, float gamma) {
  const float p = sigmoidf_(x);
  const float pt = pos ? 1.f - p : p;
  const float aw = pos ? alpha : 1.f - alpha;
  const float mod = gamma == 2.f ? pt * pt : powf(pt, gamma);
  const float bce = fmaxf(x, 0.f) - (pos ? x : 0.f) + log1pf(expf(-fabsf(x)));
  return bce * aw * mod;
}

__device__ __forceinline__ float focal_grad(float x, bool pos, float alpha, float gamma) {
  const float p = sigmoidf_(x);
  const float pt = pos ? 1.f - p : p;
  const float aw = pos ? alpha : 1.f - alpha;
  const float mod = gamma == 2.f ? pt * pt : powf(pt, gamma);
  const float dmod = gamma == 2.f ? 2.f * pt : (gamma == 0.f ? 0.f : gamma * powf(pt, gamma - 1.f));
  const float bce = fmaxf(x, 0.f) - (pos ? x : 0.f) + log1pf(expf(-fabsf(x)));
  const float dpt = (pos ? -1.f : 1.f) * p * (1.f - p);
  return aw * (mod * (p - (pos ? 1.f : 0.f)) + bce * dmod * dpt);
}

struct Giou {
  float loss;               // 1 - giou
  float dcx, dcy, dw, dh;   // d loss / d (normalised cxcywh prediction)
};

// GIoU loss of the prediction (normalised cxcywh) against the target, both scaled to pixels (iou2d_calculator.py:233-260)
template <bool kGrad>
__device__ __forceinline__ Giou giou_loss(float cx, float cy, float w, float h, const BoxTarget& t, float eps) {
  const float px1 = (cx - 0.5f * w) * t.fw, py1 = (cy - 0.5f * h) * t.fh;
  const float px2 = (cx + 0.5f * w) * t.fw, py2 = (cy + 0.5f * h) * t.fh;
  const float tx1 = (t.cx - 0.5f * t.w) * t.fw, ty1 = (t.cy - 0.5f * t.h) * t.fh;
  const float tx2 = (t.cx + 0.5f * t.w) * t.fw, ty2 = (t.cy + 0.5f * t.h) * t.fh;
  const float area_a = (px2 - px1) * (py2 - py1), area_b = (tx2 - tx1) * (ty2 - ty1);
  const float iwr = fminf(px2, tx2) - fmaxf(px1, tx1), ihr = fminf(py2, ty2) - fmaxf(py1, ty1);
  const float iw = fmaxf(iwr, 0.f), ih = fmaxf(ihr, 0.f);
  const float overlap = iw * ih;
  const float union_r = area_a + area_b - overlap;
  const float uni = fmaxf(union_r, eps);
  const float ewr = fmaxf(px2, tx2) - fminf(px1, tx1), ehr = fmaxf(py2, ty2) - fminf(py1, ty1);
  const float ew = fmaxf(ewr, 0.f), eh = fmaxf(ehr, 0.f);
  const float earea_r = ew * eh;
  const float earea = fmaxf(earea_r, eps);
  Giou r;
  r.loss = 1.f - (overlap / uni - (earea - uni) / earea);
  r.dcx = r.dcy = r.dw = r.dh = 0.f;
  if (kGrad) {
    // giou = overlap / uni - 1 + uni / earea
    const float g_uni = (-overlap / (uni * uni) + 1.f / earea) * (union_r >= eps ? 1.f : 0.f);
    const float g_ov = 1.f / uni - g_uni;
    const float g_area = g_uni;
    const float g_ea = (-uni / (earea * earea)) * (earea_r >= eps ? 1.f : 0.f);
    auto hi = [](float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); };   // share of `a` in max(a, b)
    auto lo = [](float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); };   // share of `a` in min(a, b)
    const float g_iw = g_ov * ih * (iwr >= 0.f ? 1.f : 0.f), g_ih = g_ov * iw * (ihr >= 0.f ? 1.f : 0.f);
    const float g_ew = g_ea * eh * (ewr >= 0.f ? 1.f : 0.f), g_eh = g_ea * ew * (ehr >= 0.f ? 1.f : 0.f);
    float gx1 = -g_iw * hi(px1, tx1) - g_ew * lo(px1, tx1) - g_area * (py2 - py1);
    float gx2 = g_iw * lo(px2, tx2) + g_ew * hi(px2, tx2) + g_area * (py2 - py1);
    float gy1 = -g_ih * hi(py1, ty1) - g_eh * lo(py1, ty1) - g_area * (px2 - px1);
    float gy2 = g_ih * lo(py2, ty2) + g_eh * hi(py2, ty2) + g_area * (px2 - px1);
    // d giou -> d loss, pixel -> normalised
    gx1 *= -t.fw; gx2 *= -t.fw; gy1 *= -t.fh; gy2 *= -t.fh;
    r.dcx = gx1 + gx2;
    r.dcy = gy1 + gy2;
    r.dw = 0.5f * (gx2 - gx1);
    r.dh = 0.5f * (gy2 - gy1);
  }
  return r;
}

__global__ void __launch_bounds__(kLossThreads)
detr_loss_fwd_kernel(const float* __restrict__ cls, const float* __restrict__ box, const int64_t* __restrict__ gt_inds,
                     const int* __restrict__ prob_seg, const int* __restrict__ seg_offsets,
                     const float* __restrict__ gt_bboxes, const int64_t* __restrict__ gt_labels,
                     const float* __restrict__ img_wh, const float* __restrict__ cls_weight, int Q, int C, float alpha,
                     float gamma, float eps, float* __restrict__ sums) {
  const int p = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kLossThreads / 32;
  const int seg = prob_seg[p];
  const int g0 = seg_offsets[seg];
  float focal = 0.f, l1xy = 0.f, l1hw = 0.f, gl = 0.f;
  for (int q = warp; q < Q; q += kWarps) {
    const long long gi = gt_inds[(long long)p * Q + q];
    long long label = -1;
    BoxTarget t;
    if (gi > 0) {
      t = load_target(gt_bboxes, gt_labels, img_wh, seg, g0 + (int)gi - 1);
      label = t.label;
    }
    const float* row = cls + ((long long)p * Q + q) * C;
    for (int c = lane; c < C; c += 32) focal += focal_term(row[c], c == label, alpha, gamma);
    if (gi > 0 && lane == 0) {
      const float4 b = *reinterpret_cast<const float4*>(box + ((long long)p * Q + q) * 4);
      l1xy += fabsf(b.x - t.cx) + fabsf(b.y - t.cy);
      l1hw += fabsf(b.z - t.w) + fabsf(b.w - t.h);
      gl += giou_loss<false>(b.x, b.y, b.z, b.w, t, eps).loss;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) focal += __shfl_xor_sync(0xffffffffu, focal, o);
  __shared__ float part[kWarps][4];
  if (lane == 0) {
    part[warp][0] = focal;
    part[warp][1] = l1xy;
    part[warp][2] = l1hw;
    part[warp][3] = gl;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int w = 0; w < kWarps; ++w) {
      s0 += part[w][0];
      s1 += part[w][1];
      s2 += part[w][2];
      s3 += part[w][3];
    }
    if (cls_weight) s0 *= cls_weight[p];
    float* o = sums + 5 * p;
    o[0] = s0;
    o[1] = s1 + s2;
    o[2] = s1;
    o[3] = s2;
    o[4] = s3;
  }
}

__global__ void __launch_bounds__(256)
detr_loss_bwd_kernel(const float* __restrict__ cls, const float* __restrict__ box, const int64_t* __restrict__ gt_inds,
                     const int* __restrict__ prob_seg, const int* __restrict__ seg_offsets,
                     const float* __restrict__ gt_bboxes, const int64_t* __restrict__ gt_labels,
                     const float* __restrict__ img_wh, const float* __restrict__ cls_weight,
                     const float* __restrict__ grad_sums, long long PQ, int Q, int C, float alpha, float gamma, float eps,
                     float* __restrict__ grad_cls, float* __restrict__ grad_box) {
  // one warp per (problem, query) row: lanes sweep the C logits; lane 0 also writes the box gradient
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < PQ; r += nwarps) {
    const int p = (int)(r / Q);
    const long long gi = gt_inds[r];
    const int seg = prob_seg[p];
    long long label = -1;
    BoxTarget t;
    if (gi > 0) {
      t = load_target(gt_bboxes, gt_labels, img_wh, seg, seg_offsets[seg] + (int)gi - 1);
      label = t.label;
    }
    const float* gs = grad_sums + 5 * p;
    const float gf = gs[0] * (cls_weight ? cls_weight[p] : 1.f);
    const float* row = cls + r * C;
    float* grow = grad_cls + r * C;
    for (int c = lane; c < C; c += 32) grow[c] = gf * focal_grad(row[c], c == label, alpha, gamma);
    if (lane == 0) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gi > 0) {
        const float4 b = *reinterpret_cast<const float4*>(box + r * 4);
        auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
        const float gxy = gs[1] + gs[2], ghw = gs[1] + gs[3];
        const Giou gi_ = giou_loss<true>(b.x, b.y, b.z, b.w, t, eps);
        g.x = gxy * sgn(b.x - t.cx) + gs[4] * gi_.dcx;
        g.y = gxy * sgn(b.y - t.cy) + gs[4] * gi_.dcy;
        g.z = ghw * sgn(b.z - t.w) + gs[4] * gi_.dw;
        g.w = ghw * sgn(b.w - t.h) + gs[4] * gi_.dh;
      }
      *reinterpret_cast<float4*>(grad_box + r * 4) = g;
    }
  }
}

}  // namespace
}  // namespace sdb

extern "C" int sdb_detr_loss_forward_f32(sdb_stream_t stream, const float* cls_scores, const float* bbox_preds,
                                         const int64_t* gt_inds, const int32_t* prob_seg, const int32_t* seg_offsets,
                                         const float* gt_bboxes, const int64_t* gt_labels, const float* img_wh,
                                         const float* cls_weight, int num_problems, int num_query, int num_classes,
                                         float alpha, float gamma, float giou_eps, float* sums) {
  using namespace sdb;
  SDB_REQUIRE(num_problems >= 0 && num_query >= 0 && num_classes > 0,
              "detr_loss_forward: bad sizes problems=%d query=%d classes=%d", num_problems, num_query, num_classes);
  if (num_problems == 0) return SDB_OK;
  SDB_REQUIRE(cls_scores && bbox_preds && gt_inds && prob_seg && seg_offsets && img_wh && sums,
              "detr_loss_forward: null pointer");
  SDB_REQUIRE((reinterpret_cast<uintptr_t>(bbox_preds) & 15) == 0, "detr_loss_forward: bbox_preds must be 16-byte aligned");
  detr_loss_fwd_kernel<<<num_problems, kLossThreads, 0, (cudaStream_t)stream>>>(
      cls_scores, bbox_preds, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight, num_query,
      num_classes, alpha, gamma, giou_eps, sums);
  SDB_LAUNCH_CHECK("detr_loss_fwd_kernel");
  return SDB_OK;
}

extern "C" int sdb_detr_loss_backward_f32(sdb_stream_t stream, const float* cls_scores, const float* bbox_preds,
                                          const int64_t* gt_inds, const int32_t* prob_seg, const int32_t* seg_offsets,
                                          const float* gt_bboxes, const int64_t* gt_labels, const float* img_wh,
                                          const float* cls_weight, const float* grad_sums, int num_problems,
                                          int num_query, int num_classes, float alpha, float gamma, float giou_eps,
                                          float* grad_cls_scores, float* grad_bbox_preds) {
  using namespace sdb;
  SDB_REQUIRE(num_problems >= 0 && num_query >= 0 && num_classes > 0,
              "detr_loss_backward: bad sizes problems=%d query=%d classes=%d", num_problems, num_query, num_classes);
  const long long rows = (long long)num_problems * num_query;
  if (rows == 0) return SDB_OK;
  SDB_REQUIRE(cls_scores && bbox_preds && gt_inds && prob_seg && seg_offsets && img_wh && grad_sums &&
              grad_cls_scores && grad_bbox_preds, "detr_loss_backward: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(bbox_preds) | reinterpret_cast<uintptr_t>(grad_bbox_preds)) & 15) == 0,
              "detr_loss_backward: box tensors must be 16-byte aligned");
  long long blocks = (rows + 7) / 8;   // 8 warps per CTA
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  detr_loss_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      cls_scores, bbox_preds, gt_inds, prob_seg, seg_offsets, gt_bboxes, gt_labels, img_wh, cls_weight, grad_sums, rows,
      num_query, num_classes, alpha, gamma, giou_eps, grad_cls_scores, grad_bbox_preds);
  SDB_LAUNCH_CHECK("detr_loss_bwd_kernel");
  return SDB_OK;
}
