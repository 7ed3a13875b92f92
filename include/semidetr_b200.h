/* semidetr_b200 -- C ABI of the B200-native (sm_100a) Semi-DETR hot path.
 *
 * This is the drop-in boundary: a plain C interface (pointers + sizes, no torch / ATen types)
 * exported by semi_detr_b200/lib/libsemidetr_b200.so.  Every entry point replaces one native
 * (or host-side Python/scipy) step of the reference, cited as file:line relative to
 * /root/reference.  The reference-side bindings a maintainer would add are shown in
 * INTEGRATION.md; the Python host code in semi_detr_b200/ binds these symbols with ctypes.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers unless the parameter is documented as "host".
 *  - `stream` is a cudaStream_t (pass NULL / 0 for the legacy default stream).  Calls only
 *    enqueue work; nothing synchronises the host (the reference's per-call D2H syncs are gone).
 *  - Return value: SDB_OK (0) or an SDB_ERR_* code; sdb_last_error() then returns a
 *    thread-local, human-readable message (the Python layer raises RuntimeError with it,
 *    matching the reference's AT_ASSERTM -> RuntimeError behaviour,
 *    detr_od/models/utils/ops/src/cuda/ms_deform_attn_cuda.cu:28-52).
 *    Unlike the reference (which only printf's launch failures, ms_deform_im2col_cuda.cuh:948-952),
 *    kernel-launch errors are returned.
 *  - Inputs are never modified; outputs must not alias inputs.
 *  - There is NO CPU implementation behind this ABI.
 */
#ifndef SEMIDETR_B200_H_
#define SEMIDETR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_ABI_VERSION 1

enum {
  SDB_OK = 0,
  SDB_ERR_INVALID_ARG = 1, /* bad sizes / null pointers / unsupported combination */
  SDB_ERR_CUDA = 2,        /* a CUDA runtime call or kernel launch failed */
  SDB_ERR_UNSUPPORTED = 3  /* shape outside what the kernels were built for */
};

typedef void* sdb_stream_t; /* cudaStream_t */

int sdb_abi_version(void);
const char* sdb_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Multi-scale deformable attention (MSDA).
 *
 * Replaces the launchers ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
 * (detr_od/models/utils/ops/src/cuda/ms_deform_im2col_cuda.cuh:923-954 and :956-1327) that
 * ms_deform_attn_cuda_forward / _backward (ms_deform_attn_cuda.cu:20-80, 83-153) call, with the
 * same argument meaning:
 *   value              (batch, spatial_size, num_heads, channels)       channels-last, head-major
 *   spatial_shapes     (num_levels, 2) int64 on device, rows (H_l, W_l)
 *   level_start_index  (num_levels,)   int64 on device
 *   sampling_loc       (batch, num_query, num_heads, num_levels, num_point, 2)   (x, y) in [0,1]
 *   attn_weight        (batch, num_query, num_heads, num_levels, num_point)
 *   out / grad_out     (batch, num_query, num_heads * channels)
 * Semantics (zero padding, pixel = loc * size - 0.5, sample kept iff -1 < pixel < size) follow
 * ms_deform_im2col_cuda.cuh:33-84, 237-299 (forward) and :87-159, 301-403 (backward).
 *
 * Differences from the reference launchers, all in the caller's favour:
 *  - `out` need not be zero-filled (every element is written);
 *  - `grad_value` is zero-filled by the call (stream-ordered), `grad_sampling_loc` and
 *    `grad_attn_weight` are fully overwritten -- the three at::zeros_like of
 *    ms_deform_attn_cuda.cu:121-123 are not needed;
 *  - im2col_step chunking is unnecessary: one launch covers the whole batch.
 * The f32 entry points take the tuned sm_100a path when channels == 32 and
 * num_levels * num_point <= 32 (every shipped config) and a generic kernel otherwise; the f64
 * entry points exist for the reference's gradcheck (ops/test.py:63-86).
 * ------------------------------------------------------------------------------------------ */
int sdb_msda_forward_f32(sdb_stream_t stream, const float* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const float* sampling_loc,
                         const float* attn_weight, int batch, int spatial_size, int num_heads,
                         int channels, int num_levels, int num_query, int num_point, float* out);

int sdb_msda_forward_f64(sdb_stream_t stream, const double* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const double* sampling_loc,
                         const double* attn_weight, int batch, int spatial_size, int num_heads,
                         int channels, int num_levels, int num_query, int num_point, double* out);

int sdb_msda_backward_f32(sdb_stream_t stream, const float* grad_out, const float* value,
                          const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const float* sampling_loc, const float* attn_weight, int batch,
                          int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, float* grad_value,
                          float* grad_sampling_loc, float* grad_attn_weight);

int sdb_msda_backward_f64(sdb_stream_t stream, const double* grad_out, const double* value,
                          const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const double* sampling_loc, const double* attn_weight, int batch,
                          int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, double* grad_value,
                          double* grad_sampling_loc, double* grad_attn_weight);

/* bf16 storage variant (BASELINE.json configs[3]: "bf16 value / output, fp32 sampling math").  The reference op has no
 * bf16 path (AT_DISPATCH_FLOATING_TYPES, ms_deform_attn_cuda.cu:64,134), so these extend the surface rather than
 * replace something.  `value`, `out` and `grad_out` hold bf16 bit patterns (uint16_t); locations, weights and ALL
 * three gradients are fp32 -- grad_value is accumulated with fp32 reductions and narrowed by the caller if needed.
 * Built for channels == 32, num_heads == 8, num_point == 4; anything else returns SDB_ERR_UNSUPPORTED.  value / out /
 * grad_out must be 8-byte aligned, the fp32 tensors 16-byte aligned. */
int sdb_msda_forward_bf16(sdb_stream_t stream, const uint16_t* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const float* sampling_loc, const float* attn_weight,
                          int batch, int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, uint16_t* out);

int sdb_msda_backward_bf16(sdb_stream_t stream, const uint16_t* grad_out, const uint16_t* value,
                           const int64_t* spatial_shapes, const int64_t* level_start_index,
                           const float* sampling_loc, const float* attn_weight, int batch,
                           int spatial_size, int num_heads, int channels, int num_levels,
                           int num_query, int num_point, float* grad_value,
                           float* grad_sampling_loc, float* grad_attn_weight);

/* Fused prologue (NOT part of the reference surface, apply() stays the compatibility
 * path).  Takes the module's RAW linear outputs instead of materialised locations / weights
 * (modules/ms_deform_attn.py:96-112):
 *   sampling_offsets (batch, num_query, num_heads, num_levels, num_point, 2)   sampling_offsets(query)
 *   attn_logits      (batch, num_query, num_heads, num_levels * num_point)     attention_weights(query), pre-softmax
 *   reference_points (batch, num_query, num_levels, ref_dim), ref_dim 2 (encoder) or 4 (decoder)
 * and evaluates softmax + location arithmetic in registers.  The backward returns gradients w.r.t. value, the raw
 * offsets and the raw logits (none w.r.t. the reference points: DINO detaches them, transformer.py:1030-1036).
 * Built for channels 32, heads 8, points 4, levels*points <= 16; other shapes return SDB_ERR_UNSUPPORTED. */
int sdb_msda_fused_forward_f32(sdb_stream_t stream, const float* value, const int64_t* spatial_shapes,
                               const int64_t* level_start_index, const float* reference_points, int ref_dim,
                               const float* sampling_offsets, const float* attn_logits, int batch,
                               int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                               int num_point, float* out);
int sdb_msda_fused_backward_f32(sdb_stream_t stream, const float* grad_out, const float* value,
                                const int64_t* spatial_shapes, const int64_t* level_start_index,
                                const float* reference_points, int ref_dim, const float* sampling_offsets,
                                const float* attn_logits, int batch, int spatial_size, int num_heads,
                                int channels, int num_levels, int num_query, int num_point, float* grad_value,
                                float* grad_offsets, float* grad_attn_logits);

/* The same prologue as a standalone pass for shapes the fused kernels do not take (levels x points > 16, e.g. the
 * 5-level model): softmax over the head's L*P logits and the location arithmetic of ms_deform_attn.py:98-112 in one
 * launch, fp32 `sampling_loc` (batch, num_query, num_heads, num_levels, num_point, 2) and `attn_weight` out; the backward
 * maps grad_sampling_loc / grad_attn_weight to the gradients of the raw offsets / logits (bf16 variants: raw tensors as
 * bf16 bit patterns).  `spatial_shapes_host`: HOST array of (H, W) per level.  num_point must be 4, levels <= 8. */
int sdb_msda_prologue_forward_f32(sdb_stream_t stream, const float* offsets, const float* logits, const float* ref,
                                  int ref_dim, const int64_t* spatial_shapes_host, int batch, int num_query,
                                  int num_heads, int num_levels, int num_point, float* sampling_loc, float* attn_weight);
int sdb_msda_prologue_forward_bf16(sdb_stream_t stream, const uint16_t* offsets, const uint16_t* logits, const float* ref,
                                   int ref_dim, const int64_t* spatial_shapes_host, int batch, int num_query,
                                   int num_heads, int num_levels, int num_point, float* sampling_loc,
                                   float* attn_weight);
int sdb_msda_prologue_backward_f32(sdb_stream_t stream, const float* grad_loc, const float* grad_attn,
                                   const float* attn_weight, const float* ref, int ref_dim,
                                   const int64_t* spatial_shapes_host, int batch, int num_query, int num_heads,
                                   int num_levels, int num_point, float* grad_offsets, float* grad_logits);
int sdb_msda_prologue_backward_bf16(sdb_stream_t stream, const float* grad_loc, const float* grad_attn,
                                    const float* attn_weight, const float* ref, int ref_dim,
                                    const int64_t* spatial_shapes_host, int batch, int num_query, int num_heads,
                                    int num_levels, int num_point, uint16_t* grad_offsets, uint16_t* grad_logits);

/* TMA-staged forward for encoder self-attention (num_query == spatial_size, channels 32, heads 8, points 4,
 * levels <= 4): per (image, head, 8x8 query tile) one bulk tensor copy per level stages that head's value box in
 * shared memory behind an mbarrier; out-of-box samples use global loads.  Same results as sdb_msda_forward_f32 /
 * sdb_msda_fused_forward_f32 (fused != 0: raw offsets / logits + reference points).  `spatial_shapes_host` /
 * `level_start_host` are HOST copies of the two index tensors (the TMA descriptors are encoded on the host). */
int sdb_msda_forward_tma_f32(sdb_stream_t stream, int fused, const float* value, const int64_t* spatial_shapes,
                             const int64_t* level_start_index, const int64_t* spatial_shapes_host,
                             const int64_t* level_start_host, const float* reference_points, int ref_dim,
                             const float* sampling_loc, const float* attn_weight, int batch, int spatial_size,
                             int num_heads, int channels, int num_levels, int num_query, int num_point,
                             float* out);

/* Tuning knob for benchmarking kernel variants (0 = default heuristic).  Not part of the
 * reference surface; see DESIGN.md "MSDA forward variants". */
int sdb_msda_set_variant(int forward_variant, int backward_variant);

/* ------------------------------------------------------------------------------------------
 * Hungarian matching, batched over P independent problems (decoder layer x image).
 *
 * Replaces, per problem, HungarianAssigner.assign
 * (thirdparty/mmdetection/mmdet/core/bbox/assigners/hungarian_assigner.py:96-148):
 *   cost = FocalLossCost(w_cls, alpha .25, gamma 2, eps 1e-12)          match_cost.py:83-99
 *        + BBoxL1Cost(w_l1, box_format='xywh')                           match_cost.py:33-50
 *        + IoUCost('giou', w_iou)  (bbox_overlaps eps 1e-6)              match_cost.py:169-185
 *   cost.cpu(); scipy.optimize.linear_sum_assignment(cost)               hungarian_assigner.py:131-140
 *   gt_inds[row] = col + 1; labels[row] = gt_labels[col]                 hungarian_assigner.py:142-148
 * with no device->host copy and no host synchronisation.
 *
 * Layout.  Problem p uses predictions cls_pred[p] (Q, C) / bbox_pred[p] (Q, 4: cx,cy,w,h in [0,1])
 * and ground-truth segment s = prob_seg[p]: boxes gt_bboxes[seg_offsets[s] .. seg_offsets[s+1])
 * (x1,y1,x2,y2 in pixels), labels gt_labels[...], image size seg_img_wh[s] = (w, h) of
 * img_meta['img_shape'].  prob_seg / seg_offsets are int32 on device, gt_labels int64.
 * cost_offsets (P+1, int64, device): element offset of each problem's matrix in `cost`
 * (Q*G_p elements each).
 * ------------------------------------------------------------------------------------------ */

/* Builds the fp32 cost matrices.  cost_qg (optional, may be NULL): reference layout (Q, G_p)
 * row-major.  cost_solver (required): the layout sdb_lsap_solve_f32 consumes -- row-major with
 * the SMALLER dimension as rows, i.e. (G_p, Q) when Q > G_p (scipy solves the transpose of a tall
 * matrix), else (Q, G_p). */
int sdb_match_cost_f32(sdb_stream_t stream, const float* cls_pred, const float* bbox_pred,
                       const float* gt_bboxes, const int64_t* gt_labels, const int32_t* prob_seg,
                       const int32_t* seg_offsets, const float* seg_img_wh,
                       const int64_t* cost_offsets, int num_problems, int num_query,
                       int num_classes, float w_cls, float w_l1, float w_iou, float* cost_qg,
                       float* cost_solver);

/* Exact rectangular linear-sum-assignment, one CTA per problem, float64 duals, reproducing
 * scipy.optimize.linear_sum_assignment's result including its tie-breaking (see oracle/lsap.c
 * for the step-for-step CPU twin).  For problem p with G_p = seg size of prob_seg[p]:
 *   gt_inds[p, q]  = 0 (background) or k+1 (matched to GT k)           int64 (P, Q)
 *   labels[p, q]   = -1 or gt_labels[k]  (gt_labels may be NULL -> labels not written)
 *   status[p]      = 0 ok, 1 invalid entry (NaN / -inf), 2 infeasible   int32 (P,)
 *                    (scipy raises ValueError in those cases; the host wrapper checks status
 *                     lazily so the hot path stays sync-free)
 * `max_gt` (host int) is an upper bound on every G_p; it sizes the kernel's shared memory.
 * max(Q, max_gt) is limited by shared memory to SDB_LSAP_MAX_DIM. */
#define SDB_LSAP_MAX_DIM 4096
int sdb_lsap_solve_f32(sdb_stream_t stream, const float* cost_solver, const int64_t* cost_offsets,
                       const int32_t* prob_seg, const int32_t* seg_offsets,
                       const int64_t* gt_labels, int num_problems, int num_query, int max_gt,
                       int64_t* gt_inds, int64_t* labels, int32_t* status);

/* Convenience: sdb_match_cost_f32 followed by sdb_lsap_solve_f32 on the same stream.
 * `workspace` must hold cost_offsets[P] floats. */
int sdb_hungarian_assign_f32(sdb_stream_t stream, const float* cls_pred, const float* bbox_pred,
                             const float* gt_bboxes, const int64_t* gt_labels,
                             const int32_t* prob_seg, const int32_t* seg_offsets,
                             const float* seg_img_wh, const int64_t* cost_offsets,
                             int num_problems, int num_query, int num_classes, int max_gt,
                             float w_cls, float w_l1, float w_iou, float* workspace, float* cost_qg,
                             int64_t* gt_inds, int64_t* labels, int32_t* status);

/* ------------------------------------------------------------------------------------------
 * Mean-teacher EMA, all parameters in ONE launch.
 *
 * Replaces MeanTeacher.momentum_update (detr_ssod/utils/hooks/mean_teacher.py:60-64), a Python
 * loop of `teacher.mul_(m).add_(student, alpha=1-m)` over ~430 tensors (~860 launches):
 *   teacher[i] = fma((float)(1-m), student[i], (float)m * teacher[i])
 * `chunks` is a device table that tiles the parameter list; entry c covers
 * count[c] floats starting at teacher_ptr[c] / student_ptr[c] (a flat parameter buffer is the
 * special case of contiguous chunks).  momentum is the Python double of mean_teacher.py:45-47.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  float* teacher;       /* in/out */
  const float* student; /* in */
  int64_t count;
} sdb_ema_chunk;

int sdb_ema_update_f32(sdb_stream_t stream, const sdb_ema_chunk* chunks, int num_chunks,
                       double momentum);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over d_model = 256 (the transformer's channel dimension), forward and backward.
 *
 * Replaces the nn.LayerNorm calls of the DINO encoder / decoder layers
 * (detr_od/models/utils/transformer.py:606-642, 762-791, 1039): y = (x - mean) * rstd * gamma + beta with torch's
 * statistics (biased variance, eps inside the square root).  `mean` / `rstd` (rows floats each) are saved by the
 * forward for the backward.  The backward needs a scratch buffer of sdb_layernorm_bwd_workspace_floats() floats.
 * cols other than 256 return SDB_ERR_UNSUPPORTED (the host layer then uses the library LayerNorm).
 * ------------------------------------------------------------------------------------------ */
int sdb_layernorm_bwd_workspace_floats(void);
int sdb_layernorm_forward_f32(sdb_stream_t stream, const float* x, const float* gamma, const float* beta,
                              int64_t rows, int cols, float eps, float* y, float* mean, float* rstd);
int sdb_layernorm_backward_f32(sdb_stream_t stream, const float* dy, const float* x, const float* gamma,
                               const float* mean, const float* rstd, int64_t rows, int cols, float* dx,
                               float* dgamma, float* dbeta, float* workspace);
/* The post-norm blocks of the DINO layers, `x = norm(x + dropout(sublayer(x)))` (transformer.py:606-642, 762-791), with
 * the residual add inside the LayerNorm pass: y = LN(x + residual) (residual may be NULL).  Optionally the kernel also
 * writes y_plus_pos = y + pos -- the next encoder layer's query (`with_pos_embed`, transformer.py:635) -- from the
 * registers that hold y.  backward: dx = d(x + residual) from dy (+ dy_plus_pos, may be NULL); the caller hands the same
 * dx to both addends.  x / residual are re-read and re-added, so no (rows, 256) sum is kept between the passes. */
int sdb_add_layernorm_forward_f32(sdb_stream_t stream, const float* x, const float* residual, const float* gamma,
                                  const float* beta, int64_t rows, int cols, float eps, float* y, float* mean,
                                  float* rstd, const float* pos, float* y_plus_pos);
int sdb_add_layernorm_backward_f32(sdb_stream_t stream, const float* dy, const float* dy_plus_pos, const float* x,
                                   const float* residual, const float* gamma, const float* mean, const float* rstd,
                                   int64_t rows, int cols, float* dx, float* dgamma, float* dbeta, float* workspace);


/* ------------------------------------------------------------------------------------------
 * Column sums of a row-major (rows, cols) fp32 matrix into out[cols] (zero-filled by the call): the bias gradient
 * grad_output.sum(0) of the linears on the path (transformer.py:596-630, 765-791; ms_deform_attn.py:52-55).
 * cols must be a multiple of 4 and x 16-byte aligned, else SDB_ERR_UNSUPPORTED.
 * ------------------------------------------------------------------------------------------ */
int sdb_colsum_f32(sdb_stream_t stream, const float* x, int64_t rows, int cols, float* out);

/* ReLU backward fused with the bias gradient of the linear layer in front of it (FFN linear1 -> ReLU,
 * transformer.py:628, 880): g = dy where y > 0 else 0, colsum[c] = sum_r g[r, c] (zero-filled by the call), one pass.
 * Same shape / alignment requirements as sdb_colsum_f32; g must not alias the inputs. */
int sdb_relu_backward_colsum_f32(sdb_stream_t stream, const float* dy, const float* y, int64_t rows, int cols,
                                 float* g, float* colsum);
/* The same two passes on bf16 storage (bit patterns as uint16_t; 8-byte aligned, cols % 4 == 0) with fp32 sums: the
 * bias gradients of the linears under the bf16 autocast step of BASELINE.json configs[3]. */
int sdb_colsum_bf16(sdb_stream_t stream, const uint16_t* x, int64_t rows, int cols, float* out);
int sdb_relu_backward_colsum_bf16(sdb_stream_t stream, const uint16_t* dy, const uint16_t* y, int64_t rows, int cols,
                                  uint16_t* g, float* colsum);

/* ------------------------------------------------------------------------------------------
 * Decoder self-attention core (the softmax(q k^T / sqrt(d) + mask) v inside nn.MultiheadAttention as the DINO decoder
 * layer calls it: transformer.py:765, 795-812; mask from dn_components.py:97-113), scores kept on chip.
 *
 * Operands are (T, B, H*D) tensors addressed through strides: element (t, b, h, c) of q lives at
 * q[t * q_tok + b * q_bat + h * D + c] (so q and k may be the two halves of one in-projection output); all pointers
 * 16-byte aligned, strides multiples of 4 floats.  D must be 32 (else SDB_ERR_UNSUPPORTED).  `mask_add` is the additive
 * (T, T) float mask (0 / -inf, row = query) or NULL; `scale` multiplies q before the product (torch's order).
 *   forward : out (T, B, H*D) contiguous, lse (B*H, T) = log-sum-exp of every score row (saved for the backward)
 *   backward: dq / dk / dv through the same kind of strided views; needs the transposed mask too (`mask_add_t`, (T, T),
 *             row = key; NULL iff mask_add is NULL) and a (B*H, T) float scratch `delta`.
 * `tile_flags` (optional, may be NULL): one byte per 64 x 64 tile of the mask, (ceil(T/64), ceil(T/64)) row = query block;
 * bit 0 = the tile holds a non-zero mask value, bit 1 = every element of the tile is -inf.  Fully masked tiles are
 * skipped and unmasked tiles never read the mask (the denoising mask is block-structured, dn_components.py:97-113).
 * The four products are TF32 tensor-core contractions with fp32 accumulation (the rounding of torch's TF32 matmul
 * mode, which is when the host layer uses this path); softmax statistics and all sums are fp32.
 * ------------------------------------------------------------------------------------------ */
int sdb_mha_forward_f32(sdb_stream_t stream, const float* q, int64_t q_tok, int64_t q_bat, const float* k,
                        int64_t k_tok, int64_t k_bat, const float* v, int64_t v_tok, int64_t v_bat,
                        const float* mask_add, const unsigned char* tile_flags, int T, int B, int H, int D, float scale,
                        float* out, float* lse);
int sdb_mha_backward_f32(sdb_stream_t stream, const float* q, int64_t q_tok, int64_t q_bat, const float* k,
                         int64_t k_tok, int64_t k_bat, const float* v, int64_t v_tok, int64_t v_bat,
                         const float* mask_add, const float* mask_add_t, const unsigned char* tile_flags,
                         const float* out, const float* dout,
                         const float* lse, int T, int B, int H, int D, float scale, float* dq, int64_t dq_tok,
                         int64_t dq_bat, float* dk, int64_t dk_tok, int64_t dk_bat, float* dv, int64_t dv_tok,
                         int64_t dv_bat, float* delta);

/* ------------------------------------------------------------------------------------------
 * Gradient clip + AdamW (+ mean-teacher EMA) in one pass over flat fp32 buffers.
 *
 * Replaces mmcv's OptimizerHook for configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128 (clip_grad_norm_ 0.1,
 * AdamW lr 1e-4 / wd 1e-4, backbone lr x0.1) and, when `teacher` is non-NULL, the next iteration's
 * MeanTeacher.momentum_update (mean_teacher.py:60-64) for the same parameters.
 *   params / grads / exp_avg / exp_avg_sq / teacher : flat device buffers, identical layout
 *   clip_coef  : device scalar min(1, max_norm / (||g|| + 1e-6)) or NULL (no clipping)
 *   step_count : device scalar, number of updates done so far (the caller increments it after the call)
 *   seg_bounds (host, 2*num_segs int64: [begin, end) element ranges, multiples of 4), seg_lr, seg_weight_decay
 *   (host, num_segs floats): parameter groups
 * Arithmetic follows torch.optim.AdamW; the EMA keeps the reference's two roundings.
 * ------------------------------------------------------------------------------------------ */
int sdb_adamw_ema_step_f32(sdb_stream_t stream, float* params, const float* grads, float* exp_avg,
                           float* exp_avg_sq, float* teacher, const float* clip_coef, const float* step_count,
                           const int64_t* seg_bounds, const float* seg_lr, const float* seg_weight_decay,
                           int num_segs, float beta1, float beta2, float eps, double ema_momentum);
/* The same step with the per-group hyper-parameters in DEVICE memory: seg_hparams_dev = num_segs x (lr, weight_decay)
 * floats.  A step captured in a CUDA graph then follows the learning-rate schedule of the config
 * (lr_config step decay, dino_detr_r50_8x2_12e_coco.py:129) by rewriting that small buffer between replays. */
int sdb_adamw_ema_step_sched_f32(sdb_stream_t stream, float* params, const float* grads, float* exp_avg,
                                 float* exp_avg_sq, float* teacher, const float* clip_coef,
                                 const float* step_count, const int64_t* seg_bounds, const float* seg_hparams_dev,
                                 int num_segs, float beta1, float beta2, float eps, double ema_momentum);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange fused with clip + AdamW over NVLink peer memory.
 *
 * Replaces DistributedDataParallel's gradient all-reduce followed by mmcv's OptimizerHook (detr_ssod/apis/train.py:84-93;
 * configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128) for one process per GPU on an NVSwitch node.  `grads` and
 * `params` are this rank's SYMMETRIC flat buffers (same size and offset on every rank, bound to a multicast object);
 * `grads_mc` / `params_mc` are the multicast addresses of the same buffers.  Rank r sums shard r of every rank's
 * gradients with `multimem.ld_reduce` (the switch adds), exchanges the shard's squared norm, clips by the norm of the
 * MEAN gradient (`grad_scale` = 1/world, `max_grad_norm` <= 0: no clipping), applies AdamW to shard r only (its slice of
 * exp_avg / exp_avg_sq) and broadcasts the new parameters into every rank's `params` with `multimem.st`.  The call
 * returns (in stream order) when every rank's parameter buffer is complete.  `ctrl_ptrs`: HOST array of `world` device
 * pointers, entry r = rank r's control block (sdb_dp_ctrl_bytes() bytes, symmetric, zero-filled once) as mapped in THIS
 * process.  `seg_bounds` (host) / `seg_hparams_dev` (device: lr, weight_decay per segment) as in
 * sdb_adamw_ema_step_sched_f32.  All launches are CUDA-graph replayable; a peer that never arrives sets the word at
 * sdb_dp_error_word_offset() of the control block instead of hanging the device.
 * sdb_dp_small_allreduce_f32: in-place sum over ranks of n <= 2 floats (the loss normalisers, dino_detr_head.py:698-723)
 * through the same control blocks; `slot` (0..3) separates independent call sites.
 * ------------------------------------------------------------------------------------------ */
int sdb_dp_ctrl_bytes(void);
int sdb_dp_error_word_offset(void);
int sdb_dp_adamw_exchange_f32(sdb_stream_t stream, int rank, int world, const void* const* ctrl_ptrs, float* grads,
                              const float* grads_mc, float* params, float* params_mc, float* exp_avg,
                              float* exp_avg_sq, const float* step_count, const int64_t* seg_bounds,
                              const float* seg_hparams_dev, int num_segs, float beta1, float beta2, float eps,
                              float max_grad_norm, float grad_scale, int64_t total);
int sdb_dp_small_allreduce_f32(sdb_stream_t stream, int rank, int world, const void* const* ctrl_ptrs, int slot,
                               float* values, int n);

/* ------------------------------------------------------------------------------------------
 * Classification + box losses of the DINO head for all (decoder layer, image) problems in one launch, targets
 * gathered from the assignment inside the kernel.
 *
 * Replaces the 13 x loss_single / get_targets chain of detr_od/models/dense_heads/dino_detr_head.py:634-736, 895-980
 * (FocalLoss mmdet losses/focal_loss.py:12-57, L1Loss smooth_l1_loss.py:34-46, GIoULoss iou_loss.py:101-116).
 *   cls_scores (P, Q, C) logits; bbox_preds (P, Q, 4) normalised cxcywh
 *   gt_inds (P, Q) int64: 0 = background, k + 1 = GT k of the problem's segment (HungarianAssigner output)
 *   prob_seg (P,) int32 segment of each problem; seg_offsets (nseg + 1,) int32; gt_bboxes (G, 4) pixel xyxy;
 *   gt_labels (G,) int64; img_wh (nseg, 2); cls_weight (P,) or NULL
 *   sums (P, 5) = { focal, l1, l1_xy, l1_hw, 1 - giou } summed over the problem's queries in a fixed order
 *   (un-normalised, un-weighted: the normalisers involve a cross-rank mean, dino_detr_head.py:698-723).
 * backward: grad_sums (P, 5) -> grad_cls_scores (P, Q, C), grad_bbox_preds (P, Q, 4), every element written.
 * ------------------------------------------------------------------------------------------ */
int sdb_detr_loss_forward_f32(sdb_stream_t stream, const float* cls_scores, const float* bbox_preds,
                              const int64_t* gt_inds, const int32_t* prob_seg, const int32_t* seg_offsets,
                              const float* gt_bboxes, const int64_t* gt_labels, const float* img_wh,
                              const float* cls_weight, int num_problems, int num_query, int num_classes, float alpha,
                              float gamma, float giou_eps, float* sums);
int sdb_detr_loss_backward_f32(sdb_stream_t stream, const float* cls_scores, const float* bbox_preds,
                               const int64_t* gt_inds, const int32_t* prob_seg, const int32_t* seg_offsets,
                               const float* gt_bboxes, const int64_t* gt_labels, const float* img_wh,
                               const float* cls_weight, const float* grad_sums, int num_problems, int num_query,
                               int num_classes, float alpha, float gamma, float giou_eps, float* grad_cls_scores,
                               float* grad_bbox_preds);

/* ------------------------------------------------------------------------------------------
 * Pseudo-label side path of the teacher-student step on the device.
 *
 * sdb_pseudo_label_nms_f32: the teacher's detections -> pseudo boxes, one CTA per image, no host round trip.
 *   Class-wise greedy NMS (mmdet multiclass_nms / mmcv batched_nms as called by
 *   detr_od/models/dense_heads/dino_detr_ssod_head.py:1371-1395: score > score_thr, IoU > iou_thr suppresses within a
 *   class, the first max_per_img survivors in descending score order), then -- if apply_mean_std_filter -- keep
 *   score >= mean + std (unbiased; NaN for a single detection keeps nothing) and w > 0, h > 0
 *   (detr_ssod/models/dino_detr_ssod.py:921-939).
 *     scores_sorted / index_sorted (batch, num_candidates): every image's candidates in DESCENDING score order,
 *         index = query * num_classes + class (e.g. torch.sort of the flattened sigmoid scores)
 *     boxes_xyxy (batch, num_query, 4) pixels
 *     out_boxes (batch, max_per_img, 4), out_scores / out_labels (batch, max_per_img): survivors in score order,
 *         zero past out_count[b]; nms_count (optional): survivors of the NMS before the filter
 *
 * sdb_gmm_threshold_f32: cost threshold from a two-component 1-D Gaussian mixture on the pooled matched costs
 *   (dino_detr_ssod.py:832-890; sklearn GaussianMixture(2, 'diag', reg_covar, means_init [min, max], weights .5/.5,
 *   precisions 1, one init), EM in float64): threshold[0] = cost of the most likely sample assigned to component 0
 *   (else component 1; a single cost -> itself; none -> 0), threshold[1] = number of pooled costs.
 *     costs: num_segs segments of seg_stride floats, segment s holding seg_counts[s] (device int32) values -- the
 *     padded all-gather buffer of the ranks, or one segment.  At most 4096 costs are pooled.
 * ------------------------------------------------------------------------------------------ */
int sdb_pseudo_label_nms_f32(sdb_stream_t stream, const float* scores_sorted, const int64_t* index_sorted,
                             const float* boxes_xyxy, int batch, int num_candidates, int num_query, int num_classes,
                             float score_thr, float iou_thr, int max_per_img, int apply_mean_std_filter,
                             float* out_boxes, float* out_scores, int64_t* out_labels, int32_t* out_count,
                             int32_t* nms_count);
int sdb_gmm_threshold_f32(sdb_stream_t stream, const float* costs, const int32_t* seg_counts, int num_segs,
                          int seg_stride, float tol, int max_iter, double reg_covar, float* threshold);

/* ------------------------------------------------------------------------------------------
 * Token-wise linear layers as tcgen05 (5th-generation tensor core) GEMMs, TF32 arithmetic on fp32 storage with
 * fp32 accumulation in tensor memory -- the three products of the nn.Linear layers on the path
 * (ms_deform_attn.py:61-65, 94-112 value_proj / sampling_offsets / attention_weights / output_proj;
 * transformer.py:626-630, 878-882 FFN; :1277-1290 enc_output), which the reference runs through cuBLAS:
 *
 *   y[m,n] (row-major) = op(a)[m,k] . op(b)[k,n] (+ bias[n]) (ReLU) (rows with row_mask != 0 -> 0)
 *
 *   a_mn_major == 0 : a is stored (m, k) row-major          a_mn_major != 0 : a is stored (k, m) row-major
 *   b_mn_major == 0 : b is stored (n, k) row-major (an nn.Linear weight)
 *   b_mn_major != 0 : b is stored (k, n) row-major
 *   forward  y = x W^T + b : (x, 0, W, 0)      grad x = dy W : (dy, 0, W, 1)      grad W = dy^T x : (dy, 1, x, 1)
 *   row_mask : optional (m,) bytes -- value.masked_fill(input_padding_mask[..., None], 0) of ms_deform_attn.py:96-97
 *   k_splits > 1 : the contraction is split into k_splits ranges whose partial products are REDUCE-ADDED into y
 *                  (y must hold the initial value, e.g. zeros or the gradient being accumulated); no epilogue then.
 *   round_mode : bit 0 / bit 1 -- round a / b to the nearest TF32 value (in shared memory, after the TMA load) before
 *                the tensor core reads it.  The tensor core itself truncates (a systematic -7e-4 relative bias);
 *                3 reproduces the unbiased round-to-nearest TF32 product of cuBLAS, 0 is the raw truncating product.
 *   a_column_sums : optional (m,) fp32, only with a_mn_major != 0 -- sum_k a[k, :] is ADDED to it (zero it first): the
 *                bias gradient grad_output.sum(0) of the same layer, taken from the tiles the grad-weight product
 *                stages anyway instead of a separate pass over grad_output.
 * Requirements: 16-byte aligned pointers, n % 4 == 0, and the contiguous dimension of every operand % 4 == 0.
 * ------------------------------------------------------------------------------------------ */
int sdb_gemm_tf32(sdb_stream_t stream, const float* a, int a_mn_major, const float* b, int b_mn_major, float* y,
                  int m, int n, int k, const float* bias, const uint8_t* row_mask, int relu, int k_splits,
                  int round_mode, float* a_column_sums);

/* Grad-input product of the linear layer BEHIND a ReLU, carrying that ReLU's backward and the bias gradient of the layer
 * in front of it in its epilogue (the FFN backward of detr_od/models/utils/transformer.py:626-630 and :878-882,
 * linear2(dropout(relu(linear1(x)))): autograd runs mm -> threshold_backward -> sum there, three passes over the
 * (tokens, d_ffn) gradient):
 *   y[i, j] = (a . op(b))[i, j] * (relu_src[i, j] > 0)          y, relu_src (m, n) row-major
 *   y_column_sums[j] = sum_i y[i, j]                            optional (n,), overwritten
 * a / b / round_mode as in sdb_gemm_tf32 (grad x = dy W : (dy, 0, W, 1)); relu_src 16-byte aligned. */
int sdb_gemm_tf32_relu_grad(sdb_stream_t stream, const float* a, int a_mn_major, const float* b, int b_mn_major,
                            float* y, int m, int n, int k, const float* relu_src, float* y_column_sums, int round_mode);

#ifdef __cplusplus
}
#endif
#endif /* SEMIDETR_B200_H_ */
