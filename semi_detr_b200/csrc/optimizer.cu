// clip-by-global-norm + AdamW (+ mean-teacher EMA) as ONE pass over flat parameter buffers.  sm_100a.
//
// Replaces, per iteration, mmcv's OptimizerHook (clip_grad_norm_(max_norm 0.1) then AdamW.step with the per-group
// learning rates of /root/reference/configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128) and, for the
// semi-supervised wrapper, the MeanTeacher update of the NEXT iteration
// (/root/reference/detr_ssod/utils/hooks/mean_teacher.py:37-64 blends the student as left by this step).
// The reference runs them as ~290 x (several) small launches; here parameters, gradients and both moments live in
// flat fp32 buffers and one kernel streams them once:
//   read  g, p, m, v (+ teacher)      write p, m, v (+ teacher)       28 B (40 B with EMA) per parameter
// The clip coefficient is read from device memory (computed by one norm reduction over the flat gradient), the
// bias corrections from a device step counter, so the launch is CUDA-graph replayable.
// Arithmetic follows torch.optim.AdamW (decoupled weight decay first, then the Adam update with
// denom = sqrt(v) / sqrt(1 - beta2^t) + eps, step = lr / (1 - beta1^t)).
#include "common.cuh"

namespace sdb {

struct AdamSeg {
  long long begin, end;  // element range in the flat buffers
  float lr, weight_decay;
};

constexpr int kOptThreads = 256;
constexpr int kMaxSegs = 8;

struct AdamArgs {
  AdamSeg seg[kMaxSegs];
  int nseg;
  float beta1, beta2, eps;
  float ema_m, ema_om;  // EMA momentum and 1 - momentum (unused when teacher == nullptr)
};

__global__ void __launch_bounds__(kOptThreads)
adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ teacher, const float* __restrict__ clip_coef, const float* __restrict__ step,
                 const float* __restrict__ dev_hparams, AdamArgs a) {
  const float t = step[0] + 1.f;             // this update's 1-based step index
  const float bc1 = 1.f - powf(a.beta1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(a.beta2, t));
  const float coef = clip_coef ? clip_coef[0] : 1.f;
  for (int s = 0; s < a.nseg; ++s) {
    AdamSeg sg = a.seg[s];
    if (dev_hparams) {   // (lr, weight_decay) per segment in device memory: a captured graph follows the schedule
      sg.lr = dev_hparams[2 * s];
      sg.weight_decay = dev_hparams[2 * s + 1];
    }
    const float step_size = sg.lr / bc1;
    const float decay = 1.f - sg.lr * sg.weight_decay;
    // segments start 16-byte aligned (the host pads them), so the body is float4
    const long long n4 = (sg.end - sg.begin) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + sg.begin);
    const float4* g4 = reinterpret_cast<const float4*>(g + sg.begin);
    float4* m4 = reinterpret_cast<float4*>(m + sg.begin);
    float4* v4 = reinterpret_cast<float4*>(v + sg.begin);
    float4* t4 = teacher ? reinterpret_cast<float4*>(teacher + sg.begin) : nullptr;
    for (long long i = (long long)blockIdx.x * kOptThreads + threadIdx.x; i < n4;
         i += (long long)gridDim.x * kOptThreads) {
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      const float4 gg = ld_stream_f4(g4 + i);
      float* pf = &pp.x; float* mf = &mm.x; float* vf = &vv.x; const float* gf = &gg.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gr = gf[k] * coef;
        float w = pf[k] * decay;
        mf[k] = a.beta1 * mf[k] + (1.f - a.beta1) * gr;          // lerp form torch uses: m + (g - m) * (1 - b1)
        vf[k] = a.beta2 * vf[k] + (1.f - a.beta2) * gr * gr;
        const float denom = sqrtf(vf[k]) / bc2_sqrt + a.eps;
        w -= step_size * (mf[k] / denom);
        pf[k] = w;
      }
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
      if (t4) {
        float4 tt = t4[i];
        tt.x = fmaf(a.ema_om, pp.x, __fmul_rn(tt.x, a.ema_m));
        tt.y = fmaf(a.ema_om, pp.y, __fmul_rn(tt.y, a.ema_m));
        tt.z = fmaf(a.ema_om, pp.z, __fmul_rn(tt.z, a.ema_m));
        tt.w = fmaf(a.ema_om, pp.w, __fmul_rn(tt.w, a.ema_m));
        t4[i] = tt;
      }
    }
  }
}

}  // namespace sdb

static int adamw_ema_step(sdb_stream_t stream, float* params, const float* grads, float* exp_avg,
                          float* exp_avg_sq, float* teacher, const float* clip_coef, const float* step_count,
                          const int64_t* seg_bounds, const float* seg_lr, const float* seg_weight_decay,
                          const float* dev_hparams, int num_segs, float beta1, float beta2, float eps,
                          double ema_momentum) {
  using namespace sdb;
  SDB_REQUIRE(num_segs >= 0 && num_segs <= kMaxSegs, "adamw_ema_step: num_segs=%d (max %d)", num_segs, kMaxSegs);
  if (num_segs == 0) return SDB_OK;
  SDB_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_count && seg_bounds &&
              (dev_hparams || (seg_lr && seg_weight_decay)), "adamw_ema_step: null pointer");
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                reinterpret_cast<uintptr_t>(teacher)) & 15) == 0, "adamw_ema_step: buffers must be 16-byte aligned");
  AdamArgs a;
  a.nseg = num_segs;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.ema_m = (float)ema_momentum; a.ema_om = (float)(1.0 - ema_momentum);
  long long total = 0;
  for (int s = 0; s < num_segs; ++s) {    // host arrays: a handful of scalars
    a.seg[s].begin = seg_bounds[2 * s];
    a.seg[s].end = seg_bounds[2 * s + 1];
    a.seg[s].lr = dev_hparams ? 0.f : seg_lr[s];
    a.seg[s].weight_decay = dev_hparams ? 0.f : seg_weight_decay[s];
    SDB_REQUIRE(a.seg[s].begin % 4 == 0 && a.seg[s].end % 4 == 0 && a.seg[s].end >= a.seg[s].begin,
                "adamw_ema_step: segment %d [%lld, %lld) must be 4-element aligned", s, a.seg[s].begin, a.seg[s].end);
    total += a.seg[s].end - a.seg[s].begin;
  }
  if (total == 0) return SDB_OK;
  long long grid = (total / 4 + kOptThreads - 1) / kOptThreads;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  adamw_ema_kernel<<<(unsigned)grid, kOptThreads, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq,
                                                                            teacher, clip_coef, step_count, dev_hparams, a);
  SDB_LAUNCH_CHECK("adamw_ema_kernel");
  return SDB_OK;
}

extern "C" int sdb_adamw_ema_step_f32(sdb_stream_t stream, float* params, const float* grads, float* exp_avg,
                                      float* exp_avg_sq, float* teacher, const float* clip_coef,
                                      const float* step_count, const int64_t* seg_bounds, const float* seg_lr,
                                      const float* seg_weight_decay, int num_segs, float beta1, float beta2, float eps,
                                      double ema_momentum) {
  return adamw_ema_step(stream, params, grads, exp_avg, exp_avg_sq, teacher, clip_coef, step_count, seg_bounds, seg_lr,
                        seg_weight_decay, nullptr, num_segs, beta1, beta2, eps, ema_momentum);
}

extern "C" int sdb_adamw_ema_step_sched_f32(sdb_stream_t stream, float* params, const float* grads, float* exp_avg,
                                            float* exp_avg_sq, float* teacher, const float* clip_coef,
                                            const float* step_count, const int64_t* seg_bounds,
                                            const float* seg_hparams_dev, int num_segs, float beta1, float beta2,
                                            float eps, double ema_momentum) {
  if (!seg_hparams_dev) {
    sdb::set_error("adamw_ema_step_sched: null seg_hparams_dev");
    return SDB_ERR_INVALID_ARG;
  }
  return adamw_ema_step(stream, params, grads, exp_avg, exp_avg_sq, teacher, clip_coef, step_count, seg_bounds, nullptr,
                        nullptr, seg_hparams_dev, num_segs, beta1, beta2, eps, ema_momentum);
}
