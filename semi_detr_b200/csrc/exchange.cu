// Data-parallel gradient exchange fused with the optimizer, over NVLink peer memory.  sm_100a.
//
// The reference trains under DistributedDataParallel: NCCL all-reduces the gradients (mean), then mmcv's OptimizerHook
// clips and steps (/root/reference/detr_ssod/apis/train.py:84-93, configs/dino_detr/dino_detr_r50_8x2_12e_coco.py:122-128).
// As separate passes that is: all-reduce of the 188 MB flat gradient (read + write everything, on every rank), a norm
// reduction over it, then clip + AdamW over parameters, gradients and both moments -- every rank repeating the same
// update on all 47 M elements.  Here the exchange IS the optimizer pass:
//
//   kernel R  rank r owns shard r (1/world of the flat buffers).  One `multimem.ld_reduce.add.v4.f32` per 16 bytes
//             returns the SUM over all ranks' gradient buffers -- the NVSwitch adds the operands in flight (NVLS), no
//             rank ever reads another rank's memory piecewise -- the sum is kept in the rank's own buffer, and its
//             squared norm goes, as one scalar per rank, to every peer's control block.
//   kernel U  clip coefficient from the world's partial norms (summed in rank order: bit-identical on every rank),
//             AdamW on the shard (gradient, parameter and both moments of the shard only), and the new parameters leave
//             through `multimem.st.v4.f32`: one store, replicated by the switch into every rank's parameter buffer.
//   kernel W  waits until every peer's parameter stores have landed and advances the epoch.
//
// Traffic per rank and step: 1/world of the gradient in, 1/world of the parameters out, optimizer state touched once
// per element per NODE instead of once per rank.  Buffers are symmetric allocations (same offset on every rank, mapped
// peer-to-peer and bound to a multicast object; the host layer gets them from torch's symmetric memory); flags and the
// norm partials live in a small symmetric control block per rank and carry a monotonically increasing epoch, so the
// three launches replay from a CUDA graph.  Every spin is bounded: a peer that never arrives sets the error word
// instead of hanging the GPU.
#include <algorithm>

#include "common.cuh"

namespace sdb {

namespace {

constexpr int kXThreads = 512;
constexpr int kMaxWorld = 16;
// control block layout (32-bit words)
constexpr int kArrive = 0;                  // [kArrive + r]   rank r has finished writing its gradients (epoch)
constexpr int kNormFlag = kMaxWorld;        // [kNormFlag + r] rank r's norm partial is in place (epoch)
constexpr int kDone = 2 * kMaxWorld;        // [kDone + r]     rank r's parameter stores are out (epoch)
constexpr int kNormVal = 3 * kMaxWorld;     // [kNormVal + 2r] rank r's squared-norm partial (double)
constexpr int kEpoch = 5 * kMaxWorld;       // local: epochs completed
constexpr int kGo = kEpoch + 1;             // local: CTA 0 has seen every rank arrive (epoch)
constexpr int kCountR = kEpoch + 2;         // local: CTAs of kernel R that have finished
constexpr int kCountU = kEpoch + 3;
constexpr int kError = kEpoch + 4;          // local: a bounded spin ran out
constexpr int kAcc = kEpoch + 6;            // local: double accumulator of the shard's squared norm (8-byte aligned)
constexpr int kSmallBase = 128;             // small all-reduce slots: [kSmallBase + slot * 64 ...]
constexpr int kSmallSlots = 4;
constexpr int kCtrlWords = kSmallBase + kSmallSlots * 128;
constexpr long long kSpinLimit = 120000000000LL;   // ~60 s at 1.9 GHz: ranks may be seconds apart (first-step autotuning)

struct Peers {
  unsigned* ctrl[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// epochs only grow; the comparison survives wrap-around
__device__ __forceinline__ bool reached(unsigned seen, unsigned want) { return (int)(seen - want) >= 0; }

__device__ bool spin_sys(const unsigned* p, unsigned want, unsigned* err) {
  const long long t0 = clock64();
  while (!reached(ld_acquire_sys(p), want)) {
    if (clock64() - t0 > kSpinLimit) {
      *err = 1u;
      return false;
    }
    __nanosleep(64);
  }
  return true;
}
__device__ bool spin_gpu(const unsigned* p, unsigned want, unsigned* err) {
  const long long t0 = clock64();
  while (!reached(ld_acquire_gpu(p), want)) {
    if (clock64() - t0 > kSpinLimit) {
      *err = 1u;
      return false;
    }
    __nanosleep(32);
  }
  return true;
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < kXThreads / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;   // valid in thread 0
}

// ---- R: wait for every rank's gradients, shard sum through the switch, squared norm of the shard ----------------------
__global__ void __launch_bounds__(kXThreads)
dp_reduce_kernel(Peers peers, int rank, int world, float* __restrict__ grads, const float* __restrict__ grads_mc,
                 long long shard_begin, long long shard_end) {
  __shared__ double sh[kXThreads / 32];
  unsigned* ctrl = peers.ctrl[rank];
  const unsigned epoch = ctrl[kEpoch] + 1u;
  if (blockIdx.x == 0) {
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(peers.ctrl[threadIdx.x] + kArrive + rank, epoch);
      spin_sys(ctrl + kArrive + threadIdx.x, epoch, ctrl + kError);
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(ctrl + kGo, epoch);
  } else {
    if (threadIdx.x == 0) spin_gpu(ctrl + kGo, epoch, ctrl + kError);
    __syncthreads();
  }
  double acc = 0.0;
  const long long n4 = (shard_end - shard_begin) >> 2;
  float4* g4 = reinterpret_cast<float4*>(grads + shard_begin);
  const float4* m4 = reinterpret_cast<const float4*>(grads_mc + shard_begin);
  for (long long i = (long long)blockIdx.x * kXThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kXThreads) {
    const float4 s = multimem_ld_reduce_add(reinterpret_cast<const float*>(m4 + i));
    g4[i] = s;
    acc += (double)(s.x * s.x + s.y * s.y) + (double)(s.z * s.z + s.w * s.w);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(reinterpret_cast<double*>(ctrl + kAcc), acc);
    __threadfence();
    if (atomicAdd(ctrl + kCountR, 1u) == gridDim.x - 1) {   // the last CTA publishes the shard's partial
      __threadfence();
      const double total = atomicAdd(reinterpret_cast<double*>(ctrl + kAcc), 0.0);
      for (int r = 0; r < world; ++r) {
        *reinterpret_cast<volatile double*>(peers.ctrl[r] + kNormVal + 2 * rank) = total;
        __threadfence_system();
        st_release_sys(peers.ctrl[r] + kNormFlag + rank, epoch);
      }
      *reinterpret_cast<double*>(ctrl + kAcc) = 0.0;
      ctrl[kCountR] = 0u;
    }
  }
}

struct XSeg {
  long long begin, end;   // element range of the optimizer segment (hyper-parameter group) in the flat buffers
};
struct XArgs {
  XSeg seg[4];
  int nseg;
  float beta1, beta2, eps, max_norm, grad_scale;
};

// ---- U: clip coefficient from the world's partials, AdamW on the shard, parameters out through the switch --------------
__global__ void __launch_bounds__(kXThreads)
dp_update_kernel(Peers peers, int rank, int world, float* __restrict__ params_mc, const float* __restrict__ params,
                 const float* __restrict__ grads, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                 const float* __restrict__ step, const float* __restrict__ hparams, long long shard_begin,
                 long long shard_end, XArgs a) {
  __shared__ float coef_sh;
  unsigned* ctrl = peers.ctrl[rank];
  const unsigned epoch = ctrl[kEpoch] + 1u;
  if (threadIdx.x < world) spin_sys(ctrl + kNormFlag + threadIdx.x, epoch, ctrl + kError);
  __syncthreads();
  if (threadIdx.x == 0) {
    double sq = 0.0;
    for (int r = 0; r < world; ++r) sq += *reinterpret_cast<volatile double*>(ctrl + kNormVal + 2 * r);
    float coef = a.grad_scale;
    if (a.max_norm > 0.f) {
      const float total_norm = (float)sqrt(sq) * a.grad_scale;     // norm of the MEAN gradient
      coef = fminf(a.max_norm / (total_norm + 1e-6f), 1.f) * a.grad_scale;
    }
    coef_sh = coef;
  }
  __syncthreads();
  const float coef = coef_sh;
  const float t = step[0] + 1.f;
  const float bc1 = 1.f - powf(a.beta1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(a.beta2, t));
  for (int s = 0; s < a.nseg; ++s) {
    const long long lo = max(a.seg[s].begin, shard_begin), hi = min(a.seg[s].end, shard_end);
    if (hi <= lo) continue;
    const float lr = hparams[2 * s], wd = hparams[2 * s + 1];
    const float step_size = lr / bc1, decay = 1.f - lr * wd;
    const long long n4 = (hi - lo) >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(params + lo);
    const float4* g4 = reinterpret_cast<const float4*>(grads + lo);
    float4* m4 = reinterpret_cast<float4*>(exp_avg + lo);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq + lo);
    float4* o4 = reinterpret_cast<float4*>(params_mc + lo);
    for (long long i = (long long)blockIdx.x * kXThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kXThreads) {
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      const float4 gg = g4[i];
      float* pf = &pp.x; float* mf = &mm.x; float* vf = &vv.x; const float* gf = &gg.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {     // the arithmetic of adamw_ema_kernel (optimizer.cu)
        const float gr = gf[k] * coef;
        float w = pf[k] * decay;
        mf[k] = a.beta1 * mf[k] + (1.f - a.beta1) * gr;
        vf[k] = a.beta2 * vf[k] + (1.f - a.beta2) * gr * gr;
        const float denom = sqrtf(vf[k]) / bc2_sqrt + a.eps;
        w -= step_size * (mf[k] / denom);
        pf[k] = w;
      }
      m4[i] = mm;
      v4[i] = vv;
      multimem_st(reinterpret_cast<float*>(o4 + i), pp);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(ctrl + kCountU, 1u) == gridDim.x - 1) {
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(peers.ctrl[r] + kDone + rank, epoch);
      ctrl[kCountU] = 0u;
    }
  }
}

// ---- W: every peer's parameter stores have landed; next epoch ----------------------------------------------------------
__global__ void dp_finish_kernel(Peers peers, int rank, int world) {
  unsigned* ctrl = peers.ctrl[rank];
  const unsigned epoch = ctrl[kEpoch] + 1u;
  if (threadIdx.x < world) spin_sys(ctrl + kDone + threadIdx.x, epoch, ctrl + kError);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    ctrl[kEpoch] = epoch;
  }
}

// ---- small all-reduce (sum) of n <= 2 floats in place: every rank writes its values into every peer's slot ----------
// (the two loss normalisers of the step, dino_detr_head.py:698-723: a NCCL launch each costs ~80 us of latency in the
// middle of the step; this is one 32-thread kernel and one NVLink round trip).  Slot layout (128 words): [0] local
// epoch; [8 + r] rank r's flag; [32 + (parity * 16 + r) * 2 ...] rank r's values, double-buffered by epoch parity so a
// fast rank's next round cannot overwrite what a slow rank still reads.
__global__ void dp_small_allreduce_kernel(Peers peers, int rank, int world, int slot, float* __restrict__ values, int n) {
  unsigned* ctrl = peers.ctrl[rank];
  unsigned* base = ctrl + kSmallBase + slot * 128;
  const unsigned epoch = base[0] + 1u;
  const int r = threadIdx.x;
  const int par = (int)(epoch & 1u);
  float mine[2] = {0.f, 0.f};
  for (int i = 0; i < n; ++i) mine[i] = values[i];
  if (r < world) {
    unsigned* dst = peers.ctrl[r] + kSmallBase + slot * 128;
    volatile float* out = reinterpret_cast<volatile float*>(dst + 32 + (par * kMaxWorld + rank) * 2);
    for (int i = 0; i < n; ++i) out[i] = mine[i];
    __threadfence_system();
    st_release_sys(dst + 8 + rank, epoch);
    spin_sys(base + 8 + r, epoch, ctrl + kError);
  }
  __syncthreads();
  if (r < n) {
    float s = 0.f;
    for (int q = 0; q < world; ++q)      // rank order: the same sum on every rank
      s += *(reinterpret_cast<volatile float*>(base + 32 + (par * kMaxWorld + q) * 2) + r);
    values[r] = s;
  }
  __syncthreads();
  if (r == 0) base[0] = epoch;
}

int check_peers(const char* what, int rank, int world, const void* const* ctrl_ptrs, Peers& peers) {
  SDB_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "%s: rank %d of %d (max %d)", what, rank, world,
              kMaxWorld);
  SDB_REQUIRE(ctrl_ptrs != nullptr, "%s: null control-block table", what);
  for (int r = 0; r < world; ++r) {
    SDB_REQUIRE(ctrl_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(ctrl_ptrs[r]) & 15) == 0,
                "%s: control block of rank %d is null or misaligned", what, r);
    peers.ctrl[r] = reinterpret_cast<unsigned*>(const_cast<void*>(ctrl_ptrs[r]));
  }
  return SDB_OK;
}

}  // namespace

}  // namespace sdb

extern "C" {

int sdb_dp_ctrl_bytes(void) { return sdb::kCtrlWords * 4; }

int sdb_dp_adamw_exchange_f32(sdb_stream_t stream, int rank, int world, const void* const* ctrl_ptrs, float* grads,
                              const float* grads_mc, float* params, float* params_mc, float* exp_avg,
                              float* exp_avg_sq, const float* step_count, const int64_t* seg_bounds,
                              const float* seg_hparams_dev, int num_segs, float beta1, float beta2, float eps,
                              float max_grad_norm, float grad_scale, int64_t total) {
  using namespace sdb;
  Peers peers{};
  int rc = check_peers("sdb_dp_adamw_exchange_f32", rank, world, ctrl_ptrs, peers);
  if (rc != SDB_OK) return rc;
  SDB_REQUIRE(grads && grads_mc && params && params_mc && exp_avg && exp_avg_sq && step_count && seg_bounds &&
                  seg_hparams_dev, "sdb_dp_adamw_exchange_f32: null pointer");
  SDB_REQUIRE(num_segs >= 1 && num_segs <= 4, "sdb_dp_adamw_exchange_f32: num_segs=%d (1..4)", num_segs);
  SDB_REQUIRE(((reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(grads_mc) |
                reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(params_mc) |
                reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
              "sdb_dp_adamw_exchange_f32: buffers must be 16-byte aligned");
  SDB_REQUIRE(total > 0 && total % 4 == 0, "sdb_dp_adamw_exchange_f32: total=%lld must be a positive multiple of 4",
              (long long)total);
  XArgs a{};
  a.nseg = num_segs;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_grad_norm; a.grad_scale = grad_scale;
  for (int s = 0; s < num_segs; ++s) {
    a.seg[s].begin = seg_bounds[2 * s];
    a.seg[s].end = seg_bounds[2 * s + 1];
    SDB_REQUIRE(a.seg[s].begin % 4 == 0 && a.seg[s].end % 4 == 0 && a.seg[s].begin <= a.seg[s].end && a.seg[s].end <= total,
                "sdb_dp_adamw_exchange_f32: segment %d out of range / misaligned", s);
  }
  // shards: equal 4-element-aligned slices of [0, total)
  const long long per = ((total / 4 + world - 1) / world) * 4;
  const long long lo = std::min<long long>((long long)rank * per, total), hi = std::min<long long>(lo + per, total);
  const int grid = sm_count();   // one CTA per SM: every CTA is resident, so the intra-kernel waits cannot deadlock
  cudaStream_t st = (cudaStream_t)stream;
  dp_reduce_kernel<<<grid, kXThreads, 0, st>>>(peers, rank, world, grads, grads_mc, lo, hi);
  SDB_LAUNCH_CHECK("dp_reduce_kernel");
  dp_update_kernel<<<grid, kXThreads, 0, st>>>(peers, rank, world, params_mc, params, grads, exp_avg, exp_avg_sq,
                                               step_count, seg_hparams_dev, lo, hi, a);
  SDB_LAUNCH_CHECK("dp_update_kernel");
  dp_finish_kernel<<<1, 32, 0, st>>>(peers, rank, world);
  SDB_LAUNCH_CHECK("dp_finish_kernel");
  return SDB_OK;
}

int sdb_dp_small_allreduce_f32(sdb_stream_t stream, int rank, int world, const void* const* ctrl_ptrs, int slot,
                               float* values, int n) {
  using namespace sdb;
  Peers peers{};
  int rc = check_peers("sdb_dp_small_allreduce_f32", rank, world, ctrl_ptrs, peers);
  if (rc != SDB_OK) return rc;
  SDB_REQUIRE(values != nullptr && n >= 1 && n <= 2, "sdb_dp_small_allreduce_f32: n=%d (1..2 floats)", n);
  SDB_REQUIRE(slot >= 0 && slot < kSmallSlots, "sdb_dp_small_allreduce_f32: slot %d (0..%d)", slot, kSmallSlots - 1);
  dp_small_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peers, rank, world, slot, values, n);
  SDB_LAUNCH_CHECK("dp_small_allreduce_kernel");
  return SDB_OK;
}

int sdb_dp_error_word_offset(void) { return sdb::kError * 4; }

}  // extern "C"
