// Mean-teacher EMA over every parameter in one launch.  sm_100a.
//
// Replaces MeanTeacher.momentum_update (/root/reference/detr_ssod/utils/hooks/mean_teacher.py:60-64): a Python
// loop issuing `teacher.mul_(m)` and `.add_(student, alpha=1-m)` per tensor (~430 tensors -> ~860 launches, each
// reading and writing the teacher twice).  Here a device-resident chunk table tiles the parameter list and one
// grid walks it: 12 bytes of HBM traffic per parameter (read student, read teacher, write teacher), 16-byte
// accesses, grid sized to the SM count.  Arithmetic keeps the reference's two roundings:
//   t <- fma((float)(1-m), s, fl((float)m * t)).
#include "common.cuh"

namespace sdb {

constexpr int kEmaThreads = 256;

__global__ void __launch_bounds__(kEmaThreads)
ema_kernel(const sdb_ema_chunk* __restrict__ chunks, int num_chunks, float m, float om) {
  for (int c = blockIdx.x; c < num_chunks; c += gridDim.x) {
    const sdb_ema_chunk ch = chunks[c];
    float* __restrict__ t = ch.teacher;
    const float* __restrict__ s = ch.student;
    const long long n = ch.count;
    if ((((uintptr_t)t | (uintptr_t)s) & 15) == 0) {
      const long long n4 = n >> 2;
      float4* t4 = reinterpret_cast<float4*>(t);
      const float4* s4 = reinterpret_cast<const float4*>(s);
      for (long long i = threadIdx.x; i < n4; i += kEmaThreads) {
        float4 a = t4[i];
        const float4 b = ld_stream_f4(s4 + i);
        a.x = fmaf(om, b.x, __fmul_rn(a.x, m));
        a.y = fmaf(om, b.y, __fmul_rn(a.y, m));
        a.z = fmaf(om, b.z, __fmul_rn(a.z, m));
        a.w = fmaf(om, b.w, __fmul_rn(a.w, m));
        t4[i] = a;
      }
      for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kEmaThreads)
        t[i] = fmaf(om, s[i], __fmul_rn(t[i], m));
    } else {
      for (long long i = threadIdx.x; i < n; i += kEmaThreads) t[i] = fmaf(om, s[i], __fmul_rn(t[i], m));
    }
  }
}

}  // namespace sdb

extern "C" int sdb_ema_update_f32(sdb_stream_t stream, const sdb_ema_chunk* chunks, int num_chunks,
                                  double momentum) {
  SDB_REQUIRE(num_chunks >= 0, "ema_update: bad num_chunks=%d", num_chunks);
  SDB_REQUIRE(momentum >= 0.0 && momentum <= 1.0, "ema_update: momentum %f outside [0,1]", momentum);
  if (num_chunks == 0) return SDB_OK;
  SDB_REQUIRE(chunks != nullptr, "ema_update: null chunk table");
  // mean_teacher.py:64 casts the python doubles m and 1-m to the tensor dtype
  const float m = (float)momentum, om = (float)(1.0 - momentum);
  int grid = sdb::sm_count() * 8;
  if (grid > num_chunks) grid = num_chunks;
  sdb::ema_kernel<<<grid, sdb::kEmaThreads, 0, (cudaStream_t)stream>>>(chunks, num_chunks, m, om);
  SDB_LAUNCH_CHECK("ema_kernel");
  return SDB_OK;
}
