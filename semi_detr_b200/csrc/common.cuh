// Shared helpers for the semidetr_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/semidetr_b200.h"

namespace sdb {

// thread-local last-error message (abi.cu)
void set_error(const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
int sm_count();

// predicated form: no branch around the reduction
__device__ __forceinline__ void red_add_f4_if(bool pred, float* p, float x, float y, float z, float w) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
      ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w), "r"((int)pred));  // no "memory" clobber: grad_value is
  // write-only in the kernel, and a clobber would pin every corner load behind the previous point's reductions
}

}  // namespace sdb

#define SDB_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      sdb::set_error(__VA_ARGS__);   \
      return SDB_ERR_INVALID_ARG;    \
    }                                \
  } while (0)

#define SDB_CUDA(call)                                      \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return sdb::cuda_error(e__, #call); \
  } while (0)

#define SDB_LAUNCH_CHECK(name)                                  \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return sdb::cuda_error(e__, name);  \
  } while (0)

namespace sdb {

// streaming (read-once) loads: keep them out of L1 so the gathered value lines stay resident
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_f2(float2* p, const float2& v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
// Four consecutive channels of one pixel / query as the tuned MSDA kernels move them: fp32 storage (16 bytes) or
// bf16 storage (8 bytes, widened to fp32 in registers -- all arithmetic stays fp32).
template <typename V>
struct Chan4;
template <>
struct Chan4<float> {
  static constexpr int kAlignMask = 15;
  static __device__ __forceinline__ float4 gather(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ float4 stream_in(const float* p) {
    return ld_stream_f4(reinterpret_cast<const float4*>(p));
  }
  static __device__ __forceinline__ void stream_out(float* p, const float4& v) {
    st_stream_f4(reinterpret_cast<float4*>(p), v);
  }
};
template <>
struct Chan4<__nv_bfloat16> {
  static constexpr int kAlignMask = 7;
  // element 0 sits in the low half of the first 32-bit word; bf16 -> fp32 is a 16-bit left shift
  static __device__ __forceinline__ float4 widen(const uint2& r) {
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u),
                       __uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
  }
  static __device__ __forceinline__ float4 gather(const __nv_bfloat16* p) {
    return widen(__ldg(reinterpret_cast<const uint2*>(p)));
  }
  static __device__ __forceinline__ float4 stream_in(const __nv_bfloat16* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return widen(r);
  }
  static __device__ __forceinline__ void stream_out(__nv_bfloat16* p, const float4& v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);  // round to nearest even
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p),
                 "r"(*reinterpret_cast<const unsigned*>(&lo)), "r"(*reinterpret_cast<const unsigned*>(&hi))
                 : "memory");
  }
};

// 16-byte vector reduction into global memory (sm_90+): one L2 atomic transaction per 4 floats
__device__ __forceinline__ void red_add_f4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace sdb
