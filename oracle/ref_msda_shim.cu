// C shim around the REFERENCE's own CUDA launchers, used only as a GPU-side parity cross-check
// and as the timing competitor ("reference kernel recompiled for sm_100a").
// TEST INFRASTRUCTURE: built by `make -C oracle ref` into oracle/_ref/, never linked by the product.
// It includes the reference header where it lies (path passed as -DREF_CUH=...); no reference
// source is copied into this repository.
#include REF_CUH

extern "C" int ref_msda_forward_f32(void* stream, const float* value, const int64_t* shapes,
                                    const int64_t* level_start, const float* loc, const float* attn,
                                    int batch, int spatial_size, int num_heads, int channels,
                                    int num_levels, int num_query, int num_point, float* out)
{
  // the reference zero-fills its output with at::zeros (ms_deform_attn_cuda.cu:54)
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)batch * num_query * num_heads * channels,
                  (cudaStream_t)stream);
  ms_deformable_im2col_cuda<float>((cudaStream_t)stream, value, shapes, level_start, loc, attn, batch,
                                   spatial_size, num_heads, channels, num_levels, num_query, num_point, out);
  return (int)cudaGetLastError();
}

extern "C" int ref_msda_backward_f32(void* stream, const float* grad_out, const float* value,
                                     const int64_t* shapes, const int64_t* level_start, const float* loc,
                                     const float* attn, int batch, int spatial_size, int num_heads,
                                     int channels, int num_levels, int num_query, int num_point,
                                     float* grad_value, float* grad_loc, float* grad_attn)
{
  // three at::zeros_like in the reference (ms_deform_attn_cuda.cu:121-123)
  cudaStream_t s = (cudaStream_t)stream;
  size_t nv = (size_t)batch * spatial_size * num_heads * channels;
  size_t na = (size_t)batch * num_query * num_heads * num_levels * num_point;
  cudaMemsetAsync(grad_value, 0, sizeof(float) * nv, s);
  cudaMemsetAsync(grad_loc, 0, sizeof(float) * na * 2, s);
  cudaMemsetAsync(grad_attn, 0, sizeof(float) * na, s);
  ms_deformable_col2im_cuda<float>(s, grad_out, value, shapes, level_start, loc, attn, batch, spatial_size,
                                   num_heads, channels, num_levels, num_query, num_point, grad_value,
                                   grad_loc, grad_attn);
  return (int)cudaGetLastError();
}
