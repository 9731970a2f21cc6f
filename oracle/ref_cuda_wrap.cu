// ref_cuda_wrap.cu — C-ABI wrapper around the REFERENCE's own CUDA kernels, for tests and the bench.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md). This file contains no reference code: it
// #includes the reference's kernel header from where it lies under /root/reference
// (detection/ops/src/cuda/ms_deform_im2col_cuda.cuh, passed with -I by oracle/Makefile) and calls its
// two launchers, ms_deformable_im2col_cuda (:923-954) and ms_deformable_col2im_cuda (:956-1327),
// exactly as the reference host code does (ms_deform_attn_cuda.cu:61-75, :131-148), including the
// im2col_step batching loop and the zero-filled outputs (:54, :121-123). The result is
// oracle/_ref/libmsda_refcuda.so: "the reference CUDA kernel, recompiled for sm_100a" — the GPU
// oracle and the kernel to beat.
#include <cstdint>
#include <algorithm>

#include "ms_deform_im2col_cuda.cuh"

template <typename T>
static int ref_forward(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc, const T* aw,
                       int N, int S, int M, int D, int L, int Lq, int P, int im2col_step, T* out,
                       cudaStream_t stream) {
  const int step = std::min(N, im2col_step);
  if (N % step != 0) return -7;
  cudaMemsetAsync(out, 0, sizeof(T) * (size_t)N * Lq * M * D, stream);  // at::zeros, .cu:54
  const size_t per_value = (size_t)S * M * D, per_loc = (size_t)Lq * M * L * P * 2, per_aw = (size_t)Lq * M * L * P;
  for (int n = 0; n < N / step; ++n) {
    ms_deformable_im2col_cuda<T>(stream, value + n * step * per_value, shapes, lsi, loc + n * step * per_loc,
                                 aw + n * step * per_aw, step, S, M, D, L, Lq, P,
                                 out + (size_t)n * step * Lq * M * D);
  }
  return (int)cudaGetLastError();
}

template <typename T>
static int ref_backward(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc, const T* aw,
                        const T* grad_out, int N, int S, int M, int D, int L, int Lq, int P, int im2col_step,
                        T* grad_value, T* grad_loc, T* grad_aw, cudaStream_t stream) {
  const int step = std::min(N, im2col_step);
  if (N % step != 0) return -7;
  const size_t per_value = (size_t)S * M * D, per_loc = (size_t)Lq * M * L * P * 2, per_aw = (size_t)Lq * M * L * P;
  cudaMemsetAsync(grad_value, 0, sizeof(T) * N * per_value, stream);  // zeros_like x3, .cu:121-123
  cudaMemsetAsync(grad_loc, 0, sizeof(T) * N * per_loc, stream);
  cudaMemsetAsync(grad_aw, 0, sizeof(T) * N * per_aw, stream);
  for (int n = 0; n < N / step; ++n) {
    ms_deformable_col2im_cuda<T>(stream, grad_out + (size_t)n * step * Lq * M * D, value + n * step * per_value,
                                 shapes, lsi, loc + n * step * per_loc, aw + n * step * per_aw, step, S, M, D, L,
                                 Lq, P, grad_value + n * step * per_value, grad_loc + n * step * per_loc,
                                 grad_aw + n * step * per_aw);
  }
  return (int)cudaGetLastError();
}

extern "C" {
int refcuda_forward_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                        const float* aw, int N, int S, int M, int D, int L, int Lq, int P, int step, float* out,
                        void* stream) {
  return ref_forward<float>(value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, step, out, (cudaStream_t)stream);
}
int refcuda_forward_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                        const double* aw, int N, int S, int M, int D, int L, int Lq, int P, int step, double* out,
                        void* stream) {
  return ref_forward<double>(value, shapes, lsi, loc, aw, N, S, M, D, L, Lq, P, step, out, (cudaStream_t)stream);
}
int refcuda_backward_f32(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                         const float* aw, const float* grad_out, int N, int S, int M, int D, int L, int Lq, int P,
                         int step, float* grad_value, float* grad_loc, float* grad_aw, void* stream) {
  return ref_backward<float>(value, shapes, lsi, loc, aw, grad_out, N, S, M, D, L, Lq, P, step, grad_value,
                             grad_loc, grad_aw, (cudaStream_t)stream);
}
int refcuda_backward_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                         const double* aw, const double* grad_out, int N, int S, int M, int D, int L, int Lq,
                         int P, int step, double* grad_value, double* grad_loc, double* grad_aw, void* stream) {
  return ref_backward<double>(value, shapes, lsi, loc, aw, grad_out, N, S, M, D, L, Lq, P, step, grad_value,
                              grad_loc, grad_aw, (cudaStream_t)stream);
}
}
