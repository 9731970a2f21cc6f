// adapter_colsum.cu — column sums of a [rows, C] matrix: the bias gradient of the adapter's Linears.
//
// Reference: every nn.Linear of MSDeformAttn (ms_deform_attn.py:57-60) and ConvFFN (adapter_modules.py:56,60) has a bias,
// so its backward reduces grad_output [B*tokens, C_out] over the rows. torch does that with its generic reduce kernel; on
// B200 at the adapter's shapes (86 016 x 768 bf16) that kernel runs ~10x off the HBM roofline and the eight bias
// gradients of one interaction cost 0.6 ms - as much as both deformable-attention backward kernels
// (profiles/r1_block_profile_*). Here: a thread owns 16 bytes of columns (8 bf16 / 4 fp32) and walks rows with four
// independent loads in flight, fp32 accumulation; the CTA's row slots are reduced in shared memory into one partial row
// per CTA, and a second kernel sums the <= 592 partial rows in a fixed order (deterministic, no atomics).
// Compulsory traffic: rows * C * e.
#include <cuda_fp16.h>

#include "msda_common.cuh"

namespace msda {

template <typename T> struct CsVec;
template <> struct CsVec<float> {
  static constexpr int kN = 4;
  static __device__ __forceinline__ void add(const float* p, float (&a)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    a[0] += q.x; a[1] += q.y; a[2] += q.z; a[3] += q.w;
  }
};
template <> struct CsVec<__nv_bfloat16> {
  static constexpr int kN = 8;
  static __device__ __forceinline__ void add(const __nv_bfloat16* p, float (&a)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const unsigned u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[2 * i] += __uint_as_float(u[i] << 16);
      a[2 * i + 1] += __uint_as_float(u[i] & 0xffff0000u);
    }
  }
};

template <> struct CsVec<__half> {
  static constexpr int kN = 8;
  static __device__ __forceinline__ void add(const __half* p, float (&a)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const unsigned u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[i]));
      a[2 * i] += f.x;
      a[2 * i + 1] += f.y;
    }
  }
};

// blockDim.x = cg * rs: thread -> (column group = tid % cg, row slot = tid / cg); partial [gridDim.x][C]
template <typename T>
__global__ void __launch_bounds__(256) adapter_colsum_kernel(const T* __restrict__ x, long long rows, int C, int cg, int rs,
                                                             float* __restrict__ partial) {
  constexpr int kN = CsVec<T>::kN;
  extern __shared__ __align__(16) float cs_red[];  // [rs][C]
  const int g = threadIdx.x % cg, slot = threadIdx.x / cg;
  float a0[kN], a1[kN], a2[kN], a3[kN];
#pragma unroll
  for (int i = 0; i < kN; ++i) { a0[i] = 0.f; a1[i] = 0.f; a2[i] = 0.f; a3[i] = 0.f; }
  const long long stride = (long long)gridDim.x * rs;
  const T* col = x + g * kN;
  long long r = (long long)blockIdx.x * rs + slot;
  for (; r + 3 * stride < rows; r += 4 * stride) {
    CsVec<T>::add(col + r * C, a0);
    CsVec<T>::add(col + (r + stride) * C, a1);
    CsVec<T>::add(col + (r + 2 * stride) * C, a2);
    CsVec<T>::add(col + (r + 3 * stride) * C, a3);
  }
  for (; r < rows; r += stride) CsVec<T>::add(col + r * C, a0);
#pragma unroll
  for (int i = 0; i < kN; ++i) cs_red[slot * C + g * kN + i] = (a0[i] + a1[i]) + (a2[i] + a3[i]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < rs; ++s) t += cs_red[s * C + c];
    partial[(size_t)blockIdx.x * C + c] = t;
  }
}

// second stage: 32 columns x 32 row slices per CTA, four independent partial sums per thread (latency, not bytes, is the cost)
__global__ void __launch_bounds__(1024) adapter_colsum_final_kernel(const float* __restrict__ partial, int prow, int C,
                                                                    float* __restrict__ out) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  if (c < C) {
    int r = threadIdx.y;
    for (; r + 96 < prow; r += 128) {
      t0 += partial[(size_t)r * C + c];
      t1 += partial[(size_t)(r + 32) * C + c];
      t2 += partial[(size_t)(r + 64) * C + c];
      t3 += partial[(size_t)(r + 96) * C + c];
    }
    for (; r < prow; r += 32) t0 += partial[(size_t)r * C + c];
  }
  red[threadIdx.y][threadIdx.x] = (t0 + t1) + (t2 + t3);
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int yy = 0; yy < 32; ++yy) t += red[yy][threadIdx.x];
    out[c] = t;
  }
}

static int cs_vec(int dtype) { return dtype == MSDA_F32 ? 4 : 8; }

bool colsum_supported(int dtype, int C) {
  if (dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F16) return false;
  const int v = cs_vec(dtype);
  return C > 0 && C % v == 0 && C / v <= 256 && (size_t)C * sizeof(float) * (256 / (C / v)) <= 48 * 1024;
}

static unsigned cs_grid(long long rows, int rs) {
  const long long need = (rows + rs - 1) / rs;
  return (unsigned)(need < 592 ? (need ? need : 1) : 592);  // 4 CTAs per SM: ~64 KB of loads in flight per SM
}

size_t colsum_workspace_bytes(int dtype, long long rows, int C) {
  if (!colsum_supported(dtype, C)) return 0;
  return (size_t)cs_grid(rows, 256 / (C / cs_vec(dtype))) * C * sizeof(float);
}

cudaError_t launch_colsum(int dtype, const void* x, long long rows, int C, float* out, float* partial, cudaStream_t s) {
  const int cg = C / cs_vec(dtype), rs = 256 / cg;
  const unsigned grid = cs_grid(rows, rs);
  const size_t smem = (size_t)rs * C * sizeof(float);
  if (dtype == MSDA_F32)
    adapter_colsum_kernel<float><<<grid, cg * rs, smem, s>>>(reinterpret_cast<const float*>(x), rows, C, cg, rs, partial);
  else if (dtype == MSDA_BF16)
    adapter_colsum_kernel<__nv_bfloat16><<<grid, cg * rs, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), rows, C, cg, rs, partial);
  else
    adapter_colsum_kernel<__half><<<grid, cg * rs, smem, s>>>(reinterpret_cast<const __half*>(x), rows, C, cg, rs, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  adapter_colsum_final_kernel<<<(C + 31) / 32, dim3(32, 32), 0, s>>>(partial, (int)grid, C, out);
  return cudaGetLastError();
}

}  // namespace msda
