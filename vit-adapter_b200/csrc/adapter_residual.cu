// adapter_residual.cu — the residual epilogues of the adapter's Extractor.
//
// Reference: `query = query + attn` and `query = query + self.drop_path(self.ffn(...))` (adapter_modules.py:113-116). Under
// AMP the stream `query` is fp32 and the branch (a Linear's output) is bf16 (or fp16); torch's mixed-dtype add falls off its
// vectorised path (measured: 216 us for 86 016 x 768 on B200, 0.47 of the HBM roofline). This is the same add with
// 16-byte loads on both operands: out[i] = res[i] + (float)branch[i].
// Compulsory traffic: n * (4 + e_branch + 4).
#include <cuda_fp16.h>

#include <type_traits>

#include "msda_common.cuh"

namespace msda {

// thread = 8 consecutive elements, two of them in flight
template <typename TB>
__global__ void __launch_bounds__(256) adapter_residual_add_kernel(const float* __restrict__ res, const TB* __restrict__ branch,
                                                                   float* __restrict__ out, long long n8) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(res) + 2 * i);
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(res) + 2 * i + 1);
    float b[8];
    if constexpr (sizeof(TB) == 2) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(branch) + i);
      const unsigned u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (std::is_same<TB, __half>::value) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
          b[2 * k] = f.x; b[2 * k + 1] = f.y;
        } else {
          b[2 * k] = __uint_as_float(u[k] << 16); b[2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
        }
      }
    } else {
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(branch) + 2 * i);
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(branch) + 2 * i + 1);
      b[0] = q0.x; b[1] = q0.y; b[2] = q0.z; b[3] = q0.w; b[4] = q1.x; b[5] = q1.y; b[6] = q1.z; b[7] = q1.w;
    }
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(a0.x + b[0], a0.y + b[1], a0.z + b[2], a0.w + b[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(a1.x + b[4], a1.y + b[5], a1.z + b[6], a1.w + b[7]);
  }
}

cudaError_t launch_residual_add(int branch_dtype, const float* res, const void* branch, float* out, long long n, cudaStream_t s) {
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (branch_dtype == MSDA_BF16)
    adapter_residual_add_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(res, reinterpret_cast<const __nv_bfloat16*>(branch), out, n8);
  else if (branch_dtype == MSDA_F16)
    adapter_residual_add_kernel<__half><<<(unsigned)blocks, 256, 0, s>>>(res, reinterpret_cast<const __half*>(branch), out, n8);
  else if (branch_dtype == MSDA_F32)
    adapter_residual_add_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(res, reinterpret_cast<const float*>(branch), out, n8);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace msda
