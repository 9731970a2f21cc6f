// adapter_dwconv.cu — depth-wise 3x3 convolution of the adapter's ConvFFN, directly on the token layout.
//
// Reference: DWConv.forward, */mm*_custom/models/backbones/adapter_modules.py:73-87 —
// the [B, 21n, C] token sequence holds three maps (16n tokens at 2H x 2W, 4n at H x W, n at H/2 x W/2); the
// reference slices it, transposes each slice to NCHW (.contiguous()), runs nn.Conv2d(C, C, 3, 1, 1,
// groups=C) three times, transposes back and concatenates: >= 9 kernels and ~5x the compulsory traffic.
// Here one kernel reads the tokens where they lie (channels-last: the C channels of a token are contiguous,
// so a warp's 16-byte lanes are coalesced), applies the 3x3 taps with zero padding inside each map and
// writes the output tokens once. Compulsory traffic = read x + write y; the 9-tap reuse is served by L1/L2.
// The backward reuses the same gather with flipped taps for grad_x and reduces grad_weight / grad_bias in
// registers -> shared memory -> one fp32 atomicAdd per (CTA, channel, tap).
#include "msda_common.cuh"

namespace msda {

struct DwParams {
  const void* x;       // [B, Ntok, C]
  const void* w;       // [C, 1, 3, 3]
  const void* bias;    // [C] or nullptr
  void* y;             // [B, Ntok, C]
  int B, Ntok, C, H, W;  // H, W: the H/16 grid (middle map); maps are (2H,2W), (H,W), (H/2,W/2)
};

template <typename T> struct DwAcc { using type = float; };
template <> struct DwAcc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename DwAcc<T>::type dw_ld(const T* p) { return (typename DwAcc<T>::type)(*p); }
template <> __device__ __forceinline__ float dw_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T, typename A> __device__ __forceinline__ void dw_st(T* p, A v) { *p = (T)v; }
template <> __device__ __forceinline__ void dw_st<__nv_bfloat16, float>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// token index inside the sequence -> (map origin token, map height, map width, h, w)
__device__ __forceinline__ void dw_locate(int t, int H, int W, int& t0, int& mh, int& mw, int& h, int& w) {
  const int n0 = 4 * H * W, n1 = H * W;
  if (t < n0) { t0 = 0; mh = 2 * H; mw = 2 * W; }
  else if (t < n0 + n1) { t0 = n0; mh = H; mw = W; }
  else { t0 = n0 + n1; mh = H / 2; mw = W / 2; }
  const int r = t - t0;
  h = r / mw; w = r - h * mw;
}

// One thread = one token x VEC consecutive channels. FLIP = correlate with the 180-degree rotated taps
// (that is grad_x of a stride-1, pad-1 depth-wise convolution).
template <typename T, int VEC, bool FLIP>
__global__ void __launch_bounds__(256) adapter_dwconv_kernel(const DwParams p) {
  using A = typename DwAcc<T>::type;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  A* sw = reinterpret_cast<A*>(dw_smem);  // [9][C] taps, transposed so that a thread's VEC channels are contiguous
  for (int i = threadIdx.x; i < p.C * 9; i += blockDim.x) {
    const int c = i / 9, k = i - c * 9;
    sw[(FLIP ? 8 - k : k) * p.C + c] = dw_ld<T>(reinterpret_cast<const T*>(p.w) + i);
  }
  __syncthreads();
  const int cvec = p.C / VEC;
  const size_t total = (size_t)p.B * p.Ntok * cvec;
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  T* __restrict__ y = reinterpret_cast<T*>(p.y);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const size_t bt = i / cvec;
    const int t = (int)(bt % p.Ntok);
    const size_t b = bt / p.Ntok;
    int t0, mh, mw, h, w;
    dw_locate(t, p.H, p.W, t0, mh, mw, h, w);
    A acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = (!FLIP && p.bias) ? dw_ld<T>(reinterpret_cast<const T*>(p.bias) + cv * VEC + v) : (A)0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int hh = h + dy;
      if (hh < 0 || hh >= mh) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int ww = w + dx;
        if (ww < 0 || ww >= mw) continue;
        const T* src = x + ((b * p.Ntok + t0 + hh * mw + ww) * p.C + cv * VEC);
        const A* tap = sw + ((dy + 1) * 3 + (dx + 1)) * p.C + cv * VEC;
        if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(src));
          acc[0] = fmaf(q.x, tap[0], acc[0]); acc[1] = fmaf(q.y, tap[1], acc[1]);
          acc[2] = fmaf(q.z, tap[2], acc[2]); acc[3] = fmaf(q.w, tap[3], acc[3]);
        } else if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 2) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
          const unsigned u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            acc[2 * v] = fmaf(__uint_as_float(u[v] << 16), tap[2 * v], acc[2 * v]);
            acc[2 * v + 1] = fmaf(__uint_as_float(u[v] & 0xffff0000u), tap[2 * v + 1], acc[2 * v + 1]);
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += dw_ld<T>(src + v) * tap[v];
        }
      }
    }
    T* dst = y + ((b * p.Ntok + t) * p.C + cv * VEC);
    if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else if constexpr (sizeof(T) * VEC == 16 && sizeof(T) == 2) {
      uint4 o;
      o.x = Vec<__nv_bfloat16>::pack2(acc[0], acc[1]); o.y = Vec<__nv_bfloat16>::pack2(acc[2], acc[3]);
      o.z = Vec<__nv_bfloat16>::pack2(acc[4], acc[5]); o.w = Vec<__nv_bfloat16>::pack2(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) dw_st<T, A>(dst + v, acc[v]);
    }
  }
}

// grad_weight[c][k] = sum_{b,t} grad_y[b,t,c] * x[b, nbr_k(t), c] ; grad_bias[c] = sum grad_y[b,t,c].
// CTA = (a chunk of tokens) x (all channels): thread (tl, c) walks tokens tl, tl+TL, ... of the chunk for channel c,
// keeps 10 partial sums in registers, then the TL partials per channel are summed through shared memory and one
// atomicAdd per (channel, tap) goes out. blockDim = (CX, TL): CX channels handled per CTA pass (threadIdx.x fastest =>
// coalesced channel reads).
template <typename T>
__global__ void __launch_bounds__(256) adapter_dwconv_wgrad_kernel(const DwParams p, const void* grad_y_, double* gw64, float* gw32,
                                                                   double* gb64, float* gb32, int tokens_per_cta) {
  using A = typename DwAcc<T>::type;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  A* red = reinterpret_cast<A*>(dw_smem);  // [blockDim.y][blockDim.x][10]
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  const T* __restrict__ gy = reinterpret_cast<const T*>(grad_y_);
  const size_t bt_total = (size_t)p.B * p.Ntok;
  const size_t bt_begin = (size_t)blockIdx.x * tokens_per_cta;
  const size_t bt_end = bt_begin + tokens_per_cta < bt_total ? bt_begin + tokens_per_cta : bt_total;
  for (int cbase = blockIdx.y * blockDim.x; cbase < p.C; cbase += gridDim.y * blockDim.x) {  // uniform: barriers inside
    const int c = cbase + threadIdx.x;
    const bool live = c < p.C;
    A s[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) s[k] = (A)0;
    for (size_t bt = bt_begin + threadIdx.y; live && bt < bt_end; bt += blockDim.y) {
      const int t = (int)(bt % p.Ntok);
      const size_t b = bt / p.Ntok;
      int t0, mh, mw, h, w;
      dw_locate(t, p.H, p.W, t0, mh, mw, h, w);
      const A g = dw_ld<T>(gy + bt * p.C + c);
      s[9] += g;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int hh = h + dy;
        if (hh < 0 || hh >= mh) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int ww = w + dx;
          if (ww < 0 || ww >= mw) continue;
          s[(dy + 1) * 3 + (dx + 1)] += g * dw_ld<T>(x + ((b * p.Ntok + t0 + hh * mw + ww) * p.C + c));
        }
      }
    }
    A* mine = red + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 10;
#pragma unroll
    for (int k = 0; k < 10; ++k) mine[k] = s[k];
    __syncthreads();
    if (threadIdx.y == 0 && live) {
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        A tot = (A)0;
        for (int yy = 0; yy < (int)blockDim.y; ++yy) tot += red[((size_t)yy * blockDim.x + threadIdx.x) * 10 + k];
        if constexpr (sizeof(A) == 8) {
          if (k < 9) atomicAdd(gw64 + c * 9 + k, tot); else atomicAdd(gb64 + c, tot);
        } else {
          if (k < 9) atomicAdd(gw32 + c * 9 + k, tot); else atomicAdd(gb32 + c, tot);
        }
      }
    }
    __syncthreads();
  }
}

template <typename T, bool FLIP>
static cudaError_t launch_dw(const DwParams& p, cudaStream_t s) {
  using A = typename DwAcc<T>::type;
  const size_t smem = (size_t)p.C * 9 * sizeof(A);
  constexpr int kVec = 16 / (int)sizeof(T);
  const bool vec = sizeof(T) <= 4 && p.C % kVec == 0 && (reinterpret_cast<uintptr_t>(p.x) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(p.y) % 16 == 0);
  const size_t work = (size_t)p.B * p.Ntok * (vec ? p.C / kVec : p.C);
  size_t blocks = (work + 255) / 256;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  if (blocks < 1) blocks = 1;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  if (vec) {
    adapter_dwconv_kernel<T, kVec, FLIP><<<(unsigned)blocks, 256, smem, s>>>(p);
  } else {
    adapter_dwconv_kernel<T, 1, FLIP><<<(unsigned)blocks, 256, smem, s>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_dwconv(const DwParams& p, int dtype, bool flip, cudaStream_t s) {
  switch (dtype) {
    case MSDA_F32: return flip ? launch_dw<float, true>(p, s) : launch_dw<float, false>(p, s);
    case MSDA_BF16: return flip ? launch_dw<__nv_bfloat16, true>(p, s) : launch_dw<__nv_bfloat16, false>(p, s);
    case MSDA_F64: return flip ? launch_dw<double, true>(p, s) : launch_dw<double, false>(p, s);
    default: return cudaErrorInvalidValue;
  }
}

// grad_weight / grad_bias accumulators: fp32 for f32/bf16 inputs, fp64 for f64 (both [C*9] / [C], zero-filled here).
cudaError_t launch_dwconv_wgrad(const DwParams& p, int dtype, const void* grad_y, void* gw, void* gb, cudaStream_t s) {
  const size_t asz = dtype == MSDA_F64 ? 8 : 4;
  cudaError_t e = cudaMemsetAsync(gw, 0, (size_t)p.C * 9 * asz, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(gb, 0, (size_t)p.C * asz, s);
  if (e != cudaSuccess) return e;
  const dim3 block(32, 8);
  const size_t bt_total = (size_t)p.B * p.Ntok;
  int ctas_x = 148 * 2;
  int tokens_per_cta = (int)((bt_total + ctas_x - 1) / ctas_x);
  if (tokens_per_cta < 64) tokens_per_cta = 64;
  ctas_x = (int)((bt_total + tokens_per_cta - 1) / tokens_per_cta);
  const int cgroups = (p.C + 31) / 32;
  const dim3 grid((unsigned)ctas_x, (unsigned)(cgroups < 8 ? cgroups : 8));
  const size_t smem = (size_t)block.x * block.y * 10 * asz;
  switch (dtype) {
    case MSDA_F32:
      adapter_dwconv_wgrad_kernel<float><<<grid, block, smem, s>>>(p, grad_y, nullptr, (float*)gw, nullptr, (float*)gb, tokens_per_cta);
      break;
    case MSDA_BF16:
      adapter_dwconv_wgrad_kernel<__nv_bfloat16><<<grid, block, smem, s>>>(p, grad_y, nullptr, (float*)gw, nullptr, (float*)gb, tokens_per_cta);
      break;
    case MSDA_F64:
      adapter_dwconv_wgrad_kernel<double><<<grid, block, smem, s>>>(p, grad_y, (double*)gw, nullptr, (double*)gb, nullptr, tokens_per_cta);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace msda
