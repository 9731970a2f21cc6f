// msda_bwd.cu — backward of multi-scale deformable attention for sm_100a.
//
// Replaces the reference's col2im kernels (detection/ops/src/cuda/ms_deform_im2col_cuda.cuh:
// 301-403 `…blocksize_aware_reduce_v1` used at D=32, :406-510 `…reduce_v2` used at D=64, and the
// dynamic/global variants :513-920) plus the launcher switch (:956-1327). Not a port:
//   * the reference launches one D-thread block per (b,q,m), stages 3*D partials in shared memory
//     per point and reduces them serially (v1) or with a barrier tree (v2); here a group of G lanes
//     owns a (b,q,m), each lane reduces its 4/8 channels in registers and the group finishes with
//     warp shuffles — no shared memory, no __syncthreads in the point loop;
//   * grad_value is scattered with one 16-byte vector reduction (REDG.E.ADD.F32x4) per lane per
//     corner instead of one scalar atomicAdd per channel per corner (:121,130,139,148);
//   * grad_sampling_loc / grad_attn_weight are written exactly once, so only grad_value is
//     zero-filled (the reference memsets all three, ms_deform_attn_cuda.cu:121-123).
// bf16 I/O accumulates grad_value in an fp32 scratch and converts once at the end.
#include "msda_common.cuh"

namespace msda {

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int s = G / 2; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s, G);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Vector kernel (same work decomposition as the forward).
// ---------------------------------------------------------------------------------------------
template <typename T, int G, int LT, int PT>
__global__ void __launch_bounds__(kThreads) msda_bwd_vec_kernel(const Params p) {
  using V = Vec<T>;
  constexpr int kCpl = V::kCpl;
  constexpr int kGpw = 32 / G;
  constexpr bool kStatic = (LT > 0);

  const int L = kStatic ? LT : p.L;
  const int P = kStatic ? PT : p.P;
  const int LP = L * P;
  const int MD = p.M * p.D;

  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)p.shapes[2 * threadIdx.x];
    sW[threadIdx.x] = (int)p.shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)p.lsi[threadIdx.x];
  }
  __syncthreads();

  const BlockCoord bc = block_coord(p);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / G, j = lane % G;

  const size_t img = (size_t)bc.b * p.S * MD + bc.m * p.D + j * kCpl;
  const T* __restrict__ vbase = reinterpret_cast<const T*>(p.value) + img;
  float* __restrict__ gvbase = reinterpret_cast<float*>(p.grad_value) + img;  // fp32 (scratch for bf16)
  const float* __restrict__ loc = reinterpret_cast<const float*>(p.loc);
  const float* __restrict__ aw = reinterpret_cast<const float*>(p.aw);
  const T* __restrict__ gout = reinterpret_cast<const T*>(p.grad_out);
  float* __restrict__ gloc = reinterpret_cast<float*>(p.grad_loc);
  float* __restrict__ gaw = reinterpret_cast<float*>(p.grad_aw);

  for (int qw = bc.q_begin + warp * kGpw; qw < bc.q_end; qw += kWarps * kGpw) {
    const int q = qw + grp;
    const bool active = q < bc.q_end;
    const size_t pair = ((size_t)bc.b * p.Lq + (active ? q : bc.q_begin)) * p.M + bc.m;
    const float* __restrict__ loc_pair = loc + pair * LP * 2;
    const float* __restrict__ aw_pair = aw + pair * LP;

    const V go = active ? V::load(gout + pair * p.D + j * kCpl) : V::zero();

#pragma unroll
    for (int r0 = 0; r0 < (kStatic ? LT * PT : LP); r0 += G) {
      // ---- producer ---------------------------------------------------------------------------
      const int pi = r0 + j;
      const bool mine = (pi < LP) && active;
      int off = 0, rowstride = 0;
      unsigned mask = 0u;
      float lh = 0.f, lw = 0.f, a = 0.f, fH = 0.f, fW = 0.f;
      if (mine) {
        const int l = pi / P;
        const int H = sH[l], W = sW[l];
        const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_pair) + pi);
        a = __ldg(aw_pair + pi);
        const PointGeom<float> g = point_geom<float>(xy.x, xy.y, H, W);
        lh = g.lh; lw = g.lw; mask = g.mask;
        off = (sStart[l] + g.h_low * W + g.w_low) * MD;
        rowstride = W * MD;
        fH = (float)H; fW = (float)W;
      }
      float my_ga = 0.f, my_gw = 0.f, my_gh = 0.f;
      // ---- consumers --------------------------------------------------------------------------
#pragma unroll
      for (int jj = 0; jj < G; ++jj) {
        if (r0 + jj < LP) {  // uniform
          const unsigned mk = __shfl_sync(0xffffffffu, mask, jj, G);
          const int o = __shfl_sync(0xffffffffu, off, jj, G);
          const int rs = __shfl_sync(0xffffffffu, rowstride, jj, G);
          const float flh = __shfl_sync(0xffffffffu, lh, jj, G);
          const float flw = __shfl_sync(0xffffffffu, lw, jj, G);
          const float fa = __shfl_sync(0xffffffffu, a, jj, G);
          float s_a = 0.f, s_w = 0.f, s_h = 0.f;
          if (mk != 0u) {
            const float hh = 1.f - flh, hw = 1.f - flw;
            const float w1 = hh * hw, w2 = hh * flw, w3 = flh * hw, w4 = flh * flw;
            const T* p1 = vbase + o;
            const V v1 = (mk & 1u) ? V::load(p1) : V::zero();
            const V v2 = (mk & 2u) ? V::load(p1 + MD) : V::zero();
            const V v3 = (mk & 4u) ? V::load(p1 + rs) : V::zero();
            const V v4 = (mk & 8u) ? V::load(p1 + rs + MD) : V::zero();
            float t[kCpl];
#pragma unroll
            for (int c = 0; c < kCpl; ++c) {
              const float g = go.v[c];
              const float val = w1 * v1.v[c] + w2 * v2.v[c] + w3 * v3.v[c] + w4 * v4.v[c];
              const float dw = hh * (v2.v[c] - v1.v[c]) + flh * (v4.v[c] - v3.v[c]);
              const float dh = hw * (v3.v[c] - v1.v[c]) + flw * (v4.v[c] - v2.v[c]);
              s_a = fmaf(g, val, s_a);
              s_w = fmaf(g, dw, s_w);
              s_h = fmaf(g, dh, s_h);
              t[c] = g * fa;  // top_grad_value
            }
            float* g1 = gvbase + o;
#pragma unroll
            for (int c0 = 0; c0 < kCpl; c0 += 4) {
              if (mk & 1u) red_add_v4(g1 + c0, w1 * t[c0], w1 * t[c0 + 1], w1 * t[c0 + 2], w1 * t[c0 + 3]);
              if (mk & 2u) red_add_v4(g1 + MD + c0, w2 * t[c0], w2 * t[c0 + 1], w2 * t[c0 + 2], w2 * t[c0 + 3]);
              if (mk & 4u) red_add_v4(g1 + rs + c0, w3 * t[c0], w3 * t[c0 + 1], w3 * t[c0 + 2], w3 * t[c0 + 3]);
              if (mk & 8u) red_add_v4(g1 + rs + MD + c0, w4 * t[c0], w4 * t[c0 + 1], w4 * t[c0 + 2], w4 * t[c0 + 3]);
            }
          }
          s_a = group_sum<G>(s_a);
          s_w = group_sum<G>(s_w);
          s_h = group_sum<G>(s_h);
          if (j == jj) { my_ga = s_a; my_gw = s_w; my_gh = s_h; }
        }
      }
      if (mine) {
        gaw[pair * LP + pi] = my_ga;
        reinterpret_cast<float2*>(gloc)[pair * LP + pi] = make_float2(fW * my_gw * a, fH * my_gh * a);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: any D / dtype / L,P. One warp per (b,q,m), lanes stride over channels, scalar
// atomics. AT = type of the grad_value accumulator (float scratch for bf16).
// ---------------------------------------------------------------------------------------------
template <typename T, typename F, typename AT>
__global__ void __launch_bounds__(kThreads) msda_bwd_generic_kernel(const Params p) {
  const int L = p.L, P = p.P, LP = L * P, MD = p.M * p.D;
  const BlockCoord bc = block_coord(p);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t img = (size_t)bc.b * p.S * MD + bc.m * p.D;
  const T* __restrict__ vb = reinterpret_cast<const T*>(p.value) + img;
  AT* __restrict__ gvb = reinterpret_cast<AT*>(p.grad_value) + img;
  const F* __restrict__ loc = reinterpret_cast<const F*>(p.loc);
  const F* __restrict__ aw = reinterpret_cast<const F*>(p.aw);
  const T* __restrict__ gout = reinterpret_cast<const T*>(p.grad_out);
  F* __restrict__ gloc = reinterpret_cast<F*>(p.grad_loc);
  F* __restrict__ gaw = reinterpret_cast<F*>(p.grad_aw);

  for (int q = bc.q_begin + warp; q < bc.q_end; q += kWarps) {
    const size_t pair = ((size_t)bc.b * p.Lq + q) * p.M + bc.m;
    for (int l = 0; l < L; ++l) {
      const int H = (int)p.shapes[2 * l], W = (int)p.shapes[2 * l + 1];
      const int start = (int)p.lsi[l];
      for (int k = 0; k < P; ++k) {
        const size_t pi = pair * LP + l * P + k;
        const F x = loc[2 * pi], y = loc[2 * pi + 1], a = aw[pi];
        const PointGeom<F> g = point_geom<F>(x, y, H, W);
        F s_a = 0, s_w = 0, s_h = 0;
        if (g.mask != 0u) {
          const F hh = 1 - g.lh, hw = 1 - g.lw;
          const F w1 = hh * hw, w2 = hh * g.lw, w3 = g.lh * hw, w4 = g.lh * g.lw;
          const size_t o = (size_t)(start + g.h_low * W + g.w_low) * MD;
          const size_t rs = (size_t)W * MD;
          for (int c = lane; c < p.D; c += 32) {
            const F go = (F)ld_scalar(gout + pair * p.D + c);
            const T* p1 = vb + o + c;
            AT* g1 = gvb + o + c;
            const F tg = go * a;
            F v1 = 0, v2 = 0, v3 = 0, v4 = 0;
            if (g.mask & 1u) { v1 = (F)ld_scalar(p1); atomicAdd(g1, (AT)(w1 * tg)); }
            if (g.mask & 2u) { v2 = (F)ld_scalar(p1 + MD); atomicAdd(g1 + MD, (AT)(w2 * tg)); }
            if (g.mask & 4u) { v3 = (F)ld_scalar(p1 + rs); atomicAdd(g1 + rs, (AT)(w3 * tg)); }
            if (g.mask & 8u) { v4 = (F)ld_scalar(p1 + rs + MD); atomicAdd(g1 + rs + MD, (AT)(w4 * tg)); }
            s_a += go * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
            s_w += go * (hh * (v2 - v1) + g.lh * (v4 - v3));
            s_h += go * (hw * (v3 - v1) + g.lw * (v4 - v2));
          }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          s_a += __shfl_xor_sync(0xffffffffu, s_a, s);
          s_w += __shfl_xor_sync(0xffffffffu, s_w, s);
          s_h += __shfl_xor_sync(0xffffffffu, s_h, s);
        }
        if (lane == 0) {
          gaw[pi] = s_a;
          gloc[2 * pi] = (F)W * s_w * a;
          gloc[2 * pi + 1] = (F)H * s_h * a;
        }
      }
    }
  }
}

// fp32 scratch -> bf16 grad_value (n is a multiple of 8 on the vector path; tail handled scalar).
__global__ void __launch_bounds__(kThreads) msda_cvt_f32_bf16_kernel(const float* __restrict__ src,
                                                                     __nv_bfloat16* __restrict__ dst,
                                                                     size_t n) {
  const size_t nvec = n / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = Vec<__nv_bfloat16>::pack2(a.x, a.y);
    o.y = Vec<__nv_bfloat16>::pack2(a.z, a.w);
    o.z = Vec<__nv_bfloat16>::pack2(b.x, b.y);
    o.w = Vec<__nv_bfloat16>::pack2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
template <typename T, int G>
static cudaError_t launch_vec_g(const Params& p, dim3 grid, cudaStream_t s) {
  if (p.L == 3 && p.P == 4) {
    msda_bwd_vec_kernel<T, G, 3, 4><<<grid, kThreads, 0, s>>>(p);
  } else if (p.L == 1 && p.P == 4) {
    msda_bwd_vec_kernel<T, G, 1, 4><<<grid, kThreads, 0, s>>>(p);
  } else {
    msda_bwd_vec_kernel<T, G, 0, 0><<<grid, kThreads, 0, s>>>(p);
  }
  return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_vec(const Params& p, int G, dim3 grid, cudaStream_t s) {
  switch (G) {
    case 2: return launch_vec_g<T, 2>(p, grid, s);
    case 4: return launch_vec_g<T, 4>(p, grid, s);
    case 8: return launch_vec_g<T, 8>(p, grid, s);
    case 16: return launch_vec_g<T, 16>(p, grid, s);
    case 32: return launch_vec_g<T, 32>(p, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

// `p.grad_value` must point at the ACCUMULATOR (T storage for f32/f64, fp32 scratch for bf16),
// already zero-filled on `s`.
cudaError_t launch_backward(const Params& p, int dtype, bool vec_ok, int G, cudaStream_t s) {
  const dim3 grid((unsigned)((size_t)p.N * p.nchunk * p.M));
  if (vec_ok) {
    if (dtype == MSDA_F32) return launch_vec<float>(p, G, grid, s);
    return launch_vec<__nv_bfloat16>(p, G, grid, s);
  }
  switch (dtype) {
    case MSDA_F32: msda_bwd_generic_kernel<float, float, float><<<grid, kThreads, 0, s>>>(p); break;
    case MSDA_BF16: msda_bwd_generic_kernel<__nv_bfloat16, float, float><<<grid, kThreads, 0, s>>>(p); break;
    case MSDA_F64: msda_bwd_generic_kernel<double, double, double><<<grid, kThreads, 0, s>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_cvt_f32_bf16(const float* src, void* dst, size_t n, cudaStream_t s) {
  size_t blocks = (n / 8 + kThreads - 1) / kThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  msda_cvt_f32_bf16_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return cudaGetLastError();
}

}  // namespace msda
