// msda_common.cuh — shared device helpers for the sm_100a multi-scale deformable attention kernels.
//
// Arithmetic contract (what the reference computes):
//   reference forward  : detection/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299 (+ bilinear :33-84)
//   reference backward : detection/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-510 (+ bilinear :87-159)
// The pixel coordinate is ONE fused multiply-add, fmaf(loc, (float)H, -0.5f): that is how nvcc
// compiles the reference's `loc_h * spatial_h - 0.5` (SASS: FFMA R, (float)H, loc, -0.5), and
// bit-exact index parity with the reference kernel depends on it.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msda_b200.h"

namespace msda {

constexpr int kThreads = 256;            // threads per CTA for every kernel in this library
constexpr int kWarps = kThreads / 32;
constexpr int kMaxLevels = MSDA_MAX_LEVELS;

// Kernel parameter block (passed by value; lives in constant bank).
struct Params {
  const void* value;
  const int64_t* shapes;  // [L,2] (H,W) on device
  const int64_t* lsi;     // [L]       on device
  const void* loc;
  const void* aw;
  void* out;               // fwd: out ; bwd: unused
  const void* grad_out;    // bwd
  void* grad_value;        // bwd: dtype T, or fp32 scratch when T = bf16
  void* grad_loc;          // bwd
  void* grad_aw;           // bwd
  int N, S, M, D, L, Lq, P;
  int qc;                  // queries per CTA chunk
  int nchunk;              // ceil(Lq / qc)
};

// ---------------------------------------------------------------------------------------------
// Geometry of one sampling point. Mirrors ms_deform_im2col_cuda.cuh:285-288 (coordinate + bounds
// test) and :38-78 (floor, fractions, per-corner validity).
// mask bit0: corner1 (h_low ,w_low ) readable     bit1: corner2 (h_low ,w_high)
//      bit2: corner3 (h_high,w_low )              bit3: corner4 (h_high,w_high)
// mask == 0 when the sample fails the bounds test (the reference skips it entirely).
// ---------------------------------------------------------------------------------------------
template <typename F>
struct PointGeom {
  int h_low, w_low;
  F lh, lw;
  unsigned mask;
};

__device__ __forceinline__ float msda_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double msda_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float msda_floor(float a) { return floorf(a); }
__device__ __forceinline__ double msda_floor(double a) { return floor(a); }

template <typename F>
__device__ __forceinline__ PointGeom<F> point_geom(F loc_w, F loc_h, int H, int W) {
  PointGeom<F> g;
  const F h_im = msda_fma(loc_h, (F)H, (F)-0.5);
  const F w_im = msda_fma(loc_w, (F)W, (F)-0.5);
  const bool inside = (h_im > (F)-1) && (w_im > (F)-1) && (h_im < (F)H) && (w_im < (F)W);
  const F hf = msda_floor(h_im);
  const F wf = msda_floor(w_im);
  g.h_low = inside ? (int)hf : 0;
  g.w_low = inside ? (int)wf : 0;
  g.lh = inside ? (h_im - hf) : (F)0;
  g.lw = inside ? (w_im - wf) : (F)0;
  unsigned m = 0;
  if (inside) {
    const bool t = g.h_low >= 0, b = g.h_low + 1 <= H - 1;
    const bool l = g.w_low >= 0, r = g.w_low + 1 <= W - 1;
    m = (unsigned)(t && l) | ((unsigned)(t && r) << 1) | ((unsigned)(b && l) << 2) |
        ((unsigned)(b && r) << 3);
  }
  g.mask = m;
  return g;
}

// ---------------------------------------------------------------------------------------------
// A sampling point prepared for the vector kernels: the four corner rows are CLAMPED into the level
// so they can be gathered without predicates; corners the reference does not read get weight 0
// (forward) / mask bit 0 (backward), which reproduces its zero padding exactly for finite inputs.
//   offf    byte offset of the clamped (h_low, w_low) row inside the (b, m) slab, 16-byte aligned,
//           with flags in the low bits: bit0 = the right-hand corner is a different token,
//           bit1 = the lower corners are a different row, bit2 = the sample passed the bounds test
//   rowstep bytes between rows of this level (W * M * D * sizeof(T))
// ---------------------------------------------------------------------------------------------
struct PointTap {
  unsigned offf, rowstep;
  float w[4];
};

__device__ __forceinline__ unsigned tap_offset(const PointGeom<float>& g, int H, int W, int start,
                                               unsigned MDb) {
  const int hl = max(g.h_low, 0), hh = min(g.h_low + 1, H - 1);
  const int wl = max(g.w_low, 0), wh = min(g.w_low + 1, W - 1);
  return (unsigned)(start + hl * W + wl) * MDb | (unsigned)(wh != wl) | ((unsigned)(hh != hl) << 1) |
         ((unsigned)(g.mask != 0u) << 2);
}

__device__ __forceinline__ PointTap point_tap(float x, float y, float a, int H, int W, int start,
                                              unsigned MDb) {
  const PointGeom<float> g = point_geom<float>(x, y, H, W);
  const float hh = 1.f - g.lh, hw = 1.f - g.lw;
  PointTap t;
  t.w[0] = (g.mask & 1u) ? (hh * hw) * a : 0.f;
  t.w[1] = (g.mask & 2u) ? (hh * g.lw) * a : 0.f;
  t.w[2] = (g.mask & 4u) ? (g.lh * hw) * a : 0.f;
  t.w[3] = (g.mask & 8u) ? (g.lh * g.lw) * a : 0.f;
  t.offf = tap_offset(g, H, W, start, MDb);
  t.rowstep = (unsigned)W * MDb;
  return t;
}

// ---------------------------------------------------------------------------------------------
// 16-byte channel vectors. One lane owns kCpl consecutive channels of a head:
//   float          : 4 channels  (LDG.E.128)
//   __nv_bfloat16  : 8 channels  (LDG.E.128), widened to fp32 in registers
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Vec;

template <>
struct Vec<float> {
  static constexpr int kCpl = 4;
  float v[4];
  __device__ __forceinline__ static Vec load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Vec r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  }
  __device__ __forceinline__ static Vec zero() {
    Vec r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.f;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Vec<__nv_bfloat16> {
  static constexpr int kCpl = 8;
  float v[8];
  __device__ __forceinline__ static Vec load(const __nv_bfloat16* p) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    Vec r;
    r.v[0] = __uint_as_float(t.x << 16); r.v[1] = __uint_as_float(t.x & 0xffff0000u);
    r.v[2] = __uint_as_float(t.y << 16); r.v[3] = __uint_as_float(t.y & 0xffff0000u);
    r.v[4] = __uint_as_float(t.z << 16); r.v[5] = __uint_as_float(t.z & 0xffff0000u);
    r.v[6] = __uint_as_float(t.w << 16); r.v[7] = __uint_as_float(t.w & 0xffff0000u);
    return r;
  }
  __device__ __forceinline__ static Vec zero() {
    Vec r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
    return r;
  }
  __device__ __forceinline__ static unsigned pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
    return *reinterpret_cast<const unsigned*>(&h);
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 t;
    t.x = pack2(v[0], v[1]); t.y = pack2(v[2], v[3]);
    t.z = pack2(v[4], v[5]); t.w = pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// Vector reduction into global memory: one REDG.E.ADD.F32x4 per 16 bytes (sm_90+ PTX
// `red.global.add.v4.f32`) instead of the reference's one scalar atomicAdd per channel per corner
// (ms_deform_im2col_cuda.cuh:121,130,139,148).
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// scalar load / convert helpers for the generic (any-D) kernels
__device__ __forceinline__ float ld_scalar(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_scalar(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
__device__ __forceinline__ double ld_scalar(const double* p) { return __ldg(p); }

__device__ __forceinline__ void st_scalar(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_scalar(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st_scalar(double* p, double v) { *p = v; }

// Block -> (batch, query chunk, head) decomposition shared by forward and backward.
// Heads vary fastest so concurrently running CTAs cover the same spatial neighbourhood
// (their sampling_loc / attn_weight sectors and value rows are shared in L2).
struct BlockCoord {
  int b, m, q_begin, q_end;
};
__device__ __forceinline__ BlockCoord block_coord(const Params& p) {
  BlockCoord c;
  const int bid = blockIdx.x;
  c.m = bid % p.M;
  const int t = bid / p.M;
  const int chunk = t % p.nchunk;
  c.b = t / p.nchunk;
  c.q_begin = chunk * p.qc;
  c.q_end = min(p.Lq, c.q_begin + p.qc);
  return c;
}

}  // namespace msda
