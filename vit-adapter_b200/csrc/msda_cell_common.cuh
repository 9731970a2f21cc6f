// msda_cell_common.cuh — device helpers shared by the two cell-ordered backward kernels (msda_bwd_cell.cu: cells bucketed
// per CTA in shared memory; msda_bwd_sorted.cu: cells sorted per (batch, head) slab in global memory): explicit
// shared-memory accesses, channel loads widened to fp32, integer shared-memory atomics, L2 prefetch.
#pragma once
#include "msda_common.cuh"

namespace msda {

constexpr int kScrStride = 33;  // uint4 slots per corner row of a warp's broadcast scratch: 32 points + 1 pad, so the four
                                // corner records of one point fall into four different 16-byte bank groups (one wavefront)
constexpr unsigned kScrBytesPerWarp = 4u * kScrStride * 16u;
constexpr unsigned kUbufBytesPerWarp = 32u * 16u;
constexpr unsigned kNoCell = 0xFFFFFFFFu;

// Cell word of a sample: bottom-right corner token (20 bits, always >= 0) | corner mask << 20 | level << 24. Two samples
// are in the same bilinear cell exactly when their cell words are equal (the mask separates the wrap-around aliases at
// the left / right border), and the four corner tokens follow from it with the level's width.

template <typename T>
__device__ __forceinline__ void unpack2(unsigned x, float& lo, float& hi);
template <>
__device__ __forceinline__ void unpack2<float>(unsigned, float&, float&) {}
template <>
__device__ __forceinline__ void unpack2<__nv_bfloat16>(unsigned x, float& lo, float& hi) {
  lo = __uint_as_float(x << 16);
  hi = __uint_as_float(x & 0xffff0000u);
}
template <>
__device__ __forceinline__ void unpack2<__half>(unsigned x, float& lo, float& hi) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&x));
  lo = f.x;
  hi = f.y;
}

// ---- explicit shared-memory accesses (32-bit shared addresses: no generic-to-shared conversion in the hot loop) -----
__device__ __forceinline__ uint4 lds128(unsigned a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint2 lds64(unsigned a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ unsigned lds32(unsigned a) {
  unsigned r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts128(unsigned a, unsigned x, unsigned y, unsigned z, unsigned w) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts32(unsigned a, unsigned x) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}

// CPL consecutive channels of T from shared memory -> fp32 registers
template <typename T, int CPL>
__device__ __forceinline__ void ch_load_shared(unsigned a, float (&v)[CPL]) {
  static_assert(CPL == 4 || CPL == 8, "4 or 8 channels per lane");
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int c = 0; c < CPL; c += 4) {
      const uint4 t = lds128(a + c * 4);
      v[c] = __uint_as_float(t.x); v[c + 1] = __uint_as_float(t.y); v[c + 2] = __uint_as_float(t.z); v[c + 3] = __uint_as_float(t.w);
    }
  } else if constexpr (CPL == 4) {
    const uint2 t = lds64(a);
    unpack2<T>(t.x, v[0], v[1]);
    unpack2<T>(t.y, v[2], v[3]);
  } else {
    const uint4 t = lds128(a);
    unpack2<T>(t.x, v[0], v[1]);
    unpack2<T>(t.y, v[2], v[3]);
    unpack2<T>(t.z, v[4], v[5]);
    unpack2<T>(t.w, v[6], v[7]);
  }
}

// CPL consecutive channels of T from global memory (read-only path) -> fp32 registers
template <typename T, int CPL>
__device__ __forceinline__ void ch_load_global(const char* p, float (&v)[CPL]) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int c = 0; c < CPL; c += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p) + c / 4);
      v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
    }
  } else if constexpr (CPL == 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    unpack2<T>(t.x, v[0], v[1]);
    unpack2<T>(t.y, v[2], v[3]);
  } else {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    unpack2<T>(t.x, v[0], v[1]);
    unpack2<T>(t.y, v[2], v[3]);
    unpack2<T>(t.z, v[4], v[5]);
    unpack2<T>(t.w, v[6], v[7]);
  }
}

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

__device__ __forceinline__ unsigned atoms_add(unsigned a, unsigned v) {
  unsigned r;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
  return r;
}
__device__ __forceinline__ void reds_add(unsigned a, unsigned v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_min(unsigned a, int v) { asm volatile("red.shared.min.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_max(unsigned a, int v) { asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

}  // namespace msda
