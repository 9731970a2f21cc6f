// msda_fwd_smem.cu — forward with whole-level value maps staged in shared memory (sm_100a).
//
// OPT-IN (tuning key "fwd_smem" = 2); needs the level shapes on the host (msda_forward_ex).
// Why: the L1 serves this op's gather pattern (8 lanes x 16 B per 128-byte row, rows at unrelated
// addresses) at ~0.49 rows per SM-cycle (142 rows/ns chip-wide; the texture path is no better), and
// the L1-path forward (msda_fwd.cu) sits on that ceiling; from shared memory the same pattern costs one
// LSU wavefront (~1 cycle) per row.
// Measured outcome, round 1 (profiles/r1_fwd_smem_vs_l1.md): with the gathers at 1 cycle per row the 5 broadcast
// shuffles per point (also LSU wavefronts) become 1/4 of the pipe load, and the kernel ends up at ~79 % of
// the same LSU data pipe: ViT-Adapter-B bs16 Extractor 115.6 vs 121.8 us fp32, 77.8 vs 97.3 us bf16;
// Injector (levels 1+2 staged, level 0 still through L1) 98-115 vs 97 us. Round 2 hands the taps over through
// shared-memory records (one LDS.128 + one LDS.32 per point) instead of shuffles; see profiles/r2_fwd_smem_*.
//
// Design. A CTA owns (batch b, head m, a long chunk of queries). For every level that fits
// (plan made on the host, plan_forward_smem in msda_abi.cu) it copies that head's [H*W, D] map into
// shared memory ONCE with 16-byte cp.async (LDGSTS; one 1-D TMA bulk copy per 128-byte row was
// measured 3x slower — the bulk engine wants bigger boxes than a strided token row), into a map
// that is ZERO-PADDED by one token on every side. The halo makes the op's zero-padding rule
// (ms_deform_im2col_cuda.cuh:56-78: corners outside the level read as 0) a property of the data:
// no corner masks, no clamping, no predicates — a point is just base + 4 LDS.128 at fixed deltas
// (+rowB, +rowstride, +both) and 16 FFMAs. Samples that fail the bounds test (:288) point at an
// all-zero block with zero weights. Levels that do not fit keep the L1 path of msda_fwd.cu
// (clamped rows, packed flags) inside the same kernel.
// Results are bit-identical to msda_fwd.cu (tests/test_op_gpu.py::test_smem_forward_bit_identical):
// a corner the L1 path multiplies by a zeroed weight is here a zero value times the weight.
#include "msda_common.cuh"

namespace msda {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// 16-byte asynchronous copy global -> shared (SASS: LDGSTS.E.BYPASS.128), L1-bypassing.
__device__ __forceinline__ void cp_async_16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <typename T>
__device__ __forceinline__ Vec<T> lds_vec(unsigned addr);
template <>
__device__ __forceinline__ Vec<float> lds_vec<float>(unsigned addr) {
  Vec<float> r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr));
  return r;
}
template <>
__device__ __forceinline__ Vec<__nv_bfloat16> lds_vec<__nv_bfloat16>(unsigned addr) {
  uint4 t;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr));
  Vec<__nv_bfloat16> r;
  r.v[0] = __uint_as_float(t.x << 16); r.v[1] = __uint_as_float(t.x & 0xffff0000u);
  r.v[2] = __uint_as_float(t.y << 16); r.v[3] = __uint_as_float(t.y & 0xffff0000u);
  r.v[4] = __uint_as_float(t.z << 16); r.v[5] = __uint_as_float(t.z & 0xffff0000u);
  r.v[6] = __uint_as_float(t.w << 16); r.v[7] = __uint_as_float(t.w & 0xffff0000u);
  return r;
}

// T, G, LT, PT as in msda_fwd.cu; NT = threads per CTA (one CTA per SM: the maps use most of the smem).
// Shared-memory layout (bytes): [0, plan.null_bytes) all zeros; then per staged level l at plan.smem_off[l]
// a (H+2) x (W+2) grid of kRowB-byte rows, interior = the level's tokens, border = zeros.
template <typename T, int G, int LT, int PT, int NT>
__global__ void __launch_bounds__(NT, 1) msda_fwd_smem_kernel(const Params p, const SmemPlan plan) {
  using V = Vec<T>;
  constexpr int kCpl = V::kCpl;
  constexpr int kGpw = 32 / G;
  constexpr int kWarpsNT = NT / 32;
  constexpr int LP = LT * PT;
  constexpr unsigned kRowB = (unsigned)(G * 16);  // bytes of one head-row (D * sizeof(T))

  extern __shared__ __align__(128) unsigned char smem[];
  // per-warp tap scratch (as in msda_fwd.cu): 32-byte records (4 weights | offset), rows padded by 16 bytes
  constexpr int kTapRow = 2 * G + 1;
  __shared__ uint4 s_tap[kWarpsNT * kGpw * kTapRow];

  const int MD = p.M * p.D;
  const unsigned MDb = (unsigned)MD * (unsigned)sizeof(T);
  const BlockCoord bc = block_coord(p);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / G, j = lane % G;
  const unsigned tap_row = smem_u32(s_tap) + (unsigned)((warp * kGpw + grp) * kTapRow) * 16u;
  const unsigned tap_mine = tap_row + (unsigned)j * 32u;

  const char* __restrict__ slab = reinterpret_cast<const char*>(p.value) +
                                  ((size_t)bc.b * p.S * MD + (size_t)bc.m * p.D) * sizeof(T);

  // ---- zero everything (null block + halos), then stage the interiors with cp.async -----------------
  for (unsigned o = threadIdx.x * 16u; o < plan.total_bytes; o += NT * 16u)
    *reinterpret_cast<uint4*>(smem + o) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
#pragma unroll
  for (int l = 0; l < LT; ++l) {
    if (plan.staged & (1u << l)) {
      const int W = plan.W[l], rows = plan.H[l] * W;
      const char* src = slab + (size_t)plan.start[l] * MDb + (threadIdx.x % G) * 16;
      const unsigned dst = smem_u32(smem) + plan.smem_off[l] + (threadIdx.x % G) * 16;
      for (int r = threadIdx.x / G; r < rows; r += NT / G) {
        const int h = r / W, w = r - h * W;
        cp_async_16(dst + (unsigned)((h + 1) * (W + 2) + (w + 1)) * kRowB, src + (size_t)r * MDb);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  unsigned rsl[LT];  // bytes between rows: padded smem map (staged) or global (not staged)
#pragma unroll
  for (int l = 0; l < LT; ++l)
    rsl[l] = (plan.staged & (1u << l)) ? (unsigned)(plan.W[l] + 2) * kRowB : (unsigned)plan.W[l] * MDb;

  const char* __restrict__ vb = slab + j * 16;
  const unsigned sb = smem_u32(smem) + j * 16;
  const float* __restrict__ loc = reinterpret_cast<const float*>(p.loc);
  const float* __restrict__ aw = reinterpret_cast<const float*>(p.aw);
  T* __restrict__ out = reinterpret_cast<T*>(p.out);

  // software pipeline: fetch the next iteration's locations / weights before working on this one
  constexpr int kRounds = (LP + G - 1) / G;
  float2 nxy[kRounds];
  float na[kRounds];
  auto fetch = [&](int qw_, float2 (&xy_)[kRounds], float (&a_)[kRounds]) {
    const int q_ = qw_ + grp;
    const bool act_ = q_ < bc.q_end;
    const size_t pair_ = ((size_t)bc.b * p.Lq + (act_ ? q_ : bc.q_begin)) * p.M + bc.m;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int pi_ = r * G + j;
      xy_[r] = make_float2(0.f, 0.f);
      a_[r] = 0.f;
      if (pi_ < LP && act_) {
        xy_[r] = __ldg(reinterpret_cast<const float2*>(loc + pair_ * LP * 2) + pi_);
        a_[r] = __ldg(aw + pair_ * LP + pi_);
      }
    }
  };
  fetch(bc.q_begin + warp * kGpw, nxy, na);

  // every warp waits for its own copies, then for everyone's (warps with no query in a clipped last
  // chunk must still reach the barrier, so it sits outside the query loop)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  for (int qw = bc.q_begin + warp * kGpw; qw < bc.q_end; qw += kWarpsNT * kGpw) {
    const int q = qw + grp;
    const bool active = q < bc.q_end;
    const size_t pair = ((size_t)bc.b * p.Lq + (active ? q : bc.q_begin)) * p.M + bc.m;

    float2 cxy[kRounds];
    float ca[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) { cxy[r] = nxy[r]; ca[r] = na[r]; }
    if (qw + kWarpsNT * kGpw < bc.q_end) fetch(qw + kWarpsNT * kGpw, nxy, na);

    V acc = V::zero();
#pragma unroll
    for (int r0 = 0; r0 < LP; r0 += G) {
      // ---- producer ------------------------------------------------------------------------------------
      const int pi = r0 + j;
      unsigned offf = 0u;  // staged: smem byte offset of the (h_low, w_low) row of the padded map (null block if skipped)
                           // not staged: packed clamped global offset | flags, as in msda_fwd.cu
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (pi < LP && active) {
        const int l = pi / PT;
        const int H = plan.H[l], W = plan.W[l];
        const float2 xy = cxy[r0 / G];
        const float a = ca[r0 / G];
        if (plan.staged & (1u << l)) {
          const float h_im = fmaf(xy.y, (float)H, -0.5f), w_im = fmaf(xy.x, (float)W, -0.5f);
          if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
            const float hf = floorf(h_im), wf = floorf(w_im);
            const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
            w0 = (hh * hw) * a; w1 = (hh * lw) * a; w2 = (lh * hw) * a; w3 = (lh * lw) * a;
            offf = plan.smem_off[l] + (unsigned)(((int)hf + 1) * (W + 2) + ((int)wf + 1)) * kRowB;
          }
        } else {
          const PointGeom<float> g = point_geom<float>(xy.x, xy.y, H, W);
          const float hh = 1.f - g.lh, hw = 1.f - g.lw;
          w0 = (g.mask & 1u) ? (hh * hw) * a : 0.f;
          w1 = (g.mask & 2u) ? (hh * g.lw) * a : 0.f;
          w2 = (g.mask & 4u) ? (g.lh * hw) * a : 0.f;
          w3 = (g.mask & 8u) ? (g.lh * g.lw) * a : 0.f;
          offf = tap_offset(g, H, W, plan.start[l], MDb);
        }
      }
      // ---- hand-over through the warp's scratch (one LDS.128 + one LDS.32 per point instead of 5 shuffles) ------
      __syncwarp();
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(tap_mine), "f"(w0), "f"(w1), "f"(w2), "f"(w3) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(tap_mine + 16u), "r"(offf) : "memory");
      __syncwarp();
      // ---- consumers -----------------------------------------------------------------------------------
#pragma unroll
      for (int jj = 0; jj < G; ++jj) {
        if (r0 + jj < LP) {
          const int l = (r0 + jj) / PT;  // compile-time after unrolling
          float a1, a2, a3, a4;
          unsigned of;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4) : "r"(tap_row + jj * 32u));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(of) : "r"(tap_row + jj * 32u + 16u));
          if (plan.staged & (1u << l)) {  // uniform
            const unsigned s1 = sb + of, s3 = s1 + rsl[l];
            const V v1 = lds_vec<T>(s1);
            const V v2 = lds_vec<T>(s1 + kRowB);
            const V v3 = lds_vec<T>(s3);
            const V v4 = lds_vec<T>(s3 + kRowB);
#pragma unroll
            for (int c = 0; c < kCpl; ++c) {
              acc.v[c] = fmaf(a1, v1.v[c], acc.v[c]);
              acc.v[c] = fmaf(a2, v2.v[c], acc.v[c]);
              acc.v[c] = fmaf(a3, v3.v[c], acc.v[c]);
              acc.v[c] = fmaf(a4, v4.v[c], acc.v[c]);
            }
          } else {
            const char* p1 = ptr_add(vb, of & ~15u);
            const char* p2 = ptr_madd(p1, of & 1u, MDb);
            const char* p3 = ptr_madd(p1, (of >> 1) & 1u, rsl[l]);
            const char* p4 = ptr_madd(p3, of & 1u, MDb);
            const V v1 = V::load(reinterpret_cast<const T*>(p1));
            const V v2 = V::load(reinterpret_cast<const T*>(p2));
            const V v3 = V::load(reinterpret_cast<const T*>(p3));
            const V v4 = V::load(reinterpret_cast<const T*>(p4));
#pragma unroll
            for (int c0 = 0; c0 < kCpl; c0 += 4)
              fma4x4_if(of & 4u, &acc.v[c0], a1, a2, a3, a4, &v1.v[c0], &v2.v[c0], &v3.v[c0], &v4.v[c0]);
          }
        }
      }
    }
    if (active) acc.store(out + pair * p.D + j * kCpl);
  }
}

// ---------------------------------------------------------------------------------------------
// Launcher. Returns cudaErrorNotSupported when no instantiation exists (caller falls back to the
// L1-path kernel).
// ---------------------------------------------------------------------------------------------
template <typename T, int G, int LT, int PT, int NT>
static cudaError_t launch_one(const Params& p, const SmemPlan& plan, dim3 grid, cudaStream_t s) {
  auto kern = msda_fwd_smem_kernel<T, G, LT, PT, NT>;
  // static (tap scratch) + dynamic (level maps) shared memory share the 227 KB of a CTA: ask for what this plan needs
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total_bytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, NT, plan.total_bytes, s>>>(p, plan);
  return cudaGetLastError();
}

template <typename T, int G>
static cudaError_t launch_g(const Params& p, const SmemPlan& plan, int nt, dim3 grid, cudaStream_t s) {
  if (p.L == 3 && p.P == 4) {
    return nt == 1024 ? launch_one<T, G, 3, 4, 1024>(p, plan, grid, s) : launch_one<T, G, 3, 4, 512>(p, plan, grid, s);
  }
  if (p.L == 1 && p.P == 4) {
    return nt == 1024 ? launch_one<T, G, 1, 4, 1024>(p, plan, grid, s) : launch_one<T, G, 1, 4, 512>(p, plan, grid, s);
  }
  return cudaErrorNotSupported;
}

cudaError_t launch_forward_smem(const Params& p, const SmemPlan& plan, int dtype, int G, int nt, cudaStream_t s) {
  const dim3 grid((unsigned)((size_t)p.N * p.nchunk * p.M));
  if (dtype == MSDA_F32) {
    if (G == 8) return launch_g<float, 8>(p, plan, nt, grid, s);
    if (G == 16) return launch_g<float, 16>(p, plan, nt, grid, s);
  } else if (dtype == MSDA_BF16) {
    if (G == 4) return launch_g<__nv_bfloat16, 4>(p, plan, nt, grid, s);
    if (G == 8) return launch_g<__nv_bfloat16, 8>(p, plan, nt, grid, s);
  }
  return cudaErrorNotSupported;
}

}  // namespace msda
