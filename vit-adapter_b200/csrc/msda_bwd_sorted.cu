// msda_bwd_sorted.cu — slab-sorted backward of multi-scale deformable attention for sm_100a.
//
// Replaces the reference's col2im kernels (detection/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-510, bilinear backward
// :87-159). The reference - and msda_bwd.cu, the query-order path of this library - issue four row-wide atomics into
// grad_value per sampled point: 16.5 M 128-byte RED rows for the ViT-Adapter-B Extractor at bs 16, which is exactly what the
// L2's atomic units serve in 320 us. But grad_value of one (batch, head) SLAB has only S rows (1 024 for
// that call), so ~84 contributions go to every row: the scatter is a segmented reduction in disguise.
//
// Pass 1 (msda_sort_*_kernel): every slab's sampled points are SORTED BY THE BILINEAR CELL they fall in - a counting sort whose
// histograms live in shared memory (integer ATOMS only); one 4-byte sample index per point is written in cell order.
// Pass 2 (msda_bwd_sorted_kernel): a group of D/4 lanes walks a contiguous range of the sorted list. A lane owns ONE corner
// of the bilinear cell and 16 channels (four 4-channel vectors, interleaved with the other lanes of the corner so that every
// vector access of a corner covers whole 32-byte sectors):
//   * the lane's 16 channels of the current cell's value row and of the partial grad_value row stay in REGISTERS for the
//     whole run of samples of that cell (16 + 16 floats);
//   * per sample: ONE read of the query's grad_out row (the query-order kernel gathers four value rows and scatters four),
//     staged through shared memory with cp.async one batch of 32 samples ahead; 8 packed FMAs (FFMA2) for the lane's share
//     of u_k = <grad_out, v_k> and 8 for the partial row; the corner weight comes as one 8-byte broadcast from the warp's
//     scratch, prepared with one sample per lane (index load, gather of location / weight, geometry, corner weights);
//   * u_k needs one add across the 2 (D = 32) or 4 (D = 64) lanes of the corner; the lane that prepared the sample turns
//     (u_1..u_4) into grad_attn_weight and grad_sampling_loc (the linear-in-the-corners algebra of msda_bwd.cu);
//   * when the cell changes, the partial row leaves with four REDG.E.ADD.F32x4 per lane: row atomics per sample drop
//     from 4 to 4 / (samples per cell run) - ~20 samples per cell at the ViT-Adapter-B Extractor.
// (An earlier lane layout - every lane 4 channels of all four corners - needed a transposing shuffle reduction of four
// values over the whole group per sample, 19 of its 47 instructions per step, and four predicated row flushes / loads where a
// cell starts: 108 M instead of 83 M instructions at ViT-Adapter-B bs 16. Same time in fp32, 5-10 % slower in bf16 / D = 64.)
// Nothing depends on where the reference points are: the sort is by the actual sampling location, so every input is
// handled; an adversarial input (all points in one cell) degenerates to long runs, still correct.
// The summation ORDER differs from the query-order kernel (and is not deterministic: positions inside a cell run come from
// an atomic cursor), within the same tolerance class as any atomic scatter.
#include "msda_cell_common.cuh"

namespace msda {

// Workspace carved by the host (launch_backward_sorted below); all pointers are device pointers into it.
struct SortedPlan {
  unsigned* cnt;  // [N*M][parts][S] histograms of clamped top-left tokens -> per key, the exclusive prefix over the parts
  unsigned* tot;  // [N*M][S]   per key, the samples of all parts
  unsigned* nin;  // [N*M]      samples of the slab that pass the bounds test
  unsigned* idx;  // [N*M][cap] the slab's samples in cell order: query * L*P + point
  int cap;        // Lq * L * P: samples per slab
  int ppw;        // sorted positions per warp of the walker (a multiple of 32)
  int ctas_per_slab;
  int parts;      // the sort cuts a slab into this many runs of consecutive queries
  int qpp;        // queries per part
};

#ifndef MSDA_TU
#define MSDA_TU 0
#endif

#if MSDA_TU == 0  // the sort passes do not depend on the value dtype: compiled once
// ---- pass 1: counting sort of every slab's samples by bilinear cell ---------------------------------------------------------
// Sort key = clamped top-left token of the cell (< S). Border cells that clamp to the same token share a key; the walker
// compares cell WORDS, so they only shorten runs.
// A slab is cut into sp.parts PARTS of consecutive queries so that the grid covers every SM several times over (192 slabs on
// 148 SMs would otherwise run as two waves). Three kernels:
//   hist    CTA (slab, part): histogram of the part's samples in shared memory (integer ATOMS), stored to cnt[slab][part][S]
//   prefix  CTA (slab, 32 keys): per key, the exclusive prefix of cnt over the parts (in place) and the total -> tot[slab][S]
//   scatter CTA (slab, part): scans tot[slab] over the keys (12.5 KB at ViT-Adapter-L, read from L2; every CTA of the slab
//           repeats it - cheaper than a kernel of its own) and adds its row of cnt: the cursors of (part, key); every sample
//           takes its position with one ATOMS and writes its 4-byte index there (a first version wrote 20-byte records -
//           fractions, weight, cell word - to those random positions: 75 us instead of 23 us at ViT-Adapter-B bs 16; the
//           walker now gathers location and weight by index and redoes the geometry)
constexpr int kSortThreads = 256;
constexpr int kSortUnroll = 4;

template <int LT, int PT, bool SCATTER>
__global__ void __launch_bounds__(kSortThreads) msda_sort_part_kernel(const Params p, const SortedPlan sp) {
  constexpr bool kStatic = (LT > 0);
  const int L = kStatic ? LT : p.L;
  const int P = kStatic ? PT : p.P;
  const int LP = L * P;
  extern __shared__ unsigned s_cnt[];
  __shared__ int s_H[kMaxLevels], s_W[kMaxLevels], s_start[kMaxLevels];
  const int tid = threadIdx.x;
  const int slab = blockIdx.x / sp.parts, part = blockIdx.x - slab * sp.parts;
  const int b = slab / p.M, m = slab - b * p.M;
  const int S = p.S;
  unsigned* __restrict__ gcnt = sp.cnt + ((size_t)slab * sp.parts + part) * S;
  if (tid < L) {
    s_H[tid] = (int)p.shapes[2 * tid];
    s_W[tid] = (int)p.shapes[2 * tid + 1];
    s_start[tid] = (int)p.lsi[tid];
  }
  if constexpr (SCATTER) {
    // cursors: first position of (key, this part) = samples of the slab's smaller keys + samples of this key in earlier parts.
    // Thread t scans keys [t * run, (t + 1) * run) of tot[slab] straight from global memory (the second read hits L1).
    __shared__ unsigned s_warp[kSortThreads / 32];
    const unsigned* __restrict__ tot = sp.tot + (size_t)slab * S;
    const int run = (S + kSortThreads - 1) / kSortThreads;
    const int k0 = min(S, tid * run), k1 = min(S, k0 + run);
    unsigned sum = 0;
    for (int i = k0; i < k1; ++i) sum += tot[i];
    unsigned incl = sum;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = incl - sum;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) before += w < warp ? s_warp[w] : 0u;
    if (part == 0 && tid == kSortThreads - 1) sp.nin[slab] = before + sum;
    for (int i = k0; i < k1; ++i) {
      s_cnt[i] = before + gcnt[i];
      before += tot[i];
    }
  } else {
    for (int i = tid; i < S; i += kSortThreads) s_cnt[i] = 0u;
  }
  __syncthreads();
  const size_t pair0 = ((size_t)b * p.Lq * p.M + (size_t)m) * LP;  // (b, query 0, m, point 0)
  const unsigned qstride = (unsigned)p.M * (unsigned)LP;
  const float2* __restrict__ loc = reinterpret_cast<const float2*>(p.loc) + pair0;
  const unsigned s_base = (unsigned)__cvta_generic_to_shared(s_cnt);
  const unsigned i_begin = (unsigned)part * (unsigned)sp.qpp * (unsigned)LP;
  const unsigned i_end = min((unsigned)sp.cap, i_begin + (unsigned)sp.qpp * (unsigned)LP);
  unsigned* __restrict__ sidx = sp.idx + (size_t)slab * sp.cap;

  for (unsigned i0 = i_begin + tid; i0 < i_end; i0 += kSortThreads * kSortUnroll) {
    float2 xy[kSortUnroll];
    unsigned o[kSortUnroll];
#pragma unroll
    for (int k = 0; k < kSortUnroll; ++k) {
      const unsigned i = i0 + k * kSortThreads;
      const unsigned q = i / (unsigned)LP;
      o[k] = q * qstride + (i - q * (unsigned)LP);
      xy[k] = i < i_end ? __ldg(loc + o[k]) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kSortUnroll; ++k) {
      const unsigned i = i0 + k * kSortThreads;
      if (i >= i_end) break;
      const int l = (int)(i % (unsigned)LP) / P;
      const int W = s_W[l];
      const PointGeom<float> g = point_geom<float>(xy[k].x, xy[k].y, s_H[l], W);
      if (g.mask != 0u) {
        const unsigned ka = s_base + 4u * (unsigned)(s_start[l] + max(g.h_low, 0) * W + max(g.w_low, 0));
        if constexpr (SCATTER) {
          const unsigned pos = atoms_add(ka, 1u);
          sidx[pos] = i;
        } else {
          reds_add(ka, 1u);
        }
      } else if constexpr (!SCATTER) {
        // the reference skips the sample (ms_deform_im2col_cuda.cuh:365): its gradients are the zeros of the memset there
        (reinterpret_cast<float*>(p.grad_aw) + pair0)[o[k]] = 0.f;
        (reinterpret_cast<float2*>(p.grad_loc) + pair0)[o[k]] = make_float2(0.f, 0.f);
      }
    }
  }
  if constexpr (!SCATTER) {
    __syncthreads();
    for (int i = tid; i < S; i += kSortThreads) gcnt[i] = s_cnt[i];
  }
}

// Per key, the exclusive prefix of cnt[slab][part][key] over the parts (in place) and the total of the key -> tot[slab][key].
// A CTA of 256 threads serves (slab, 256 / groups consecutive keys); a thread owns one key and a group of up to 8 consecutive
// parts, so that a warp access covers 32 consecutive keys of one part and a thread's loads all go out at once; groups =
// 1 / 2 / 4 / 8 for up to 8 / 16 / 32 / 64 parts (ViT-Adapter-B bs 16: 4 parts, 256 keys per CTA, 768 CTAs; L bs 1: 37 parts,
// 32 keys per CTA). (The first version scanned in (key, part) order with two kernels over (slab, 256 keys) whose threads
// walked the parts eight at a time: 9 + 17 us at ViT-Adapter-L bs 1 - 37 parts of 3 136 keys per slab -, against 8.6 us for
// the histogram itself.)
constexpr int kPrefixThreads = 256, kPrefixPer = 8, kMaxParts = 64;
static int prefix_groups(int parts) { return parts <= 8 ? 1 : parts <= 16 ? 2 : parts <= 32 ? 4 : 8; }

__global__ void __launch_bounds__(kPrefixThreads) msda_sort_prefix_kernel(const SortedPlan sp, int S, int groups) {
  __shared__ unsigned s_g[kPrefixThreads];          // [groups][keys per CTA]
  const int slab = blockIdx.y, parts = sp.parts;
  const int kpc = kPrefixThreads / groups;          // keys per CTA: a multiple of 32, so a warp has one group
  const int grp = (int)threadIdx.x / kpc, kl = (int)threadIdx.x - grp * kpc;
  const int key = blockIdx.x * kpc + kl;
  const int per = (parts + groups - 1) / groups;    // parts per group, <= 8
  const int p0 = grp * per;
  unsigned* __restrict__ c = sp.cnt + ((size_t)slab * parts + p0) * S + key;
  unsigned x[kPrefixPer];
#pragma unroll
  for (int k = 0; k < kPrefixPer; ++k) x[k] = (key < S && k < per && p0 + k < parts) ? c[(size_t)k * S] : 0u;
  unsigned sum = 0;
#pragma unroll
  for (int k = 0; k < kPrefixPer; ++k) {
    const unsigned t = x[k];
    x[k] = sum;
    sum += t;
  }
  unsigned before = 0, total = sum;
  if (groups > 1) {                                  // uniform
    s_g[threadIdx.x] = sum;
    __syncthreads();
    total = 0;
    for (int g = 0; g < groups; ++g) {
      const unsigned t = s_g[g * kpc + kl];
      before += g < grp ? t : 0u;
      total += t;
    }
  }
  if (key < S) {
#pragma unroll
    for (int k = 0; k < kPrefixPer; ++k)
      if (k < per && p0 + k < parts) c[(size_t)k * S] = before + x[k];
    if (grp == 0) sp.tot[(size_t)slab * S + key] = total;
  }
}
#endif  // MSDA_TU == 0

// ---- pass 2: walk the sorted list ---------------------------------------------------------------------------------------------
// T = value dtype, G = lanes per group (D = 4 * G), LT/PT = compile-time levels / points (0,0 = runtime).
//
// A warp owns sp.ppw consecutive sorted positions, split into 32 / G contiguous group ranges. A batch is G steps; in step t
// every group works on position t of its batch. Lane (group, j) PREPARES position j of its group's batch (one point per
// lane: record load, corner weights, row offsets, start-of-cell flag) into the warp's scratch, and FINISHES it after the
// dot products of the batch are reduced. The grad_out rows of a batch are copied into shared memory with cp.async one batch
// ahead (a first version read them from global memory at the point of use: 45 % of all stall samples sat on that load).
// Scratch per warp (32 points; group rows padded so that the groups' broadcasts hit different banks):
//   wq  2 x float4  corner weight * attention weight, each weight twice (the operand pairs of the packed FMAs); 0 for a
//                   corner the reference does not read
//   xq  uint4       byte offset of each corner's row inside the (b, m) value slab, + 2;  1 = the reference does not read it
//   uq  float4      the reduced (u_1..u_4) of the point, written by the consumer for the finishing lane
//   gbuf            the grad_out rows of the batch (kBufs batches deep), [group][point][D], each group's rows skewed by kSkew
// (the start-of-cell flags of a batch travel as one ballot register, not through the scratch)

// Packed fp32 pairs (FFMA2 on 64-bit register pairs; the sm_100 intrinsic __ffma2_rn): per lane and sample the walker needs
// 16 FMAs for its share of the dot product and 16 for its share of the partial row; as pairs these are 8 + 8 instructions.
// A vector of 4 channels is two pairs.
struct Row4 {
  float2 lo, hi;
};
template <typename T>
__device__ __forceinline__ Row4 row_from_global(const char* p) {
  Row4 r;
  if constexpr (sizeof(T) == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.lo = make_float2(t.x, t.y);
    r.hi = make_float2(t.z, t.w);
  } else {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    unpack2<T>(t.x, r.lo.x, r.lo.y);
    unpack2<T>(t.y, r.hi.x, r.hi.y);
  }
  return r;
}
template <typename T>
__device__ __forceinline__ Row4 row_from_shared(unsigned a) {
  Row4 r;
  if constexpr (sizeof(T) == 4) {
    const uint4 t = lds128(a);
    r.lo = make_float2(__uint_as_float(t.x), __uint_as_float(t.y));
    r.hi = make_float2(__uint_as_float(t.z), __uint_as_float(t.w));
  } else {
    const uint2 t = lds64(a);
    unpack2<T>(t.x, r.lo.x, r.lo.y);
    unpack2<T>(t.y, r.hi.x, r.hi.y);
  }
  return r;
}

// one predicated vector reduction (a branch around it costs 4 more instructions on the start-of-cell path, four times)
__device__ __forceinline__ void red_add_v4_if(unsigned on, float* p, float a, float b, float c, float d) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "f"(a), "f"(b),
               "f"(c), "f"(d), "r"(on)
               : "memory");
}

#ifndef WALK_THREADS
#define WALK_THREADS 256
#endif
constexpr int kWalkThreads = WALK_THREADS;  // threads per walker CTA
constexpr int kWalkWarps = kWalkThreads / 32;

template <typename T, int G>
struct WalkScratch {
  static constexpr int kGroups = 32 / G;
  static constexpr int kStride = G + 1;                           // records per group row (one pad)
  static constexpr unsigned kRowB = 4u * G * sizeof(T);           // one grad_out row
#ifndef WALK_BUFS
#define WALK_BUFS 1
#endif
#ifndef WALK_MINB
#define WALK_MINB 3
#endif
  static constexpr int kBufs = kRowB <= 128 ? WALK_BUFS : 1;      // batches of grad_out rows in flight
  static constexpr unsigned kW = 0;                                // 2 x float4 [kGroups][kStride]
  static constexpr unsigned kX = kW + kGroups * kStride * 32u;    // uint4
  static constexpr unsigned kU = kX + kGroups * kStride * 16u;    // float4
  static constexpr unsigned kG = kU + kGroups * kStride * 16u;    // grad_out rows: [kBufs][group][G rows + skew]
  static constexpr unsigned kSkew = (G / 4) * 4u * sizeof(T);  // bytes one vector covers across the lanes of a corner: shifts each
                                                               // group's rows so that the groups' reads hit different banks
  static constexpr unsigned kGroupB = G * kRowB + kSkew;  // group g's rows start at g * kGroupB, i.e. skewed by g * kSkew
  static constexpr unsigned kBytes = kG + kBufs * kGroups * kGroupB;
};

template <typename T, int G, int LT, int PT, int MINB>
__global__ void __launch_bounds__(kWalkThreads, MINB) msda_bwd_sorted_kernel(const Params p, const SortedPlan sp) {
  constexpr bool kStatic = (LT > 0);
  constexpr int D = 4 * G;
  constexpr int NG = 32 / G;                         // groups per warp
  constexpr unsigned kRowB = D * sizeof(T);          // bytes of one head row of value / grad_out
  constexpr int kAccShift = sizeof(T) == 2 ? 1 : 0;  // fp32 accumulator rows are 4 / sizeof(T) times as long
  using SC = WalkScratch<T, G>;
  constexpr int kBufs = SC::kBufs;

  const int L = kStatic ? LT : p.L;
  const int P = kStatic ? PT : p.P;
  const int LP = L * P;
  const unsigned MDb = (unsigned)p.M * kRowB;  // bytes between neighbouring tokens of value / neighbouring queries of grad_out

  __shared__ int s_lvH[kMaxLevels], s_lvW[kMaxLevels], s_lvStart[kMaxLevels];
  extern __shared__ __align__(16) char s_scr[];  // kWalkWarps * SC::kBytes

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slab = blockIdx.x / sp.ctas_per_slab, chunk = blockIdx.x - slab * sp.ctas_per_slab;
  const int b = slab / p.M, m = slab - b * p.M;
  if (tid < L) {
    s_lvH[tid] = (int)p.shapes[2 * tid];
    s_lvW[tid] = (int)p.shapes[2 * tid + 1];
    s_lvStart[tid] = (int)p.lsi[tid];
  }
  __syncthreads();  // the only block-wide barrier
  const int n_in = (int)sp.nin[slab];
  const int r0 = (chunk * kWalkWarps + warp) * sp.ppw;
  if (r0 >= n_in) return;
  const int rend = min(n_in, r0 + sp.ppw);
  const int gi = lane / G, j = lane % G;
  const int glen = (rend - r0 + NG - 1) / NG;          // positions per group (the last group may get fewer, even none)
  const int g0 = r0 + gi * glen, gend = min(rend, g0 + glen);
  const int nb = (glen + G - 1) / G;                    // batches: the same count for every group of the warp

  const unsigned* __restrict__ sidx = sp.idx + (size_t)slab * sp.cap;

  const unsigned scr = (unsigned)__cvta_generic_to_shared(s_scr) + warp * SC::kBytes;
  const unsigned my = (unsigned)(gi * SC::kStride + j);     // this lane's prepared point
  const unsigned grp = (unsigned)(gi * SC::kStride);        // step 0 of this lane's group
  const size_t slab_v = ((size_t)b * p.S * p.M + (size_t)m) * D;
  const size_t pair0 = ((size_t)b * p.Lq * p.M + (size_t)m);  // (b, query 0, m)
  const unsigned qstride = (unsigned)p.M * (unsigned)LP;      // points between consecutive queries of this head
  // Consumer lane mapping: lane j of a group owns ONE corner (kc) and a quarter / half of the channels as four 4-channel
  // vectors: vector c covers channels 4 * (c * NH + hq) ..+4, so that the NH lanes of a corner cover whole 32-byte sectors
  // with every vector access. The dot product <grad_out, v_kc> then needs one add across NH lanes (NH = 2 / 4) instead of a
  // transposing reduction of four values over all G lanes, and the start-of-cell path handles one row per lane with
  // immediate offsets instead of four rows with four predicates and addresses.
  constexpr int NH = G / 4;
  const int kc = j / NH, hq = j % NH;
  constexpr unsigned kVecT = 4u * sizeof(T);          // one vector of value / grad_out
  constexpr unsigned kVecStepT = NH * kVecT;          // between a lane's consecutive vectors
  // value / accumulator bases are biased by the "+ 2" of the offset words
  const char* __restrict__ vb2 = reinterpret_cast<const char*>(p.value) + slab_v * sizeof(T) + hq * kVecT - 2;
  char* __restrict__ gvb2 = reinterpret_cast<char*>(p.grad_value) + slab_v * 4u + hq * 16 - (2 << kAccShift);  // fp32 accumulator
  const char* __restrict__ gob_row = reinterpret_cast<const char*>(p.grad_out) + pair0 * kRowB;
  float* __restrict__ gaw = reinterpret_cast<float*>(p.grad_aw) + pair0 * LP;
  float2* __restrict__ gloc = reinterpret_cast<float2*>(p.grad_loc) + pair0 * LP;
  const float* __restrict__ aw = reinterpret_cast<const float*>(p.aw) + pair0 * LP;
  const float2* __restrict__ loc = reinterpret_cast<const float2*>(p.loc) + pair0 * LP;

  unsigned cur = 1u;  // offset word of the current cell's row of this lane's corner (1 = the reference does not read it)
  Row4 v[4], acc[4];  // this lane's four vectors of that value row and of the partial grad_value row
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    v[c].lo = v[c].hi = make_float2(0.f, 0.f);
    acc[c].lo = acc[c].hi = make_float2(0.f, 0.f);
  }
  auto flush_row = [&]() {
    if (cur > 1u) {
      float* dst = reinterpret_cast<float*>(gvb2 + ((size_t)cur << kAccShift));
#pragma unroll
      for (int c = 0; c < 4; ++c) red_add_v4(dst + c * (NH * 4), acc[c].lo.x, acc[c].lo.y, acc[c].hi.x, acc[c].hi.y);
    }
  };
  unsigned lastcw = 0xFFFFFFFEu;  // cell word of the group's previous position

  // The grad_out rows of a batch go to shared memory: the lanes of a group copy row after row, 16 bytes per lane
  // (consecutive lanes -> consecutive addresses on both sides); `pidx` is the point index of THIS lane's point, whose row
  // offset the other lanes receive by shuffle.
  auto stage_rows = [&](unsigned pidx, int buf) {
    constexpr int kLanesPerRow = kRowB / 16;          // 8 (fp32, D = 32), 4 (16-bit, D = 32), 16 / 8 (D = 64)
    constexpr int kRowsPerIter = G / kLanesPerRow;    // 1 (fp32) or 2 (16-bit)
    const unsigned my_off = (pidx / (unsigned)LP) * MDb;
    const unsigned dst0 = scr + SC::kG + (unsigned)(buf * NG + gi) * SC::kGroupB + (unsigned)(j % kLanesPerRow) * 16u;
    const char* src0 = gob_row + (j % kLanesPerRow) * 16;
#pragma unroll
    for (int t = 0; t < G; t += kRowsPerIter) {
      const int row = t + j / kLanesPerRow;
      const unsigned off = __shfl_sync(0xffffffffu, my_off, row, G);
      cp_async16(dst0 + (unsigned)row * kRowB, ptr_add(src0, off));
    }
  };

  // Software pipeline over batches, one sample per lane: the sorted INDEX is read two batches ahead, location and
  // attention weight are gathered by it one batch ahead (the sort stores only the index), the grad_out rows are staged one
  // batch ahead.
  constexpr unsigned kNoIdx = 0xFFFFFFFFu;
  auto point_offset = [&](unsigned pidx) {  // sample index (query * LP + point) -> offset in the [Lq, M, LP] arrays of this head
    const unsigned ql = pidx / (unsigned)LP;
    return ql * qstride + (pidx - ql * (unsigned)LP);
  };
  unsigned idxC = kNoIdx, idxN = kNoIdx;
  float2 xyC = make_float2(0.f, 0.f), xyN = make_float2(0.f, 0.f);
  float aC = 0.f, aN = 0.f;
  if (g0 + j < gend) idxC = __ldg(sidx + g0 + j);
  if (g0 + G + j < gend) idxN = __ldg(sidx + g0 + G + j);
  if (idxC != kNoIdx) { const unsigned o = point_offset(idxC); xyC = __ldg(loc + o); aC = __ldg(aw + o); }
  if constexpr (kBufs == 2) {
    stage_rows(idxC == kNoIdx ? 0u : idxC, 0);  // (a lane without a sample names the row of query 0: finite data, weight 0)
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  for (int bi = 0; bi < nb; ++bi) {
    const int pos = g0 + bi * G + j;
    const bool valid = pos < gend;
    unsigned idxNN = kNoIdx;
    if (pos + 2 * G < gend) idxNN = __ldg(sidx + pos + 2 * G);
    if (idxN != kNoIdx) { const unsigned o = point_offset(idxN); xyN = __ldg(loc + o); aN = __ldg(aw + o); }
    if constexpr (kBufs == 2) {
      stage_rows(idxN == kNoIdx ? 0u : idxN, (bi + 1) & 1);  // rows of the NEXT batch (none: a harmless copy nobody reads)
    } else {
      __syncwarp();                     // the previous batch's readers are done with the only buffer
      stage_rows(idxC == kNoIdx ? 0u : idxC, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // the geometry of this lane's sample, redone from its location exactly as the sort and the forward do it
    unsigned cwC = kNoCell, lvl = 0u;
    float lh = 0.f, lw = 0.f;
    if (valid) {
      lvl = (idxC % (unsigned)LP) / (unsigned)P;
      const int W = s_lvW[lvl];
      const PointGeom<float> gm = point_geom<float>(xyC.x, xyC.y, s_lvH[lvl], W);
      lh = gm.lh;
      lw = gm.lw;
      cwC = (unsigned)(s_lvStart[lvl] + (gm.h_low + 1) * W + (gm.w_low + 1)) | (gm.mask << 20) | (lvl << 24);
    }
    unsigned heads = 0u;
    // ---- prepare: one point per lane ----------------------------------------------------------------------------------------------
    {
      unsigned prev = __shfl_up_sync(0xffffffffu, cwC, 1, G);
      if (j == 0) prev = lastcw;
      lastcw = __shfl_sync(0xffffffffu, cwC, G - 1, G);
      const bool head = valid && cwC != prev;
      unsigned xw[4] = {1u, 1u, 1u, 1u};
      float cf[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const unsigned W = (unsigned)s_lvW[(cwC >> 24) & 15u];
        const unsigned br = cwC & 0xFFFFFu;
        xw[0] = (cwC & (1u << 20)) ? (br - W - 1u) * MDb + 2u : 1u;
        xw[1] = (cwC & (2u << 20)) ? (br - W) * MDb + 2u : 1u;
        xw[2] = (cwC & (4u << 20)) ? (br - 1u) * MDb + 2u : 1u;
        xw[3] = (cwC & (8u << 20)) ? br * MDb + 2u : 1u;
        const float a = aC;
        const float hh = 1.f - lh, hw = 1.f - lw;
        cf[0] = xw[0] > 1u ? (hh * hw) * a : 0.f;
        cf[1] = xw[1] > 1u ? (hh * lw) * a : 0.f;
        cf[2] = xw[2] > 1u ? (lh * hw) * a : 0.f;
        cf[3] = xw[3] > 1u ? (lh * lw) * a : 0.f;
      }
      if constexpr (kBufs == 2) __syncwarp();  // the previous batch's readers are done with the scratch
      sts128(scr + SC::kW + my * 32u, __float_as_uint(cf[0]), __float_as_uint(cf[0]), __float_as_uint(cf[1]), __float_as_uint(cf[1]));
      sts128(scr + SC::kW + my * 32u + 16u, __float_as_uint(cf[2]), __float_as_uint(cf[2]), __float_as_uint(cf[3]), __float_as_uint(cf[3]));
      sts128(scr + SC::kX + my * 16u, xw[0], xw[1], xw[2], xw[3]);
      // start-of-cell flags of the whole batch as one register: bit (lane) = that lane's sample starts a cell
      heads = __ballot_sync(0xffffffffu, head);
      if constexpr (kBufs == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");  // this batch's rows have landed
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    // ---- consume: G steps; in step t every group works on point t of its batch ---------------------------------------------------
    const unsigned gbuf = scr + SC::kG + (unsigned)((kBufs == 2 ? (bi & 1) : 0) * NG + gi) * SC::kGroupB + hq * kVecT;
#pragma unroll 2
    for (int t = 0; t < G; ++t) {
      const unsigned st = grp + (unsigned)t;
      if ((heads >> (gi * G + t)) & 1u) {  // group-uniform: a new cell starts here - flush the partial row, fetch the new cell's row
        flush_row();
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c].lo = acc[c].hi = make_float2(0.f, 0.f);
        cur = lds32(scr + SC::kX + st * 16u + 4u * kc);
        if (cur > 1u) {
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = row_from_global<T>(vb2 + cur + c * kVecStepT);
        }
      }
      const uint2 Wk = lds64(scr + SC::kW + st * 32u + 8u * kc);  // (weight, weight) of this lane's corner
      const float2 cfk = make_float2(__uint_as_float(Wk.x), __uint_as_float(Wk.y));
      float2 dd = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const Row4 g = row_from_shared<T>(gbuf + (unsigned)t * kRowB + c * kVecStepT);
        dd = __ffma2_rn(g.lo, v[c].lo, dd);
        dd = __ffma2_rn(g.hi, v[c].hi, dd);
        acc[c].lo = __ffma2_rn(cfk, g.lo, acc[c].lo);
        acc[c].hi = __ffma2_rn(cfk, g.hi, acc[c].hi);
      }
      float r = dd.x + dd.y;  // <grad_out, v_kc> over this lane's 16 channels; the other NH - 1 lanes of the corner hold the rest
#pragma unroll
      for (int s = NH / 2; s > 0; s >>= 1) r += __shfl_xor_sync(0xffffffffu, r, s, NH);
      if (hq == 0) sts32(scr + SC::kU + st * 16u + 4u * kc, __float_as_uint(r));
    }
    __syncwarp();
    // ---- finish: the lane that holds the record turns (u_1..u_4) into the gradients of its point -------------------------------------
    if (valid) {
      const unsigned cw = cwC;
      const float a = aC;
      const float fW = (float)s_lvW[lvl], fH = (float)s_lvH[lvl];
      const unsigned o32 = point_offset(idxC);
      const uint4 uu = lds128(scr + SC::kU + my * 16u);
      // corners the reference does not read count as zero rows (their registers held whatever row was loaded last)
      const float u1 = (cw & (1u << 20)) ? __uint_as_float(uu.x) : 0.f, u2 = (cw & (2u << 20)) ? __uint_as_float(uu.y) : 0.f;
      const float u3 = (cw & (4u << 20)) ? __uint_as_float(uu.z) : 0.f, u4 = (cw & (8u << 20)) ? __uint_as_float(uu.w) : 0.f;
      const float hh = 1.f - lh, hw = 1.f - lw;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      const float s_a = w1 * u1 + w2 * u2 + w3 * u3 + w4 * u4;
      const float s_w = hh * (u2 - u1) + lh * (u4 - u3);
      const float s_h = hw * (u3 - u1) + lw * (u4 - u2);
      gaw[o32] = s_a;
      gloc[o32] = make_float2(fW * s_w * a, fH * s_h * a);
    }
    idxC = idxN; xyC = xyN; aC = aN;
    idxN = idxNN;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // (the last iteration's look-ahead copy)
  flush_row();
}

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
template <typename T, int G, int LT, int PT>
static cudaError_t launch_sorted_k(const Params& p, const SortedPlan& sp, dim3 grid, cudaStream_t s) {
  auto kern = msda_bwd_sorted_kernel<T, G, LT, PT, WALK_MINB>;
  constexpr unsigned smem = kWalkWarps * WalkScratch<T, G>::kBytes;
  if (smem > 40u * 1024u) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, kWalkThreads, smem, s>>>(p, sp);
  return cudaGetLastError();
}

template <typename T, int G>
static cudaError_t launch_sorted_c(const Params& p, const SortedPlan& sp, dim3 grid, cudaStream_t s) {
  if (p.L == 3 && p.P == 4) return launch_sorted_k<T, G, 3, 4>(p, sp, grid, s);
  if (p.L == 1 && p.P == 4) return launch_sorted_k<T, G, 1, 4>(p, sp, grid, s);
  return launch_sorted_k<T, G, 0, 0>(p, sp, grid, s);
}

template <typename T>
static cudaError_t launch_sorted_t(const Params& p, const SortedPlan& sp, cudaStream_t s) {
  const dim3 grid((unsigned)((size_t)p.N * p.M * sp.ctas_per_slab));
  if (p.D == 32) return launch_sorted_c<T, 8>(p, sp, grid, s);
  if (p.D == 64) return launch_sorted_c<T, 16>(p, sp, grid, s);
  return cudaErrorNotSupported;
}

cudaError_t bwd_sorted_bf16(const Params& p, const SortedPlan& sp, cudaStream_t s);
cudaError_t bwd_sorted_f16(const Params& p, const SortedPlan& sp, cudaStream_t s);

#if MSDA_TU == 1
cudaError_t bwd_sorted_bf16(const Params& p, const SortedPlan& sp, cudaStream_t s) { return launch_sorted_t<__nv_bfloat16>(p, sp, s); }
#elif MSDA_TU == 2
cudaError_t bwd_sorted_f16(const Params& p, const SortedPlan& sp, cudaStream_t s) { return launch_sorted_t<__half>(p, sp, s); }
#else
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Parts per slab of the sort: enough CTAs for ~4 per SM, at most 64, at least 64 queries each. (2 / 3 per SM measured slower;
// picking the count that spreads the CTAs most evenly over the SMs - 6 parts instead of 4 at ViT-Adapter-B - measured the same.)
static int sort_parts(size_t slabs, int Lq, int sm_count) {
#ifndef SORT_PARTS_PER_SM
#define SORT_PARTS_PER_SM 4
#endif
  long long parts = ((long long)SORT_PARTS_PER_SM * sm_count + (long long)slabs - 1) / (long long)slabs;
  if (parts > kMaxParts) parts = kMaxParts;
  while (parts > 1 && Lq / parts < 64) --parts;
  return (int)(parts < 1 ? 1 : parts);
}

// Bytes of workspace the sorted backward needs for these dimensions (0 = the shape is outside its domain).
size_t backward_sorted_workspace_bytes(int N, int S, int M, int D, int L, int Lq, int P, int sm_count) {
  if (!(D == 32 || D == 64)) return 0;
  if (S >= (1 << 19)) return 0;                        // cell words keep the corner token in 20 bits
  if ((size_t)S * 4 > kSmemBudget - 1024u) return 0;   // the sort's histograms live in shared memory
  const long long cap = (long long)Lq * L * P;
  if (cap * M >= (1ll << 31) || (long long)Lq * M * D * 4 >= (1ll << 31)) return 0;  // 32-bit in-image point / row offsets
  const size_t slabs = (size_t)N * M;
  if (slabs > 65535) return 0;                         // the prefix kernel puts the slab in gridDim.y
  const int parts = sort_parts(slabs, Lq, sm_count);
  return align16(slabs * parts * S * 4) + align16(slabs * S * 4) + align16(slabs * 4) + align16(slabs * (size_t)cap * 4);
}

template <int LT, int PT>
static cudaError_t launch_sort(const Params& p, const SortedPlan& sp, cudaStream_t s) {
  const unsigned smem = (unsigned)p.S * 4u;
  if (smem > 40u * 1024u) {
    cudaError_t e = cudaFuncSetAttribute(msda_sort_part_kernel<LT, PT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(msda_sort_part_kernel<LT, PT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const unsigned grid = (unsigned)((size_t)p.N * p.M * sp.parts);
  msda_sort_part_kernel<LT, PT, false><<<grid, kSortThreads, smem, s>>>(p, sp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int groups = prefix_groups(sp.parts), kpc = kPrefixThreads / groups;
  const dim3 gprefix((unsigned)((p.S + kpc - 1) / kpc), (unsigned)((size_t)p.N * p.M));
  msda_sort_prefix_kernel<<<gprefix, kPrefixThreads, 0, s>>>(sp, p.S, groups);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  msda_sort_part_kernel<LT, PT, true><<<grid, kSortThreads, smem, s>>>(p, sp);
  return cudaGetLastError();
}

// Walker CTAs one SM holds: the launch bound, or fewer where the scratch of a CTA is large (fp32 rows of 64 channels: 2).
static int walker_resident_ctas(int dtype, int D) {
  const unsigned per_warp = dtype == MSDA_F32 ? (D == 64 ? WalkScratch<float, 16>::kBytes : WalkScratch<float, 8>::kBytes)
                                              : (D == 64 ? WalkScratch<__half, 16>::kBytes : WalkScratch<__half, 8>::kBytes);
  const int by_smem = (int)(228u * 1024u / (kWalkWarps * per_warp + 1024u + 256u));
  return by_smem < 1 ? 1 : (by_smem < WALK_MINB ? by_smem : WALK_MINB);
}

// Counting sort + walk. `p.grad_value` must point at the zero-filled fp32 ACCUMULATOR (grad_value itself for f32, the scratch
// for bf16 / f16); `ws` at backward_sorted_workspace_bytes() bytes, 16-byte aligned. `launches` counts kernel launches.
cudaError_t launch_backward_sorted(const Params& p, int dtype, void* ws, int sm_count, int* launches, cudaStream_t s) {
  const size_t slabs = (size_t)p.N * p.M;
  const long long cap = (long long)p.Lq * p.L * p.P;
  SortedPlan sp;
  sp.parts = sort_parts(slabs, p.Lq, sm_count);
  sp.qpp = (p.Lq + sp.parts - 1) / sp.parts;
  char* w = reinterpret_cast<char*>(ws);
  sp.cnt = reinterpret_cast<unsigned*>(w);
  w += align16(slabs * sp.parts * p.S * 4);
  sp.tot = reinterpret_cast<unsigned*>(w);
  w += align16(slabs * p.S * 4);
  sp.nin = reinterpret_cast<unsigned*>(w);
  w += align16(slabs * 4);
  sp.idx = reinterpret_cast<unsigned*>(w);
  sp.cap = (int)cap;
  // positions per warp: as long as possible (every group flushes its last run when its range ends) while the grid is at least
  // two full rounds of resident CTAs and fills its last round: the largest multiple of 32 whose CTA count is >= 90 % (else
  // >= 80 %) of a whole number of rounds; if none, the first of 256 / 128 / 64 / 32 with six CTAs per SM.
  // (ViT-Adapter-S bs 16: 1 056 CTAs at 256 positions per warp are 2.4 rounds of 3 x 148, 1 632 at 160 are 3.7: 235 -> 224 us
  // fp32, 208 -> 198 bf16; T 155 -> 149; B - 2 112 CTAs, 4.8 rounds - and L are unchanged: profiles/r2_walker_ppw_{old,new}.jsonl)
  const long long round_ctas = (long long)walker_resident_ctas(dtype, p.D) * sm_count;
  auto ctas_at = [&](int ppw) { return (long long)slabs * ((cap + (long long)kWalkWarps * ppw - 1) / ((long long)kWalkWarps * ppw)); };
  sp.ppw = 0;
  for (int pass = 0; pass < 2 && sp.ppw == 0; ++pass) {
    for (int ppw = 256; ppw >= 32; ppw -= 32) {
      const long long ctas = ctas_at(ppw), rounds = (ctas + round_ctas - 1) / round_ctas;
      if (rounds >= 2 && ctas * 10 >= rounds * round_ctas * (pass == 0 ? 9 : 8)) { sp.ppw = ppw; break; }
    }
  }
  if (sp.ppw == 0) {
    sp.ppw = 32;
    for (int ppw = 256; ppw >= 32; ppw >>= 1)
      if (ctas_at(ppw) >= 6ll * sm_count) { sp.ppw = ppw; break; }
  }
  sp.ctas_per_slab = (int)((cap + (long long)kWalkWarps * sp.ppw - 1) / ((long long)kWalkWarps * sp.ppw));

  cudaError_t e;
  if (p.L == 3 && p.P == 4) e = launch_sort<3, 4>(p, sp, s);
  else if (p.L == 1 && p.P == 4) e = launch_sort<1, 4>(p, sp, s);
  else e = launch_sort<0, 0>(p, sp, s);
  if (e != cudaSuccess) return e;
  if (dtype == MSDA_F32) e = launch_sorted_t<float>(p, sp, s);
  else if (dtype == MSDA_BF16) e = bwd_sorted_bf16(p, sp, s);
  else if (dtype == MSDA_F16) e = bwd_sorted_f16(p, sp, s);
  else e = cudaErrorNotSupported;
  if (e == cudaSuccess && launches) *launches = 4;
  return e;
}
#endif  // MSDA_TU

}  // namespace msda
