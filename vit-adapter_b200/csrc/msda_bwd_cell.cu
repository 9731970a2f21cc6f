// msda_bwd_cell.cu — cell-bucketed backward of multi-scale deformable attention for sm_100a.
//
// Replaces the reference's col2im kernels (detection/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-510, bilinear
// backward :87-159). The reference - and msda_bwd.cu, the general path of this library - walk the queries in order
// and, per sampled point, gather four value rows and issue four row-wide atomics into grad_value. That formulation
// is bound by the L2's atomic units (16.5 M 128-byte RED rows for the ViT-Adapter-B Extractor at bs 16).
//
// Here a CTA owns (batch b, head m, chunk of consecutive queries) and first BUCKETS its sampled points by the
// bilinear cell they fall in (counting sort in shared memory, integer ATOMS only; the histogram covers just the band
// of rows the chunk touches on each level, in windows if that band is larger than the histogram). Every point of one
// cell reads the same four value rows and updates the same four grad_value rows, so a warp that walks the sorted list
// keeps both in REGISTERS for the whole run of a cell:
//   * lane layout: 4 corner groups x 8 lanes; group k holds value row k of the current cell (D/8 channels per lane)
//     and the partial sum of grad_value row k;
//   * per point: one 16-byte broadcast record from shared memory (corner token, corner weight * attention weight,
//     look-ahead token, offset of the query's grad_out row), one read of the grad_out row (staged in shared memory once
//     per chunk with cp.async), D/8 FMAs for u_k = <grad_out, v_k> and D/8 FMAs for the grad_value partial;
//   * the 8 x 4 partial dot products of 8 points are reduced with ONE transposed shuffle network (7 shuffles per 8
//     points) and handed through shared memory to the lane that prepared the point, which turns (u_1..u_4) into
//     grad_attn_weight and grad_sampling_loc (the same linear-in-the-corners algebra as msda_bwd.cu);
//   * when the cell changes the four partial rows leave with ONE vector reduction per lane (REDG.E.ADD.F32x4): row
//     atomics per point drop from 4 to 4 / (points per cell run), value-row gathers likewise; the value rows of the
//     NEXT run are requested as soon as a run starts (one look-ahead buffer in registers, tagged by token, so a wrong
//     or missing look-ahead only costs a direct load).
// Nothing depends on where the reference points are: bucketing is by the actual sampling location, so any
// distribution is handled; spatially coherent queries (the adapter's raster order) just make the runs longer.
#include "msda_cell_common.cuh"

namespace msda {

// Shared-memory plan made on the host (plan_backward_cell below). Sorted records start at byte 0.
struct CellPlan {
  unsigned off_tl, off_go, off_scr, off_ubuf, off_hist;  // byte offsets into dynamic shared memory
  unsigned total;                                        // dynamic shared memory bytes
  int hist_cap;                                          // histogram entries (one key window)
};

constexpr int kCellPPT = 8;                         // sampled points per thread a chunk may hold (registers carry them)
constexpr int kCellMaxPts = kCellPPT * kThreads;    // 2048 points per chunk
// Level table in shared memory (int each): H[16] | W[16] | start[16] | rmin[16] | rmax[16] | off[16]
constexpr unsigned kLvH = 0, kLvW = 64, kLvStart = 128, kLvRmin = 192, kLvRmax = 256, kLvOff = 320, kLvBytes = 384;

// T = value dtype, CPL = channels per lane (D = 8 * CPL), LT/PT = compile-time levels / points (0,0 = runtime).
//
// Corner record (16 bytes, one per corner of every point of a batch, in the warp's broadcast scratch):
//   x  byte offset of the corner's row inside the (b, m) value slab, + 2;  1 = the reference does not read this corner
//   y  (only where a cell starts) the x-word of the same corner of the NEXT cell of this warp's range: the row to request
//      now;  1 = unknown / unread
//   z  bilinear weight of the corner * attention weight (0 for an unread corner)
//   w  byte offset of the query's grad_out row in the staged tile | 1 when the point STARTS a cell
// (z, w) is all a point needs while its cell continues; (x, y) is read only where a cell starts.
template <typename T, int CPL, int LT, int PT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) msda_bwd_cell_kernel(const Params p, const CellPlan cp) {
  constexpr bool kStatic = (LT > 0);
  constexpr int D = 8 * CPL;
  constexpr unsigned kRowB = D * sizeof(T);        // bytes of one head row of value / grad_out
  constexpr int kAccShift = sizeof(T) == 2 ? 1 : 0;  // fp32 accumulator rows are 4 / sizeof(T) times as long
  extern __shared__ __align__(16) char smem[];

  const int L = kStatic ? LT : p.L;
  const int P = kStatic ? PT : p.P;
  const int LP = L * P;
  const unsigned MDb = (unsigned)p.M * kRowB;  // bytes between neighbouring tokens of the value tensor

  __shared__ __align__(16) int s_lv[kLvBytes / 4];
  __shared__ unsigned s_warp_tot[kWarps];
  __shared__ int s_nin, s_R;

  const BlockCoord bc = block_coord(p);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = bc.q_end - bc.q_begin;
  const int npts = nq * LP;

  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned lv = (unsigned)__cvta_generic_to_shared(s_lv);
  const unsigned rec_s = sbase;                 // sorted: (lh, lw, attn, point index)  16 B
  const unsigned tl_s = sbase + cp.off_tl;      // sorted: cell word                     4 B
  const unsigned hist_s = sbase + cp.off_hist;
  // (b, q_begin, m) bases; everything inside the chunk is a 32-bit offset from them
  const size_t pair0 = ((size_t)bc.b * p.Lq + bc.q_begin) * p.M + bc.m;
  const unsigned qstride = (unsigned)p.M * (unsigned)LP;  // points between consecutive queries of this head

  // ---- phase 0: level table, grad_out rows of the chunk on their way into shared memory ----------------------------------
  if (tid < L) {
    s_lv[kLvH / 4 + tid] = (int)p.shapes[2 * tid];
    s_lv[kLvW / 4 + tid] = (int)p.shapes[2 * tid + 1];
    s_lv[kLvStart / 4 + tid] = (int)p.lsi[tid];
    s_lv[kLvRmin / 4 + tid] = 0x7fffffff;
    s_lv[kLvRmax / 4 + tid] = -1;
  }
  {
    constexpr int kVecPerRow = kRowB / 16;
    const char* __restrict__ gsrc = reinterpret_cast<const char*>(p.grad_out) + pair0 * kRowB;
    const unsigned gdst = sbase + cp.off_go;
    for (int i = tid; i < nq * kVecPerRow; i += kThreads) {
      const int r = i / kVecPerRow, c = i % kVecPerRow;
      cp_async16(gdst + r * kRowB + c * 16, gsrc + (size_t)r * MDb + c * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __syncthreads();

  // ---- phase 1: classify every sample ONCE; its cell word stays in a register until the scatter ---------------------------
  // kNoCell = not a sample of this chunk, or out of range (the reference's bounds test, ms_deform_im2col_cuda.cuh:365 -
  // its gradients are zero and are written here).
  unsigned c_cw[kCellPPT];
  {
    const float2* __restrict__ loc0 = reinterpret_cast<const float2*>(p.loc) + pair0 * LP;
    float2* __restrict__ gloc0 = reinterpret_cast<float2*>(p.grad_loc) + pair0 * LP;
    float* __restrict__ gaw0 = reinterpret_cast<float*>(p.grad_aw) + pair0 * LP;
#pragma unroll
    for (int kk = 0; kk < kCellPPT; ++kk) {
      const int i = tid + kk * kThreads;
      c_cw[kk] = kNoCell;
      if (i < npts) {
        const int ql = i / LP, pi = i - ql * LP;
        const unsigned o = (unsigned)ql * qstride + (unsigned)pi;
        const float2 xy = __ldg(loc0 + o);
        const int l = pi / P;
        const int H = (int)lds32(lv + kLvH + 4 * l), W = (int)lds32(lv + kLvW + 4 * l);
        const PointGeom<float> g = point_geom<float>(xy.x, xy.y, H, W);
        if (g.mask != 0u) {
          c_cw[kk] = (unsigned)((int)lds32(lv + kLvStart + 4 * l) + (g.h_low + 1) * W + (g.w_low + 1)) | (g.mask << 20) | ((unsigned)l << 24);
          const int row = max(g.h_low, 0);
          reds_min(lv + kLvRmin + 4 * l, row);
          reds_max(lv + kLvRmax + 4 * l, row);
        } else {
          gaw0[o] = 0.f;
          gloc0[o] = make_float2(0.f, 0.f);
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int l = 0; l < L; ++l) {
      const int rmin = s_lv[kLvRmin / 4 + l], rmax = s_lv[kLvRmax / 4 + l], W = s_lv[kLvW / 4 + l];
      const int rows = rmax >= rmin ? rmax - rmin + 1 : 0;
      s_lv[kLvOff / 4 + l] = acc - (rows ? rmin * W : 0) - s_lv[kLvStart / 4 + l];  // key = off[l] + clamped top-left token
      acc += rows * W;
    }
    s_R = acc;
  }
  __syncthreads();
  const int R = s_R;
  // sort key of a cell word: the clamped top-left token, counted from the first touched row of its level
  auto sort_key = [&](unsigned cw) -> int {
    const unsigned l4 = (cw >> 22) & 60u;  // 4 * level
    const int W = (int)lds32(lv + kLvW + l4);
    int t = (int)(cw & 0xFFFFFu) - W - 1;        // unclamped top-left token
    if ((cw & (3u << 20)) == 0u) t += W;         // no readable corner in the top row: h_low = -1
    if ((cw & (5u << 20)) == 0u) t += 1;         // no readable corner in the left column: w_low = -1
    return t + (int)lds32(lv + kLvOff + l4);
  };

  // ---- per-lane constants of the consumer ----------------------------------------------------------------------------------
  const int k = lane >> 3, j = lane & 7;  // corner group, channel slice
  const unsigned scr_w = sbase + cp.off_scr + warp * kScrBytesPerWarp;     // this warp's broadcast scratch
  const unsigned scr_k = scr_w + k * (kScrStride * 16u);                    // ... its corner row
  const unsigned ub_w = sbase + cp.off_ubuf + warp * kUbufBytesPerWarp;
  const unsigned go_s = sbase + cp.off_go + j * (CPL * (unsigned)sizeof(T));
  const size_t slab = ((size_t)bc.b * p.S * p.M + (size_t)bc.m) * D;
  // both bases are biased by the "+ 2" of the record's offset words
  const char* __restrict__ vb2 = reinterpret_cast<const char*>(p.value) + slab * sizeof(T) + j * (CPL * sizeof(T)) - 2;
  char* __restrict__ gvb2 = reinterpret_cast<char*>(p.grad_value) + slab * 4u + j * (CPL * 4) - (2 << kAccShift);  // fp32 accumulator

  for (int w0 = 0; w0 < R; w0 += cp.hist_cap) {
    const int wn = min(cp.hist_cap, R - w0);
    // ---- phase 2: histogram of this key window, exclusive scan -------------------------------------------------------------
    if (w0 > 0) __syncthreads();  // the previous window's consumers are done with the sorted records and the histogram
    for (int i = tid; i < wn; i += kThreads) sts32(hist_s + 4 * i, 0u);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kCellPPT; ++kk) {
      if (c_cw[kk] != kNoCell) {
        const unsigned key = (unsigned)(sort_key(c_cw[kk]) - w0);
        if (key < (unsigned)wn) reds_add(hist_s + 4 * key, 1u);
      }
    }
    __syncthreads();
    {
      const int seg = ((wn + kThreads - 1) / kThreads) | 1;  // odd stride: conflict-free thread-serial segments
      const int lo = min(tid * seg, wn), hi = min(lo + seg, wn);
      unsigned sum = 0;
      for (int i = lo; i < hi; ++i) sum += lds32(hist_s + 4 * i);
      unsigned incl = sum;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += o;
      }
      if (lane == 31) s_warp_tot[warp] = incl;
      __syncthreads();
      unsigned before = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) before += (w < warp) ? s_warp_tot[w] : 0u;
      unsigned run = before + incl - sum;
      for (int i = lo; i < hi; ++i) {
        const unsigned c = lds32(hist_s + 4 * i);
        sts32(hist_s + 4 * i, run);
        run += c;
      }
      if (tid == kThreads - 1) s_nin = (int)run;
    }
    __syncthreads();
    // ---- phase 3: scatter the window's samples into cell order ------------------------------------------------------------
    {
      const float2* __restrict__ loc0 = reinterpret_cast<const float2*>(p.loc) + pair0 * LP;
      const float* __restrict__ aw0 = reinterpret_cast<const float*>(p.aw) + pair0 * LP;
#pragma unroll
      for (int kk = 0; kk < kCellPPT; ++kk) {
        if (c_cw[kk] != kNoCell) {
          const unsigned key = (unsigned)(sort_key(c_cw[kk]) - w0);
          if (key < (unsigned)wn) {
            // the fractions again (two FFMAs and two floors; the bounds test and the corner mask are in the cell word)
            const int i = tid + kk * kThreads;
            const int ql = i / LP, pi = i - ql * LP;
            const unsigned o = (unsigned)ql * qstride + (unsigned)pi;
            const float2 xy = __ldg(loc0 + o);
            const float a = __ldg(aw0 + o);
            const unsigned l4 = (c_cw[kk] >> 22) & 60u;
            const float h_im = fmaf(xy.y, (float)(int)lds32(lv + kLvH + l4), -0.5f);
            const float w_im = fmaf(xy.x, (float)(int)lds32(lv + kLvW + l4), -0.5f);
            const unsigned pos = atoms_add(hist_s + 4 * key, 1u);
            sts128(rec_s + 16 * pos, __float_as_uint(h_im - floorf(h_im)), __float_as_uint(w_im - floorf(w_im)), __float_as_uint(a),
                   (unsigned)ql | ((unsigned)pi << 16));
            sts32(tl_s + 4 * pos, c_cw[kk]);
          }
        }
      }
    }
    if (w0 == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- phase 4: walk the sorted list; one warp = one contiguous range, all 32 lanes on one point at a time -----------
    const int n_in = s_nin;
    const int nb = (n_in + 31) >> 5;
    const int b_begin = (int)(((long long)warp * nb) / kWarps), b_end = (int)(((long long)(warp + 1) * nb) / kWarps);
    const int rend = min(n_in, b_end * 32);  // end of this warp's range (positions)

    unsigned lastcw = 0xFFFFFFFEu;  // cell word of the previous position (warp-uniform)
    unsigned cur = 1u, tagn = 0u;   // x-word of the current cell's corner; x-word whose row sits in vn (0 = none)
    float v[CPL], vn[CPL], acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { v[c] = 0.f; vn[c] = 0.f; acc[c] = 0.f; }

    // x-words of the four corners of a cell word (see the record layout above)
    auto corner_words = [&](unsigned cw, unsigned (&xw)[4]) {
      const unsigned W = lds32(lv + kLvW + ((cw >> 22) & 60u));
      const unsigned br = cw & 0xFFFFFu;
      xw[0] = (cw & (1u << 20)) ? (br - W - 1u) * MDb + 2u : 1u;
      xw[1] = (cw & (2u << 20)) ? (br - W) * MDb + 2u : 1u;
      xw[2] = (cw & (4u << 20)) ? (br - 1u) * MDb + 2u : 1u;
      xw[3] = (cw & (8u << 20)) ? br * MDb + 2u : 1u;
    };

    for (int bi = b_begin; bi < b_end; ++bi) {
      const int bbase = bi * 32;
      const int pos = bbase + lane;
      const bool valid = pos < rend;
      // ---- prepare: lane i turns sorted record i of the batch into four corner records ------------------------------------
      {
        const unsigned cw0 = valid ? lds32(tl_s + 4 * pos) : kNoCell;
        const unsigned cw1 = (pos + 32 < rend) ? lds32(tl_s + 4 * pos + 128) : kNoCell;  // the next batch, for the look-ahead only
        unsigned prev0 = __shfl_up_sync(0xffffffffu, cw0, 1);
        if (lane == 0) prev0 = lastcw;
        const unsigned last0 = __shfl_sync(0xffffffffu, cw0, 31);
        unsigned prev1 = __shfl_up_sync(0xffffffffu, cw1, 1);
        if (lane == 0) prev1 = last0;
        const bool head = cw0 != prev0;
        const bool head1 = (cw1 != prev1) && (cw1 != kNoCell);
        const unsigned H0 = __ballot_sync(0xffffffffu, head && valid);
        const unsigned H1 = __ballot_sync(0xffffffffu, head1);
        lastcw = last0;
        // the cells the NEXT batch starts: ask the L2 for their rows now (they are 32 points away)
        if (head1 || (bi == b_begin && head && valid)) {
          unsigned xw[4];
          corner_words(head1 ? cw1 : cw0, xw);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (xw[kk] > 1u) prefetch_l2(vb2 + xw[kk]);
        }
        unsigned xw[4] = {1u, 1u, 1u, 1u}, zw[4] = {0u, 0u, 0u, 0u};
        float cf[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned gooff = 0u;
        if (valid) {
          const uint4 r = lds128(rec_s + 16 * pos);
          const float lh = __uint_as_float(r.x), lw = __uint_as_float(r.y), a = __uint_as_float(r.z);
          gooff = (r.w & 0xFFFFu) * kRowB;
          corner_words(cw0, xw);
          const float hh = 1.f - lh, hw = 1.f - lw;
          cf[0] = xw[0] > 1u ? (hh * hw) * a : 0.f;
          cf[1] = xw[1] > 1u ? (hh * lw) * a : 0.f;
          cf[2] = xw[2] > 1u ? (lh * hw) * a : 0.f;
          cf[3] = xw[3] > 1u ? (lh * lw) * a : 0.f;
        }
        if (head) {  // (a lane past the range end also starts a "cell": it makes the consumer flush)
          zw[0] = zw[1] = zw[2] = zw[3] = 1u;
          const unsigned lo = H0 & (0xFFFFFFFEu << lane);  // heads after this lane, this batch
          int npos = -1;
          if (lo != 0u) npos = bbase + __ffs(lo) - 1;
          else if (H1 != 0u) npos = bbase + 32 + __ffs(H1) - 1;
          if (npos >= 0) corner_words(lds32(tl_s + 4 * npos), zw);
        }
        __syncwarp();  // the previous batch's readers are done with the scratch
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          sts128(scr_w + (kk * kScrStride + lane) * 16u, xw[kk], zw[kk], __float_as_uint(cf[kk]), gooff | (head ? 1u : 0u));
        __syncwarp();
      }
      // ---- consume: 8 points per round ------------------------------------------------------------------------------------------
      const int nsb = min(4, (rend - bbase + 7) >> 3);
      for (int sb = 0; sb < nsb; ++sb) {
        float d[8];
        // software pipeline: the steady half of the next record (weight, grad_out offset) and the grad_out row it names
        // are requested before the current point is worked on
        uint2 Rn = lds64(scr_k + (sb * 8) * 16u + 8u);
        float gn[CPL];
        ch_load_shared<T, CPL>(go_s + (Rn.y & ~1u), gn);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint2 Rc = Rn;
          float g[CPL];
#pragma unroll
          for (int c = 0; c < CPL; ++c) g[c] = gn[c];
          if (t < 7) {
            Rn = lds64(scr_k + (sb * 8 + t + 1) * 16u + 8u);
            ch_load_shared<T, CPL>(go_s + (Rn.y & ~1u), gn);
          }
          if (Rc.y & 1u) {  // warp-uniform: a new cell starts here
            const uint2 Hh = lds64(scr_k + (sb * 8 + t) * 16u);  // (row of this cell, row of the next cell)
            if (cur > 1u) {
              float* dst = reinterpret_cast<float*>(gvb2 + ((size_t)cur << kAccShift));
#pragma unroll
              for (int c = 0; c < CPL; c += 4) red_add_v4(dst + c, acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
            }
            cur = Hh.x;
            if (cur == tagn) {  // the row was requested when the previous cell started
#pragma unroll
              for (int c = 0; c < CPL; ++c) v[c] = vn[c];
            } else if (cur > 1u) {
              ch_load_global<T, CPL>(vb2 + cur, v);
            }
            tagn = Hh.y;
            if (tagn > 1u) ch_load_global<T, CPL>(vb2 + tagn, vn);
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
          }
          const float cfk = __uint_as_float(Rc.x);
          float dd0 = 0.f, dd1 = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; c += 2) {
            dd0 = fmaf(g[c], v[c], dd0);
            dd1 = fmaf(g[c + 1], v[c + 1], dd1);
            acc[c] = fmaf(cfk, g[c], acc[c]);
            acc[c + 1] = fmaf(cfk, g[c + 1], acc[c + 1]);
          }
          d[t] = dd0 + dd1;
        }
        // transposed reduction over the 8 lanes of a corner group: afterwards lane (k, j) holds u_k of point sb*8 + j
        {
          const bool up = (j & 4) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = up ? d[i] : d[i + 4];
            const float keep = up ? d[i + 4] : d[i];
            d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4, 8);
          }
        }
        {
          const bool up = (j & 2) != 0;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = up ? d[i] : d[i + 2];
            const float keep = up ? d[i + 2] : d[i];
            d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2, 8);
          }
        }
        {
          const bool up = (j & 1) != 0;
          const float send = up ? d[0] : d[1];
          const float keep = up ? d[1] : d[0];
          d[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1, 8);
        }
        sts32(ub_w + ((sb * 8 + j) * 4 + k) * 4u, __float_as_uint(d[0]));
      }
      __syncwarp();
      // ---- finish: the preparing lane turns (u_1..u_4) into the gradients of its point --------------------------------------
      if (valid) {  // (its record is read again rather than kept in registers across the rounds)
        const uint4 r = lds128(rec_s + 16 * pos);
        const unsigned cw = lds32(tl_s + 4 * pos);
        const unsigned l4 = (cw >> 22) & 60u;
        const float lh = __uint_as_float(r.x), lw = __uint_as_float(r.y), a = __uint_as_float(r.z);
        const float fW = (float)(int)lds32(lv + kLvW + l4), fH = (float)(int)lds32(lv + kLvH + l4);
        const unsigned o32 = (r.w & 0xFFFFu) * qstride + (r.w >> 16);
        const uint4 uu = lds128(ub_w + lane * 16u);
        // corners the reference does not read count as zero rows (their lanes held whatever row was loaded last)
        const float u1 = (cw & (1u << 20)) ? __uint_as_float(uu.x) : 0.f, u2 = (cw & (2u << 20)) ? __uint_as_float(uu.y) : 0.f;
        const float u3 = (cw & (4u << 20)) ? __uint_as_float(uu.z) : 0.f, u4 = (cw & (8u << 20)) ? __uint_as_float(uu.w) : 0.f;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float s_a = w1 * u1 + w2 * u2 + w3 * u3 + w4 * u4;
        const float s_w = hh * (u2 - u1) + lh * (u4 - u3);
        const float s_h = hw * (u3 - u1) + lw * (u4 - u2);
        reinterpret_cast<float*>(p.grad_aw)[pair0 * LP + o32] = s_a;
        reinterpret_cast<float2*>(p.grad_loc)[pair0 * LP + o32] = make_float2(fW * s_w * a, fH * s_h * a);
      }
    }
    if (cur > 1u) {
      float* dst = reinterpret_cast<float*>(gvb2 + ((size_t)cur << kAccShift));
#pragma unroll
      for (int c = 0; c < CPL; c += 4) red_add_v4(dst + c, acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    }
  }
  if (R == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");  // nothing in range: still drain the staged copy
}

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
template <typename T, int CPL, int LT, int PT>
static cudaError_t launch_cell_k(const Params& p, const CellPlan& cp, dim3 grid, cudaStream_t s) {
  auto kern = msda_bwd_cell_kernel<T, CPL, LT, PT, (CPL == 4 ? 3 : 2)>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp.total);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, cp.total, s>>>(p, cp);
  return cudaGetLastError();
}

template <typename T, int CPL>
static cudaError_t launch_cell_c(const Params& p, const CellPlan& cp, dim3 grid, cudaStream_t s) {
  if (p.L == 3 && p.P == 4) return launch_cell_k<T, CPL, 3, 4>(p, cp, grid, s);
  if (p.L == 1 && p.P == 4) return launch_cell_k<T, CPL, 1, 4>(p, cp, grid, s);
  return launch_cell_k<T, CPL, 0, 0>(p, cp, grid, s);
}

template <typename T>
static cudaError_t launch_cell_t(const Params& p, const CellPlan& cp, cudaStream_t s) {
  const dim3 grid((unsigned)((size_t)p.N * p.nchunk * p.M));
  if (p.D == 32) return launch_cell_c<T, 4>(p, cp, grid, s);
  if (p.D == 64) return launch_cell_c<T, 8>(p, cp, grid, s);
  return cudaErrorNotSupported;
}

#ifndef MSDA_TU
#define MSDA_TU 0
#endif
cudaError_t bwd_cell_bf16(const Params& p, const CellPlan& cp, cudaStream_t s);
cudaError_t bwd_cell_f16(const Params& p, const CellPlan& cp, cudaStream_t s);

#if MSDA_TU == 1
cudaError_t bwd_cell_bf16(const Params& p, const CellPlan& cp, cudaStream_t s) { return launch_cell_t<__nv_bfloat16>(p, cp, s); }
#elif MSDA_TU == 2
cudaError_t bwd_cell_f16(const Params& p, const CellPlan& cp, cudaStream_t s) { return launch_cell_t<__half>(p, cp, s); }
#else
// Largest chunk the kernel accepts: its samples live in kCellPPT registers per thread between classification and scatter.
int backward_cell_max_points() { return kCellMaxPts; }

// Shared-memory plan for (queries per chunk, histogram entries). Returns false when it does not fit `budget`.
bool plan_backward_cell(int D, int LP, int esize, int qc, int hist_cap, unsigned budget, CellPlan* cp) {
  if ((long long)qc * LP > kCellMaxPts || hist_cap < 1) return false;
  const unsigned npts = (unsigned)qc * LP;
  const unsigned rec = npts * 16u;
  const unsigned tl = (npts * 4u + 15u) & ~15u;
  const unsigned go = (unsigned)qc * D * esize;  // multiple of 16 (D in {32, 64})
  const unsigned scr = kWarps * kScrBytesPerWarp, ubuf = kWarps * kUbufBytesPerWarp;
  const unsigned hist = ((unsigned)hist_cap * 4u + 15u) & ~15u;
  cp->off_tl = rec;
  cp->off_go = cp->off_tl + tl;
  cp->off_scr = cp->off_go + go;
  cp->off_ubuf = cp->off_scr + scr;
  cp->off_hist = cp->off_ubuf + ubuf;
  cp->total = cp->off_hist + hist;
  cp->hist_cap = hist_cap;
  return cp->total <= budget && go < (1u << 31);
}

// `p.grad_value` must point at the zero-filled fp32 ACCUMULATOR (grad_value itself for f32, the scratch for bf16 / f16).
cudaError_t launch_backward_cell(const Params& p, const CellPlan& cp, int dtype, cudaStream_t s) {
  if (dtype == MSDA_F32) return launch_cell_t<float>(p, cp, s);
  if (dtype == MSDA_BF16) return bwd_cell_bf16(p, cp, s);
  if (dtype == MSDA_F16) return bwd_cell_f16(p, cp, s);
  return cudaErrorNotSupported;
}
#endif  // MSDA_TU

}  // namespace msda
