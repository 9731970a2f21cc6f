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
#include <cuda_fp16.h>

#include <type_traits>

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
template <> __device__ __forceinline__ float dw_ld<__half>(const __half* p) { return __half2float(*p); }
template <typename T, typename A> __device__ __forceinline__ void dw_st(T* p, A v) { *p = (T)v; }
template <> __device__ __forceinline__ void dw_st<__nv_bfloat16, float>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void dw_st<__half, float>(__half* p, float v) { *p = __float2half_rn(v); }

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
        } else if constexpr (sizeof(T) * VEC == 16 && std::is_same<T, __nv_bfloat16>::value) {
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
    } else if constexpr (sizeof(T) * VEC == 16 && std::is_same<T, __nv_bfloat16>::value) {
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

// ---------------------------------------------------------------------------------------------------------------
// Sliding-window kernels (fp32 / bf16, C % 4 == 0): the fast path.
//
// The kernels above read nine neighbours per output token; at 16 bytes per lane that is nine L1 row gathers per
// 128 output bytes, and the measured L1 gather ceiling (profiles/r1_microbench.md, ~0.49 rows/cycle/SM) caps them at
// ~1/3 of the HBM roofline. Here a thread owns 4 channels and a RUN of kRun consecutive tokens of one map row and
// carries the 3x3 window in registers: per output it loads only the three tokens of the next column (3.75 loads per
// output with the two halo columns of a run, instead of 9), and the taps live in registers, so there is no shared-
// memory traffic at all in the forward / grad_x kernels. A work item is (batch, map, row, run); lanes of a warp are
// adjacent channel quads of the same item, so each load instruction covers one contiguous token row.
constexpr int kRun = 8;

struct DwRunGeom {
  int seg[3];       // runs per row of each map
  int items[3];     // rows * runs of each map (per batch)
  int per_batch;    // sum of items
  int total;        // per_batch * B
};

static DwRunGeom dw_run_geom(const DwParams& p) {
  DwRunGeom g;
  const int mh[3] = {2 * p.H, p.H, p.H / 2}, mw[3] = {2 * p.W, p.W, p.W / 2};
  g.per_batch = 0;
  for (int i = 0; i < 3; ++i) {
    g.seg[i] = (mw[i] + kRun - 1) / kRun;
    g.items[i] = mh[i] * g.seg[i];
    g.per_batch += g.items[i];
  }
  g.total = g.per_batch * p.B;
  return g;
}

// work item -> batch, first token of the map, map height / width, row, first column of the run
__device__ __forceinline__ void dw_run_locate(int item, const DwParams& p, const DwRunGeom& g, int& b, int& t0, int& mh, int& mw,
                                              int& h, int& w0) {
  b = item / g.per_batch;
  int r = item - b * g.per_batch;
  int seg;
  if (r < g.items[0]) { t0 = 0; mh = 2 * p.H; mw = 2 * p.W; seg = g.seg[0]; }
  else if (r < g.items[0] + g.items[1]) { r -= g.items[0]; t0 = 4 * p.H * p.W; mh = p.H; mw = p.W; seg = g.seg[1]; }
  else { r -= g.items[0] + g.items[1]; t0 = 5 * p.H * p.W; mh = p.H / 2; mw = p.W / 2; seg = g.seg[2]; }
  h = r / seg;
  w0 = (r - h * seg) * kRun;
}

template <typename T> struct Quad;  // 4 consecutive channels of one token: raw 8/16-byte load, fp32 view
template <> struct Quad<float> {
  using Raw = float4;
  static __device__ __forceinline__ Raw ld(const float* p, bool ok) {
    Raw q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) q = __ldg(reinterpret_cast<const float4*>(p));
    return q;
  }
  static __device__ __forceinline__ void unpack(const Raw& q, float (&v)[4]) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Quad<__nv_bfloat16> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw ld(const __nv_bfloat16* p, bool ok) {
    Raw q = make_uint2(0u, 0u);
    if (ok) q = __ldg(reinterpret_cast<const uint2*>(p));
    return q;
  }
  static __device__ __forceinline__ void unpack(const Raw& q, float (&v)[4]) {
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 o;
    o.x = Vec<__nv_bfloat16>::pack2(v[0], v[1]); o.y = Vec<__nv_bfloat16>::pack2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = o;
  }
};

template <> struct Quad<__half> {
  using Raw = uint2;
  static __device__ __forceinline__ Raw ld(const __half* p, bool ok) {
    Raw q = make_uint2(0u, 0u);
    if (ok) q = __ldg(reinterpret_cast<const uint2*>(p));
    return q;
  }
  static __device__ __forceinline__ void unpack(const Raw& q, float (&v)[4]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st(__half* p, const float (&v)[4]) {
    uint2 o;
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    o.x = *reinterpret_cast<const unsigned*>(&a); o.y = *reinterpret_cast<const unsigned*>(&b);
    *reinterpret_cast<uint2*>(p) = o;
  }
};

// Packed fp32 pairs: Blackwell issues two IEEE fp32 FMAs per instruction (FFMA2, PTX fma.rn.f32x2) on 64-bit register
// pairs. The run kernels below are issue-bound in the 16-bit types (36 FMAs + 15 unpack instructions per 8 output
// bytes), so the 36 FMAs of an output quad go out as 18 FFMA2; each lane's result is bit-identical to the scalar form.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 p, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ void ffma2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ void fadd2(f32x2& acc, f32x2 a) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a)); }
template <typename T>
__device__ __forceinline__ void unpack_pairs(const typename Quad<T>::Raw& q, f32x2 (&v)[2]) {
  float f[4];
  Quad<T>::unpack(q, f);
  v[0] = pk2(f[0], f[1]);
  v[1] = pk2(f[2], f[3]);
}

// the kRun + 2 tokens (columns w0-1 .. w0+kRun) of one map row, zero outside the map: all loads issued back to back
template <typename T>
__device__ __forceinline__ void dw_load_row(const T* row, int C, int w0, int mw, bool row_ok, typename Quad<T>::Raw (&raw)[kRun + 2]) {
#pragma unroll
  for (int jj = 0; jj < kRun + 2; ++jj) {
    const int w = w0 - 1 + jj;
    raw[jj] = Quad<T>::ld(row + (ptrdiff_t)w * C, row_ok && w >= 0 && w < mw);
  }
}

// y = dwconv3x3(x) (+ bias), or with FLIP grad_x = dwconv3x3(grad_y, taps rotated by 180 degrees).
// blockDim.x = cvec * ty_count (cvec = C / 4): thread -> (channel quad cv, item slot ty). Per item the three input
// rows are processed one after the other, each as kRun + 2 independent loads (memory-level parallelism: 10 loads of
// 512 B per warp in flight instead of 3 when walking column by column); bf16 keeps the next row in flight too.
template <typename T, bool FLIP>
__global__ void __launch_bounds__(256, 2) adapter_dwconv_run_kernel(const DwParams p, const DwRunGeom g, int cvec, int ty_count) {
  using Raw = typename Quad<T>::Raw;
  constexpr bool kAhead = sizeof(T) == 2;
  const int cv = threadIdx.x % cvec, ty = threadIdx.x / cvec;
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  T* __restrict__ y = reinterpret_cast<T*>(p.y);
  f32x2 tp[9][2], bias[2];
  {
    float tf[9][4], bf[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const T* wv = reinterpret_cast<const T*>(p.w) + (cv * 4 + v) * 9;
#pragma unroll
      for (int k = 0; k < 9; ++k) tf[FLIP ? 8 - k : k][v] = dw_ld<T>(wv + k);
      bf[v] = (!FLIP && p.bias) ? dw_ld<T>(reinterpret_cast<const T*>(p.bias) + cv * 4 + v) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) { tp[k][0] = pk2(tf[k][0], tf[k][1]); tp[k][1] = pk2(tf[k][2], tf[k][3]); }
    bias[0] = pk2(bf[0], bf[1]);
    bias[1] = pk2(bf[2], bf[3]);
  }
  for (int item = blockIdx.x * ty_count + ty; item < g.total; item += gridDim.x * ty_count) {
    int b, t0, mh, mw, h, w0;
    dw_run_locate(item, p, g, b, t0, mh, mw, h, w0);
    const size_t row_tok = (size_t)b * p.Ntok + t0 + h * mw;  // token index of (h, 0)
    const T* xr = x + row_tok * p.C + cv * 4;
    T* yr = y + row_tok * p.C + cv * 4;
    const ptrdiff_t rs = (ptrdiff_t)mw * p.C;  // one map row in elements
    const bool rok[3] = {h > 0, true, h + 1 < mh};
    f32x2 acc[kRun][2];
#pragma unroll
    for (int j = 0; j < kRun; ++j) { acc[j][0] = bias[0]; acc[j][1] = bias[1]; }
    Raw cur[kRun + 2], nxt[kRun + 2];
    dw_load_row<T>(xr - rs, p.C, w0, mw, rok[0], cur);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (kAhead && r < 2) dw_load_row<T>(xr + (r == 0 ? 0 : rs), p.C, w0, mw, rok[r + 1], nxt);
#pragma unroll
      for (int jj = 0; jj < kRun + 2; ++jj) {
        f32x2 val[2];
        unpack_pairs<T>(cur[jj], val);
#pragma unroll
        for (int dx = 2; dx >= 0; --dx) {   // token jj is column dx of output j = jj - dx; dx descending keeps the
          const int j = jj - dx;            // per-output order (row-major over the taps)
          if (j < 0 || j >= kRun) continue;
          ffma2(acc[j][0], val[0], tp[r * 3 + dx][0]);
          ffma2(acc[j][1], val[1], tp[r * 3 + dx][1]);
        }
      }
      if (r < 2) {
        if (kAhead) {
#pragma unroll
          for (int jj = 0; jj < kRun + 2; ++jj) cur[jj] = nxt[jj];
        } else {
          dw_load_row<T>(xr + (r == 0 ? 0 : rs), p.C, w0, mw, rok[r + 1], cur);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kRun; ++j)
      if (w0 + j < mw) {
        float o[4];
        upk2(acc[j][0], o[0], o[1]);
        upk2(acc[j][1], o[2], o[3]);
        Quad<T>::st(yr + (ptrdiff_t)(w0 + j) * p.C, o);
      }
  }
}

// grad_weight / grad_bias with the same run decomposition: per item the grad_y quads of the run stay in registers and
// each of the three x rows is loaded as kRun + 2 independent quads; 40 partial sums per thread, reduced over the CTA's
// item slots in shared memory and written as one partial row per CTA (no atomics: the second kernel sums the rows in a
// fixed order, so grad_weight is deterministic).
template <typename T>
__global__ void __launch_bounds__(256, 2) adapter_dwconv_wgrad_run_kernel(const DwParams p, const DwRunGeom g, const void* grad_y_,
                                                                          float* partial, int cvec, int ty_count) {
  using Raw = typename Quad<T>::Raw;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  float* red = reinterpret_cast<float*>(dw_smem);  // [ty_count][40][cvec]
  const int cv = threadIdx.x % cvec, ty = threadIdx.x / cvec;
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  const T* __restrict__ gy = reinterpret_cast<const T*>(grad_y_);
  f32x2 s[10][2];
#pragma unroll
  for (int k = 0; k < 10; ++k) { s[k][0] = pk2(0.f, 0.f); s[k][1] = pk2(0.f, 0.f); }
  for (int item = blockIdx.x * ty_count + ty; item < g.total; item += gridDim.x * ty_count) {
    int b, t0, mh, mw, h, w0;
    dw_run_locate(item, p, g, b, t0, mh, mw, h, w0);
    const size_t row_tok = (size_t)b * p.Ntok + t0 + h * mw;
    const T* xr = x + row_tok * p.C + cv * 4;
    const T* gr = gy + row_tok * p.C + cv * 4;
    const ptrdiff_t rs = (ptrdiff_t)mw * p.C;
    const bool rok[3] = {h > 0, true, h + 1 < mh};
    f32x2 gq[kRun][2];
#pragma unroll
    for (int j = 0; j < kRun; ++j) {
      const Raw graw = Quad<T>::ld(gr + (ptrdiff_t)(w0 + j) * p.C, w0 + j < mw);
      unpack_pairs<T>(graw, gq[j]);
      fadd2(s[9][0], gq[j][0]);
      fadd2(s[9][1], gq[j][1]);
    }
    Raw cur[kRun + 2];
    dw_load_row<T>(xr - rs, p.C, w0, mw, rok[0], cur);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      f32x2 val[kRun + 2][2];
#pragma unroll
      for (int jj = 0; jj < kRun + 2; ++jj) unpack_pairs<T>(cur[jj], val[jj]);
      if (r < 2) dw_load_row<T>(xr + (r == 0 ? 0 : rs), p.C, w0, mw, rok[r + 1], cur);   // the next row is in flight during the FMAs
#pragma unroll
      for (int j = 0; j < kRun; ++j) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          ffma2(s[r * 3 + dx][0], gq[j][0], val[j + dx][0]);
          ffma2(s[r * 3 + dx][1], gq[j][1], val[j + dx][1]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    float f[4];
    upk2(s[k][0], f[0], f[1]);
    upk2(s[k][1], f[2], f[3]);
#pragma unroll
    for (int v = 0; v < 4; ++v) red[(ty * 40 + k * 4 + v) * cvec + cv] = f[v];
  }
  __syncthreads();
  float* mine = partial + (size_t)blockIdx.x * 40 * cvec;
  for (int i = threadIdx.x; i < 40 * cvec; i += blockDim.x) {
    float tot = 0.f;
    for (int yy = 0; yy < ty_count; ++yy) tot += red[yy * 40 * cvec + i];
    mine[i] = tot;
  }
}

// second stage: sum the per-CTA partial rows; i = (tap * 4 + v) * cvec + cv -> grad_weight[(cv*4+v)*9 + tap] / grad_bias.
// block = (32 columns, 32 row slices), four independent partial sums per thread
__global__ void __launch_bounds__(1024) adapter_dwconv_wgrad_sum_kernel(const float* __restrict__ partial, int rows, int cvec,
                                                                        float* __restrict__ gw, float* __restrict__ gb) {
  __shared__ float red[32][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int n = 40 * cvec;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  if (i < n) {
    int r = threadIdx.y;
    for (; r + 96 < rows; r += 128) {
      t0 += partial[(size_t)r * n + i];
      t1 += partial[(size_t)(r + 32) * n + i];
      t2 += partial[(size_t)(r + 64) * n + i];
      t3 += partial[(size_t)(r + 96) * n + i];
    }
    for (; r < rows; r += 32) t0 += partial[(size_t)r * n + i];
  }
  red[threadIdx.y][threadIdx.x] = (t0 + t1) + (t2 + t3);
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float tot = 0.f;
#pragma unroll
    for (int yy = 0; yy < 32; ++yy) tot += red[yy][threadIdx.x];
    const int kv = i / cvec, cvi = i - kv * cvec;
    const int k = kv >> 2, ch = cvi * 4 + (kv & 3);
    if (k < 9) gw[ch * 9 + k] = tot; else gb[ch] = tot;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Tile kernels (round 2). The run kernels above keep their loads in registers, so the bytes a thread has in flight are
// bounded by its register file: 16 warps x 20 loads x 8 bytes per lane in bf16 - half of what fp32 gets with the same
// code, which is why bf16 ran in the same TIME as fp32 (0.35 vs 0.61 of the HBM roofline) and why packing the FMAs
// (FFMA2) changed nothing. Here a CTA owns (batch, map, band of output rows, block of channels) and brings the band's
// input rows (+1 halo row above and below, +1 zero token left and right) into shared memory with cp.async - all of
// it in flight at once, no registers held - then every thread computes runs of 4 output tokens for its channel quad
// from shared memory (18 LDS per 4 outputs) with the taps in registers.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kTileRun = 4;

struct DwTileGeom {
  int rb[3];        // output rows per band, per map
  int bands[3];     // bands per map
  int nblk;         // channel blocks (C / (4 * CBQ))
  int items[3];     // bands * nblk per map (per batch)
  int per_batch, total;
  unsigned row_bytes_max;  // bytes of one padded tile row of the widest map
  unsigned smem;           // dynamic shared memory: (max rb + 2) rows of the widest configuration
};

template <typename T>
static bool dw_tile_geom(const DwParams& p, int cbq, DwTileGeom* g) {
  const int CB = 4 * cbq;
  if (p.C % CB != 0) return false;
  const int mh[3] = {2 * p.H, p.H, p.H / 2}, mw[3] = {2 * p.W, p.W, p.W / 2};
  g->nblk = p.C / CB;
  g->per_batch = 0;
  g->smem = 0;
  g->row_bytes_max = 0;
  for (int i = 0; i < 3; ++i) {
    if (mw[i] % kTileRun != 0 || mw[i] < kTileRun) return false;   // whole runs only (the adapter's maps are multiples of 4 wide)
    const unsigned row_bytes = (unsigned)(mw[i] + 2) * CB * sizeof(T);
    // band height: as many rows as keep the tile at ~70 KB (3 resident CTAs per SM, which is also what the registers
    // allow), at least 2, at most the map
    int rb = (int)(70u * 1024u / row_bytes) - 2;
    if (rb < 2) rb = 2;
    if (rb > 8) rb = 8;
    if (rb > mh[i]) rb = mh[i];
    g->rb[i] = rb;
    g->bands[i] = (mh[i] + rb - 1) / rb;
    g->items[i] = g->bands[i] * g->nblk;
    g->per_batch += g->items[i];
    const unsigned need = (unsigned)(rb + 2) * row_bytes + (unsigned)CB * 9u * (unsigned)sizeof(T) + 16u;   // + the block's taps
    if (need > g->smem) g->smem = need;
    if (row_bytes > g->row_bytes_max) g->row_bytes_max = row_bytes;
  }
  g->total = g->per_batch * p.B;
  return g->smem <= 200u * 1024u;
}

template <typename T, bool FLIP, int CBQ>
__global__ void __launch_bounds__(256, 3) adapter_dwconv_tile_kernel(const DwParams p, const DwTileGeom g) {
  using Raw = typename Quad<T>::Raw;
  constexpr int CB = 4 * CBQ;                       // channels per block
  constexpr unsigned kTokB = CB * sizeof(T);        // bytes of one token of the block
  constexpr int kVecPerTok = kTokB / 16;            // 16-byte chunks per token
  extern __shared__ __align__(16) unsigned char dw_tile[];
  const int tid = threadIdx.x;

  // ---- work item -> (batch, map, band, channel block) ----------------------------------------------------------------
  int item = blockIdx.x;
  const int b = item / g.per_batch;
  item -= b * g.per_batch;
  int mi = 0;
  if (item >= g.items[0]) { item -= g.items[0]; mi = 1; }
  if (mi == 1 && item >= g.items[1]) { item -= g.items[1]; mi = 2; }
  const int mh = mi == 0 ? 2 * p.H : mi == 1 ? p.H : p.H / 2;
  const int mw = mi == 0 ? 2 * p.W : mi == 1 ? p.W : p.W / 2;
  const int t0 = mi == 0 ? 0 : mi == 1 ? 4 * p.H * p.W : 5 * p.H * p.W;
  const int band = item / g.nblk, cblk = item - band * g.nblk;
  const int rb = g.rb[mi];
  const int h0 = band * rb;
  const int nrow = min(rb, mh - h0);                // output rows of this band
  const unsigned row_bytes = (unsigned)(mw + 2) * kTokB;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(dw_tile);
  const size_t map_tok = (size_t)b * p.Ntok + t0;
  const T* __restrict__ xin = reinterpret_cast<const T*>(p.x) + map_tok * p.C + cblk * CB;
  T* __restrict__ yout = reinterpret_cast<T*>(p.y) + map_tok * p.C + cblk * CB;

  // ---- stage rows h0-1 .. h0+nrow of the map: interior by cp.async, everything outside the map is zero. Two commit
  //      groups: the rows the first half of the output rows needs, then the rest - the first half is computed while the
  //      second is still on its way. The block's 9 x CB taps (contiguous in the [C,1,3,3] weight) ride in the first group.
  const int nrowA = (nrow + 1) / 2;                 // output rows of the first half: needs tile rows 0 .. nrowA+1
  const unsigned taps_s = sbase + (unsigned)(rb + 2) * row_bytes;
  {
    constexpr int kTapChunks = (CB * 9 * (int)sizeof(T)) / 16;   // CB in {32, 64}: a whole number of 16-byte chunks
    const char* wsrc = reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.w) + (size_t)cblk * CB * 9);
    for (int i = tid; i < kTapChunks; i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(taps_s + i * 16), "l"(wsrc + i * 16) : "memory");
  }
  for (int r = 0; r < nrow + 2; ++r) {
    if (r == nrowA + 2) asm volatile("cp.async.commit_group;" ::: "memory");
    const int h = h0 - 1 + r;
    const unsigned srow = sbase + r * row_bytes;
    if (h >= 0 && h < mh) {
      // thread -> (16-byte chunk c of a token, tokens wq, wq + 256 / kVecPerTok, ...): constant strides, no division in the loop
      const int c = tid % kVecPerTok, wq = tid / kVecPerTok;
      const char* src = reinterpret_cast<const char*>(xin + (size_t)h * mw * p.C) + (size_t)wq * p.C * sizeof(T) + c * 16;
      unsigned dst = srow + (unsigned)(wq + 1) * kTokB + c * 16;
      const size_t sstep = (size_t)(256 / kVecPerTok) * p.C * sizeof(T);
      for (int w = wq; w < mw; w += 256 / kVecPerTok, src += sstep, dst += (256 / kVecPerTok) * kTokB)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      if (tid < 2 * kVecPerTok) {   // the zero token left of column 0 and right of column mw-1
        const unsigned dst = srow + (tid < kVecPerTok ? 0u : (unsigned)(mw + 1) * kTokB) + (tid % kVecPerTok) * 16;
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
      }
    } else {
      for (unsigned o = tid * 16u; o < row_bytes; o += 256u * 16u)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(srow + o), "r"(0u) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const bool two_groups = (nrow + 2) > (nrowA + 2);

  // ---- this thread's channel quad: taps and bias in registers (packed pairs) --------------------------------------------
  const int quad = tid % CBQ;
  if (two_groups) asm volatile("cp.async.wait_group 1;" ::: "memory");
  else asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  f32x2 tp[9][2], bias[2];
  {
    float tf[9][4], bf[4];
    const T* ws = reinterpret_cast<const T*>(dw_tile + (size_t)(rb + 2) * row_bytes) + quad * 36;   // 4 channels x 9 taps, contiguous
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
      for (int k = 0; k < 9; ++k) tf[FLIP ? 8 - k : k][v] = dw_ld<T>(ws + v * 9 + k);
      bf[v] = (!FLIP && p.bias) ? dw_ld<T>(reinterpret_cast<const T*>(p.bias) + cblk * CB + quad * 4 + v) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) { tp[k][0] = pk2(tf[k][0], tf[k][1]); tp[k][1] = pk2(tf[k][2], tf[k][3]); }
    bias[0] = pk2(bf[0], bf[1]);
    bias[1] = pk2(bf[2], bf[3]);
  }

  // ---- compute: unit = (output row, run of 4 tokens, quad); quads fastest so a warp reads whole tokens -----------------
  const int nrun = mw / kTileRun;
  const unsigned qoff = (unsigned)quad * (4u * sizeof(T));
  for (int half = 0; half < 2; ++half) {
    const int r_begin = half == 0 ? 0 : nrowA, r_end = half == 0 ? nrowA : nrow;
    if (half == 1) {
      if (r_begin >= r_end) break;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
    const int units = (r_end - r_begin) * nrun * CBQ;
    for (int u = tid; u < units; u += 256) {
      const int rr = r_begin + u / (nrun * CBQ);
      const int run = (u % (nrun * CBQ)) / CBQ;
      const int w0 = run * kTileRun;
      f32x2 acc[kTileRun][2];
#pragma unroll
      for (int j = 0; j < kTileRun; ++j) { acc[j][0] = bias[0]; acc[j][1] = bias[1]; }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const unsigned srow = sbase + (unsigned)(rr + r) * row_bytes + (unsigned)w0 * kTokB + qoff;   // padded column w0 = map column w0 - 1
#pragma unroll
        for (int jj = 0; jj < kTileRun + 2; ++jj) {
          Raw raw;
          if constexpr (sizeof(T) == 4) {
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(raw.x), "=f"(raw.y), "=f"(raw.z), "=f"(raw.w) : "r"(srow + jj * kTokB));
          } else {
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(srow + jj * kTokB));
          }
          f32x2 val[2];
          unpack_pairs<T>(raw, val);
#pragma unroll
          for (int dx = 2; dx >= 0; --dx) {
            const int j = jj - dx;
            if (j < 0 || j >= kTileRun) continue;
            ffma2(acc[j][0], val[0], tp[r * 3 + dx][0]);
            ffma2(acc[j][1], val[1], tp[r * 3 + dx][1]);
          }
        }
      }
      T* dst = yout + ((size_t)(h0 + rr) * mw + w0) * p.C + quad * 4;
#pragma unroll
      for (int j = 0; j < kTileRun; ++j) {
        float o[4];
        upk2(acc[j][0], o[0], o[1]);
        upk2(acc[j][1], o[2], o[3]);
        Quad<T>::st(dst + (size_t)j * p.C, o);
      }
    }
  }
}

template <typename T, bool FLIP>
static cudaError_t launch_dw_tile(const DwParams& p, cudaStream_t s, bool* taken) {
  *taken = false;
  if (p.H % 2 != 0 || p.W % 2 != 0) return cudaSuccess;
  const uintptr_t al = 15;
  if ((reinterpret_cast<uintptr_t>(p.x) & al) || (reinterpret_cast<uintptr_t>(p.y) & al) || (reinterpret_cast<uintptr_t>(p.w) & al)) return cudaSuccess;
  DwTileGeom g;
  int cbq = 0;
  // 16 bytes per cp.async: a token of the block must be a multiple of 16 bytes -> CB * sizeof(T) % 16 == 0
  if (p.C % 64 == 0 && dw_tile_geom<T>(p, 16, &g)) cbq = 16;
  else if (p.C % 32 == 0 && dw_tile_geom<T>(p, 8, &g)) cbq = 8;
  if (!cbq || (size_t)p.C * sizeof(T) % 16 != 0) return cudaSuccess;
  *taken = true;
  cudaError_t e;
  if (cbq == 16) {
    auto k = adapter_dwconv_tile_kernel<T, FLIP, 16>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)g.total, 256, g.smem, s>>>(p, g);
  } else {
    auto k = adapter_dwconv_tile_kernel<T, FLIP, 8>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    k<<<(unsigned)g.total, 256, g.smem, s>>>(p, g);
  }
  return cudaGetLastError();
}

// fast-path eligibility + launch shape: blockDim = cvec * ty_count <= 256
template <typename T>
static bool dw_run_shape(const DwParams& p, const void* extra, int& cvec, int& ty_count) {
  if (p.C % 4 != 0 || p.C / 4 > 256) return false;
  const uintptr_t al = 4 * sizeof(T) - 1;
  if ((reinterpret_cast<uintptr_t>(p.x) & al) || (reinterpret_cast<uintptr_t>(p.y) & al) || (reinterpret_cast<uintptr_t>(extra) & al))
    return false;
  cvec = p.C / 4;
  ty_count = 256 / cvec;
  return true;
}

static unsigned dw_run_grid(const DwRunGeom& g, int ty_count) {
  const unsigned need = (unsigned)((g.total + ty_count - 1) / ty_count);
  const unsigned resident = 148u * 2u;
  return need < resident ? (need ? need : 1u) : resident;
}

template <typename T, bool FLIP>
static cudaError_t launch_dw(const DwParams& p, cudaStream_t s) {
  using A = typename DwAcc<T>::type;
  if constexpr (sizeof(T) <= 4) {
    if constexpr (sizeof(T) == 2) {   // fp32 measured faster on the run kernel (bytes in flight are not its limit)
      bool taken = false;
      const cudaError_t et = launch_dw_tile<T, FLIP>(p, s, &taken);
      if (taken) return et;
    }
    int cvec, ty_count;
    if (dw_run_shape<T>(p, nullptr, cvec, ty_count)) {
      const DwRunGeom g = dw_run_geom(p);
      adapter_dwconv_run_kernel<T, FLIP><<<dw_run_grid(g, ty_count), cvec * ty_count, 0, s>>>(p, g, cvec, ty_count);
      return cudaGetLastError();
    }
  }
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
    case MSDA_F16: return flip ? launch_dw<__half, true>(p, s) : launch_dw<__half, false>(p, s);
    case MSDA_F64: return flip ? launch_dw<double, true>(p, s) : launch_dw<double, false>(p, s);
    default: return cudaErrorInvalidValue;
  }
}

// bytes of the per-CTA partial rows the deterministic two-stage reduction needs (0: the generic atomic path is used)
size_t dwconv_wgrad_workspace_bytes(const DwParams& p, int dtype) {
  if ((dtype != MSDA_F32 && dtype != MSDA_BF16 && dtype != MSDA_F16) || p.C % 4 != 0 || p.C / 4 > 256) return 0;
  const int cvec = p.C / 4;
  return (size_t)dw_run_grid(dw_run_geom(p), 256 / cvec) * 40 * cvec * sizeof(float);
}

// grad_weight / grad_bias accumulators: fp32 for f32/bf16 inputs, fp64 for f64 (both [C*9] / [C]).
cudaError_t launch_dwconv_wgrad(const DwParams& p, int dtype, const void* grad_y, void* gw, void* gb, void* workspace,
                                size_t workspace_bytes, int* launches, cudaStream_t s) {
  const size_t asz = dtype == MSDA_F64 ? 8 : 4;
  if (dtype == MSDA_F32 || dtype == MSDA_BF16 || dtype == MSDA_F16) {
    int cvec, ty_count;
    const bool ok = dtype == MSDA_F32 ? dw_run_shape<float>(p, grad_y, cvec, ty_count) : dw_run_shape<__nv_bfloat16>(p, grad_y, cvec, ty_count);
    if (ok && workspace && workspace_bytes >= dwconv_wgrad_workspace_bytes(p, dtype)) {
      const DwRunGeom g = dw_run_geom(p);
      const unsigned grid = dw_run_grid(g, ty_count);
      const size_t smem = (size_t)ty_count * 40 * cvec * sizeof(float);  // <= 256 * 40 * 4 = 40 KB
      float* partial = reinterpret_cast<float*>(workspace);
      if (dtype == MSDA_F32)
        adapter_dwconv_wgrad_run_kernel<float><<<grid, cvec * ty_count, smem, s>>>(p, g, grad_y, partial, cvec, ty_count);
      else if (dtype == MSDA_BF16)
        adapter_dwconv_wgrad_run_kernel<__nv_bfloat16><<<grid, cvec * ty_count, smem, s>>>(p, g, grad_y, partial, cvec, ty_count);
      else
        adapter_dwconv_wgrad_run_kernel<__half><<<grid, cvec * ty_count, smem, s>>>(p, g, grad_y, partial, cvec, ty_count);
      cudaError_t e2 = cudaGetLastError();
      if (e2 != cudaSuccess) return e2;
      adapter_dwconv_wgrad_sum_kernel<<<(40 * cvec + 31) / 32, dim3(32, 32), 0, s>>>(partial, (int)grid, cvec, (float*)gw, (float*)gb);
      *launches = 2;
      return cudaGetLastError();
    }
  }
  *launches = 1;
  cudaError_t e = cudaMemsetAsync(gw, 0, (size_t)p.C * 9 * asz, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(gb, 0, (size_t)p.C * asz, s);
  if (e != cudaSuccess) return e;
  const dim3 block(32, 8);
  const size_t bt_total = (size_t)p.B * p.Ntok;
  int ctas_x = device_sm_count() * 2;
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
    case MSDA_F16:
      adapter_dwconv_wgrad_kernel<__half><<<grid, block, smem, s>>>(p, grad_y, nullptr, (float*)gw, nullptr, (float*)gb, tokens_per_cta);
      break;
    case MSDA_F64:
      adapter_dwconv_wgrad_kernel<double><<<grid, block, smem, s>>>(p, grad_y, (double*)gw, nullptr, (double*)gb, nullptr, tokens_per_cta);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace msda
