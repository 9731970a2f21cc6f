// msda_abi.cu — extern "C" boundary declared in include/msda_b200.h.
//
// Host-side role of the reference's ms_deform_attn_cuda.cu:20-153 (argument checks, dimension
// extraction, launches on the caller's stream) without any ATen type: raw pointers in, error code
// out. No allocation, no synchronisation, no global mutable state besides two tuning knobs and a
// launch counter (all atomics).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "msda_common.cuh"

namespace msda {
bool vec_supported(int dtype, int D, int* G_out);
cudaError_t launch_forward(const Params& p, int dtype, bool vec_ok, int G, int minb, cudaStream_t s);
cudaError_t launch_backward(const Params& p, int dtype, bool vec_ok, int G, int minb, cudaStream_t s);
cudaError_t launch_cvt_f32_bf16(const float* src, void* dst, size_t n, cudaStream_t s);

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_qc_fwd{0}, g_qc_bwd{0}, g_minb_fwd{0}, g_minb_bwd{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

static size_t elem_size(int dtype) {
  return dtype == MSDA_F32 ? 4 : dtype == MSDA_BF16 ? 2 : dtype == MSDA_F64 ? 8 : 0;
}

static int check_dims(const msda_dims* d, int dtype) {
  if (!d) return fail(MSDA_E_NULL, "dims is NULL");
  if (elem_size(dtype) == 0) return fail(MSDA_E_DTYPE, "unknown dtype %d", dtype);
  if (d->batch <= 0 || d->spatial_size <= 0 || d->num_heads <= 0 || d->channels <= 0 ||
      d->num_levels <= 0 || d->num_query <= 0 || d->num_point <= 0)
    return fail(MSDA_E_DIMS, "non-positive dimension (N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d)", d->batch,
                d->spatial_size, d->num_heads, d->channels, d->num_levels, d->num_query, d->num_point);
  if (d->num_levels > MSDA_MAX_LEVELS)
    return fail(MSDA_E_LEVELS, "num_levels=%d exceeds MSDA_MAX_LEVELS=%d", d->num_levels, MSDA_MAX_LEVELS);
  // in-image byte offsets are uint32 in the kernels (the reference is int32 element offsets throughout);
  // the token index shares a register with 4 flag bits.
  const int64_t per_image = (int64_t)d->spatial_size * d->num_heads * d->channels;
  const int64_t per_pair = (int64_t)d->num_levels * d->num_point;
  if (per_image >= (1ll << 29) || per_pair >= (1ll << 20) || d->spatial_size >= (1 << 27) ||
      (int64_t)d->num_heads * d->channels >= (1ll << 24))
    return fail(MSDA_E_DIMS, "S*M*D=%lld does not fit the 32-bit in-image byte offset", (long long)per_image);
  return 0;
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// queries per CTA chunk: enough CTAs to fill 148 SMs x ~8 resident CTAs, chunks as long as possible
// otherwise (longer chunk = more L1 reuse of the value rows around neighbouring queries).
static int pick_chunk(const msda_dims* d, int per_iter, int override_qc) {
  int qc;
  if (override_qc > 0) {
    qc = override_qc;
  } else {
    const int64_t bm = (int64_t)d->batch * d->num_heads;
    const int64_t target_ctas = 148 * 8;
    int64_t chunks = (target_ctas + bm - 1) / bm;
    if (chunks < 1) chunks = 1;
    qc = (int)((d->num_query + chunks - 1) / chunks);
    if (qc > 256) qc = 256;
  }
  qc = ((qc + per_iter - 1) / per_iter) * per_iter;
  if (qc < per_iter) qc = per_iter;
  return qc;
}

static void fill_params(Params& p, const msda_dims* d) {
  memset(&p, 0, sizeof(p));
  p.N = d->batch; p.S = d->spatial_size; p.M = d->num_heads; p.D = d->channels;
  p.L = d->num_levels; p.Lq = d->num_query; p.P = d->num_point;
}

// ---- debug: per-point index dump (same point_geom the kernels use) -----------------------------
__global__ void msda_debug_index_kernel(const int64_t* shapes, const int64_t* lsi, const float* loc,
                                        int32_t* idx, int N, int M, int D, int L, int Lq, int P) {
  const size_t total = (size_t)N * Lq * M * L * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int l = (int)((i / P) % L);
    const int m = (int)((i / ((size_t)P * L)) % M);
    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
    const PointGeom<float> g = point_geom<float>(loc[2 * i], loc[2 * i + 1], H, W);
    idx[4 * i + 0] = g.h_low;
    idx[4 * i + 1] = g.w_low;
    idx[4 * i + 2] = (int)g.mask;
    idx[4 * i + 3] = ((int)lsi[l] + g.h_low * W + g.w_low) * (M * D) + m * D;
  }
}
}  // namespace msda

using namespace msda;

extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

const char* msda_last_error(void) { return g_err; }

uint64_t msda_launch_count(void) { return g_launches.load(); }

void msda_set_tuning(int32_t fwd_chunk, int32_t bwd_chunk, int32_t fwd_min_ctas, int32_t bwd_min_ctas) {
  g_qc_fwd.store(fwd_chunk);
  g_qc_bwd.store(bwd_chunk);
  g_minb_fwd.store(fwd_min_ctas);
  g_minb_bwd.store(bwd_min_ctas);
}

int msda_check_im2col_step(int32_t batch, int32_t im2col_step) {
  if (batch <= 0 || im2col_step <= 0) return fail(MSDA_E_STEP, "batch(%d) / im2col_step(%d) must be positive", batch, im2col_step);
  const int step = batch < im2col_step ? batch : im2col_step;
  if (batch % step != 0) return fail(MSDA_E_STEP, "batch(%d) must divide im2col_step(%d)", batch, step);
  return 0;
}

int msda_forward(const msda_dims* dims, int dtype, const void* value, const int64_t* spatial_shapes,
                 const int64_t* level_start_index, const void* sampling_loc, const void* attn_weight,
                 void* out, void* stream) {
  if (int e = check_dims(dims, dtype)) return e;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !out)
    return fail(MSDA_E_NULL, "msda_forward: NULL tensor pointer");
  const size_t es = elem_size(dtype), fs = dtype == MSDA_F64 ? 8 : 4;
  if (!aligned(value, es) || !aligned(out, es) || !aligned(sampling_loc, fs) || !aligned(attn_weight, fs) ||
      !aligned(spatial_shapes, 8) || !aligned(level_start_index, 8))
    return fail(MSDA_E_ALIGN, "msda_forward: pointer not aligned to its element size");

  Params p;
  fill_params(p, dims);
  p.value = value; p.shapes = spatial_shapes; p.lsi = level_start_index;
  p.loc = sampling_loc; p.aw = attn_weight; p.out = out;

  int G = 0;
  bool vec = vec_supported(dtype, p.D, &G) && aligned(value, 16) && aligned(out, 16) &&
             aligned(sampling_loc, 8);
  const int per_iter = vec ? kWarps * (32 / G) : kWarps;
  p.qc = pick_chunk(dims, per_iter, g_qc_fwd.load());
  p.nchunk = (p.Lq + p.qc - 1) / p.qc;

  const cudaError_t e = launch_forward(p, dtype, vec, G, g_minb_fwd.load(), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "msda_forward launch");
  g_launches.fetch_add(1);
  return 0;
}

size_t msda_backward_workspace_bytes(const msda_dims* dims, int dtype) {
  if (!dims || dtype != MSDA_BF16) return 0;
  return (size_t)dims->batch * dims->spatial_size * dims->num_heads * dims->channels * sizeof(float);
}

int msda_backward(const msda_dims* dims, int dtype, const void* value, const int64_t* spatial_shapes,
                  const int64_t* level_start_index, const void* sampling_loc, const void* attn_weight,
                  const void* grad_out, void* grad_value, void* grad_sampling_loc, void* grad_attn_weight,
                  void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_dims(dims, dtype)) return e;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_out ||
      !grad_value || !grad_sampling_loc || !grad_attn_weight)
    return fail(MSDA_E_NULL, "msda_backward: NULL tensor pointer");
  const size_t es = elem_size(dtype), fs = dtype == MSDA_F64 ? 8 : 4;
  if (!aligned(value, es) || !aligned(grad_out, es) || !aligned(grad_value, es) ||
      !aligned(sampling_loc, fs) || !aligned(attn_weight, fs) || !aligned(grad_sampling_loc, fs) ||
      !aligned(grad_attn_weight, fs) || !aligned(spatial_shapes, 8) || !aligned(level_start_index, 8))
    return fail(MSDA_E_ALIGN, "msda_backward: pointer not aligned to its element size");
  const size_t need = msda_backward_workspace_bytes(dims, dtype);
  if (need > 0 && (!workspace || workspace_bytes < need || !aligned(workspace, 16)))
    return fail(MSDA_E_WORKSPACE, "msda_backward: workspace of %zu bytes (16-byte aligned) required, got %zu",
                need, workspace ? workspace_bytes : (size_t)0);

  cudaStream_t s = (cudaStream_t)stream;
  Params p;
  fill_params(p, dims);
  p.value = value; p.shapes = spatial_shapes; p.lsi = level_start_index;
  p.loc = sampling_loc; p.aw = attn_weight; p.grad_out = grad_out;
  p.grad_loc = grad_sampling_loc; p.grad_aw = grad_attn_weight;

  const size_t nvalue = (size_t)p.N * p.S * p.M * p.D;
  void* accum = (dtype == MSDA_BF16) ? workspace : grad_value;
  const size_t accum_bytes = (dtype == MSDA_BF16) ? nvalue * sizeof(float) : nvalue * es;
  p.grad_value = accum;

  int G = 0;
  bool vec = vec_supported(dtype, p.D, &G) && aligned(value, 16) && aligned(grad_out, 16) &&
             aligned(accum, 16) && aligned(sampling_loc, 8) && aligned(grad_sampling_loc, 8);
  const int per_iter = vec ? kWarps * (32 / G) : kWarps;
  p.qc = pick_chunk(dims, per_iter, g_qc_bwd.load());
  p.nchunk = (p.Lq + p.qc - 1) / p.qc;

  cudaError_t e = cudaMemsetAsync(accum, 0, accum_bytes, s);
  if (e != cudaSuccess) return cuda_fail(e, "msda_backward memset(grad_value)");
  e = launch_backward(p, dtype, vec, G, g_minb_bwd.load(), s);
  if (e != cudaSuccess) return cuda_fail(e, "msda_backward launch");
  g_launches.fetch_add(1);
  if (dtype == MSDA_BF16) {
    e = launch_cvt_f32_bf16(reinterpret_cast<const float*>(accum), grad_value, nvalue, s);
    if (e != cudaSuccess) return cuda_fail(e, "msda_backward bf16 convert launch");
    g_launches.fetch_add(1);
  }
  return 0;
}

int msda_debug_point_index(const msda_dims* dims, const int64_t* spatial_shapes,
                           const int64_t* level_start_index, const float* sampling_loc, int32_t* idx,
                           void* stream) {
  if (int e = check_dims(dims, MSDA_F32)) return e;
  if (!spatial_shapes || !level_start_index || !sampling_loc || !idx)
    return fail(MSDA_E_NULL, "msda_debug_point_index: NULL pointer");
  const size_t total = (size_t)dims->batch * dims->num_query * dims->num_heads * dims->num_levels * dims->num_point;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  msda_debug_index_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      spatial_shapes, level_start_index, sampling_loc, idx, dims->batch, dims->num_heads, dims->channels,
      dims->num_levels, dims->num_query, dims->num_point);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "msda_debug_point_index launch");
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
