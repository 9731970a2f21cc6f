// mb_tile.cu — can a tiled backward beat the L2 RED ceiling? Throughput of WARP-PRIVATE shared-memory accumulation:
// each warp owns a region of R rows x 128 B; per iteration each of its four 8-lane groups does the four corner updates of
// one sampling point: 4 x (LDS.128, 4 FADD, STS.128) on random rows of the region (+ optionally 4 LDS.128 "value" reads
// from a CTA-shared read-only tile). WARPS warps per CTA, one CTA per SM. Reports SM cycles per corner row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_tile mb_tile.cu && ./mb_tile
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>  // 0: RMW only; 1: RMW + value LDS; 2: RMW + value LDS + conflict check (match_any on the row index)
__global__ void tile_kernel(int rows_per_warp, int value_rows, int iters, const int* __restrict__ rnd, float* out, long long* cycles) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  float* vtile = sm;                                            // [value_rows][32] shared by the CTA, read-only
  float* gtile = sm + value_rows * 32 + warp * rows_per_warp * 32;  // this warp's private accumulation region
  for (int i = threadIdx.x; i < (value_rows + nwarps * rows_per_warp) * 32; i += blockDim.x) sm[i] = 0.25f;
  __syncthreads();
  const int* r = rnd + ((blockIdx.x * nwarps + warp) * 4 + grp) * 64;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int base = r[it & 63];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // four corners of one point: distinct rows (base, base+1, base+W, base+W+1 pattern)
      const int row = (base + (c & 1) + (c >> 1) * 12) & (rows_per_warp - 1);  // rows_per_warp is a power of two
      float4 w = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
      if (MODE >= 1) {
        const float4 v = *reinterpret_cast<const float4*>(vtile + ((base * 7 + c * 13) & (value_rows - 1)) * 32 + j * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        w.x += v.x;
      }
      bool go = true;
      if (MODE >= 2) {
        const unsigned m = __match_any_sync(0xffffffffu, row);
        go = (__ffs(m) - 1) / 8 == grp || true;  // measure the cost of the match; conflicts resolved by a (rare) retry loop in a real kernel
        acc.x += (float)(m & 1);
      }
      if (go) {
        float4* p = reinterpret_cast<float4*>(gtile + row * 32 + j * 4);
        float4 x = *p;
        x.x += w.x; x.y += w.y; x.z += w.z; x.w += w.w;
        *p = x;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + gtile[lane];
}

int main() {
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  float* out; long long* cyc; int* rnd;
  CHECK(cudaMalloc(&out, nsm * 1024 * 4)); CHECK(cudaMalloc(&cyc, 8));
  const int nr = nsm * 32 * 4 * 64;
  int* h = (int*)malloc(nr * 4);
  unsigned s = 12345u;
  for (int i = 0; i < nr; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) & 0xffff; }
  CHECK(cudaMalloc(&rnd, nr * 4)); CHECK(cudaMemcpy(rnd, h, nr * 4, cudaMemcpyHostToDevice));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\",\n", prop.name);
  struct Cfg { int warps, rows, vrows; } cfgs[] = {{4, 128, 256}, {6, 256, 256}, {8, 128, 256}, {12, 128, 256}, {16, 64, 256}, {24, 64, 64}, {32, 32, 256}};
  for (auto c : cfgs) {
    const size_t smem = (size_t)(c.vrows + c.warps * c.rows) * 128;
    if (smem > 227 * 1024) continue;
    long long hc[3];
#define RUN(M) CHECK(cudaFuncSetAttribute(tile_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    for (int rep = 0; rep < 2; ++rep) tile_kernel<M><<<nsm, c.warps * 32, smem>>>(c.rows, c.vrows, iters, rnd, out, cyc); \
    CHECK(cudaDeviceSynchronize()); CHECK(cudaMemcpy(&hc[M], cyc, 8, cudaMemcpyDeviceToHost));
    RUN(0) RUN(1) RUN(2)
    // rows processed per SM = warps * iters * 4 groups * 4 corners
    const double rows = (double)c.warps * iters * 16;
    printf(" \"warps%d_rows%d_smem%zuKB\": {\"rmw_cyc_per_row\": %.3f, \"rmw_plus_value_lds\": %.3f, \"plus_match_any\": %.3f},\n", c.warps, c.rows,
           smem >> 10, hc[0] / rows, hc[1] / rows, hc[2] / rows);
  }
  printf(" \"done\": 1}\n");
  return 0;
}
