// mb_tex.cu — can the texture path add row-gather bandwidth on top of the LSU path? (B200)
// Pattern of the MSDeformAttn forward: every 8-lane group reads one random 128-byte row (16 B per lane).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// mode 0: all rows via LDG.128   mode 1: all rows via tex1Dfetch<float4>   mode 2: alternate LDG / TEX per row
template <int MODE>
__global__ void gather(const float* __restrict__ table, cudaTextureObject_t tex, int nrows_mask, const int* __restrict__ rowidx,
                       float* out, int iters) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 8
  for (int i = 0; i < iters; ++i) {
    const int row = __ldg(&rowidx[((gw * 4 + grp) * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
    float4 x;
    if (MODE == 0 || (MODE == 2 && (i & 1) == 0)) {
      x = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * 32 + j * 4));
    } else {
      x = tex1Dfetch<float4>(tex, row * 8 + j);
    }
    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  float *table, *out; int* rowidx;
  const size_t tbytes = 64u << 20;  // 64 MB table: 524288 rows
  CHECK(cudaMalloc(&table, tbytes)); CHECK(cudaMemset(table, 0, tbytes));
  CHECK(cudaMalloc(&out, sizeof(float) * nsm * 8 * 256)); CHECK(cudaMalloc(&rowidx, 4 << 20));
  { int* h = (int*)malloc(4 << 20); unsigned s = 12345; for (int i = 0; i < (1 << 20); ++i) { s = s * 1664525u + 1013904223u; h[i] = (int)(s >> 8); }
    CHECK(cudaMemcpy(rowidx, h, 4 << 20, cudaMemcpyHostToDevice)); free(h); }
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = table;
  rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = tbytes;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = cudaAddressModeClamp;
  cudaTextureObject_t tex; CHECK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  printf("{\"gpu\": \"%s\",\n", prop.name);
  struct { const char* n; int mask; } tabs[2] = {{"l1_64KB", 511}, {"l2_32MB", (1 << 18) - 1}};
  const char* mnames[3] = {"ldg128", "tex1dfetch_f4", "alternate_ldg_tex"};
  for (int t = 0; t < 2; ++t) for (int mode = 0; mode < 3; ++mode) {
    const int iters = 1024, ctas = nsm * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1) cudaEventRecord(e0);
      if (mode == 0) gather<0><<<ctas, 256>>>(table, tex, tabs[t].mask, rowidx, out, iters);
      if (mode == 1) gather<1><<<ctas, 256>>>(table, tex, tabs[t].mask, rowidx, out, iters);
      if (mode == 2) gather<2><<<ctas, 256>>>(table, tex, tabs[t].mask, rowidx, out, iters);
    }
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)ctas * 8 * iters * 4;
    printf(" \"gather_%s_%s\": {\"ms\": %.4f, \"rows_per_ns\": %.2f},\n", mnames[mode], tabs[t].n, ms, rows / (ms * 1e6));
  }
  printf(" \"done\": 1}\n");
  return 0;
}
