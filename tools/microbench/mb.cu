// mb.cu — pipe-throughput microbenchmarks that size the MSDeformAttn kernel design on B200.
// Each test runs 1 CTA of 256 threads per SM (148 CTAs), a dependent-free loop of N identical
// instructions per warp, and reports SM cycles per warp-instruction (clock64 on one SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// mode 0: LDS.128, every lane of an 8-lane group reads the SAME 16 B (4 distinct addresses per warp)
// mode 1: LDS.32 broadcast per group (4 distinct words per warp)
// mode 2: SHFL.IDX width 8
// mode 3: LDS.128, each group reads its own 128-B row (8 lanes x 16 B), rows random  (smem value tile gather)
// mode 4: LDS.64 broadcast per group
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ float lds32(unsigned a) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
  return r;
}
__global__ void smem_kernel(int mode, float* out, long long* cycles, const int* rowidx) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  float4 acc = make_float4(0, 0, 0, 0);
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = (float)(lane + k);
  int rows[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) rows[k] = rowidx[(blockIdx.x * 32 + warp * 4 + grp) * 16 + k] & 255;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    const unsigned a0 = sbase + (((warp * 4 + grp) * 100) & 4095) * 4;
    for (int i = 0; i < ITERS / 16; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { const float4 x = lds128(a0 + k * 16); acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
    }
  } else if (mode == 1) {
    const unsigned a0 = sbase + (((warp * 4 + grp) * 101) & 4095) * 4;
    for (int i = 0; i < ITERS / 16; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) acc.x += lds32(a0 + k * 4);
    }
  } else if (mode == 2) {
    for (int i = 0; i < ITERS / 8; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __shfl_sync(0xffffffffu, v[k], (k + i) & 7, 8);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.x += v[k];
  } else if (mode == 3) {
    for (int i = 0; i < ITERS / 16; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { const float4 x = lds128(sbase + rows[k] * 128 + j * 16); acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
    }
  } else if (mode == 4) {
    const unsigned a0 = sbase + (((warp * 4 + grp) * 102) & 4095) * 4;
    for (int i = 0; i < ITERS / 16; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { const float2 x = lds64(a0 + k * 8); acc.x += x.x; acc.y += x.y; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// global gathers served from L1: each 8-lane group reads one 128-B row (LDG.128); rows drawn from a
// small (L1-resident, 64 KB) or medium (L2-resident, 32 MB) table. mode 0: LDG.128 4 rows/warp-instr
// mode 1: LDG.32 one row per warp-instr (32 lanes x 4 B, the reference's pattern)
__global__ void gather_kernel(int mode, const float* __restrict__ table, int nrows_mask, const int* __restrict__ rowidx,
                              float* out, long long* cycles, int iters) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
      const int row = __ldg(&rowidx[((gw * 4 + grp) * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
      const float4 x = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * 32 + j * 4));
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  } else {
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
      const int row = __ldg(&rowidx[(gw * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
      acc.x += __ldg(table + (size_t)row * 32 + lane);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// RED.ADD.F32x4 into an L2-resident table: each 8-lane group adds one 128-B row.
__global__ void red_kernel(float* table, int nrows_mask, const int* __restrict__ rowidx, int iters, int contig) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < iters; ++i) {
    int row = __ldg(&rowidx[((gw * 4 + grp) * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
    if (contig) row = ((gw * 4 + grp) * iters + i) & nrows_mask;
    float* p = table + (size_t)row * 32 + j * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
  }
}

// shared-memory accumulation patterns: each 8-lane group updates one 128-B row (4 floats / ints per lane),
// rows random in a 256-row (32 KB) tile. mode 0: atomicAdd(float) x4 (CAS loop)   mode 1: atomicAdd(int) x4
// mode 2: atomicAdd(unsigned long long) x2 (64-bit)   mode 3: plain LDS.128 + FADD + STS.128 (racy; upper bound)
// mode 4: red.shared.add.f32 via PTX x4
__global__ void smem_atomic_kernel(int mode, float* out, long long* cycles, const int* rowidx) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 0.f;
  int rows[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) rows[k] = rowidx[(blockIdx.x * 32 + warp * 4 + grp) * 16 + k] & 255;
  __syncthreads();
  const int ITER2 = 512;
  long long t0 = clock64();
  for (int i = 0; i < ITER2 / 16; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float* p = sm + rows[k] * 32 + j * 4;
      if (mode == 0) {
        atomicAdd(p + 0, 1.f); atomicAdd(p + 1, 1.f); atomicAdd(p + 2, 1.f); atomicAdd(p + 3, 1.f);
      } else if (mode == 1) {
        int* q = reinterpret_cast<int*>(p);
        atomicAdd(q + 0, lane + k); atomicAdd(q + 1, lane + 2 * k); atomicAdd(q + 2, lane + 3); atomicAdd(q + 3, lane + 4);
      } else if (mode == 2) {
        unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
        atomicAdd(q + 0, 1ull); atomicAdd(q + 1, 1ull);
      } else if (mode == 3) {
        float4 x = *reinterpret_cast<float4*>(p);
        x.x += 1.f; x.y += 1.f; x.z += 1.f; x.w += 1.f;
        *reinterpret_cast<float4*>(p) = x;
      } else {
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(1.f) : "memory");
        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 4), "f"(1.f) : "memory");
        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 8), "f"(1.f) : "memory");
        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 12), "f"(1.f) : "memory");
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x];
}

int main() {
  int dev = 0; cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, dev));
  const int nsm = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, nsm, prop.clockRate);
  float* out; long long* cyc; int* rowidx;
  CHECK(cudaMalloc(&out, sizeof(float) * 4096 * 256)); CHECK(cudaMalloc(&cyc, 8)); CHECK(cudaMalloc(&rowidx, 4 << 20));
  { int* h = (int*)malloc(4 << 20); unsigned s = 12345; for (int i = 0; i < (1 << 20); ++i) { s = s * 1664525u + 1013904223u; h[i] = (int)(s >> 8); }
    CHECK(cudaMemcpy(rowidx, h, 4 << 20, cudaMemcpyHostToDevice)); free(h); }
  long long hc;
  const char* names[5] = {"lds128_group_broadcast", "lds32_group_broadcast", "shfl_idx_w8", "lds128_row_gather", "lds64_group_broadcast"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) smem_kernel<<<nsm, 256, 32768>>>(mode, out, cyc, rowidx);
    CHECK(cudaDeviceSynchronize()); CHECK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf(" \"%s_cycles_per_warp_instr_per_sm\": %.3f,\n", names[mode], (double)hc / (ITERS * 8.0));
  }
  const char* anames[5] = {"smem_atomicadd_f32x4_row", "smem_atomicadd_s32x4_row", "smem_atomicadd_u64x2_row", "smem_plain_rmw128_row", "smem_red_f32x4_row"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) smem_atomic_kernel<<<nsm, 256, 32768>>>(mode, out, cyc, rowidx);
    CHECK(cudaDeviceSynchronize()); CHECK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    printf(" \"%s_cycles_per_4rows_per_sm\": %.3f,\n", anames[mode], (double)hc / (512 * 8.0));
  }
  // gathers
  float* table; const size_t tbytes = 256u << 20; CHECK(cudaMalloc(&table, tbytes)); CHECK(cudaMemset(table, 0, tbytes));
  struct { const char* n; int mask; } tabs[3] = {{"l1_64KB", 511}, {"l2_32MB", (1 << 18) - 1}, {"hbm_256MB", (1 << 21) - 1}};
  for (int mode = 0; mode < 2; ++mode) for (int t = 0; t < 3; ++t) {
    const int iters = 1024; const int ctas = nsm * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<<<ctas, 256>>>(mode, table, tabs[t].mask, rowidx, out, cyc, iters);
    cudaEventRecord(e0);
    gather_kernel<<<ctas, 256>>>(mode, table, tabs[t].mask, rowidx, out, cyc, iters);
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)ctas * 8 * iters * (mode == 0 ? 4 : 1);
    printf(" \"gather_%s_%s\": {\"ms\": %.4f, \"rows_per_ns\": %.2f, \"GBps\": %.1f},\n", mode == 0 ? "ldg128x4rows" : "ldg32x1row", tabs[t].n, ms,
           rows / (ms * 1e6), rows * 128 / (ms * 1e6));
  }
  // reds
  for (int t = 0; t < 3; ++t) for (int contig = 0; contig < 2; ++contig) {
    const int iters = 512; const int ctas = nsm * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    red_kernel<<<ctas, 256>>>(table, tabs[t].mask, rowidx, iters, contig);
    cudaEventRecord(e0);
    red_kernel<<<ctas, 256>>>(table, tabs[t].mask, rowidx, iters, contig);
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)ctas * 8 * iters * 4;
    printf(" \"red_f32x4_%s_%s\": {\"ms\": %.4f, \"rows_per_ns\": %.2f, \"GBps\": %.1f},\n", tabs[t].n, contig ? "contig" : "random", ms, rows / (ms * 1e6),
           rows * 128 / (ms * 1e6));
  }
  printf(" \"done\": 1}\n");
  return 0;
}
