// mb_red.cu — where is the 49 rows/ns ceiling of `red.global.add.v4.f32`? And is the TMA bulk reduce a faster way to L2?
//  (1) RED.v4 on random 128-byte rows of a 32 MB table with all SMs, half, a quarter of them: a per-SM limit scales with the
//      SM count, an L2-side limit does not.
//  (2) the same rows through `cp.reduce.async.bulk.global.shared::cta.add.f32` (row staged in shared memory by the 8-lane
//      group, one lane issues the 128-byte bulk reduce): the async-proxy path instead of the LSU RED path.
//  (3) bulk reduce of 256 / 512 contiguous bytes per request (what a head-major grad_value layout would allow).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_red mb_red.cu && ./mb_red
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void red_v4_kernel(float* table, int nrows_mask, const int* __restrict__ rowidx, int iters) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int i = 0; i < iters; ++i) {
    const int row = __ldg(&rowidx[((gw * 4 + grp) * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
    float* p = table + (size_t)row * 32 + j * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
  }
}

// packed bf16x2 atomics: 16 bytes per lane carry 8 channels; LANES lanes cover one (b, token, head) row of 8*LANES channels
template <int LANES>
__global__ void red_bf16x2_kernel(unsigned* table, int nrows_mask, const int* __restrict__ rowidx, int iters) {
  const int lane = threadIdx.x & 31, grp = lane / LANES, j = lane % LANES;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned one2 = 0x3f803f80u;  // (1.0bf16, 1.0bf16)
  for (int i = 0; i < iters; ++i) {
    const int row = __ldg(&rowidx[((gw * (32 / LANES) + grp) * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
    unsigned* p = table + (size_t)row * (LANES * 4) + j * 4;
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(one2), "r"(one2), "r"(one2), "r"(one2) : "memory");
  }
}

// rows_per_req: 1 -> every group reduces its own 128-byte row; 2 / 4 -> lane 0 of group 0 (and 2) reduces 256 / 512 contiguous bytes
template <int ROWS_PER_REQ>
__global__ void bulk_red_kernel(float* table, int nrows_mask, const int* __restrict__ rowidx, int iters) {
  extern __shared__ __align__(128) float sm[];  // [warps][4 slots deep][4 rows][32 floats]
  const int lane = threadIdx.x & 31, grp = lane >> 3, j = lane & 7, warp = threadIdx.x >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr int kDepth = 4;
  float* mine = sm + (size_t)warp * kDepth * 4 * 32;
  for (int i = 0; i < iters; ++i) {
    const int slot = i & (kDepth - 1);
    if (i >= kDepth) {  // the slot's previous bulk reduce must have finished READING shared memory
      if (lane == 0 || (ROWS_PER_REQ == 1 && j == 0) || (ROWS_PER_REQ == 2 && lane == 16))
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kDepth - 1) : "memory");
      __syncwarp();
    }
    float* s = mine + (slot * 4 + grp) * 32 + j * 4;
    *reinterpret_cast<float4*>(s) = make_float4(1.f, 1.f, 1.f, 1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int reqgrp = (ROWS_PER_REQ == 1) ? grp : (ROWS_PER_REQ == 2 ? (grp & 2) : 0);
    int row = __ldg(&rowidx[((gw * 4 + reqgrp) * 64 + (i & 63)) & 0xfffff]) & nrows_mask;
    row &= ~(ROWS_PER_REQ - 1);
    const bool issuer = (ROWS_PER_REQ == 1) ? (j == 0) : (ROWS_PER_REQ == 2 ? (lane == 0 || lane == 16) : lane == 0);
    if (issuer) {
      const unsigned src = (unsigned)__cvta_generic_to_shared(mine + (slot * 4 + reqgrp) * 32);
      float* dst = table + (size_t)row * 32;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(128 * ROWS_PER_REQ) : "memory");
    }
    if (lane == 0 || (ROWS_PER_REQ == 1 && j == 0) || (ROWS_PER_REQ == 2 && lane == 16)) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  float* table; const size_t tbytes = 256u << 20; CHECK(cudaMalloc(&table, tbytes)); CHECK(cudaMemset(table, 0, tbytes));
  const int nr = 1 << 20;
  int* h = (int*)malloc(nr * 4);
  unsigned s = 777u;
  for (int i = 0; i < nr; ++i) { s = s * 1664525u + 1013904223u; h[i] = (int)(s >> 4); }
  int* rowidx; CHECK(cudaMalloc(&rowidx, nr * 4)); CHECK(cudaMemcpy(rowidx, h, nr * 4, cudaMemcpyHostToDevice));
  const int mask = (1 << 18) - 1;  // 32 MB of 128-byte rows: L2 resident
  const int iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  printf("{\"gpu\": \"%s\", \"sms\": %d,\n", prop.name, nsm);
  for (int div = 1; div <= 8; div *= 2) {
    const int ctas = nsm * 8 / div;   // 8 CTAs of 256 threads per SM when div == 1; fewer SMs' worth of CTAs otherwise
    red_v4_kernel<<<ctas, 256>>>(table, mask, rowidx, iters);
    cudaEventRecord(e0);
    red_v4_kernel<<<ctas, 256>>>(table, mask, rowidx, iters);
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    printf(" \"red_v4_ctas_%d\": {\"rows_per_ns\": %.2f},\n", ctas, (double)ctas * 8 * iters * 4 / (ms * 1e6));
  }
  // one CTA per SM (grid = nsm / div): per-SM rate when few SMs issue
  for (int div = 1; div <= 16; div *= 4) {
    const int ctas = nsm / div;
    red_v4_kernel<<<ctas, 1024>>>(table, mask, rowidx, iters);
    cudaEventRecord(e0);
    red_v4_kernel<<<ctas, 1024>>>(table, mask, rowidx, iters);
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)ctas * 32 * iters * 4;
    printf(" \"red_v4_one_cta_of_1024_per_sm_on_%d_sms\": {\"rows_per_ns\": %.2f, \"sm_cycles_per_row\": %.2f},\n", ctas, rows / (ms * 1e6),
           ms * 1e6 * 1.965 * ctas / rows);
  }
  {
    const int ctas = nsm * 8;
    // 4 lanes x 16 B = 64-byte rows = 32 bf16 channels (D = 32): 8 rows per warp instruction
    red_bf16x2_kernel<4><<<ctas, 256>>>((unsigned*)table, (1 << 19) - 1, rowidx, iters);
    cudaEventRecord(e0);
    red_bf16x2_kernel<4><<<ctas, 256>>>((unsigned*)table, (1 << 19) - 1, rowidx, iters);
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    printf(" \"red_v4_bf16x2_64B_rows_32ch\": {\"rows_per_ns\": %.2f},\n", (double)ctas * 8 * iters * 8 / (ms * 1e6));
    // 8 lanes x 16 B = 128-byte rows = 64 bf16 channels (D = 64)
    red_bf16x2_kernel<8><<<ctas, 256>>>((unsigned*)table, mask, rowidx, iters);
    cudaEventRecord(e0);
    red_bf16x2_kernel<8><<<ctas, 256>>>((unsigned*)table, mask, rowidx, iters);
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
    printf(" \"red_v4_bf16x2_128B_rows_64ch\": {\"rows_per_ns\": %.2f},\n", (double)ctas * 8 * iters * 4 / (ms * 1e6));
  }
  const size_t smem = 8 * 4 * 4 * 32 * 4;
#define BULK(R)                                                                                                      \
  for (int div = 1; div <= 4; div *= 4) {                                                                            \
    const int ctas = nsm * 4 / div;                                                                                  \
    bulk_red_kernel<R><<<ctas, 256, smem>>>(table, mask, rowidx, iters);                                             \
    cudaEventRecord(e0);                                                                                             \
    bulk_red_kernel<R><<<ctas, 256, smem>>>(table, mask, rowidx, iters);                                             \
    cudaEventRecord(e1); CHECK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);                          \
    printf(" \"bulk_reduce_%dB_per_request_ctas_%d\": {\"rows_per_ns\": %.2f},\n", 128 * R, ctas, (double)ctas * 8 * iters * 4 / (ms * 1e6)); \
  }
  BULK(1) BULK(2) BULK(4)
  float hsum[32]; CHECK(cudaMemcpy(hsum, table, 128, cudaMemcpyDeviceToHost));
  printf(" \"check_row0\": %.1f, \"done\": 1}\n", hsum[0]);
  return 0;
}
