/* kat_client.c — a plain C11 host program that drives the library through include/msda_b200.h only (no Python, no torch,
 * no C++): the header compiles as C, and the entry points are usable from the kind of host code the reference's
 * ms_deform_attn_cuda.cu:20-153 is. Reads one operator case from raw little-endian files written by the test
 * (tests/test_cabi_c_client.py), runs msda_forward + msda_backward on the GPU, writes the results back.
 *
 *   kat_client <dir>      with <dir>/meta.txt : N S M D L Lq P
 *                              <dir>/{value,loc,aw,grad_out}.f32, <dir>/{shapes,lsi}.i64
 *                         ->   <dir>/{out,grad_value,grad_loc,grad_aw}.f32
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "msda_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

static void* slurp(const char* dir, const char* name, size_t bytes) {
  char path[1024];
  snprintf(path, sizeof(path), "%s/%s", dir, name);
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(3); }
  void* p = malloc(bytes);
  if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read %s\n", path); exit(3); }
  fclose(f);
  return p;
}

static void dump(const char* dir, const char* name, const void* p, size_t bytes) {
  char path[1024];
  snprintf(path, sizeof(path), "%s/%s", dir, name);
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "cannot write %s\n", path); exit(3); }
  fclose(f);
}

static void* to_device(const void* h, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    fprintf(stderr, "device upload failed\n");
    exit(2);
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: kat_client <dir>\n"); return 1; }
  const char* dir = argv[1];
  char path[1024];
  snprintf(path, sizeof(path), "%s/meta.txt", dir);
  FILE* f = fopen(path, "r");
  msda_dims d;
  if (!f || fscanf(f, "%d %d %d %d %d %d %d", &d.batch, &d.spatial_size, &d.num_heads, &d.channels, &d.num_levels, &d.num_query,
                   &d.num_point) != 7) { fprintf(stderr, "bad meta.txt\n"); return 3; }
  fclose(f);
  if (msda_abi_version() != MSDA_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 4; }

  const size_t nval = (size_t)d.batch * d.spatial_size * d.num_heads * d.channels;
  const size_t npts = (size_t)d.batch * d.num_query * d.num_heads * d.num_levels * d.num_point;
  const size_t nout = (size_t)d.batch * d.num_query * d.num_heads * d.channels;
  float* h_value = slurp(dir, "value.f32", nval * 4);
  float* h_loc = slurp(dir, "loc.f32", npts * 8);
  float* h_aw = slurp(dir, "aw.f32", npts * 4);
  float* h_go = slurp(dir, "grad_out.f32", nout * 4);
  int64_t* h_shapes = slurp(dir, "shapes.i64", (size_t)d.num_levels * 16);
  int64_t* h_lsi = slurp(dir, "lsi.i64", (size_t)d.num_levels * 8);

  float* value = to_device(h_value, nval * 4);
  float* loc = to_device(h_loc, npts * 8);
  float* aw = to_device(h_aw, npts * 4);
  float* go = to_device(h_go, nout * 4);
  int64_t* shapes = to_device(h_shapes, (size_t)d.num_levels * 16);
  int64_t* lsi = to_device(h_lsi, (size_t)d.num_levels * 8);
  float *out, *gv, *gl, *ga;
  CK(cudaMalloc((void**)&out, nout * 4));
  CK(cudaMalloc((void**)&gv, nval * 4));
  CK(cudaMalloc((void**)&gl, npts * 8));
  CK(cudaMalloc((void**)&ga, npts * 4));
  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));

  /* argument errors come back as codes + a message, never as a crash */
  if (msda_forward(&d, 99, value, shapes, lsi, loc, aw, out, stream) != MSDA_E_DTYPE || strlen(msda_last_error()) == 0) {
    fprintf(stderr, "expected MSDA_E_DTYPE for dtype 99\n");
    return 5;
  }
  if (msda_check_im2col_step(d.batch, 64) != 0) { fprintf(stderr, "%s\n", msda_last_error()); return 5; }

  const uint64_t l0 = msda_launch_count();
  int rc = msda_forward(&d, MSDA_F32, value, shapes, lsi, loc, aw, out, stream);
  if (rc) { fprintf(stderr, "msda_forward: %d %s\n", rc, msda_last_error()); return 6; }
  const size_t ws = msda_backward_workspace_bytes(&d, MSDA_F32);
  void* wsp = NULL;
  if (ws) CK(cudaMalloc(&wsp, ws));
  rc = msda_backward(&d, MSDA_F32, value, shapes, lsi, loc, aw, go, gv, gl, ga, wsp, ws, stream);
  if (rc) { fprintf(stderr, "msda_backward: %d %s\n", rc, msda_last_error()); return 6; }
  CK(cudaStreamSynchronize(stream));
  if (msda_launch_count() - l0 < 2) { fprintf(stderr, "no kernels launched?\n"); return 7; }

  float* h = malloc(nval * 4 > npts * 8 ? nval * 4 : npts * 8);
  float* ho = malloc(nout * 4);
  CK(cudaMemcpy(ho, out, nout * 4, cudaMemcpyDeviceToHost)); dump(dir, "out.f32", ho, nout * 4);
  CK(cudaMemcpy(h, gv, nval * 4, cudaMemcpyDeviceToHost)); dump(dir, "grad_value.f32", h, nval * 4);
  CK(cudaMemcpy(h, gl, npts * 8, cudaMemcpyDeviceToHost)); dump(dir, "grad_loc.f32", h, npts * 8);
  CK(cudaMemcpy(h, ga, npts * 4, cudaMemcpyDeviceToHost)); dump(dir, "grad_aw.f32", h, npts * 4);
  printf("ok launches=%llu\n", (unsigned long long)(msda_launch_count() - l0));
  return 0;
}
