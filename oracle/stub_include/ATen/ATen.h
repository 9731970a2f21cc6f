// Empty stand-in so the reference's ms_deform_im2col_cuda.cuh compiles without a torch install:
// the header includes <ATen/ATen.h> but uses nothing from it (only cudaStream_t, atomicAdd, printf).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
