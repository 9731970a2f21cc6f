// Empty stand-in: float/double atomicAdd are native on sm_100a (see ../ATen/ATen.h).
#pragma once
#include <cuda_runtime.h>
