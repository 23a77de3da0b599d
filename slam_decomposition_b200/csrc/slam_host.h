// slam_host.h -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "slam_core.cuh"

namespace slam {

void set_cuda_error(cudaError_t e, const char* where);

#define SLAM_CUDA_CHECK(expr)                   \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) {                    \
      ::slam::set_cuda_error(_e, #expr);        \
      return SLAM_ERR_CUDA;                     \
    }                                           \
  } while (0)

// Validate a SlamTemplateDesc and lower it to the kernel-side KTemplate.
//   allow_bound_smush: parameter-bound smush gates are accepted (forward-only kernels)
int compile_template(const SlamTemplateDesc* d, KTemplate* kt, bool allow_bound_smush);

// Host evaluation of a *constant* smush gate is not done on the CPU: callers lower constant smush
// gates to GM_DENSE by running the device smush kernel once (see slam_smush.cu).
int lower_const_smush(const SlamTemplateDesc* d, KTemplate* kt, cudaStream_t stream);

// cudaMallocAsync scratch is used for small per-call work areas.  With the default release threshold (0) every
// synchronisation returns the pool to the OS and the next call pays for re-mapping it (measured on B200: stalls of
// 0.1-1 s per sweep step); raise the threshold so freed blocks stay in the device's default pool.
int keep_async_pool(int device);

}  // namespace slam
