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
//   allow_bound_smush: parameter-bound smush gates are accepted (forward evaluators and the slice adjoint)
//   allow_ties: one Xk entry may be bound to several slots (tied pulse parameters, pd_playground / parallel_drive_swap);
//               only for entry points that never STORE a per-slot partial derivative (forward evaluation, Nelder-Mead,
//               finite differences, coverage)
int compile_template(const SlamTemplateDesc* d, KTemplate* kt, bool allow_bound_smush, bool allow_ties = false);

// Host evaluation of a *constant* smush gate is not done on the CPU: callers lower constant smush
// gates to GM_DENSE by running the device smush kernel once (see slam_smush.cu).
int lower_const_smush(const SlamTemplateDesc* d, KTemplate* kt, cudaStream_t stream);

// Stream-ordered scratch (work counters, solved flags, optimiser workspaces) comes from a per-device memory pool OWNED BY THE
// LIBRARY -- the device's default pool (shared with the rest of the process, torch included) is left alone.  The pool keeps up
// to 1 GiB of freed blocks across synchronisations (with the CUDA default of 0 every synchronisation returns the memory to the
// driver and the next call pays for re-mapping it: measured stalls of 0.1-1 s per sweep step on B200) and gives larger
// workspaces back.  Scratch frees what it handed out when it leaves scope, on the success and on every error path.
int scratch_pool(int device, cudaMemPool_t* pool);

class Scratch {
 public:
  explicit Scratch(cudaStream_t st) : st_(st), n_(0) {}
  ~Scratch() {
    for (int i = 0; i < n_; ++i) cudaFreeAsync(p_[i], st_);
  }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  // SLAM status; *p is valid on st_ (and streams ordered after it) until this object dies
  template <class T>
  int alloc(T** p, size_t bytes, bool zero = false) {
    *p = nullptr;
    if (n_ >= 8) return SLAM_ERR_INVALID;
    int dev = 0;
    SLAM_CUDA_CHECK(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    if (int rc = scratch_pool(dev, &pool)) return rc;
    void* q = nullptr;
    SLAM_CUDA_CHECK(cudaMallocFromPoolAsync(&q, bytes ? bytes : 1, pool, st_));
    p_[n_++] = q;
    if (zero) SLAM_CUDA_CHECK(cudaMemsetAsync(q, 0, bytes, st_));
    *p = static_cast<T*>(q);
    return SLAM_OK;
  }

 private:
  cudaStream_t st_;
  void* p_[8];
  int n_;
};

}  // namespace slam
