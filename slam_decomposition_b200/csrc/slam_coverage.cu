// slam_coverage.cu -- K6: fused coverage-set Monte-Carlo.
//
// Replaces the N-iteration Python loop of src/slam/utils/gates/parallel_drive_volume.py:209-225 (random
// parameters -> qiskit Operator -> c1c2c3) and the mirror fold at :292-307 by one kernel per (gate, k):
//   Philox4x32-10(seed, sample, param) -> template parameters -> 4x4 chain (closed-form or smush gates)
//   -> Weyl coordinates (joint Jacobi) -> fold c1 -> bin -> one 64-bit atomic per sample.
// Nothing but the histogram (and, optionally, the coordinates) is written to HBM; there is no parameter
// array at all.  The RNG is counter-based, so a rank regenerates exactly its own [first, first+n) slice.
#include "slam_fwd1.cuh"
#include "slam_host.h"
#include "slam_philox.cuh"
#include "slam_weyl.cuh"

namespace slam {

struct PhiloxParams {
  uint64_t seed, sample;
  double lo, span;
  __device__ __forceinline__ double get(int j) const { return philox_param(seed, sample, j, lo, span); }
};

__global__ void __launch_bounds__(128) coverage_kernel(uint64_t seed, int64_t first, int64_t n, double lo, double span,
                                                       int nbins, unsigned long long* __restrict__ hist,
                                                       double* __restrict__ coords, const __grid_constant__ KTemplate kt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double scale = 2.0 * (double)nbins;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    PhiloxParams ps{seed, (uint64_t)(first + i), lo, span};
    cd R[4][4];
    fwd1_chain(kt, ps, R);
    cd M[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) M[r][c] = R[c][r];
    double c3[3];
    weyl_makhlin(M, SLAM_WEYL_FOLD, c3, nullptr);
    if (coords) {
      coords[3 * i] = c3[0];
      coords[3 * i + 1] = c3[1];
      coords[3 * i + 2] = c3[2];
    }
    if (hist) {
      int b0 = min(max((int)floor(c3[0] * scale), 0), nbins - 1);
      int b1 = min(max((int)floor(c3[1] * scale), 0), nbins - 1);
      int b2 = min(max((int)floor(c3[2] * scale), 0), nbins - 1);
      atomicAdd(hist + ((size_t)b0 * nbins + b1) * nbins + b2, 1ULL);
    }
  }
}

}  // namespace slam

extern "C" int slam_coverage_mc(const SlamTemplateDesc* desc, uint64_t seed, int64_t first_sample, int64_t n_samples,
                                double lo, double hi, int32_t nbins, unsigned long long* hist, double* coords,
                                void* stream) {
  using namespace slam;
  if (!desc || n_samples < 0 || first_sample < 0 || (!hist && !coords) || (hist && (nbins < 1 || nbins > 1024)))
    return SLAM_ERR_INVALID;
  if (n_samples == 0) return SLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KTemplate kt;
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/true);
  if (rc != SLAM_OK) return rc;
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  int dev = 0, sms = 0;
  SLAM_CUDA_CHECK(cudaGetDevice(&dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t want = (n_samples + 127) / 128;
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * 16);  // grid-stride, a multiple of the SM count
  coverage_kernel<<<grid, 128, 0, st>>>(seed, first_sample, n_samples, lo, hi - lo, nbins, hist, coords, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}
