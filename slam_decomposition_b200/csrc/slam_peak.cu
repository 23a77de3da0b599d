// slam_peak.cu -- register-resident DFMA microbenchmark: the FP64 roofline denominator.
// (MEASURED_PEAKS.json carries HBM and bf16 peaks only; SURVEY.md 8(d) asks for a measured DFMA peak.)
#include "slam_host.h"

namespace slam {

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double v[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) v[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < kChains; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += v[i];
  if (s == 123.456) out[0] = s;  // keep the chains live
}

}  // namespace slam

extern "C" int slam_fp64_peak(int32_t iters, double* flops, double* ms_out) {
  using namespace slam;
  if (iters <= 0 || !flops) return SLAM_ERR_INVALID;
  int dev = 0, sms = 0;
  SLAM_CUDA_CHECK(cudaGetDevice(&dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  SLAM_CUDA_CHECK(cudaMalloc(&d, sizeof(double)));
  const int grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  SLAM_CUDA_CHECK(cudaEventCreate(&e0));
  SLAM_CUDA_CHECK(cudaEventCreate(&e1));
  dfma_kernel<<<grid, block>>>(d, iters / 8 + 1, 0.999999, 1e-9);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    SLAM_CUDA_CHECK(cudaEventRecord(e0));
    dfma_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-9);
    SLAM_CUDA_CHECK(cudaEventRecord(e1));
    SLAM_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    SLAM_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  SLAM_CUDA_CHECK(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double n_fma = (double)grid * block * (double)iters * 8.0 * kChains;
  *flops = 2.0 * n_fma / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return SLAM_OK;
}

extern "C" int slam_selftest_sincos(const double* x, int64_t n, double* s, double* c) {
  if (!x || !s || !c || n < 0) return SLAM_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i) slam::fast_sincos(x[i], s + i, c + i);
  return SLAM_OK;
}
