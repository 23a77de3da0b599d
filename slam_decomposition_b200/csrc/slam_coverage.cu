// slam_coverage.cu -- K6: fused coverage-set Monte-Carlo.
//
// Replaces the N-iteration Python loop of src/slam/utils/gates/parallel_drive_volume.py:209-225 (random
// parameters -> qiskit Operator -> c1c2c3) and the mirror fold at :292-307 by one kernel per (gate, k):
//   Philox4x32-10(seed, sample, param) -> template parameters -> 4x4 chain (closed-form or smush gates)
//   -> Weyl coordinates (joint Jacobi) -> fold c1 -> bin -> one 64-bit atomic per sample.
// Nothing but the histogram (and, optionally, the coordinates) is written to HBM; there is no parameter
// array at all.  The RNG is counter-based, so a rank regenerates exactly its own [first, first+n) slice.
#include <algorithm>
#include <cstdlib>

#include "slam_fwd1.cuh"
#include "slam_host.h"
#include "slam_philox.cuh"
#include "slam_weyl.cuh"

namespace slam {

// The sample's P parameters are generated once, two per Philox block, into the thread's column of a shared-memory
// table (entry j of thread t at xs[j * CTA + t]: conflict-free) and read from there by the chain.  Calling the generator
// per parameter access ran every Philox block twice; at P = 12 that was a third of the kernel's instructions.
template <int CTA>
struct StagedParams {
  const double* col;
  __device__ __forceinline__ double get(int j) const { return col[j * CTA]; }
};

// GMT:  gate mode fixed at compile time (-1 = runtime dispatch)
// MINB: 128-thread units per SM the register allocation targets (2: 255 registers, 3: 168, 4: 128)
// SYNC: phase-locked variant -- ONE CTA of 128 * MINB threads per SM with barriers at the phase boundaries (slam_fwd1.cuh)
template <int GMT, int MINB, bool SYNC>
__global__ void __launch_bounds__(SYNC ? 128 * MINB : 128, SYNC ? 1 : MINB)
coverage_kernel(uint64_t seed, int64_t first, int64_t n, double lo, double span, int nbins,
                unsigned long long* __restrict__ hist, double* __restrict__ coords, const __grid_constant__ KTemplate kt) {
  constexpr int CTA = SYNC ? 128 * MINB : 128;
  extern __shared__ double xs[];
  double* col = xs + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * CTA;
  const double scale = 2.0 * (double)nbins;
  // uniform trip count per CTA (the barriers of the SYNC variant sit inside the loop); out-of-range lanes recompute the
  // last sample and drop the result
  for (int64_t base = (int64_t)blockIdx.x * CTA; base < n; base += stride) {
    const int64_t i = min(base + (int64_t)threadIdx.x, n - 1);
    const bool valid = base + (int64_t)threadIdx.x < n;
    for (int jj = 0; 2 * jj < kt.P; ++jj) {
      double u0, u1;
      philox_u53_pair(seed, (uint64_t)(first + i), jj, &u0, &u1);
      // lo + span * u with a separate multiply and add (numpy does not fuse; keeps the stream bit-identical)
      col[(2 * jj) * CTA] = __dadd_rn(lo, __dmul_rn(span, u0));
      if (2 * jj + 1 < kt.P) col[(2 * jj + 1) * CTA] = __dadd_rn(lo, __dmul_rn(span, u1));
    }
    StagedParams<CTA> ps{col};
    cd R[4][4];
    fwd1_chain<StagedParams<CTA>, GMT, SYNC>(kt, ps, R);
    if (SYNC) __syncthreads();
    cd M[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) M[r][c] = R[c][r];
    double c3[3];
    weyl_makhlin(M, SLAM_WEYL_FOLD, c3, nullptr);
    if (valid && coords) {
      coords[3 * i] = c3[0];
      coords[3 * i + 1] = c3[1];
      coords[3 * i + 2] = c3[2];
    }
    if (valid && hist) {
      int b0 = min(max((int)floor(c3[0] * scale), 0), nbins - 1);
      int b1 = min(max((int)floor(c3[1] * scale), 0), nbins - 1);
      int b2 = min(max((int)floor(c3[2] * scale), 0), nbins - 1);
      atomicAdd(hist + ((size_t)b0 * nbins + b1) * nbins + b2, 1ULL);
    }
    if (SYNC) __syncthreads();
  }
}

template <int GMT, int MINB, bool SYNC>
static int launch_coverage(const KTemplate& kt, uint64_t seed, int64_t first, int64_t n, double lo, double span, int nbins,
                           unsigned long long* hist, double* coords, int sms, cudaStream_t st) {
  constexpr int CTA = SYNC ? 128 * MINB : 128;
  const int64_t want = (n + CTA - 1) / CTA;
  // grid-stride, a multiple of the SM count
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * (SYNC ? 8 : 8 * MINB));
  const size_t smem = (size_t)std::max(kt.P, 1) * CTA * sizeof(double);
  auto kern = coverage_kernel<GMT, MINB, SYNC>;
  if (smem > 48 * 1024) SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, CTA, smem, st>>>(seed, first, n, lo, span, nbins, hist, coords, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
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
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/true, /*allow_ties=*/true);
  if (rc != SLAM_OK) return rc;
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  int dev = 0, sms = 0;
  SLAM_CUDA_CHECK(cudaGetDevice(&dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const double span = hi - lo;
  // 128-thread units per SM the kernels are compiled for (register cap 255 / 168 / 128).  Measured on B200, Msamples/s at
  // 2 / 3 / 4 units: sqrt(iSWAP) k=3 plain 2355 / 2771 / 2928, CNOT k=3 plain 2338 / 2745 / 2903 (latency bound: 16 warps/SM
  // win despite 160 B of spills); sqrt(iSWAP) k=3 smush, phase-locked, 542 / 542 / 515 (scripts/cov_ab.py).
  // Phase-locked variant (one CTA per SM, barriers at the layer / gate / slice boundaries): on for the smush templates,
  // whose kernel is 75 KB of code (sqrt(iSWAP) k=3 smush 458 -> 542, CNOT k=2 smush 424 -> 454 Msamples/s); off for the
  // closed-form gates, whose hot code fits the 32 KB instruction cache (2928 vs 2862 Msamples/s).
  const bool sync = kt.gmode == GM_SMUSH;
#define SLAM_COV(GM, MB)                                                                                                  \
  return (sync && (size_t)kt.P * 128 * MB * sizeof(double) <= 200 * 1024)                                                  \
             ? launch_coverage<GM, MB, true>(kt, seed, first_sample, n_samples, lo, span, nbins, hist, coords, sms, st)   \
             : launch_coverage<GM, MB, false>(kt, seed, first_sample, n_samples, lo, span, nbins, hist, coords, sms, st)
  switch (kt.gmode) {  // closed-form gates: 4 units (16 warps/SM); smush: 2 units (255 registers)
    case GM_SYM: SLAM_COV(GM_SYM, 4);
    case GM_BLOCK: SLAM_COV(GM_BLOCK, 4);
    case GM_DENSE: SLAM_COV(GM_DENSE, 4);
    default: SLAM_COV(GM_SMUSH, 2);
  }
#undef SLAM_COV
  return SLAM_OK;
}
