// slam_smush.cu -- K4: templates whose 2Q gate is a time-sliced smush Hamiltonian (forward evaluation),
// constant-gate lowering, and K4b: the parallel-drive Weyl trajectory.
#include <cstdlib>

#include "slam_adj1.cuh"
#include "slam_host.h"
#include "slam_weyl.cuh"

namespace slam {

// ---- K1 for parameter-bound smush templates: one thread per parameter row ---------------------------
// 128-thread CTAs, two per SM (the kernels need ~250 registers).  The phase-locked form that pays off in the coverage kernel
// (CTA-wide barriers at the layer / gate / slice boundaries, slam_fwd1.cuh) was measured SLOWER here -- 597 vs 661 M evals/s
// (K1), 170 vs 176 M loss+grad/s (K2) at sqrt(iSWAP) k=3: these kernels read their parameter rows from global memory, and a
// barrier makes every warp wait for the slowest row; with the rows staged through shared memory first it was still slower
// (554 vs 656 and 167 vs 184: the warps of a one-shot streaming kernel start in step anyway, the barriers only add waits) --
// so SYNC stays off; padding lanes still recompute the last row so that the flag can be flipped for experiments.
// (Re-measured in round 2 for the eigen-form gradient kernel: 326 vs 330 and 211 vs 209 M rows/s at k = 2 / 3 -- no difference.)
constexpr int kSmushCta = 128;
constexpr bool kSmushSync = false;

__global__ void __launch_bounds__(kSmushCta) smush_eval_kernel(const double* __restrict__ x, int64_t ldx,
                                                                  double* __restrict__ U, int64_t B,
                                                                  const __grid_constant__ KTemplate kt) {
  const int64_t b0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = min(b0, B - 1);
  GlobalParams ps{x + b * ldx};
  cd R[4][4];
  fwd1_chain<GlobalParams, GM_SMUSH, kSmushSync>(kt, ps, R);
  if (b0 >= B) return;
  double* out = U + b * 32;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<double2*>(out + (r * 4 + c) * 2) = make_double2(R[c][r].re, R[c][r].im);
}

int smush_eval_launch(const KTemplate& kt, const double* x, int64_t ldx, double* U, int64_t B, cudaStream_t st) {
  const unsigned grid = (unsigned)((B + kSmushCta - 1) / kSmushCta);
  smush_eval_kernel<<<grid, kSmushCta, 0, st>>>(x, ldx, U, B, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

// ---- K2 for parameter-bound smush templates: loss + analytic adjoint gradient, one thread per row -------
// (replaces objective_func + scipy's (P+1)-evaluation finite-difference gradient for these templates,
//  optimizer.py:191-214, 270-278)
// EIG: eigen-form forward sweep handing (Q, lam) of every slice to the backward sweep (slam_adj1.cuh, k T <= kEigSlices)
template <bool WANT_GRAD, bool EIG = false>
__global__ void __launch_bounds__(kSmushCta)
smush_loss_grad_kernel(const double* __restrict__ x, int64_t ldx, const double* __restrict__ V, int64_t Nt,
                       const int32_t* __restrict__ tgt_idx, int cost_kind, double* __restrict__ loss,
                       double* __restrict__ grad, int64_t ldg, double* __restrict__ trace, int64_t B,
                       const __grid_constant__ KTemplate kt) {
  const int64_t b0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b0 < B;
  const int64_t b = min(b0, B - 1);
  const int64_t tgt = tgt_idx ? (int64_t)tgt_idx[b] : (b % Nt);
  GlobalParams ps{x + b * ldx};
  cd T;
  double l;
  if (WANT_GRAD) {
    RowGrad gs{valid ? grad + b * ldg : nullptr};
    if (valid)
      for (int j = 0; j < kt.P; ++j) gs.row[j] = 0.0;  // parameters bound to no slot keep a zero derivative
    if (EIG) l = adj1_loss_grad<GlobalParams, RowGrad, kSmushSync, LocalEigStore>(kt, ps, V + tgt * 32, cost_kind, gs, &T);
    else l = adj1_loss_grad<GlobalParams, RowGrad, kSmushSync>(kt, ps, V + tgt * 32, cost_kind, gs, &T);
  } else {
    l = fwd1_loss<GlobalParams, kSmushSync>(kt, ps, V + tgt * 32, cost_kind, &T);
  }
  if (!valid) return;
  loss[b] = l;
  if (trace) {
    trace[2 * b] = T.re;
    trace[2 * b + 1] = T.im;
  }
}

int smush_loss_grad_launch(const KTemplate& kt, const double* x, int64_t ldx, const double* V, int64_t Nt,
                           const int32_t* tgt_idx, int cost_kind, double* loss, double* grad, int64_t ldg, double* trace,
                           int64_t B, cudaStream_t st) {
  const unsigned grid = (unsigned)((B + kSmushCta - 1) / kSmushCta);
  if (grad && kt.k * kt.T <= kEigSlices)
    smush_loss_grad_kernel<true, true><<<grid, kSmushCta, 0, st>>>(x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, kt);
  else if (grad)
    smush_loss_grad_kernel<true><<<grid, kSmushCta, 0, st>>>(x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, kt);
  else
    smush_loss_grad_kernel<false><<<grid, kSmushCta, 0, st>>>(x, ldx, V, Nt, tgt_idx, cost_kind, loss, nullptr, 0, trace, B, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

// ---- constant smush gates -> dense matrices (device computes them; no host expm) ---------------------
struct ConstParams {
  __device__ __forceinline__ double get(int) const { return 0.0; }
};

__global__ void const_smush_kernel(double* __restrict__ out, const __grid_constant__ KTemplate kt) {
  const int g = threadIdx.x;
  if (g >= kt.k) return;
  cd R[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) R[c][r] = mkc(c == r ? 1.0 : 0.0, 0.0);
  ConstParams ps;
  fwd1_gate(kt, ps, g, R);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      out[g * 32 + (r * 4 + c) * 2] = R[c][r].re;
      out[g * 32 + (r * 4 + c) * 2 + 1] = R[c][r].im;
    }
}

int lower_const_smush(const SlamTemplateDesc* d, KTemplate* kt, cudaStream_t st) {
  (void)d;
  KTemplate tmp = *kt;
  tmp.gmode = GM_SMUSH;
  Scratch scratch(st);
  double* dev = nullptr;
  if (int rc = scratch.alloc(&dev, sizeof(double) * 32 * SLAM_MAX_K)) return rc;
  const_smush_kernel<<<1, SLAM_MAX_K, 0, st>>>(dev, tmp);
  SLAM_CUDA_CHECK(cudaGetLastError());
  SLAM_CUDA_CHECK(cudaMemcpyAsync(kt->dense, dev, sizeof(double) * 32 * kt->k, cudaMemcpyDeviceToHost, st));
  SLAM_CUDA_CHECK(cudaStreamSynchronize(st));
  kt->gmode = GM_DENSE;
  return SLAM_OK;
}

// ---- K4b: Weyl trajectory of N smush1q slices, R sub-times each (pd_playground.py:179-208) -----------
// SYNC: phase-locked form (CTA-wide barriers between the exponential and the Weyl step of every trajectory point; all threads
// run the same N x R loop, padding lanes recompute the last trajectory) -- see fwd1_gate in slam_fwd1.cuh
template <bool SYNC>
__global__ void __launch_bounds__(SYNC ? 256 : 128, SYNC ? 1 : 2)
trajectory_kernel(const double* __restrict__ gate, int gate_slice_stride, const double* __restrict__ gx,
                  const double* __restrict__ gy, int N, int Rn, double dt, int flags, double* __restrict__ coords,
                  double* __restrict__ Ufinal, int64_t B) {
  // gate rows: one per trajectory (gate_slice_stride = 0) or one per SLICE (8: multi-segment pulses, whose drive phases /
  // couplings change between segments -- ParallelDrivenGateWidget.__add__, pd_playground.py:46-58)
  const int64_t b0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b0 < B;
  if (!SYNC && !valid) return;
  const int64_t b = valid ? b0 : B - 1;
  const double* gp = gate + b * (gate_slice_stride ? (int64_t)gate_slice_stride * N : 8);
  SmushGate G = smush_gate(gp[0], gp[1], gp[2], gp[3], gp[4], gp[5], gp[6], gp[7]);
  cd P[4][4];  // prefix product of the completed slices, [col][row]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) P[c][r] = mkc(c == r ? 1.0 : 0.0, 0.0);
  for (int s = 0; s < N; ++s) {
    const double ax = gx[b * N + s], ay = gy[b * N + s];
    if (gate_slice_stride && s > 0) {
      const double* gs = gp + (int64_t)s * gate_slice_stride;
      G = smush_gate(gs[0], gs[1], gs[2], gs[3], gs[4], gs[5], gs[6], gs[7]);
    }
    for (int q = 0; q < Rn; ++q) {
      // np.linspace(0, dt, R)[q]; the last point is exactly dt
      const double t = (Rn == 1) ? 0.0 : ((q == Rn - 1) ? dt : dt * ((double)q / (double)(Rn - 1)));
      const bool last = (q == Rn - 1);
      if (!coords && !last) continue;
      if (SYNC) __syncthreads();
      cd Y[4][4];
      smush_slice(G, ax, ay, t, Y);
      cd W[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) W[c][r] = P[c][r];
      left_mul(Y, W);
      if (SYNC) __syncthreads();
      if (coords) {
        cd M[4][4];  // [row][col]
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < 4; ++r) M[r][c] = W[c][r];
        double cc[3];
        weyl_makhlin(M, flags, cc, nullptr);
        if (valid) {
          double* o = coords + ((b * N + s) * Rn + q) * 3;
          o[0] = cc[0];
          o[1] = cc[1];
          o[2] = cc[2];
        }
      }
      if (last) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int r = 0; r < 4; ++r) P[c][r] = W[c][r];
      }
    }
  }
  if (Ufinal && valid) {
    double* out = Ufinal + b * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<double2*>(out + (r * 4 + c) * 2) = make_double2(P[c][r].re, P[c][r].im);
  }
}

}  // namespace slam

static int launch_trajectory(const double* gate, int gate_slice_stride, const double* gx, const double* gy, int32_t N, int32_t R,
                             double dt, int32_t flags, double* coords, double* Ufinal, int64_t B, void* stream) {
  using namespace slam;
  if (!gate || !gx || !gy || N < 1 || R < 1 || B < 0 || (!coords && !Ufinal)) return SLAM_ERR_INVALID;
  if (B == 0) return SLAM_OK;
  {  // phase-locked CTA (barriers at the slice boundaries): 2.08 vs 1.75 G trajectory points/s without them
    const unsigned grid = (unsigned)((B + 255) / 256);
    trajectory_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(gate, gate_slice_stride, gx, gy, N, R, dt, flags, coords,
                                                                     Ufinal, B);
  }
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

extern "C" int slam_pd_trajectory(const double* gate, const double* gx, const double* gy, int32_t N, int32_t R, double dt,
                                  int32_t flags, double* coords, double* Ufinal, int64_t B, void* stream) {
  return launch_trajectory(gate, 0, gx, gy, N, R, dt, flags, coords, Ufinal, B, stream);
}

extern "C" int slam_pd_trajectory_slices(const double* gate, const double* gx, const double* gy, int32_t N, int32_t R, double dt,
                                         int32_t flags, double* coords, double* Ufinal, int64_t B, void* stream) {
  return launch_trajectory(gate, 8, gx, gy, N, R, dt, flags, coords, Ufinal, B, stream);
}
