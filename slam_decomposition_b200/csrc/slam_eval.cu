// slam_eval.cu -- K1 (template_eval) and K2 (loss + analytic gradient) streaming kernels.
//
// One CTA of 128 threads handles 128/LPP problems.  Parameters are staged through shared memory
// with coalesced loads (x is [B, P] row-major, a thread-per-row read would be strided), the
// per-problem (cos, sin) cache lives in shared memory, and the gradient is staged back the same way.
#include <cstdlib>

#include "slam_host.h"

namespace slam {

constexpr int kThreads = 128;

template <int LPP, int GM, bool WANT_GRAD>
__global__ void __launch_bounds__(kThreads)
loss_grad_kernel(const double* __restrict__ x, int64_t ldx, const double* __restrict__ V, int64_t Nt,
                 const int32_t* __restrict__ tgt_idx, int cost_kind, double* __restrict__ loss, double* __restrict__ grad,
                 int64_t ldg, double* __restrict__ trace, int64_t B, const __grid_constant__ KTemplate kt) {
  constexpr int CPL = 4 / LPP;
  const int nthr = blockDim.x;
  const int TPB = nthr / LPP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = kt.P;
  const int XS = P | 1;  // odd row stride (in doubles): conflict-free 64-bit access across teams
  double* xs = reinterpret_cast<double*>(smem_raw);
  double* gs = xs + TPB * XS;
  double2* tg = reinterpret_cast<double2*>(gs + (WANT_GRAD ? TPB * XS : 0) + ((TPB * XS * (WANT_GRAD ? 2 : 1)) & 1));

  const int tid = threadIdx.x;
  const int64_t first = (int64_t)blockIdx.x * TPB;
  const int nrows = (int)min((int64_t)TPB, B - first);
  for (int e = tid; e < TPB * P; e += nthr) {
    const int row = e / P, col = e - row * P;
    xs[row * XS + col] = row < nrows ? x[(first + row) * ldx + col] : 0.0;
  }
  __syncthreads();

  const int team = tid / LPP, sub = tid % LPP;
  const bool valid = team < nrows;
  const int64_t p = first + team;
  int64_t tgt = 0;
  if (valid) tgt = tgt_idx ? (int64_t)tgt_idx[p] : (p % Nt);
  cd vcol[CPL][4];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double2 v = *reinterpret_cast<const double2*>(V + tgt * 32 + (a * 4 + sub * CPL + c) * 2);
      vcol[c][a] = mkc(v.x, v.y);
    }
  cd T;
  const double l = loss_grad_team<LPP, GM, WANT_GRAD>(kt, xs + team * XS, tg + team * kt.n_trig, gs + team * XS, vcol,
                                                      cost_kind, sub, &T);
  if (valid && sub == 0) {
    loss[p] = l;
    if (trace) {
      trace[2 * p] = T.re;
      trace[2 * p + 1] = T.im;
    }
  }
  if (WANT_GRAD) {
    __syncthreads();
    for (int e = tid; e < nrows * P; e += nthr) {
      const int row = e / P, col = e - row * P;
      grad[(first + row) * ldg + col] = gs[row * XS + col];
    }
  }
}

template <int LPP, int GM>
__global__ void __launch_bounds__(kThreads)
eval_kernel(const double* __restrict__ x, int64_t ldx, double* __restrict__ U, int64_t B,
            const __grid_constant__ KTemplate kt) {
  constexpr int CPL = 4 / LPP;
  const int nthr = blockDim.x;
  const int TPB = nthr / LPP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = kt.P;
  const int XS = P | 1;
  double* xs = reinterpret_cast<double*>(smem_raw);
  double2* tg = reinterpret_cast<double2*>(xs + TPB * XS + ((TPB * XS) & 1));
  const int tid = threadIdx.x;
  const int64_t first = (int64_t)blockIdx.x * TPB;
  const int nrows = (int)min((int64_t)TPB, B - first);
  for (int e = tid; e < TPB * P; e += nthr) {
    const int row = e / P, col = e - row * P;
    xs[row * XS + col] = row < nrows ? x[(first + row) * ldx + col] : 0.0;
  }
  __syncthreads();
  const int team = tid / LPP, sub = tid % LPP;
  double2* tgp = tg + team * kt.n_trig;
  fill_trig<LPP, GM>(kt, xs + team * XS, tgp, sub);
  __syncwarp();
  cd r[CPL][4];
  forward_chain<LPP, GM>(kt, tgp, sub, r);
  if (team < nrows) {
    double* out = U + (first + team) * 32;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int a = 0; a < 4; ++a)
        *reinterpret_cast<double2*>(out + (a * 4 + sub * CPL + c) * 2) = make_double2(r[c][a].re, r[c][a].im);
  }
}

static size_t smem_bytes(const KTemplate& kt, int lpp, bool grad, int threads) {
  const int TPB = threads / lpp;
  const int XS = kt.P | 1;
  size_t d = (size_t)TPB * XS * (grad ? 2 : 1);
  d += d & 1;
  return d * sizeof(double) + (size_t)TPB * kt.n_trig * sizeof(double2);
}

// largest CTA (<= kThreads) whose staging buffers fit the 227 KB shared-memory budget
static int pick_threads(const KTemplate& kt, int lpp, bool grad) {
  int t = kThreads;
  while (t > 32 && smem_bytes(kt, lpp, grad, t) > 200 * 1024) t >>= 1;
  return t;
}

// lanes per problem for the streaming kernels: measured on B200 (scripts/quick_bench.py, loss+grad, % of DFMA peak)
//   k=1: LPP 4/2/1 = 55/62/57 %   k=3: 71/79/65 %   k=6: 79/70/46 %   -> 2 lanes up to k=4, 4 lanes beyond
static int pick_lpp(const KTemplate& kt, int lanes) {
  if (lanes == 1 || lanes == 2 || lanes == 4) return lanes;
  return kt.k <= 4 ? 2 : 4;
}

template <int LPP, int GM>
static int launch_loss_grad(const KTemplate& kt, const double* x, int64_t ldx, const double* V, int64_t Nt,
                            const int32_t* tgt_idx, int cost_kind, double* loss, double* grad, int64_t ldg,
                            double* trace, int64_t B, cudaStream_t st) {
  const int threads = pick_threads(kt, LPP, grad != nullptr);
  const int TPB = threads / LPP;
  const unsigned grid = (unsigned)((B + TPB - 1) / TPB);
  if (grad) {
    const size_t sm = smem_bytes(kt, LPP, true, threads);
    auto kern = loss_grad_kernel<LPP, GM, true>;
    if (sm > 48 * 1024) SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<grid, threads, sm, st>>>(x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, kt);
  } else {
    const size_t sm = smem_bytes(kt, LPP, false, threads);
    auto kern = loss_grad_kernel<LPP, GM, false>;
    if (sm > 48 * 1024) SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<grid, threads, sm, st>>>(x, ldx, V, Nt, tgt_idx, cost_kind, loss, nullptr, 0, trace, B, kt);
  }
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

template <int LPP>
static int dispatch_loss_grad(const KTemplate& kt, const double* x, int64_t ldx, const double* V, int64_t Nt,
                              const int32_t* tgt_idx, int cost_kind, double* loss, double* grad, int64_t ldg,
                              double* trace, int64_t B, cudaStream_t st) {
  switch (kt.gmode) {
    case GM_SYM: return launch_loss_grad<LPP, GM_SYM>(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
    case GM_BLOCK: return launch_loss_grad<LPP, GM_BLOCK>(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
    case GM_DENSE: return launch_loss_grad<LPP, GM_DENSE>(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
    default: return SLAM_ERR_UNSUPPORTED;
  }
}

template <int LPP, int GM>
static int launch_eval(const KTemplate& kt, const double* x, int64_t ldx, double* U, int64_t B, cudaStream_t st) {
  const int threads = pick_threads(kt, LPP, false);
  const int TPB = threads / LPP;
  const unsigned grid = (unsigned)((B + TPB - 1) / TPB);
  const size_t sm = smem_bytes(kt, LPP, false, threads);
  auto kern = eval_kernel<LPP, GM>;
  if (sm > 48 * 1024) SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<grid, threads, sm, st>>>(x, ldx, U, B, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

int smush_eval_launch(const KTemplate& kt, const double* x, int64_t ldx, double* U, int64_t B, cudaStream_t st);
int smush_loss_grad_launch(const KTemplate& kt, const double* x, int64_t ldx, const double* V, int64_t Nt,
                           const int32_t* tgt_idx, int cost_kind, double* loss, double* grad, int64_t ldg, double* trace,
                           int64_t B, cudaStream_t st);

}  // namespace slam

using namespace slam;

extern "C" int slam_template_eval(const SlamTemplateDesc* desc, const double* x, int64_t ldx, double* U, int64_t B,
                                  void* stream) {
  if (!desc || B < 0 || ldx < desc->n_params) return SLAM_ERR_INVALID;
  if (B == 0) return SLAM_OK;
  if (!U || (desc->n_params > 0 && !x)) return SLAM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  KTemplate kt;
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/true, /*allow_ties=*/true);
  if (rc != SLAM_OK) return rc;
  if (kt.gmode == GM_SMUSH) return smush_eval_launch(kt, x, ldx, U, B, st);
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  switch (kt.gmode) {
    case GM_SYM: return launch_eval<4, GM_SYM>(kt, x, ldx, U, B, st);
    case GM_BLOCK: return launch_eval<4, GM_BLOCK>(kt, x, ldx, U, B, st);
    case GM_DENSE: return launch_eval<4, GM_DENSE>(kt, x, ldx, U, B, st);
    default: return SLAM_ERR_UNSUPPORTED;
  }
}

extern "C" int slam_loss_grad(const SlamTemplateDesc* desc, const double* x, int64_t ldx, const double* V, int64_t Nt,
                              const int32_t* tgt_idx, int32_t cost_kind, double* loss, double* grad, int64_t ldg,
                              double* trace, int64_t B, void* stream) {
  return slam_loss_grad_lanes(desc, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, 0, stream);
}

extern "C" int slam_loss_grad_lanes(const SlamTemplateDesc* desc, const double* x, int64_t ldx, const double* V, int64_t Nt,
                                    const int32_t* tgt_idx, int32_t cost_kind, double* loss, double* grad, int64_t ldg,
                                    double* trace, int64_t B, int32_t lanes, void* stream) {
  if (!desc || B < 0 || ldx < desc->n_params) return SLAM_ERR_INVALID;
  if (lanes != 0 && lanes != 1 && lanes != 2 && lanes != 4) return SLAM_ERR_INVALID;
  if (grad && ldg < desc->n_params) return SLAM_ERR_INVALID;
  if (cost_kind < SLAM_COST_BASIC || cost_kind > SLAM_COST_BASIC_INVERSE) return SLAM_ERR_INVALID;  // optimizer.py:211
  if (B == 0) return SLAM_OK;  // empty batch: nothing to read or write
  if (!V || !loss || Nt <= 0 || (desc->n_params > 0 && !x)) return SLAM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  KTemplate kt;
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/true);
  if (rc != SLAM_OK) return rc;
  if (kt.gmode == GM_SMUSH)  // parameter-bound smush gates: thread-per-row adjoint through the slice exponentials
    return smush_loss_grad_launch(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  switch (pick_lpp(kt, lanes)) {
    case 1: return dispatch_loss_grad<1>(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
    case 2: return dispatch_loss_grad<2>(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
    default: return dispatch_loss_grad<4>(kt, x, ldx, V, Nt, tgt_idx, cost_kind, loss, grad, ldg, trace, B, st);
  }
}
