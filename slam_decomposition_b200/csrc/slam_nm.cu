// slam_nm.cu -- K5b: batched derivative-free optimiser (Nelder-Mead) over (target, restart) problems with a
// GENERIC objective: any template the forward evaluator handles (closed-form, fixed, smush gates; U3 or RZ layers)
// against any in-scope cost functional (trace-based, Weyl/Makhlin-coordinate based, canonical-reduced).
//
// Replaces scipy.optimize.minimize(method="Nelder-Mead") as selected through TemplateOptimizer(override_method=...)
// (src/slam/optimizer.py:266-278; used for the Makhlin / Weyl functionals in scripts/cost_function_comparison.ipynb
// and for the drive-amplitude searches of scripts/parallel_drive_swap).  The simplex update follows scipy's
// _minimize_neldermead (rho=1, chi=2, psi=0.5, sigma=0.5, initial simplex 5% / 0.00025, xatol = fatol = 1e-4,
// termination on simplex size AND function spread, or maxiter).
//
// One thread per problem (the objective is forward-only and cheap; there is no shuffle-level parallelism to gain),
// persistent grid with a global work counter, restart-major order with early exit like K5.  The simplex lives in
// a global-memory workspace interleaved across the threads of a CTA (element e of thread t at chunk[e * 128 + t]) so
// every vector operation is a coalesced stream.  A tick-structured, phase-locked form of this kernel (one objective evaluation per
// thread per tick, as in the adjoint K5c kernel) was written and measured: bit-identical results, but 183 against 208 M
// objective evaluations/s on 131072 x 4 Makhlin problems (sqrt(iSWAP) k=3) -- a simplex step is one or two cheap evaluations
// between O(n) vector updates, so the CTA-wide barrier per evaluation costs more than the divergence it removes -- and dropped.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "slam_host.h"
#include "slam_objective.cuh"
#include "slam_philox.cuh"

namespace slam {

struct NmArgs {
  const double* V;
  const double* x0;
  int64_t ldx0;
  uint64_t seed;
  const int32_t* active;
  int64_t Nt;
  int restarts, max_iter, cost_kind, early_exit;
  double success_threshold, xatol, fatol, x0_lo, x0_span;
  double* out_loss;
  double* out_x;
  int32_t* out_iters;
  unsigned long long* out_evals;
  unsigned long long* next;
  int32_t* solved;
  double* ws;   // workspace: (n + 4) vectors of n doubles + (n + 1) function values per thread, interleaved per CTA
  int64_t T;    // threads in the grid
};

__global__ void __launch_bounds__(128) nm_kernel(const __grid_constant__ NmArgs A, const __grid_constant__ KTemplate kt) {
  const int n = kt.P;
  // CTA-local interleave: element e of thread t of this CTA at chunk[e * 128 + t] -- still one coalesced line per warp
  // access, and the element index is 32-bit arithmetic with a constant stride (with the grid-wide stride A.T every access
  // cost a 64-bit multiply; in the K5c adjoint kernel that addressing was 19 % of all executed instructions)
  constexpr int T = 128;
  const int per_thread = (n + 4) * n + (n + 1);
  double* ws = A.ws + (int64_t)blockIdx.x * ((int64_t)per_thread * T) + threadIdx.x;
  // workspace vectors (each n entries, entry j at [(vec * n + j) * T]): simplex vertices 0..n, then xr, xe/xc, sum
  auto vec = [&](int v, int j) -> double& { return ws[(v * n + j) * T]; };
  const int V_XR = n + 1, V_XT = n + 2, V_SUM = n + 3;
  double* fs = ws + (n + 4) * n * T;  // fs[v * T], v = 0..n
  const int64_t total = A.Nt * (int64_t)A.restarts;
  unsigned long long evals = 0;

  while (true) {
    const unsigned long long w = atomicAdd(A.next, 1ULL);
    if ((int64_t)w >= total) break;
    const int64_t r_idx = (int64_t)w / A.Nt, t = (int64_t)w - r_idx * A.Nt;
    const int64_t pid = t * A.restarts + r_idx;
    bool skip = A.active && A.active[t] == 0;
    if (!skip && A.early_exit) skip = *((volatile int32_t*)(A.solved + t)) != 0;
    if (skip) {
      A.out_loss[pid] = DBL_MAX;
      A.out_iters[pid] = 0;
      for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = 0.0;
      continue;
    }
    TargetInfo ti;
    target_info_init(ti, A.V + t * 32, A.cost_kind);
    auto f_at = [&](int v) -> double {
      ++evals;
      return objective_value_nl(&kt, &vec(v, 0), T, -1, 0.0, &ti, A.cost_kind);
    };
    // initial simplex (scipy: y[k] *= 1.05, or 0.00025 if zero)
    for (int j = 0; j < n; ++j) {
      const double x = A.x0 ? A.x0[pid * A.ldx0 + j] : philox_param(A.seed, (uint64_t)pid, j, A.x0_lo, A.x0_span);
      for (int v = 0; v <= n; ++v) vec(v, j) = x;
    }
    for (int k = 0; k < n; ++k) {
      const double x = vec(0, k);
      vec(k + 1, k) = (x != 0.0) ? 1.05 * x : 0.00025;
    }
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int v = 0; v <= n; ++v) s += vec(v, j);
      vec(V_SUM, j) = s;
    }
    for (int v = 0; v <= n; ++v) fs[v * T] = f_at(v);

    int iter = 0;
    int ib = 0;
    while (true) {
      // best, worst, second worst
      int iw = 0;
      double fb = DBL_MAX, fw = -DBL_MAX, f2 = -DBL_MAX;
      for (int v = 0; v <= n; ++v) {
        const double f = fs[v * T];
        if (f < fb) { fb = f; ib = v; }
        if (f > fw) { f2 = fw; fw = f; iw = v; }
        else if (f > f2) { f2 = f; }
      }
      // termination (scipy): max |sim[1:] - sim[0]| <= xatol and max |fsim[0] - fsim[1:]| <= fatol
      // (the O(n^2) simplex-diameter scan only runs once the function-value spread has passed its own test)
      double dxmax = DBL_MAX;
      if ((fw - fb) <= A.fatol) {
        dxmax = 0.0;
        for (int j = 0; j < n; ++j) {
          const double xb = vec(ib, j);
          for (int v = 0; v <= n; ++v) dxmax = fmax(dxmax, fabs(vec(v, j) - xb));
        }
      }
      bool stop = (dxmax <= A.xatol && (fw - fb) <= A.fatol) || iter >= A.max_iter || !(fb == fb);
      if (!stop && fb < A.success_threshold * 1e-3) stop = true;  // far below the success threshold: nothing left to gain
      if (!stop && A.early_exit && (iter & 15) == 0) stop = *((volatile int32_t*)(A.solved + t)) != 0;
      if (stop) break;
      ++iter;
      const double inv_n = 1.0 / (double)n;
      // reflection: xr = (1 + rho) xbar - rho xw, xbar = (sum - xw) / n
      for (int j = 0; j < n; ++j) {
        const double xw = vec(iw, j), xbar = (vec(V_SUM, j) - xw) * inv_n;
        vec(V_XR, j) = 2.0 * xbar - xw;
      }
      const double fxr = f_at(V_XR);
      int take = -1;  // which trial vector replaces the worst vertex
      double ftake = 0.0;
      bool shrink = false;
      if (fxr < fb) {
        for (int j = 0; j < n; ++j) {  // expansion: (1 + rho chi) xbar - rho chi xw
          const double xw = vec(iw, j), xbar = (vec(V_SUM, j) - xw) * inv_n;
          vec(V_XT, j) = 3.0 * xbar - 2.0 * xw;
        }
        const double fxe = f_at(V_XT);
        if (fxe < fxr) { take = V_XT; ftake = fxe; } else { take = V_XR; ftake = fxr; }
      } else if (fxr < f2) {
        take = V_XR; ftake = fxr;
      } else if (fxr < fw) {
        for (int j = 0; j < n; ++j) {  // outside contraction: (1 + psi rho) xbar - psi rho xw
          const double xw = vec(iw, j), xbar = (vec(V_SUM, j) - xw) * inv_n;
          vec(V_XT, j) = 1.5 * xbar - 0.5 * xw;
        }
        const double fxc = f_at(V_XT);
        if (fxc <= fxr) { take = V_XT; ftake = fxc; } else shrink = true;
      } else {
        for (int j = 0; j < n; ++j) {  // inside contraction: (1 - psi) xbar + psi xw
          const double xw = vec(iw, j), xbar = (vec(V_SUM, j) - xw) * inv_n;
          vec(V_XT, j) = 0.5 * xbar + 0.5 * xw;
        }
        const double fxcc = f_at(V_XT);
        if (fxcc < fw) { take = V_XT; ftake = fxcc; } else shrink = true;
      }
      if (!shrink) {
        for (int j = 0; j < n; ++j) {
          const double xn = vec(take, j), xw = vec(iw, j);
          vec(V_SUM, j) += xn - xw;
          vec(iw, j) = xn;
        }
        fs[iw * T] = ftake;
      } else {
        for (int v = 0; v <= n; ++v) {
          if (v == ib) continue;
          for (int j = 0; j < n; ++j) vec(v, j) = vec(ib, j) + 0.5 * (vec(v, j) - vec(ib, j));
          fs[v * T] = f_at(v);
        }
        for (int j = 0; j < n; ++j) {
          double s = 0.0;
          for (int v = 0; v <= n; ++v) s += vec(v, j);
          vec(V_SUM, j) = s;
        }
      }
    }
    const double fbest = fs[ib * T];
    A.out_loss[pid] = fbest;
    A.out_iters[pid] = iter;
    for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = vec(ib, j);
    if (A.early_exit && fbest < A.success_threshold) atomicExch(A.solved + t, 1);
  }
  if (A.out_evals && evals) atomicAdd(A.out_evals, evals);
}

}  // namespace slam

using namespace slam;

extern "C" void slam_nm_defaults(SlamNmOpts* o) {
  if (!o) return;
  o->max_iter = 2500;  // optimizer.py:274 options={"maxiter": 2500}
  o->cost_kind = SLAM_COST_BASIC;
  o->early_exit = 1;
  o->reserved = 0;
  o->success_threshold = 1e-10;
  o->xatol = 1e-4;  // scipy Nelder-Mead defaults
  o->fatol = 1e-4;
  o->x0_lo = 0.0;
  o->x0_hi = 6.283185307179586;
}

extern "C" int slam_nm_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts, const double* x0,
                             int64_t ldx0, uint64_t seed, const int32_t* active, const SlamNmOpts* opts, double* out_loss,
                             double* out_x, int32_t* out_iters, unsigned long long* out_evals, void* stream) {
  if (!desc || !V || !opts || !out_loss || !out_x || !out_iters || Nt < 0 || restarts < 1) return SLAM_ERR_INVALID;
  if (x0 && ldx0 < desc->n_params) return SLAM_ERR_INVALID;
  if (opts->cost_kind < SLAM_COST_BASIC || opts->cost_kind > SLAM_COST_SQUARE_REDUCED) return SLAM_ERR_INVALID;
  if (opts->max_iter < 1 || desc->n_params < 1) return SLAM_ERR_INVALID;
  if (Nt == 0) return SLAM_OK;
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
  const int n = kt.P;
  const int64_t total = Nt * (int64_t)restarts;
  const int threads = 128;
  int64_t blocks = std::min<int64_t>((int64_t)sms * 2, (total + threads - 1) / threads);
  const size_t per_thread = ((size_t)(n + 4) * n + (n + 1)) * sizeof(double);
  while (blocks > 1 && per_thread * threads * (size_t)blocks > ((size_t)4 << 30)) blocks /= 2;  // workspace <= 4 GiB
  const int64_t T = blocks * threads;

  Scratch scratch(st);
  unsigned long long* next = nullptr;
  int32_t* solved = nullptr;
  double* ws = nullptr;
  if ((rc = scratch.alloc(&next, sizeof(unsigned long long), true)) != SLAM_OK) return rc;
  if ((rc = scratch.alloc(&solved, sizeof(int32_t) * (size_t)Nt, true)) != SLAM_OK) return rc;
  if ((rc = scratch.alloc(&ws, per_thread * (size_t)T)) != SLAM_OK) return rc;

  NmArgs A;
  A.V = V; A.x0 = x0; A.ldx0 = ldx0; A.seed = seed; A.active = active; A.Nt = Nt; A.restarts = restarts;
  A.max_iter = opts->max_iter; A.cost_kind = opts->cost_kind; A.early_exit = opts->early_exit;
  A.success_threshold = opts->success_threshold; A.xatol = opts->xatol; A.fatol = opts->fatol;
  A.x0_lo = opts->x0_lo; A.x0_span = opts->x0_hi - opts->x0_lo;
  A.out_loss = out_loss; A.out_x = out_x; A.out_iters = out_iters; A.out_evals = out_evals;
  A.next = next; A.solved = solved; A.ws = ws; A.T = T;
  nm_kernel<<<(unsigned)blocks, threads, 0, st>>>(A, kt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e, "nm_kernel launch");
    return SLAM_ERR_CUDA;
  }
  return SLAM_OK;
}
