// slam_fdopt.cu -- K5c: batched thread-per-problem quasi-Newton optimiser over (target, restart) problems for the templates
// the team kernels of slam_core.cuh cannot differentiate in closed form (parameter-bound smush gates,
// hamiltonian.py:114-182) and the trace-based functionals incl. BasicCostInverse x circuit_fidelity.  Two kernels:
//   * fd_lbfgs_kernel  -- FINITE-DIFFERENCE gradients (forward or central), sequential control flow per thread;
//   * adj_lbfgs_kernel -- ANALYTIC adjoint gradients through the smush slices (slam_adj1.cuh), tick-structured and
//                         phase-locked (bottom of this file); the default for smush templates.
//
// The finite-difference form is literally the reference's algorithm class: scipy.optimize.minimize(method="BFGS", jac=None) builds the gradient
// from P forward differences with step sqrt(eps) = 1.49e-8 (src/slam/optimizer.py:270-278).  Here every (target, restart)
// pair is one thread: P + 1 forward evaluations of the whole template (time-sliced exponentials included) per gradient,
// limited-memory BFGS with the same Armijo / cautious-update / stopping rules as K5 (slam_lbfgs.cuh), optional box
// bounds by projection.  Central differences (2P evaluations, step 6e-6) are selectable for a ~1000x cleaner gradient.
//
// One thread per problem, persistent grid with a global work counter, restart-major order with early exit like K5.
// The optimiser state (x, g, trial x, trial g, direction, m (s, y) pairs) lives in a global-memory workspace
// interleaved across threads (element e of thread t at ws[e * T + t]) so every vector operation is a coalesced
// stream out of L2; the objective evaluations dominate by three orders of magnitude.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "slam_host.h"
#include "slam_adj1.cuh"
#include "slam_objective.cuh"
#include "slam_philox.cuh"

namespace slam {

constexpr int kFdHist = 8;
constexpr double kArmijoFd = 1e-4;

struct FdArgs {
  const double* V;
  const double* x0;
  int64_t ldx0;
  uint64_t seed;
  const int32_t* active;
  int64_t Nt;
  int restarts, max_iter, cost_kind, early_exit, central, debug;
  int m;  // (s, y) pairs kept by the adjoint kernels (<= kAdjHist); the finite-difference kernel keeps kFdHist
  double success_threshold, f_stop, gtol, gtol_far, f_far, x0_lo, x0_span;
  const double* lower;
  const double* upper;
  double con_max, con_mu;    // circuit-cost constraint (FD modes); con_mu = 0: none
  const double* con_lambda;  // [Nt * restarts] multipliers or null
  double* out_loss;
  double* out_x;
  int32_t* out_iters;
  unsigned long long* out_evals;
  unsigned long long* next;
  int32_t* solved;
  double* ws;  // (5 + 2 m) vectors of n doubles per thread, interleaved
  int64_t T;
};

// out-of-line, phase-locked copy of the adjoint pass (loss + gradient of one workspace vector) for the tick kernel: every
// thread of the CTA calls it in the same tick (CTA-wide barriers inside)
static __device__ __noinline__ double adj1_nl_sync_recompute(const KTemplate* kt, const double* p, int64_t stride, const double* V,
                                                             int cost_kind, double* g, int64_t gstride) {
  StridedParams ps{p, stride};
  StridedGrad gsw{g, gstride};
  return adj1_loss_grad<StridedParams, StridedGrad, true>(*kt, ps, V, cost_kind, gsw, nullptr);
}
// eigen-form forward sweep handing (Q, lam) of every slice to the backward sweep (slam_adj1.cuh); k T <= kEigSlices
static __device__ __noinline__ double adj1_nl_sync_eig(const KTemplate* kt, const double* p, int64_t stride, const double* V,
                                                       int cost_kind, double* g, int64_t gstride) {
  StridedParams ps{p, stride};
  StridedGrad gsw{g, gstride};
  return adj1_loss_grad<StridedParams, StridedGrad, true, LocalEigStore>(*kt, ps, V, cost_kind, gsw, nullptr);
}
static __device__ __forceinline__ double adj1_nl_sync(const KTemplate* kt, const double* p, int64_t stride, const double* V,
                                                      int cost_kind, double* g, int64_t gstride) {
  if (kt->k * kt->T <= kEigSlices) return adj1_nl_sync_eig(kt, p, stride, V, cost_kind, g, gstride);  // (uniform over the grid)
  return adj1_nl_sync_recompute(kt, p, stride, V, cost_kind, g, gstride);
}

// MODE 0: forward differences (scipy's jac=None), 1: central differences.  (The analytic adjoint mode, central = 2 at the
// C ABI, has its own tick-structured kernel below.)
template <int MODE>
__global__ void __launch_bounds__(128) fd_lbfgs_kernel(const __grid_constant__ FdArgs A, const __grid_constant__ KTemplate kt) {
  const int n = kt.P;
  const int m = kFdHist;
  const int64_t tidg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t T = A.T;
  double* ws = A.ws + tidg;
  auto vec = [&](int v, int j) -> double& { return ws[((int64_t)v * n + j) * T]; };
  // vector ids: (X, G) double-buffered, direction, then the history
  const int V_D = 4, V_S = 5, V_Y = 5 + m;
  const int64_t total = A.Nt * (int64_t)A.restarts;
  unsigned long long evals = 0;
  const double h_fwd = 1.4901161193847656e-08;  // scipy: sqrt(machine epsilon), absolute step
  const double h_cen = 6.0554544523933395e-06;  // cbrt(machine epsilon)

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
    if (A.con_mu > 0.0) {
      ti.pen_mu = A.con_mu;
      ti.pen_max = A.con_max;
      ti.pen_lambda = A.con_lambda ? A.con_lambda[pid] : 0.0;
    }

    auto f_at = [&](int v) -> double {
      ++evals;
      return objective_value_nl(&kt, &vec(v, 0), T, -1, 0.0, &ti, A.cost_kind);
    };
    // gradient of buffer xb into buffer gb (fx = objective at xb); returns max |projected g|
    auto grad_at = [&](int xb, int gb, double fx) -> double {
      double gmax = 0.0;
      for (int j = 0; j < n; ++j) {
        double gj;
        if (MODE == 1) {
          gj = (objective_value_nl(&kt, &vec(xb, 0), T, j, h_cen, &ti, A.cost_kind) -
                objective_value_nl(&kt, &vec(xb, 0), T, j, -h_cen, &ti, A.cost_kind)) / (2.0 * h_cen);
          evals += 2;
        } else {
          gj = (objective_value_nl(&kt, &vec(xb, 0), T, j, h_fwd, &ti, A.cost_kind) - fx) / h_fwd;
          ++evals;
        }
        vec(gb, j) = gj;
        double gp = gj;
        if (A.lower) {
          const double xj = vec(xb, j);
          if ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0)) gp = 0.0;
        }
        gmax = fmax(gmax, fabs(gp));
      }
      return gmax;
    };

    int cur = 0;  // (X, G) = vectors (2 cur, 2 cur + 1); trial = the other pair
    for (int j = 0; j < n; ++j) {
      double x = A.x0 ? A.x0[pid * A.ldx0 + j] : philox_param(A.seed, (uint64_t)pid, j, A.x0_lo, A.x0_span);
      if (A.lower) x = fmin(fmax(x, A.lower[j]), A.upper[j]);
      vec(0, j) = x;
    }
    double f = f_at(0);
    double gmax = grad_at(0, 1, f);
    int iter = 0, hcount = 0, hpos = 0;
    double rho[kFdHist], alp[kFdHist];
    double gamma = 1.0, f_chk = f;
    bool slow = false;

    int reason = 0;  // diagnostic stop code (SLAM_B200_FD_DEBUG=1 stores it in bits 24.. of out_iters)
    while (true) {
      if (f < A.f_stop) reason = 1;
      else if (gmax < A.gtol) reason = 2;
      else if (gmax < A.gtol_far && (f > A.f_far || slow)) reason = 3;
      else if (iter >= A.max_iter) reason = 4;
      else if (!(f == f)) reason = 5;
      else if (A.early_exit && (iter & 3) == 0 && *((volatile int32_t*)(A.solved + t)) != 0) reason = 6;
      if (reason) break;
      const int X = 2 * cur, G = X + 1, XT = 2 * (cur ^ 1), GT = XT + 1;
      // ---- direction: two-loop recursion on the (projected) gradient -------------------------------
      double gg = 0.0;
      for (int j = 0; j < n; ++j) {
        double gj = vec(G, j);
        if (A.lower) {
          const double xj = vec(X, j);
          if ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0)) gj = 0.0;
        }
        vec(V_D, j) = gj;
        gg = fma(gj, gj, gg);
      }
      for (int hh = 0; hh < hcount; ++hh) {
        int slot = hpos - 1 - hh;
        if (slot < 0) slot += m;
        double a = 0.0;
        for (int j = 0; j < n; ++j) a = fma(vec(V_S + slot, j), vec(V_D, j), a);
        a *= rho[slot];
        alp[slot] = a;
        for (int j = 0; j < n; ++j) vec(V_D, j) = fma(-a, vec(V_Y + slot, j), vec(V_D, j));
      }
      if (hcount > 0)
        for (int j = 0; j < n; ++j) vec(V_D, j) *= gamma;
      for (int hh = hcount - 1; hh >= 0; --hh) {
        int slot = hpos - 1 - hh;
        if (slot < 0) slot += m;
        double b = 0.0;
        for (int j = 0; j < n; ++j) b = fma(vec(V_Y + slot, j), vec(V_D, j), b);
        const double c = alp[slot] - b * rho[slot];
        for (int j = 0; j < n; ++j) vec(V_D, j) = fma(c, vec(V_S + slot, j), vec(V_D, j));
      }
      double gd = 0.0;
      for (int j = 0; j < n; ++j) {
        const double d = -vec(V_D, j);
        vec(V_D, j) = d;
        gd = fma(vec(G, j), d, gd);
      }
      double alpha = 1.0;
      if (hcount == 0 || !(gd < 0.0)) {  // first step or not a descent direction: steepest descent, unit length
        hcount = 0;
        for (int j = 0; j < n; ++j) {
          double gj = vec(G, j);
          if (A.lower) {
            const double xj = vec(X, j);
            if ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0)) gj = 0.0;
          }
          vec(V_D, j) = -gj;
        }
        gd = -gg;
        alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
      }
      if (!(gd < 0.0)) {  // zero (projected) gradient
        reason = 7;
        break;
      }
      // ---- Armijo backtracking; quadratic interpolation from (f, gd, ft), safeguarded to [0.1, 0.5] alpha -------
      double ft = f;
      bool accepted = false;
      int ls = 0;
      while (true) {
        double gde = 0.0;  // directional derivative along the (projected) segment
        for (int j = 0; j < n; ++j) {
          double v = fma(alpha, vec(V_D, j), vec(X, j));
          if (A.lower) v = fmin(fmax(v, A.lower[j]), A.upper[j]);
          vec(XT, j) = v;
          gde = fma(vec(G, j), v - vec(X, j), gde);
        }
        ft = f_at(XT);
        if (ft <= f + kArmijoFd * gde) {
          accepted = true;
          break;
        }
        const double den = 2.0 * (ft - f - gde);
        double an = 0.5 * alpha;
        if (ft == ft && den > 0.0) an = -gde * alpha / den;
        if (!(an == an)) an = 0.5 * alpha;
        alpha = fmin(fmax(an, 0.1 * alpha), 0.5 * alpha);
        if (++ls > 30) {
          if (hcount > 0) {  // curvature model is bad: restart from steepest descent
            hcount = 0;
            for (int j = 0; j < n; ++j) {
              double gj = vec(G, j);
              if (A.lower) {
                const double xj = vec(X, j);
                if ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0)) gj = 0.0;
              }
              vec(V_D, j) = -gj;
            }
            alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
            ls = 0;
          } else {
            break;  // no progress possible at the accuracy of the differences
          }
        }
      }
      if (!accepted) {
        reason = 8;
        break;
      }
      // ---- accept: gradient at the new point, history pair ---------------------------------------------
      const double gmax_t = grad_at(XT, GT, ft);
      double sy = 0.0, yy = 0.0;
      for (int j = 0; j < n; ++j) {
        const double s = vec(XT, j) - vec(X, j), y = vec(GT, j) - vec(G, j);
        vec(V_S + hpos, j) = s;
        vec(V_Y + hpos, j) = y;
        sy = fma(s, y, sy);
        yy = fma(y, y, yy);
      }
      if (sy > 1e-14 * yy && yy > 0.0) {  // cautious update
        rho[hpos] = 1.0 / sy;
        gamma = sy / yy;
        hpos = (hpos + 1 == m) ? 0 : hpos + 1;
        hcount = min(hcount + 1, m);
      } else if (hcount == m) {
        hcount = m - 1;
      }
      ++iter;
      cur ^= 1;
      f = ft;
      gmax = gmax_t;
      // progress checkpoint every 32 accepted steps.  Much looser than K5's (4x): without exterior 1Q gates these
      // landscapes have near-singular directions and the last decades towards zero are won linearly (measured on the
      // sqiSwap k=2 smush template: with K5's rule every in-basin restart stopped between 1e-6 and 1e-9)
      if ((iter & 31) == 0) {
        slow = f > 0.97 * f_chk;
        f_chk = f;
      }
    }
    A.out_loss[pid] = f;
    A.out_iters[pid] = A.debug ? (iter | (reason << 24)) : iter;
    for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = vec(2 * cur, j);
    if (A.early_exit && f < A.success_threshold) atomicExch(A.solved + t, 1);
  }
  if (A.out_evals && evals) atomicAdd(A.out_evals, evals);
}

// ------------------------------------------------------------------------------------------------------------------
// K5c, adjoint mode: the same optimiser in TICK form (as K5, slam_lbfgs.cuh).  Every tick each thread performs exactly one
// loss + adjoint-gradient evaluation at its trial point -- the warp stays convergent through the 20k-instruction
// evaluation -- and then does its own (cheap, divergent) accept / backtrack bookkeeping.  The sequential form above,
// with a separate forward evaluation per line-search trial, ran the two kinds of evaluation of a warp's lanes one after
// the other (measured: 32 M evaluations/s against 183 M/s for the streaming K2 kernel on the same template).
// Evaluating the gradient at every trial also makes the cubic interpolation of K5 available and saves the re-evaluation
// of an accepted point.
// ------------------------------------------------------------------------------------------------------------------
enum { AST_IDLE = 0, AST_INIT = 1, AST_LS = 2 };

// Bookkeeping layout.  (x, g) double-buffered and the (s, y) history stay in the interleaved global workspace; the working
// vector of the two-loop recursion / the search direction `dv` and two scratch vectors live in THREAD-LOCAL arrays.  That
// is what makes the bookkeeping cheap: with everything behind one global pointer the compiler must keep every load after
// the previous store (possible aliasing), which serialised ~1700 L2 round trips per tick and left the evaluations with 15 %
// of the time.  Every loop below either only loads from global memory (pipelined, unrolled) or only stores to it, and the
// axpy of one history pair is fused with the dot product of the next (m + 1 dependent passes per loop instead of 2 m).
// A line search that fails (31 backtracks) restarts from steepest descent.  On a landscape where only tiny steps succeed that
// cycle -- SD step accepted after many backtracks, next quasi-Newton step rejected 31 times -- costs ~60 evaluations per
// iteration for up to max_iter iterations, and the phase-locked CTA waits for it (measured: one such restart turned a 0.1 s
// launch of the configs[3] grid into 4 s).  scipy's BFGS stops at its FIRST line-search failure; a few more are allowed here.
constexpr int kMaxLsFail = 4;
// A restart at a clearly non-zero loss (f > f_far) that, over a window of 32 accepted iterations, improved the loss by < 3 %
// AND needed more than kStallEvals evaluations per iteration (i.e. nearly every line search backtracked many times) is
// crawling along a ridge: it ends with stop reason 9.  (Restarts creeping towards a zero of the loss are left alone.)  Measured on the configs[3] grid: 3 of 196608 restarts did this for all 2500 iterations at ~19 evaluations
// each (48 k evaluations, 3 - 6 s of a phase-locked CTA waiting for ONE thread, against 0.1 - 0.3 s for the whole launch).
// scipy ends such a restart at its first failed line search (maxls = 20 evaluations).
constexpr int kStallEvals = 6;
constexpr int kAdjCta = 256;  // one CTA per SM (the evaluation needs ~255 registers), phase-locked
constexpr int kAdjHist = kFdHist;  // (s, y) pairs kept, in double.  Six float pairs were measured: +16 % evaluations/s (the
                                   // history loads are what the bookkeeping waits for: 52 % of the stall samples, 109 MB of
                                   // workspace against a 126 MB L2) but +21 % evaluations on the near-singular smush
                                   // landscapes and fewer solved targets -- no net gain, so the history stays exact

// how many of the four (x0, g0, x1, g1) vectors of the register-form adjoint kernel fit in shared memory (NQ entries per
// thread each, 256 threads, 227 KB)
// all four or none: with three (NQ = 32) the mixed shared / global buffer selection costs more than it saves
// (sqrt(iSWAP) k = 3, P = 30: 1489 ms against 1310 ms with all four in the global chunk)
__host__ __device__ constexpr int adj_smem_vecs(int nq) { return 4 * nq * kAdjCta * 8 <= 227 * 1024 ? 4 : 0; }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NQ>
__global__ void __launch_bounds__(kAdjCta, 1) adj_lbfgs_kernel(const __grid_constant__ FdArgs A, const __grid_constant__ KTemplate kt) {
  const int n = kt.P;
  const int m = A.m;
  // CTA-local interleave of the workspace (entry j of a vector of thread t at [j * 256 + t]): constant stride, 32-bit indices
  constexpr int64_t T = kAdjCta;
  double* ws = A.ws + (int64_t)blockIdx.x * ((int64_t)(5 + 2 * kFdHist) * n * T) + threadIdx.x;
  auto vecp = [&](int v) -> double* { return ws + (int64_t)v * n * T; };  // entry j at [j * T]
  // history behind the four (x, g) vectors: pair slot i at hs(i) / hy(i), entry j at [j * T]
  auto hs = [&](int slot) -> double* { return ws + (int64_t)(4 + slot) * n * T; };
  auto hy = [&](int slot) -> double* { return ws + (int64_t)(4 + m + slot) * n * T; };
  const int64_t total = A.Nt * (int64_t)A.restarts;
  const bool bounded = A.lower != nullptr;

  int state = AST_IDLE, cur = 0, iter = 0, hcount = 0, hpos = 0, ls = 0;
  bool exhausted = false, slow = false, dir_sd = false;
  int pevals = 0;  // evaluations of the current problem (diagnostics)
  int nfail = 0;   // failed line searches of the current problem
  int pe_chk = 0;  // pevals at the last progress checkpoint
  int64_t pid = 0, t = 0;
  double f = 0.0, alpha = 1.0, gde = 0.0, gamma = 1.0, f_chk = 0.0;
  double rho[kAdjHist], alp[kAdjHist];
  double dv[NQ], t1[NQ], t2[NQ];
  unsigned long long evals = 0;
  for (int v = 0; v < 4; ++v) {
    double* p = vecp(v);
    for (int j = 0; j < n; ++j) p[j * T] = 0.0;  // idle lanes evaluate their (finite) trial buffer
  }

  while (true) {
    // ---------------- fetch -------------------------------------------------------------------------------
    while (state == AST_IDLE && !exhausted) {
      const unsigned long long w = atomicAdd(A.next, 1ULL);
      if ((int64_t)w >= total) {
        exhausted = true;
        break;
      }
      const int64_t r_idx = (int64_t)w / A.Nt;
      t = (int64_t)w - r_idx * A.Nt;
      pid = t * A.restarts + r_idx;
      bool skip = A.active && A.active[t] == 0;
      if (!skip && A.early_exit) skip = *((volatile int32_t*)(A.solved + t)) != 0;
      if (skip) {
        A.out_loss[pid] = DBL_MAX;
        A.out_iters[pid] = 0;
        for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = 0.0;
        continue;
      }
      cur = 0;  // trial buffer = vectors (2, 3)
      double* x1 = vecp(2);
      for (int j = 0; j < n; ++j) {
        double x = A.x0 ? A.x0[pid * A.ldx0 + j] : philox_param(A.seed, (uint64_t)pid, j, A.x0_lo, A.x0_span);
        if (bounded) x = fmin(fmax(x, A.lower[j]), A.upper[j]);
        x1[j * T] = x;
      }
      state = AST_INIT;
      iter = 0; ls = 0; hcount = 0; hpos = 0; gamma = 1.0; slow = false; pevals = 0; nfail = 0; pe_chk = 0;
    }
    // CTA-wide vote = the tick barrier: the CTA's warps enter the evaluation together and, with the barriers inside it,
    // walk through its 200 KB of code in step (one instruction stream through the 32 KB instruction cache instead of eight)
    if (__syncthreads_and(state == AST_IDLE)) break;

    // ---------------- one loss + gradient evaluation per thread (CTA-convergent) ----------------------------
    double* xt = vecp(2 * (cur ^ 1));
    double* gt = xt + (int64_t)n * T;
    for (int j = 0; j < n; ++j) gt[j * T] = 0.0;
    const double ft = adj1_nl_sync(&kt, xt, T, A.V + t * 32, A.cost_kind, gt, T);
    if (state == AST_IDLE) continue;
    ++evals;
    ++pevals;

    // ---------------- bookkeeping ---------------------------------------------------------------------------
    const double* x = vecp(2 * cur);
    const double* g = x + (int64_t)n * T;
    const bool first = state == AST_INIT;
    bool done = false;
    int reason = 0;
    if (first || ft <= f + kArmijoFd * gde) {
      // ---- accept ----
      // t1 <- projected gradient at the accepted point (also the start of the two-loop recursion); history pair
      double gmax = 0.0, gg = 0.0, sy = 0.0, yy = 0.0;
      // active set of this iteration (variables on a bound with the gradient pushing outward): the two-loop recursion runs
      // in the free subspace, see adj_lbfgs_reg_kernel
      unsigned act[(NQ + 31) / 32];
#pragma unroll
      for (int w = 0; w < (NQ + 31) / 32; ++w) act[w] = 0;
      auto is_act = [&](int j) -> bool { return bounded && ((act[j >> 5] >> (j & 31)) & 1u); };
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const double xtj = xt[j * T], gtj = gt[j * T];
        double gp = gtj;
        if (bounded && ((xtj <= A.lower[j] && gtj > 0.0) || (xtj >= A.upper[j] && gtj < 0.0))) {
          gp = 0.0;
          act[j >> 5] |= 1u << (j & 31);
        }
        t1[j] = gp;
        gmax = fmax(gmax, fabs(gp));
        gg = fma(gp, gp, gg);
        if (!first) {
          const double sv = xtj - x[j * T], yv = gtj - g[j * T];
          dv[j] = sv;  // (the old direction is dead once the step is accepted)
          t2[j] = yv;
          sy = fma(sv, yv, sy);
          yy = fma(yv, yv, yy);
        }
      }
      if (!first) {
        double* sn = hs(hpos);
        double* yn = hy(hpos);
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
          sn[j * T] = dv[j];
          yn[j * T] = t2[j];
        }
        if (sy > 1e-14 * yy && yy > 0.0) {  // cautious update
          rho[hpos] = 1.0 / sy;
          gamma = sy / yy;
          hpos = (hpos + 1 == m) ? 0 : hpos + 1;
          hcount = min(hcount + 1, m);
        } else if (hcount == m) {
          hcount = m - 1;
        }
        ++iter;
      } else {
        f_chk = ft;
      }
      cur ^= 1;  // the trial point becomes the current point
      f = ft;
      bool stalled = false;
      if (!first && (iter & 31) == 0) {  // progress checkpoint (same rule as the sequential form)
        slow = f > 0.97 * f_chk;
        f_chk = f;
        stalled = slow && f > A.f_far && pevals - pe_chk > 32 * kStallEvals;  // crawling: see kStallEvals
        pe_chk = pevals;
      }
      if (f < A.f_stop) reason = 1;
      else if (gmax < A.gtol) reason = 2;
      else if (gmax < A.gtol_far && (f > A.f_far || slow)) reason = 3;
      else if (iter >= A.max_iter) reason = 4;
      else if (!(f == f)) reason = 5;
      else if (A.early_exit && (iter & 3) == 0 && *((volatile int32_t*)(A.solved + t)) != 0) reason = 6;
      else if (stalled) reason = 9;
      done = reason != 0;
      if (!done) {
        // two-loop recursion, q = dv <- t1.  First loop, newest to oldest: a_i = rho_i s_i.q ; q -= a_i y_i -- the axpy of
        // pair i is fused with the dot product of pair i+1.
        for (int j = 0; j < n; ++j) dv[j] = t1[j];
        double a_prev = 0.0;
        const double* y_prev = hy(0);
        for (int hh = 0; hh <= hcount; ++hh) {
          int slot = hpos - 1 - hh;
          if (slot < 0) slot += m;
          const bool last = hh == hcount;  // last pass: only the pending axpy
          const double* sk = hs(last ? 0 : slot);
          if (hh + 1 < hcount) {  // next pass's vectors on their way from DRAM / far L2 while this one computes
            int nx = slot - 1;
            if (nx < 0) nx += m;
            const double* ps = hs(nx);
            const double* py = hy(slot);
            for (int j = 0; j < n; ++j) {
              prefetch_l2(ps + j * T);
              prefetch_l2(py + j * T);
            }
          }
          double acc = 0.0;
#pragma unroll 4
          for (int j = 0; j < n; ++j) {
            double qj = hh == 0 ? dv[j] : fma(-a_prev, y_prev[j * T], dv[j]);  // (no load in the first pass: the
                                                                               // history may be uninitialised)
            if (is_act(j)) qj = 0.0;
            dv[j] = qj;
            acc = fma(sk[j * T], qj, acc);
          }
          if (!last) {
            a_prev = acc * rho[slot];
            alp[slot] = a_prev;
            y_prev = hy(slot);
          }
        }
        // second loop, oldest to newest: b_i = rho_i y_i.r ; r += (alp_i - b_i) s_i, with the initial scaling r = gamma q
        // folded into the first pass
        double c_prev = 0.0;
        const double* s_prev = hs(0);
        const double scale0 = hcount > 0 ? gamma : 1.0;
        for (int hh = hcount - 1; hh >= -1; --hh) {
          int slot = hpos - 1 - hh;
          if (slot < 0) slot += m;
          if (slot >= m) slot -= m;
          const bool last = hh < 0;
          const bool firstpass = hh == hcount - 1;
          const double* yk = hy(last ? 0 : slot);
          double acc = 0.0;
#pragma unroll 4
          for (int j = 0; j < n; ++j) {
            double rj = firstpass ? scale0 * dv[j] : fma(c_prev, s_prev[j * T], dv[j]);
            if (is_act(j)) rj = 0.0;
            dv[j] = rj;
            acc = fma(yk[j * T], rj, acc);
          }
          if (!last) {
            c_prev = alp[slot] - acc * rho[slot];
            s_prev = hs(slot);
          }
        }
        // d = -r ; g.d
        double gd = 0.0;
        for (int j = 0; j < n; ++j) {
          const double d = -dv[j];
          dv[j] = d;
          gd = fma(t1[j], d, gd);  // (projected gradient; equals g.d on the free variables)
        }
        alpha = 1.0;
        dir_sd = false;
        if (hcount == 0 || !(gd < 0.0)) {  // first step or not a descent direction: steepest descent, unit length
          hcount = 0;
          dir_sd = true;
          for (int j = 0; j < n; ++j) dv[j] = -t1[j];
          alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
        }
        ls = 0;
        state = AST_LS;
      }
    } else {
      // ---- reject: backtrack along dv; cubic through (0, f, gde/alpha) and (alpha, ft, g_t.d) ----
      double gdt = 0.0;
#pragma unroll 4
      for (int j = 0; j < n; ++j) gdt = fma(gt[j * T], dv[j], gdt);
      const double gd0 = gde / alpha;
      double an = 0.5 * alpha;
      if (ft == ft && gdt == gdt) {
        const double d1 = gd0 + gdt - 3.0 * (ft - f) / alpha;
        const double disc = d1 * d1 - gd0 * gdt;
        if (disc >= 0.0) {
          const double d2 = sqrt(disc);
          const double den = gdt - gd0 + 2.0 * d2;
          if (den != 0.0) {
            const double cand = alpha - alpha * (gdt + d2 - d1) / den;
            if (cand == cand) an = cand;
          }
        }
      }
      alpha = fmin(fmax(an, 0.1 * alpha), 0.5 * alpha);
      if (++ls > 30) {
        if (hcount > 0 && nfail < kMaxLsFail) {  // curvature model is bad: restart from steepest descent
          ++nfail;
          hcount = 0;
          dir_sd = true;
          double gg = 0.0;
#pragma unroll 4
          for (int j = 0; j < n; ++j) {
            const double xj = x[j * T], gj = g[j * T];
            double gp = gj;
            if (bounded && ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0))) gp = 0.0;
            dv[j] = -gp;
            gg = fma(gp, gp, gg);
          }
          alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
          ls = 0;
        } else {
          done = true;  // no progress possible at working precision
          reason = 8;
        }
      }
    }
    if (!done) {
      // next trial point x + alpha d (clamped to the box) into the non-current buffer, and the directional derivative
      // along the (projected) segment for the Armijo test.  A quasi-Newton direction whose clamped segment is not a descent
      // segment falls back to projected steepest descent (as K5 does); only when that has no descent either is the point a
      // KKT point of the box problem.
      const double* xc = vecp(2 * cur);
      const double* gc = xc + (int64_t)n * T;
      double* xn = vecp(2 * (cur ^ 1));
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        double g_step = 0.0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
          const double xj = xc[j * T];
          double v = fma(alpha, dv[j], xj);
          if (bounded) v = fmin(fmax(v, A.lower[j]), A.upper[j]);
          t2[j] = v;
          g_step = fma(gc[j * T], v - xj, g_step);
        }
        gde = g_step;
        if (gde < 0.0) break;
        if (dir_sd) {  // zero (projected) gradient along the steepest-descent step
          done = true;
          reason = 7;
          break;
        }
        dir_sd = true;
        hcount = 0;
        ls = 0;
        double gg = 0.0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
          const double xj = xc[j * T], gj = gc[j * T];
          double gp = gj;
          if (bounded && ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0))) gp = 0.0;
          dv[j] = -gp;
          gg = fma(gp, gp, gg);
        }
        alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
      }
      if (!done) {
#pragma unroll 4
        for (int j = 0; j < n; ++j) xn[j * T] = t2[j];
      }
    }
    if (done) {
      const double* xf = vecp(2 * cur);
      A.out_loss[pid] = f;
      A.out_iters[pid] = A.debug == 2 ? (min(pevals, 0xFFFFFF) | (reason << 24)) : A.debug ? (iter | (reason << 24)) : iter;
      for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = xf[j * T];
      if (A.early_exit && f < A.success_threshold) atomicExch(A.solved + t, 1);
      state = AST_IDLE;
    }
  }
  if (A.out_evals && evals) atomicAdd(A.out_evals, evals);
}


// ------------------------------------------------------------------------------------------------------------------
// K5c, adjoint mode, REGISTER form (P <= 32): the same tick-structured optimiser with the bookkeeping written like K5's.
//   * the vector length bound NQ is a compile-time constant and every vector loop is fully unrolled over NQ padded entries, so
//     the working vector of the two-loop recursion lives in REGISTERS and every pass issues all of its loads before the
//     first use.  The local-array form above walks its vectors with a runtime bound: 4 loads in flight per thread, each an
//     L2 / DRAM round trip of the interleaved workspace -- ncu: long_scoreboard 3.25 per issue, FP64 pipe 26 %
//     (profiles/r01_k5c_adjoint_kernel_ncu_full.txt) -- and, the CTA being phase-locked, that latency is not hidden by the
//     other warps' evaluations;
//   * nothing vector-sized is carried across the evaluation: the search direction is recovered as (xt - x) / alpha while a
//     line search is in progress (as K5 does), so no thread-local arrays exist at all.
// (x, g) double-buffered and the (s, y) history stay in global memory, in a CTA-local interleaved chunk (entry j of thread t at
// [j 256 + t]: every access of a warp is one coalesced line and every offset of an unrolled pass an immediate).
// Measured and rejected on this kernel (scripts/k5c_bench.py, 131072 x 8 restarts, sqiSwap k = 2 / 3 smush templates):
//   * 2 or 4 CTAs per SM instead of one phase-locked 256-thread CTA (so that one CTA's bookkeeping overlaps another's
//     evaluation): 98 -> 81 -> 48 M evaluations/s -- the CTA-wide instruction stream through the ~300 KB of code matters more;
//   * prefetching every row of the two-loop recursion into L2 at the start of the bookkeeping: no change, and the burst of
//     prefetches stalls on the load/store queue (13 % of the stall samples); the s and y row of a pair are issued together;
//   * a persisting L2 access-policy window over the history chunk (79 MB set-aside, the rest streaming): 1.82 -> 2.05 s (P = 18)
//     and 1.37 -> 1.66 s (P = 30) -- the thread-local traffic of the evaluation needs the L2 as much as the history does.
// With 8 warps per SM at 255 registers the kernel is bound by latency per warp, not by throughput: a tick costs the same
// ~0.25 ms whether one or all 256 threads are live.  The lever would be occupancy, which the 4x4 complex R, W, Q of the
// adjoint pass (192 registers) do not leave room for.
// ------------------------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(kAdjCta, 1) adj_lbfgs_reg_kernel(const __grid_constant__ FdArgs A,
                                                                   const __grid_constant__ KTemplate kt) {
  // The two (x, g) buffers live in SHARED memory when they fit (4 NQ doubles per thread: NQ <= 28 at 256 threads), only the
  // (s, y) history in the global chunk: at bench size the workspace of all threads was 145 MB against a 126 MB L2, and ncu
  // showed the kernel moving 4.7 TB through DRAM in 2.15 s (27 % of peak) with long_scoreboard at 4.4 per issue.
  constexpr int kSmemVecs = adj_smem_vecs(NQ);
  extern __shared__ __align__(16) double xg_smem[];
  const int n = kt.P;
  const int m = A.m;
  // CTA-local workspace chunk: vector v of thread t has its entry j at chunk[(v NQ + j) T + t], T = the CTA size -- a
  // compile-time constant, so every access of an unrolled pass is ONE load / store with an immediate offset (j T 8 bytes)
  // from the row pointer, still one coalesced line per warp.  Vectors are padded to NQ entries that stay zero, so the
  // passes need no `j < n` guards either.  (With the grid-wide interleave [j T_grid + t] of the local-array form every
  // element cost a 64-bit multiply-add, a compare and a select: ncu had this one line at 19 % of ALL executed instructions.)
  constexpr int64_t T = kAdjCta;
  constexpr int kXg = 4 - kSmemVecs;  // (x, g) vectors kept in the global chunk: the last kXg of the four
  constexpr int kVecs = kXg + 2 * kAdjHist;
  double* ws = A.ws + (int64_t)blockIdx.x * (kVecs * NQ * T) + threadIdx.x;
  double* xg = xg_smem + threadIdx.x;
  auto vecp = [&](int v) -> double* {  // vector v of (x0, g0, x1, g1); entry j at [j * T]
    return v < kSmemVecs ? xg + (int64_t)v * (NQ * T) : ws + (int64_t)(v - kSmemVecs) * (NQ * T);
  };
  auto hs = [&](int slot) -> double* { return ws + (int64_t)(kXg + slot) * (NQ * T); };
  auto hy = [&](int slot) -> double* { return ws + (int64_t)(kXg + m + slot) * (NQ * T); };
  const int64_t total = A.Nt * (int64_t)A.restarts;
  const bool bounded = A.lower != nullptr;
  // box bounds padded to NQ entries (the padding is never active)
  __shared__ double s_lo[NQ], s_hi[NQ];
  if (threadIdx.x < NQ) {
    const int j = threadIdx.x;
    s_lo[j] = (bounded && j < n) ? A.lower[j] : -DBL_MAX;
    s_hi[j] = (bounded && j < n) ? A.upper[j] : DBL_MAX;
  }
  __syncthreads();

  int state = AST_IDLE, cur = 0, iter = 0, hcount = 0, hpos = 0, ls = 0;
  bool exhausted = false, slow = false;
  int pevals = 0;  // evaluations of the current problem (diagnostics)
  int nfail = 0;   // failed line searches of the current problem
  int pe_chk = 0;  // pevals at the last progress checkpoint
  int64_t pid = 0, t = 0;
  double f = 0.0, alpha = 1.0, gd = 0.0, gamma = 1.0, f_chk = 0.0;
  double rho[kAdjHist];
  unsigned long long evals = 0;
  for (int e = 0; e < kVecs * NQ; ++e) ws[e * T] = 0.0;  // padding stays zero; idle lanes evaluate their (finite) trial buffer
  for (int e = 0; e < kSmemVecs * NQ; ++e) xg[e * T] = 0.0;

  while (true) {
    // ---------------- fetch -------------------------------------------------------------------------------
    while (state == AST_IDLE && !exhausted) {
      const unsigned long long w = atomicAdd(A.next, 1ULL);
      if ((int64_t)w >= total) {
        exhausted = true;
        break;
      }
      const int64_t r_idx = (int64_t)w / A.Nt;
      t = (int64_t)w - r_idx * A.Nt;
      pid = t * A.restarts + r_idx;
      bool skip = A.active && A.active[t] == 0;
      if (!skip && A.early_exit) skip = *((volatile int32_t*)(A.solved + t)) != 0;
      if (skip) {
        A.out_loss[pid] = DBL_MAX;
        A.out_iters[pid] = 0;
        for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = 0.0;
        continue;
      }
      cur = 0;  // trial buffer = vectors (2, 3)
      double* x1 = vecp(2);
      for (int j = 0; j < n; ++j) {
        double x = A.x0 ? A.x0[pid * A.ldx0 + j] : philox_param(A.seed, (uint64_t)pid, j, A.x0_lo, A.x0_span);
        if (bounded) x = fmin(fmax(x, A.lower[j]), A.upper[j]);
        x1[j * T] = x;
      }
      state = AST_INIT;
      iter = 0; ls = 0; hcount = 0; hpos = 0; gamma = 1.0; slow = false; pevals = 0; nfail = 0; pe_chk = 0;
    }
    if (__syncthreads_and(state == AST_IDLE)) break;  // CTA-wide vote = the tick barrier (phase lock, see above)

    // ---------------- one loss + gradient evaluation per thread (CTA-convergent) ----------------------------
    double* xt = vecp(2 * (cur ^ 1));
    double* gt = vecp(2 * (cur ^ 1) + 1);
    for (int j = 0; j < n; ++j) gt[j * T] = 0.0;
    const double ft = adj1_nl_sync(&kt, xt, T, A.V + t * 32, A.cost_kind, gt, T);
    if (state == AST_IDLE) continue;
    ++evals;
    ++pevals;

    // ---------------- bookkeeping (registers; every pass loads first, then computes) -------------------------
    const double* x = vecp(2 * cur);
    const double* g = vecp(2 * cur + 1);
    const bool first = state == AST_INIT;
    bool done = false;
    int reason = 0;
    // (register budget: at most three NQ-vectors are live at any point -- the evaluation behind the call keeps 255)
    auto load_row = [&](const double* p, double* v) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) v[j] = p[j * T];
    };
    auto project = [&](double* gp, const double* xv) {  // drop gradient components pushing against an active bound
      if (bounded) {
#pragma unroll
        for (int j = 0; j < NQ; ++j)
          if ((xv[j] <= s_lo[j] && gp[j] > 0.0) || (xv[j] >= s_hi[j] && gp[j] < 0.0)) gp[j] = 0.0;
      }
    };
    if (first || ft <= f + kArmijoFd * alpha * gd) {
      // ---- accept ----
      double q[NQ];  // projected gradient at the accepted point, then the two-loop working vector
      double gmax = 0.0, gg = 0.0, sy = 0.0, yy = 0.0;
      // Active set of this iteration: variables on a bound whose gradient pushes outward.  The two-loop recursion runs in
      // the FREE subspace -- the working vector is masked after every update -- so the quasi-Newton direction never moves an
      // active variable and is a model step of the free variables only.  Projecting the full-space direction instead (the
      // clamp then discards its active components) leaves a step the free variables were never meant to take alone: measured
      // on the configs[3] grid, most restarts that reached a face of the +-2 pi amplitude box got non-descent segments and
      // degenerated to steepest descent (6 % ran into max_iter).
      unsigned act = 0;
      auto mask_q = [&]() {
        if (act) {
#pragma unroll
          for (int j = 0; j < NQ; ++j)
            if (act & (1u << j)) q[j] = 0.0;
        }
      };
      {
        double a[NQ], b[NQ];
        if (!first) {
          load_row(xt, a);
          load_row(x, b);
          double* sn = hs(hpos);
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            a[j] -= b[j];  // s
            sn[j * T] = a[j];
          }
        }
        load_row(gt, q);
        if (!first) {
          load_row(g, b);
          double* yn = hy(hpos);
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            const double yv = q[j] - b[j];
            yn[j * T] = yv;
            sy = fma(a[j], yv, sy);
            yy = fma(yv, yv, yy);
          }
        }
        if (bounded) {
          load_row(xt, b);
#pragma unroll
          for (int j = 0; j < NQ; ++j)
            if ((b[j] <= s_lo[j] && q[j] > 0.0) || (b[j] >= s_hi[j] && q[j] < 0.0)) {
              q[j] = 0.0;
              act |= 1u << j;
            }
        }
      }
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        gmax = fmax(gmax, fabs(q[j]));
        gg = fma(q[j], q[j], gg);
      }
      if (!first) {
        if (sy > 1e-14 * yy && yy > 0.0) {  // cautious update
          rho[hpos] = 1.0 / sy;
          gamma = sy / yy;
          hpos = (hpos + 1 == m) ? 0 : hpos + 1;
          hcount = min(hcount + 1, m);
        } else if (hcount == m) {
          hcount = m - 1;
        }
        ++iter;
      } else {
        f_chk = ft;
      }
      cur ^= 1;  // the trial point becomes the current point
      f = ft;
      bool stalled = false;
      if (!first && (iter & 31) == 0) {  // progress checkpoint (same rule as the local-array form)
        slow = f > 0.97 * f_chk;
        f_chk = f;
        stalled = slow && f > A.f_far && pevals - pe_chk > 32 * kStallEvals;  // crawling: see kStallEvals
        pe_chk = pevals;
      }
      if (f < A.f_stop) reason = 1;
      else if (gmax < A.gtol) reason = 2;
      else if (gmax < A.gtol_far && (f > A.f_far || slow)) reason = 3;
      else if (iter >= A.max_iter) reason = 4;
      else if (!(f == f)) reason = 5;
      else if (A.early_exit && (iter & 3) == 0 && *((volatile int32_t*)(A.solved + t)) != 0) reason = 6;
      else if (stalled) reason = 9;
      done = reason != 0;
      if (!done) {
        // two-loop recursion on the register vector: per pair one fully unrolled load pass for s and one for y
        double alp[kAdjHist];
        for (int hh = 0; hh < hcount; ++hh) {
          int slot = hpos - 1 - hh;
          if (slot < 0) slot += m;
          double v[NQ], v2[NQ];
          load_row(hs(slot), v);
          load_row(hy(slot), v2);  // (issued with the s row: one exposed memory latency per pair instead of two)
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int j = 0; j < NQ; j += 2) {
            a0 = fma(v[j], q[j], a0);
            if (j + 1 < NQ) a1 = fma(v[j + 1], q[j + 1], a1);
          }
          const double a = (a0 + a1) * rho[slot];
          alp[slot] = a;
#pragma unroll
          for (int j = 0; j < NQ; ++j) q[j] = fma(-a, v2[j], q[j]);
          mask_q();
        }
        if (hcount > 0) {
#pragma unroll
          for (int j = 0; j < NQ; ++j) q[j] *= gamma;
        }
        for (int hh = hcount - 1; hh >= 0; --hh) {
          int slot = hpos - 1 - hh;
          if (slot < 0) slot += m;
          double v[NQ], v2[NQ];
          load_row(hy(slot), v);
          load_row(hs(slot), v2);
          double b0 = 0.0, b1 = 0.0;
#pragma unroll
          for (int j = 0; j < NQ; j += 2) {
            b0 = fma(v[j], q[j], b0);
            if (j + 1 < NQ) b1 = fma(v[j + 1], q[j + 1], b1);
          }
          const double c = alp[slot] - (b0 + b1) * rho[slot];
#pragma unroll
          for (int j = 0; j < NQ; ++j) q[j] = fma(c, v2[j], q[j]);
          mask_q();
        }
        // d = -q ; g.d with the projected gradient (equals g.d on the free variables)
        double xa[NQ], ga[NQ];
        load_row(xt, xa);
        load_row(gt, ga);  // (un-projected: the directional derivative along the clamped segment uses it)
        auto gproj = [&](int j) -> double {  // projected gradient component j at the accepted point
          if (bounded && ((xa[j] <= s_lo[j] && ga[j] > 0.0) || (xa[j] >= s_hi[j] && ga[j] < 0.0))) return 0.0;
          return ga[j];
        };
        double gdn = 0.0;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          q[j] = -q[j];
          gdn = fma(gproj(j), q[j], gdn);
        }
        bool use_sd = hcount == 0 || !(gdn < 0.0);  // first step or not a descent direction: steepest descent, unit length
        ls = 0;
        state = AST_LS;
        double* xn = vecp(2 * (cur ^ 1));
        // next trial point xt + alpha d (clamped to the box) into the other buffer; directional derivative along the
        // (projected) segment per unit alpha, as the Armijo test and the cubic use it.  A quasi-Newton direction whose
        // clamped segment is not a descent segment falls back to projected steepest descent (as K5 does); only when that
        // has no descent either is the point a KKT point of the box problem.
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
          alpha = 1.0;
          if (use_sd) {
            hcount = 0;
#pragma unroll
            for (int j = 0; j < NQ; ++j) q[j] = -gproj(j);
            alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
          }
          double g_step = 0.0;
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            double v = fma(alpha, q[j], xa[j]);
            if (bounded) v = fmin(fmax(v, s_lo[j]), s_hi[j]);
            xn[j * T] = v;
            g_step = fma(ga[j], v - xa[j], g_step);
          }
          gd = g_step / alpha;
          if (gd < 0.0) break;
          if (use_sd) {  // zero (projected) gradient along the steepest-descent step
            done = true;
            reason = 7;
            break;
          }
          use_sd = true;
        }
      }
    } else {
      // ---- reject: backtrack along d = (xt - x) / alpha; cubic through (0, f, gd) and (alpha, ft, g_t.d) ----
      double xv[NQ], dx[NQ];
      load_row(x, xv);
      load_row(xt, dx);
      {
        double gv[NQ];
        load_row(gt, gv);
        double gdt = 0.0;
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          dx[j] -= xv[j];
          gdt = fma(gv[j], dx[j], gdt);
        }
        gdt /= alpha;
        double an = 0.5 * alpha;
        if (ft == ft && gdt == gdt) {
          const double d1 = gd + gdt - 3.0 * (ft - f) / alpha;
          const double disc = d1 * d1 - gd * gdt;
          if (disc >= 0.0) {
            const double d2 = sqrt(disc);
            const double den = gdt - gd + 2.0 * d2;
            if (den != 0.0) {
              const double cand = alpha - alpha * (gdt + d2 - d1) / den;
              if (cand == cand) an = cand;
            }
          }
        }
        an = fmin(fmax(an, 0.1 * alpha), 0.5 * alpha);
        const double ratio = an / alpha;
        bool restart_sd = false;
        if (++ls > 30) {
          if (hcount > 0 && nfail < kMaxLsFail) {  // curvature model is bad: restart from steepest descent at the current point
            ++nfail;
            hcount = 0;
            restart_sd = true;
            ls = 0;
          } else {
            done = true;  // no progress possible at working precision
            reason = 8;
          }
        }
        if (!done && restart_sd) {
          load_row(g, gv);
          project(gv, xv);
          double gg = 0.0, g_step = 0.0;
#pragma unroll
          for (int j = 0; j < NQ; ++j) gg = fma(gv[j], gv[j], gg);
          an = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
          load_row(g, dx);  // (un-projected, for the directional derivative)
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            double v = fma(-an, gv[j], xv[j]);
            if (bounded) v = fmin(fmax(v, s_lo[j]), s_hi[j]);
            xt[j * T] = v;
            g_step = fma(dx[j], v - xv[j], g_step);
          }
          alpha = an;
          gd = g_step / alpha;
          if (!(gd < 0.0)) {
            done = true;
            reason = 7;
          }
        } else if (!done) {
          // the segment x -> xt was already inside the box (both ends are), so shrinking it needs no projection
#pragma unroll
          for (int j = 0; j < NQ; ++j) xt[j * T] = fma(ratio, dx[j], xv[j]);
          alpha = an;
        }
      }
    }
    if (done) {
      const double* xf = vecp(2 * cur);
      A.out_loss[pid] = f;
      A.out_iters[pid] = A.debug == 2 ? (min(pevals, 0xFFFFFF) | (reason << 24)) : A.debug ? (iter | (reason << 24)) : iter;
      for (int j = 0; j < n; ++j) A.out_x[pid * n + j] = xf[j * T];
      if (A.early_exit && f < A.success_threshold) atomicExch(A.solved + t, 1);
      state = AST_IDLE;
    }
  }
  if (A.out_evals && evals) atomicAdd(A.out_evals, evals);
}

template <int NQ>
static int launch_adj_reg(const FdArgs& A, const KTemplate& kt, unsigned blocks, cudaStream_t st) {
  const size_t smem = (size_t)adj_smem_vecs(NQ) * NQ * kAdjCta * 8;
  auto kern = adj_lbfgs_reg_kernel<NQ>;
  if (smem) SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<blocks, kAdjCta, smem, st>>>(A, kt);
  return SLAM_OK;
}

}  // namespace slam

using namespace slam;

extern "C" int slam_fd_lbfgs_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts,
                                   const double* x0, int64_t ldx0, uint64_t seed, const int32_t* active,
                                   const SlamOptOpts* opts, int32_t central, double* out_loss, double* out_x,
                                   int32_t* out_iters, unsigned long long* out_evals, void* stream) {
  if (!desc || !V || !opts || !out_loss || !out_x || !out_iters || Nt < 0 || restarts < 1) return SLAM_ERR_INVALID;
  if (x0 && ldx0 < desc->n_params) return SLAM_ERR_INVALID;
  if (opts->cost_kind != SLAM_COST_BASIC && opts->cost_kind != SLAM_COST_SQUARE && opts->cost_kind != SLAM_COST_BASIC_INVERSE)
    return SLAM_ERR_UNSUPPORTED;  // the coordinate-based functionals are piecewise constant (8-dp rounding): no gradient
  if (opts->max_iter < 1 || desc->n_params < 1 || central < 0 || central > 2) return SLAM_ERR_INVALID;
  if (opts->history < 0 || opts->history > kFdHist) return SLAM_ERR_INVALID;
  if ((opts->lower == nullptr) != (opts->upper == nullptr)) return SLAM_ERR_INVALID;  // box = both arrays (+-inf allowed)
  if (Nt == 0) return SLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KTemplate kt;
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/true, /*allow_ties=*/central != 2);
  if (rc != SLAM_OK) return rc;
  if (central == 2 && kt.gmode != GM_SMUSH) return SLAM_ERR_UNSUPPORTED;  // closed-form gates: slam_lbfgs_solve
  if (central == 2 && opts->con_mu != 0.0) return SLAM_ERR_UNSUPPORTED;   // the constraint term is differenced, not adjoint
  if (opts->con_mu < 0.0) return SLAM_ERR_INVALID;
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  int dev = 0, sms = 0;
  SLAM_CUDA_CHECK(cudaGetDevice(&dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n = kt.P;
  const int64_t total = Nt * (int64_t)restarts;
  const int threads = central == 2 ? kAdjCta : 128;
  int64_t blocks = std::min<int64_t>((int64_t)sms * (256 / threads), (total + threads - 1) / threads);
  // (5 + 2 m) double vectors per thread (the adjoint mode keeps its direction in thread-local memory and uses 4 + 2 m)
  // the register form of the adjoint kernel (P <= 32) pads its vectors to the instantiated bound NQ
  const int nq = (central == 2 && n <= 32) ? std::max(16, (n + 3) / 4 * 4) : n;  // instantiated: 16, 20, 24, 28, 32
  const size_t per_thread = (size_t)(5 + 2 * kFdHist) * nq * sizeof(double);
  while (blocks > 1 && per_thread * threads * (size_t)blocks > ((size_t)4 << 30)) blocks /= 2;  // workspace <= 4 GiB
  const int64_t T = blocks * threads;

  Scratch scratch(st);
  unsigned long long* next = nullptr;
  int32_t* solved = nullptr;
  double* ws = nullptr;
  if ((rc = scratch.alloc(&next, sizeof(unsigned long long), true)) != SLAM_OK) return rc;
  if ((rc = scratch.alloc(&solved, sizeof(int32_t) * (size_t)Nt, true)) != SLAM_OK) return rc;
  if ((rc = scratch.alloc(&ws, per_thread * (size_t)T)) != SLAM_OK) return rc;

  FdArgs A;
  A.V = V; A.x0 = x0; A.ldx0 = ldx0; A.seed = seed; A.active = active; A.Nt = Nt; A.restarts = restarts;
  A.max_iter = opts->max_iter; A.cost_kind = opts->cost_kind; A.early_exit = opts->early_exit; A.central = central;
  // smush gates carry no circuit_fidelity factor, so 1 - BasicCostInverse x 1 is BasicCost (optimizer.py:200-201)
  if (central == 2 && A.cost_kind == SLAM_COST_BASIC_INVERSE) A.cost_kind = SLAM_COST_BASIC;
  // history of the adjoint kernels, measured (scripts/k5c_bench.py, 131072 x 8 restarts; scripts/k5c_basin_check.py;
  // scripts/smush_training_grid.py): at P = 18 eight pairs are needed -- with six, 2 of the 12 targets of the
  // singular-Hessian in-basin problem of the tests no longer reach 1e-10 and the run is no faster; at P = 22 .. 30 six pairs
  // are 10 - 14 % faster (7 - 9 % more evaluations, each cheaper; the history is what misses L2) at unchanged solved
  // fractions (0.994 on sqrt(iSWAP) k = 3, 0.98 / 0.99 on the configs[3] grid); the local-array form (P = 42) is 16 % faster
  // with five than with eight
  A.m = opts->history > 0 ? std::min<int>(opts->history, kAdjHist) : (n <= 20 ? kAdjHist : n <= 32 ? 6 : 5);
  A.debug = opts->diag;  // stop-reason / evaluation-count packing in out_iters (see SlamOptOpts.diag)
  A.success_threshold = opts->success_threshold; A.f_stop = opts->f_stop; A.gtol = opts->gtol;
  A.gtol_far = opts->gtol_far; A.f_far = opts->f_far; A.x0_lo = opts->x0_lo; A.x0_span = opts->x0_hi - opts->x0_lo;
  A.lower = opts->lower;
  A.upper = opts->upper;
  A.con_max = opts->con_max; A.con_mu = opts->con_mu; A.con_lambda = opts->con_lambda;
  A.out_loss = out_loss; A.out_x = out_x; A.out_iters = out_iters; A.out_evals = out_evals;
  A.next = next; A.solved = solved; A.ws = ws; A.T = T;
  if (central == 2) {
    if (n <= 16) rc = launch_adj_reg<16>(A, kt, (unsigned)blocks, st);
    else if (n <= 20) rc = launch_adj_reg<20>(A, kt, (unsigned)blocks, st);
    else if (n <= 24) rc = launch_adj_reg<24>(A, kt, (unsigned)blocks, st);
    else if (n <= 28) rc = launch_adj_reg<28>(A, kt, (unsigned)blocks, st);
    else if (n <= 32) rc = launch_adj_reg<32>(A, kt, (unsigned)blocks, st);
    if (rc != SLAM_OK) return rc;
    // (a 48-entry register form was measured: 3.8 KB of spills around the 255-register evaluation, 6.3 vs 7.6 M evaluations/s
    //  on the sqCNOT k = 4 template, P = 42 -- the local-array form keeps the larger templates)
    else if (n <= 96) adj_lbfgs_kernel<96><<<(unsigned)blocks, threads, 0, st>>>(A, kt);
    else adj_lbfgs_kernel<SLAM_MAX_PARAMS><<<(unsigned)blocks, threads, 0, st>>>(A, kt);
  }
  else if (central == 1) fd_lbfgs_kernel<1><<<(unsigned)blocks, threads, 0, st>>>(A, kt);
  else fd_lbfgs_kernel<0><<<(unsigned)blocks, threads, 0, st>>>(A, kt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e, "fd_lbfgs_kernel launch");
    return SLAM_ERR_CUDA;
  }
  return SLAM_OK;
}
