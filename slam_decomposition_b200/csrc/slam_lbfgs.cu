// slam_lbfgs.cu -- K5: device-resident batched L-BFGS over (target, restart) problems.
//
// Replaces the restart loop around scipy.optimize.minimize(method="BFGS") (src/slam/optimizer.py:253-295):
// instead of one Python-driven BFGS with (P+1)-evaluation finite-difference gradients per restart, every
// (target, restart) pair is an independent problem solved by a TEAM of 4 lanes with analytic adjoint
// gradients (slam_core.cuh).  All optimiser state of a problem -- x, g, trial point, direction, the L-BFGS
// (s, y) history and the (cos, sin) cache -- lives in that team's slice of shared memory; nothing but the
// final result touches HBM.  A persistent grid pulls problems from a global counter, so early exits and
// uneven iteration counts do not leave SMs idle.
//
// Warp-level structure: each "tick" every team of the warp performs exactly one loss+gradient evaluation
// (the expensive, fully convergent part), then runs its own cheap, possibly divergent, line-search /
// history bookkeeping.  In the bookkeeping each lane owns the vector entries j = sub + 4 i (i < NPL) and keeps
// its slice of the two-loop working vector in registers, so the recursion is a stream of independent
// shared-memory loads and FMAs with one 2-stage shuffle reduction per history pair.
#include <cfloat>
#include <cstdlib>

#include "slam_host.h"
#include "slam_philox.cuh"

namespace slam {

constexpr int LPP = 4;
constexpr int kMaxHist = 8;
constexpr double kArmijo = 1e-4;
#ifndef SLAM_LBFGS_LOCKSTEP
#define SLAM_LBFGS_LOCKSTEP 0
#endif
#ifndef SLAM_LBFGS_MAX_THREADS
#define SLAM_LBFGS_MAX_THREADS 384  // register cap 168/thread -> up to 12 warps per SM
#endif
constexpr int kLbfgsMaxThreads = SLAM_LBFGS_MAX_THREADS;

struct LbfgsArgs {
  const double* V;
  const double* x0;
  int64_t ldx0;
  uint64_t seed;
  const int32_t* active;
  int64_t Nt;
  int restarts;
  int m;        // history length
  int RS;       // doubles of shared memory per team (vectors, rho, alp, trig cache)
  int HS;       // history elements (of HT) per team; the history slices follow the RS slices of all teams
  int max_iter;
  int cost_kind;
  int early_exit;
  double success_threshold, f_stop, gtol, gtol_far, f_far, x0_lo, x0_span;
  int trace_cap;
  double* trace_loss;
  double* trace_x;
  const double* lower;  // box bounds [P] or null (projected L-BFGS; the reference switches to L-BFGS-B, optimizer.py:257-258)
  const double* upper;
  double* out_loss;
  double* out_x;
  int32_t* out_iters;
  unsigned long long* out_evals;
  unsigned long long* next;  // work counter
  int32_t* solved;           // per-target flag (early exit)
};

__device__ __forceinline__ double tsum(double v, unsigned mask) {
  v += __shfl_xor_sync(mask, v, 1);
  v += __shfl_xor_sync(mask, v, 2);
  return v;
}
__device__ __forceinline__ double tmax(double v, unsigned mask) {
  v = fmax(v, __shfl_xor_sync(mask, v, 1));
  v = fmax(v, __shfl_xor_sync(mask, v, 2));
  return v;
}

enum { ST_IDLE = 0, ST_INIT = 1, ST_LS = 2 };

// NPL = max vector entries per lane held in registers (Pp <= 4 * NPL); HT = storage type of the (s, y) history.
//
// Shared-memory slice of a team (doubles): [x0 | g0 | x1 | g1] 4 Pp, rho[m], alp[m], (cos, sin) cache; the history
// S, Y (2 m Pp elements of HT per team) is a second region behind the slices of all teams.  Vectors are padded to Pp = 4 ceil(P/4) entries that stay zero, so every lane owns
// exactly npl = Pp/4 entries (j = sub + 4 i) and the vector loops need no per-lane bounds checks.
// The search direction is not stored: while a line search is in progress it is (xt - x) / alpha.
// EXTRAS = box bounds and/or per-iteration trace requested: compiled out of the common kernel so that the hot tick body
// stays compact (the tick is instruction-cache bound: adding these paths inline cost 30 % even when unused).
template <int GM, int NPL, typename HT, bool EXTRAS>
__global__ void __launch_bounds__(kLbfgsMaxThreads, 1) lbfgs_kernel(const __grid_constant__ LbfgsArgs A,
                                                                    const __grid_constant__ KTemplate kt) {
  extern __shared__ __align__(16) double smem[];
  const int P = kt.P, m = A.m;
  const int Pp = (P + 3) & ~3;
  const int npl = Pp >> 2;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int team = tid >> 2, sub = tid & 3;
  const unsigned tmask = 0xFu << (lane & ~3);
  double* base = smem + (size_t)team * A.RS;
  // history slices live in their own region with a stride that tiles the 32 banks for HT-sized accesses
  HT* S = reinterpret_cast<HT*>(smem + (size_t)(blockDim.x / LPP) * A.RS) + (size_t)team * A.HS;
  HT* Y = S + m * Pp;
  double* rho = base + 4 * Pp;
  double* alp = rho + m;
  double2* tg = reinterpret_cast<double2*>(alp + m);

  // zero the slice once: the padding entries of every vector must stay zero
  for (int j = sub; j < A.RS; j += LPP) base[j] = 0.0;
  for (int j = sub; j < A.HS; j += LPP) S[j] = (HT)0;
  __syncwarp();

  const int64_t total = A.Nt * (int64_t)A.restarts;
  // team-uniform scalars
  int state = ST_IDLE, cur = 0, iter = 0, ls = 0, hcount = 0, hpos = 0;
  int64_t pid = -1, tgt = 0;
  double f = 0.0, alpha = 1.0, gd = 0.0, gamma = 1.0, f_chk = 0.0;
  bool slow = false;
  unsigned long long evals = 0;
  bool exhausted = false;
  cd vcol[1][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) vcol[0][a] = mkc(0.0, 0.0);

  // The whole tick is WARP-CONVERGENT: every lane executes every section and per-team decisions only predicate
  // stores / scalar updates.  Team reductions can therefore use full-mask xor-1/xor-2 shuffles (which never leave a
  // 4-lane team); team-masked shuffles inside divergent branches compile to a WARPSYNC.COLLECTIVE + BSSY/BSYNC
  // sequence of ~10 instructions each and made up a fifth of the kernel before this restructuring.
  constexpr unsigned FULL = 0xffffffffu;
  while (true) {
    // ---------------- fetch work for idle teams (convergent loop) --------------------------------
    while (true) {
      const bool need = (state == ST_IDLE && !exhausted);
      if (!__any_sync(FULL, need)) break;
      unsigned long long w = 0;
      if (need && sub == 0) w = atomicAdd(A.next, 1ULL);
      w = __shfl_sync(FULL, w, lane & ~3);
      if (need) {
        if ((int64_t)w >= total) {
          exhausted = true;
        } else {
          // restart-major order: all targets' restart 0 first, then restart 1, ...  With many more targets than
          // teams in flight this reproduces the reference's sequential restart loop with its break on first success
          // (optimizer.py:253-295): restart r of a target is skipped once an earlier restart has solved it.
          const int64_t r_idx = (int64_t)w / A.Nt;
          const int64_t t = (int64_t)w - r_idx * A.Nt;
          pid = t * A.restarts + r_idx;  // index of the (target, restart) pair in the output tables / the x0 stream
          tgt = t;
          bool skip = A.active && A.active[t] == 0;
          if (!skip && A.early_exit) skip = *((volatile int32_t*)(A.solved + t)) != 0;
          if (skip) {
            if (sub == 0) {
              A.out_loss[pid] = DBL_MAX;
              A.out_iters[pid] = 0;
            }
            for (int j = sub; j < P; j += LPP) A.out_x[pid * P + j] = 0.0;
          } else {
            // initial point into the trial buffer (buffer 1), target columns into registers
            cur = 0;
            double* x1 = base + 2 * Pp;
            for (int j = sub; j < P; j += LPP)
              x1[j] = A.x0 ? A.x0[pid * A.ldx0 + j] : philox_param(A.seed, (uint64_t)pid, j, A.x0_lo, A.x0_span);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const double2 v = *reinterpret_cast<const double2*>(A.V + t * 32 + (a * 4 + sub) * 2);
              vcol[0][a] = mkc(v.x, v.y);
            }
            state = ST_INIT;
            iter = 0;
            ls = 0;
            hcount = 0;
            hpos = 0;
            gamma = 1.0;
            slow = false;
          }
        }
      }
    }
    if (__all_sync(FULL, state == ST_IDLE)) break;

    // ---------------- one loss+grad evaluation per team -------------------------------------------
    double* xt = base + 2 * (cur ^ 1) * Pp;
    double* gt = xt + Pp;
    const double ft = loss_grad_team<LPP, GM, true>(kt, xt, tg, gt, vcol, A.cost_kind, sub, nullptr);
    const bool live = (state != ST_IDLE);
    if (live) ++evals;

    const double* x = base + 2 * cur * Pp;
    const double* g = x + Pp;
    const bool first = live && (state == ST_INIT);
    const bool accepted = live && (first || (ft <= f + kArmijo * alpha * gd));  // Armijo; NaN compares false
    const bool rejected = live && !accepted;
    bool done = false;

    // ---------------- accept path (all lanes execute; effects predicated on `accepted`) -----------
    double q[NPL];  // this lane's slice of the working vector (entries j = sub + 4 i)
    {
      HT* s_new = S + hpos * Pp;
      HT* y_new = Y + hpos * Pp;
      const bool pair = accepted && !first;
      double sy = 0.0, yy = 0.0, gmax = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        q[i] = 0.0;
        if (i < npl) {
          const int j = sub + LPP * i;
          const double gj = gt[j];
          double gp = gj;  // projected gradient: components pushing against an active bound are dropped
          if (EXTRAS && A.lower && j < P) {
            const double xj = xt[j];
            if ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0)) gp = 0.0;
          }
          q[i] = gp;
          gmax = fmax(gmax, fabs(gp));
          const HT sf = (HT)(xt[j] - x[j]), yf = (HT)(gj - g[j]);
          if (pair) {  // history pair s = xt - x, y = gt - g (rounded to HT; the curvature uses the rounded values)
            s_new[j] = sf;
            y_new[j] = yf;
          }
          sy = fma((double)sf, (double)yf, sy);
          yy = fma((double)yf, (double)yf, yy);
        }
      }
      gmax = team_max<LPP>(gmax);
      sy = team_sum<LPP>(sy);
      yy = team_sum<LPP>(yy);
      if (accepted) {
        if (pair) {
          if (sy > 1e-14 * yy && yy > 0.0) {  // cautious update: keep only positive-curvature pairs
            if (sub == 0) rho[hpos] = 1.0 / sy;
            gamma = sy / yy;
            hpos = (hpos + 1 == m) ? 0 : hpos + 1;
            hcount = min(hcount + 1, m);
          } else if (hcount == m) {
            hcount = m - 1;  // the rejected pair overwrote the oldest slot
          }
          ++iter;
          if (EXTRAS && iter <= A.trace_cap) {  // per-iteration trace (the reference's callbackF)
            const int64_t e = pid * A.trace_cap + (iter - 1);
            if (sub == 0) A.trace_loss[e] = ft;
            if (A.trace_x)
              for (int j = sub; j < P; j += LPP) A.trace_x[e * P + j] = xt[j];
          }
        }
        cur ^= 1;  // trial point becomes the current point
        f = ft;
        // progress checkpoint every 32 accepted steps: "slow" = less than 4x reduction since the last one
        if ((iter & 31) == 0) {
          slow = iter > 0 && f > 0.25 * f_chk;
          f_chk = f;
        }
        // gtol_far is scipy's BFGS default gtol (1e-5), where the reference stops unconditionally.  Here it only
        // ends restarts that sit at a non-zero local minimum (f > f_far) or have stopped making real progress;
        // restarts still converging towards zero loss run on to f_stop / gtol.
        done = (f < A.f_stop) || (gmax < A.gtol) || (gmax < A.gtol_far && (f > A.f_far || slow)) ||
               (iter >= A.max_iter) || !(f == f);
        if (!done && A.early_exit && (iter & 3) == 0) done = *((volatile int32_t*)(A.solved + tgt)) != 0;
      }
    }
    const bool step = accepted && !done;  // teams that need a new search direction
    __syncwarp();                         // s_new / y_new / rho visible to the team
    {
      // two-loop recursion on the register slice: q <- H g.  Loop bound = the largest history in the warp; teams
      // with a shorter history (or not stepping) run with a zero coefficient.
      const int hmax = __reduce_max_sync(FULL, step ? hcount : 0);
      for (int h = 0; h < hmax; ++h) {
        const bool on = step && h < hcount;
        int slot = hpos - 1 - h;
        if (slot < 0) slot += m;
        if (!on) slot = 0;
        const HT* s = S + slot * Pp + sub;
        const HT* y = Y + slot * Pp + sub;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; i += 2) {
          if (i < npl) a0 = fma((double)s[LPP * i], q[i], a0);
          if (i + 1 < NPL && i + 1 < npl) a1 = fma((double)s[LPP * (i + 1)], q[i + 1], a1);
        }
        const double asum = team_sum<LPP>(a0 + a1);
        const double a = on ? asum * rho[slot] : 0.0;
        if (on && sub == 0) alp[slot] = a;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (i < npl) q[i] = fma(-a, (double)y[LPP * i], q[i]);
      }
      __syncwarp();
      if (step) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) q[i] *= gamma;
      }
      for (int h = hmax - 1; h >= 0; --h) {
        const bool on = step && h < hcount;
        int slot = hpos - 1 - h;
        if (slot < 0) slot += m;
        if (!on) slot = 0;
        const HT* s = S + slot * Pp + sub;
        const HT* y = Y + slot * Pp + sub;
        double b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; i += 2) {
          if (i < npl) b0 = fma((double)y[LPP * i], q[i], b0);
          if (i + 1 < NPL && i + 1 < npl) b1 = fma((double)y[LPP * (i + 1)], q[i + 1], b1);
        }
        const double bsum = team_sum<LPP>(b0 + b1);
        const double c = on ? alp[slot] - bsum * rho[slot] : 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (i < npl) q[i] = fma(c, (double)s[LPP * i], q[i]);
      }
      // d = -q ; gd = g.d ; gg = g.g  (g = gt: the accepted gradient)
      double gdn = 0.0, gg = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i)
        if (i < npl) {
          const double gj = gt[sub + LPP * i];
          q[i] = -q[i];
          gdn = fma(gj, q[i], gdn);
          gg = fma(gj, gj, gg);
        }
      gdn = team_sum<LPP>(gdn);
      gg = team_sum<LPP>(gg);
      double* xn = base + 2 * (cur ^ 1) * Pp;  // for stepping teams: the old current buffer
      if (step) {
        if (hcount == 0 || !(gdn < 0.0)) {  // first step or not a descent direction: steepest descent, unit length
          hcount = 0;
#pragma unroll
          for (int i = 0; i < NPL; ++i)
            if (i < npl) q[i] = -gt[sub + LPP * i];
          gdn = -gg;
          alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
        } else {
          alpha = 1.0;
        }
        gd = gdn;
        ls = 0;
        state = ST_LS;
        if (!EXTRAS || !A.lower) {  // next trial point x_new = xt + alpha d
#pragma unroll
          for (int i = 0; i < NPL; ++i)
            if (i < npl) xn[sub + LPP * i] = fma(alpha, q[i], xt[sub + LPP * i]);
        }
      }
      if (EXTRAS && A.lower) {
        // box constraints: project the trial point; the line search then runs along the projected segment, whose
        // directional derivative is g.(x_new - x)/alpha
        double gde = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (i < npl) {
            const int j = sub + LPP * i;
            double v = fma(alpha, q[i], xt[j]);
            if (j < P) v = fmin(fmax(v, A.lower[j]), A.upper[j]);
            if (step) xn[j] = v;
            gde = fma(gt[j], v - xt[j], gde);
          }
        gde = team_sum<LPP>(gde) / alpha;
        const bool sd = step && !(gde < 0.0);  // projection killed the descent: projected steepest descent
        const double asd = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
        double gds = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (i < npl) {
            const int j = sub + LPP * i;
            double v = fma(-asd, gt[j], xt[j]);
            if (j < P) v = fmin(fmax(v, A.lower[j]), A.upper[j]);
            if (sd) xn[j] = v;
            gds = fma(gt[j], v - xt[j], gds);
          }
        gds = team_sum<LPP>(gds) / asd;
        if (step) {
          gd = gde;
          if (sd) {
            hcount = 0;
            alpha = asd;
            gd = gds;
            if (!(gds < 0.0)) done = true;  // no feasible descent direction: a KKT point of the box problem
          }
        }
      }
    }

    // ---------------- backtrack path (all lanes execute; effects predicated on `rejected`) ---------
    {
      // cubic through (0, f, gd) and (alpha, ft, gdt), safeguarded to [0.1, 0.5] alpha; the direction is recovered
      // from the failed trial point: d = (xt - x) / alpha
      double dx[NPL];
      double gdt = 0.0, gg = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        dx[i] = 0.0;
        if (i < npl) {
          const int j = sub + LPP * i;
          dx[i] = xt[j] - x[j];
          gdt = fma(gt[j], dx[i], gdt);
          gg = fma(g[j], g[j], gg);
        }
      }
      gdt = team_sum<LPP>(gdt);
      gg = team_sum<LPP>(gg);
      if (rejected) {
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
        double ratio = an / alpha;
        alpha = an;
        ++ls;
        if (ls > 30) {
          if (hcount > 0) {  // curvature model is bad: restart from steepest descent
            hcount = 0;
#pragma unroll
            for (int i = 0; i < NPL; ++i)
              if (i < npl) dx[i] = -g[sub + LPP * i];
            gd = -gg;
            alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
            ratio = alpha;
            ls = 0;
          } else {
            done = true;  // no progress possible at working precision
          }
        }
        if (!done) {
#pragma unroll
          for (int i = 0; i < NPL; ++i)
            if (i < npl) xt[sub + LPP * i] = fma(ratio, dx[i], x[sub + LPP * i]);
        }
      }
    }
    if (done) {
      const double* xf = base + 2 * cur * Pp;
      if (sub == 0) {
        A.out_loss[pid] = f;
        A.out_iters[pid] = iter;
        if (A.early_exit && f < A.success_threshold) atomicExch(A.solved + tgt, 1);
      }
      for (int j = sub; j < P; j += LPP) A.out_x[pid * P + j] = xf[j];
      state = ST_IDLE;
    }
    __syncwarp();
  }
  if (A.out_evals && sub == 0 && evals) atomicAdd(A.out_evals, evals);
}

static int team_doubles(const KTemplate& kt, int m) {
  const int Pp = (kt.P + 3) & ~3;
  int rs = 4 * Pp + 2 * m + 2 * kt.n_trig;
  while ((rs & 15) != 4) ++rs;  // 32 B (mod 128 B): the 4 teams of a half-warp hit disjoint banks on 64-bit accesses
  return rs;
}

static int team_hist_elems(const KTemplate& kt, int m, int hist_bytes) {
  const int Pp = (kt.P + 3) & ~3;
  int hs = 2 * m * Pp;
  // stride = 4 elements (mod 128 B): 8 teams x 16 B (float) or 4 teams x 32 B (double) tile the 32 banks
  while ((hs * hist_bytes) % 128 != 4 * hist_bytes) ++hs;
  return hs;
}

static size_t team_bytes(const KTemplate& kt, int m, int hist_bytes) {
  return (size_t)team_doubles(kt, m) * 8 + (size_t)team_hist_elems(kt, m, hist_bytes) * hist_bytes;
}

template <int GM, int NPL, typename HT, bool EXTRAS>
static int launch_lbfgs_x(const KTemplate& kt, const LbfgsArgs& A, int grid, int threads, size_t smem, cudaStream_t st) {
  auto kern = lbfgs_kernel<GM, NPL, HT, EXTRAS>;
  SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, threads, smem, st>>>(A, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

template <int GM, int NPL, typename HT>
static int launch_lbfgs(const KTemplate& kt, const LbfgsArgs& A, int grid, int threads, size_t smem, cudaStream_t st) {
  if (A.lower || A.trace_cap > 0) return launch_lbfgs_x<GM, NPL, HT, true>(kt, A, grid, threads, smem, st);
  return launch_lbfgs_x<GM, NPL, HT, false>(kt, A, grid, threads, smem, st);
}

template <int GM, typename HT>
static int dispatch_npl(const KTemplate& kt, const LbfgsArgs& A, int grid, int threads, size_t smem, cudaStream_t st) {
  const int Pp = (kt.P + 3) & ~3;
  if (Pp <= 32) return launch_lbfgs<GM, 8, HT>(kt, A, grid, threads, smem, st);
  if (Pp <= 56) return launch_lbfgs<GM, 14, HT>(kt, A, grid, threads, smem, st);
  if (Pp <= 96) return launch_lbfgs<GM, 24, HT>(kt, A, grid, threads, smem, st);
  return SLAM_ERR_UNSUPPORTED;  // the device optimiser keeps 4*NPL <= 96 parameters in registers
}

template <typename HT>
static int dispatch_gm(const KTemplate& kt, const LbfgsArgs& A, int grid, int threads, size_t smem, cudaStream_t st) {
  switch (kt.gmode) {
    case GM_SYM: return dispatch_npl<GM_SYM, HT>(kt, A, grid, threads, smem, st);
    case GM_BLOCK: return dispatch_npl<GM_BLOCK, HT>(kt, A, grid, threads, smem, st);
    case GM_DENSE: return dispatch_npl<GM_DENSE, HT>(kt, A, grid, threads, smem, st);
    default: return SLAM_ERR_UNSUPPORTED;
  }
}

}  // namespace slam

using namespace slam;

extern "C" void slam_opt_defaults(SlamOptOpts* o) {
  if (!o) return;
  o->max_iter = 2500;            // optimizer.py:274 options={"maxiter": 2500}
  o->history = 0;
  o->cost_kind = SLAM_COST_BASIC;
  o->early_exit = 1;
  o->success_threshold = 1e-10;  // optimizer.py:18
  o->f_stop = 1e-13;
  o->gtol = 1e-9;
  o->gtol_far = 1e-5;            // scipy BFGS default gtol, applied to restarts stuck at a non-zero local minimum
  o->f_far = 1e-6;
  o->x0_lo = 0.0;                // basis.py:111: np.random.random(P) * 2 pi
  o->x0_hi = 6.283185307179586;
  o->trace_cap = 0;
  o->reserved = 0;
  o->trace_loss = nullptr;
  o->trace_x = nullptr;
  o->lower = nullptr;
  o->upper = nullptr;
}

extern "C" int slam_lbfgs_solve(const SlamTemplateDesc* desc, const double* V, int64_t Nt, int32_t restarts,
                                const double* x0, int64_t ldx0, uint64_t seed, const int32_t* active,
                                const SlamOptOpts* opts, double* out_loss, double* out_x, int32_t* out_iters,
                                unsigned long long* out_evals, void* stream) {
  if (!desc || !V || !opts || !out_loss || !out_x || !out_iters || Nt < 0 || restarts < 1) return SLAM_ERR_INVALID;
  if (x0 && ldx0 < desc->n_params) return SLAM_ERR_INVALID;
  if (opts->cost_kind != SLAM_COST_BASIC && opts->cost_kind != SLAM_COST_SQUARE) return SLAM_ERR_UNSUPPORTED;
  if (opts->max_iter < 1 || opts->history < 0 || opts->history > kMaxHist) return SLAM_ERR_INVALID;
  if (desc->n_params < 1) return SLAM_ERR_INVALID;
  if (Nt == 0) return SLAM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  KTemplate kt;
  int rc = compile_template(desc, &kt, /*allow_bound_smush=*/false);
  if (rc != SLAM_OK) return rc;
  if (kt.P > 96) return SLAM_ERR_UNSUPPORTED;
  if (kt.gmode == GM_DENSE && desc->gate_kind != SLAM_GATE_FIXED) {
    rc = lower_const_smush(desc, &kt, st);
    if (rc != SLAM_OK) return rc;
  }
  int dev = 0, sms = 0, max_smem = 0;
  SLAM_CUDA_CHECK(cudaGetDevice(&dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SLAM_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // history storage: float by default (halves the dominant shared-memory consumer -> 1.5x the resident teams; the
  // curvature scalars are computed from the rounded pairs, so the two-loop recursion stays self-consistent);
  // SLAM_B200_HIST_F64=1 keeps the pairs in double for A/B comparisons.
  const char* henv = getenv("SLAM_B200_HIST_F64");
  const bool hist64 = henv && henv[0] == '1';
  const int hb = hist64 ? 8 : 4;
  // history length and teams per CTA from the shared-memory budget (one persistent CTA per SM)
  const int max_teams = kLbfgsMaxThreads / LPP;
  int m = opts->history ? opts->history : 6;
  int teams = max_teams;
  if (!opts->history)  // prefer a full complement of teams (occupancy) over a longer history, down to m = 4
    while (m > 4 && team_bytes(kt, m, hb) * max_teams > (size_t)max_smem) --m;
  while (teams > 8 && team_bytes(kt, m, hb) * teams > (size_t)max_smem) teams -= 8;
  const int RS = team_doubles(kt, m);
  const int HS = team_hist_elems(kt, m, hb);
  const size_t smem = team_bytes(kt, m, hb) * teams;
  if (smem > (size_t)max_smem) return SLAM_ERR_UNSUPPORTED;

  // stream-ordered scratch: work counter + per-target early-exit flags
  rc = keep_async_pool(dev);
  if (rc != SLAM_OK) return rc;
  unsigned long long* next = nullptr;
  int32_t* solved = nullptr;
  SLAM_CUDA_CHECK(cudaMallocAsync((void**)&next, sizeof(unsigned long long), st));
  SLAM_CUDA_CHECK(cudaMallocAsync((void**)&solved, sizeof(int32_t) * (size_t)Nt, st));
  SLAM_CUDA_CHECK(cudaMemsetAsync(next, 0, sizeof(unsigned long long), st));
  SLAM_CUDA_CHECK(cudaMemsetAsync(solved, 0, sizeof(int32_t) * (size_t)Nt, st));

  LbfgsArgs A;
  A.V = V; A.x0 = x0; A.ldx0 = ldx0; A.seed = seed; A.active = active; A.Nt = Nt; A.restarts = restarts;
  A.m = m; A.RS = RS; A.HS = HS; A.max_iter = opts->max_iter; A.cost_kind = opts->cost_kind; A.early_exit = opts->early_exit;
  A.success_threshold = opts->success_threshold; A.f_stop = opts->f_stop; A.gtol = opts->gtol;
  A.gtol_far = opts->gtol_far; A.f_far = opts->f_far; A.x0_lo = opts->x0_lo; A.x0_span = opts->x0_hi - opts->x0_lo;
  A.trace_cap = (opts->trace_loss && opts->trace_cap > 0) ? opts->trace_cap : 0;
  A.trace_loss = opts->trace_loss; A.trace_x = opts->trace_x;
  A.lower = (opts->lower && opts->upper) ? opts->lower : nullptr;
  A.upper = A.lower ? opts->upper : nullptr;
  A.out_loss = out_loss; A.out_x = out_x; A.out_iters = out_iters; A.out_evals = out_evals;
  A.next = next; A.solved = solved;

  const int64_t total = Nt * (int64_t)restarts;
  const int grid = (int)std::min<int64_t>((int64_t)sms, (total + teams - 1) / teams);
  const int threads = teams * LPP;
  rc = hist64 ? dispatch_gm<double>(kt, A, grid, threads, smem, st) : dispatch_gm<float>(kt, A, grid, threads, smem, st);
  cudaFreeAsync(next, st);
  cudaFreeAsync(solved, st);
  return rc;
}
