// slam_lbfgs.cuh -- K5: device-resident batched L-BFGS over (target, restart) problems (kernel template).
//
// Replaces the restart loop around scipy.optimize.minimize(method="BFGS") (src/slam/optimizer.py:253-295):
// instead of one Python-driven BFGS with (P+1)-evaluation finite-difference gradients per restart, every
// (target, restart) pair is an independent problem solved by a TEAM of LPP lanes (2 or 4) with analytic adjoint
// gradients (slam_core.cuh).  All optimiser state of a problem -- x, g, trial point, the L-BFGS (s, y) history
// and the (cos, sin) cache -- lives in that team's slice of shared memory; nothing but the final result touches
// HBM.  A persistent grid pulls problems from a global counter, so early exits and uneven iteration counts do
// not leave SMs idle.
//
// Warp-level structure: each "tick" every team of the warp performs exactly one loss+gradient evaluation
// (the expensive, fully convergent part), then runs its own cheap line-search / history bookkeeping with the
// per-team decisions turned into predicates.  In the bookkeeping each lane owns the vector entries
// j = sub + LPP i (i < NPL) and keeps its slice of the two-loop working vector in registers, so the recursion is
// a stream of independent shared-memory loads and FMAs with one shuffle reduction per history pair.
//
// The kernel is instantiated per gate mode in slam_lbfgs_{sym_hi32,sym,block,dense}.cu (parallel compilation); the C ABI
// entry points and the launch configuration live in slam_lbfgs.cu.
#pragma once
#include <cfloat>

#include "slam_host.h"
#include "slam_philox.cuh"

namespace slam {

constexpr int kMaxHist = 8;
constexpr double kArmijo = 1e-4;

struct LbfgsArgs {
  const double* V;
  const double* x0;
  int64_t ldx0;
  uint64_t seed;
  const int32_t* active;
  int64_t Nt;
  int restarts;
  int m;        // history length
  int RS;       // doubles of shared memory per team (vectors, alp, trig cache)
  int HS;       // history elements (of HT) per team; the history slices follow the RS slices of all teams
  int Pp;       // padded vector length (multiple of LPP; exactly LPP * NPL for EXACT kernels)
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
  const int32_t* solved_in;  // flags of the launch for the previous template size (may still be written), or null
  unsigned long long* best_key;  // per-target packed (success class, k, loss, restart) minimum (SlamOptOpts.best_key), or null
};

// launch configuration chosen on the host (slam_lbfgs.cu)
struct LbfgsCfg {
  int lpp;      // lanes per problem: 2 or 4
  int maxt;     // CTA size the kernel is compiled for (4 lanes: 384 or 512)
  int npl;      // vector entries per lane held in registers
  int exact;    // canonical template (P = 6(k+1), every layer present, one constant symmetric gate) with npl == Pp / lpp
                // known at compile time: no per-entry guards, no index tables, vectors kept in circuit creation order
  int extras;   // bounds and/or trace
  int grid, threads;
  size_t smem;
};

// history element: the upper half of the double (sign, exponent, 20 mantissa bits) halves the dominant shared-memory consumer
// (the curvature scalars are computed from the rounded pairs, so the two-loop recursion stays self-consistent) and
// unpacking is a register move instead of an F2F conversion on the quarter-rate pipe
struct HistHi32 {
  typedef int T;
  static __device__ __forceinline__ T pack(double v) {
    const long long b = __double_as_longlong(v) + 0x80000000LL;  // round to nearest on the kept bits
    return (int)(b >> 32);
  }
  static __device__ __forceinline__ double unpack(T v) { return __hiloint2double(v, 0); }
};

enum { ST_IDLE = 0, ST_INIT = 1, ST_LS = 2 };

// Packed per-target reduction key (SlamOptOpts.best_key): a 64-bit unsigned minimum over all restarts of all chained launches
// implements the merge rule of optimizer.py:283-303 -- the smallest template size that reached the threshold, else the lowest
// loss seen -- with ONE fire-and-forget atomic per retired restart (no lock, nothing for the warp to wait on):
//   bit 63      0 = loss below the success threshold, 1 = not
//   bits 62-59  k for successes (smaller k wins), 0 otherwise
//   bits 58-12  the top 47 bits of the (non-negative) loss: sign-less exponent + 36 mantissa bits, order preserving
//   bits 11-8   k          bits 7-0  restart       (where the winner's row is: slam_best_gather reads it back)
// Non-finite losses never enter the reduction.
__host__ __device__ __forceinline__ unsigned long long best_key_pack(double f, bool success, int k, int restart) {
#ifdef __CUDA_ARCH__
  const unsigned long long fb = (unsigned long long)__double_as_longlong(f);
#else
  unsigned long long fb;
  memcpy(&fb, &f, sizeof(fb));
#endif
  return (success ? 0ULL : (1ULL << 63)) | ((unsigned long long)(success ? (k & 15) : 0) << 59) | (((fb >> 16) & ((1ULL << 47) - 1)) << 12) |
         ((unsigned long long)(k & 15) << 8) | (unsigned long long)(restart & 255);
}

// LPP  = lanes per problem; MAXT = CTA size the kernel is compiled for (register cap = 64K / MAXT);
// NPL  = vector entries per lane held in registers (Pp <= LPP * NPL; Pp == LPP * NPL when EXACT);
// EXACT also selects the canonical parameter layout of slam_core.cuh (CANON): x, g and the history are kept in circuit
// creation order and permuted to/from the API (name-sorted) order only when a problem is fetched or retired;
// H    = storage policy of the (s, y) history.
//
// Shared-memory slice of a team (doubles): [x0 | g0 | x1 | g1] 4 Pp, alp[m], (cos, sin) cache; the history
// S, Y, rho (2 m Pp + m elements per team) is a second region behind the slices of all teams.  Vectors are padded to Pp entries
// that stay zero, so every lane owns exactly npl = Pp/LPP entries (j = sub + LPP i) and the vector loops need no
// per-lane bounds checks.  The search direction is not stored: while a line search is in progress it is (xt - x)/alpha.
// EXTRAS = box bounds and/or per-iteration trace requested: compiled out of the common kernel so that the hot tick body
// stays compact (the tick is instruction-cache bound: adding these paths inline cost 30 % even when unused).
template <int LPP, int MAXT, int GM, int NPL, bool EXACT, class H, bool EXTRAS>
__global__ void __launch_bounds__(MAXT, 1) lbfgs_kernel(const __grid_constant__ LbfgsArgs A,
                                                        const __grid_constant__ KTemplate kt) {
  typedef typename H::T HT;
  constexpr int CPL = 4 / LPP;
  extern __shared__ __align__(16) double smem[];
  const int P = kt.P, m = A.m;
  const int Pp = EXACT ? LPP * NPL : A.Pp;
  const int npl = EXACT ? NPL : Pp / LPP;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int team = tid / LPP, sub = tid % LPP;
  double* base = smem + (size_t)team * A.RS;
  // history slices live in their own region with a stride that tiles the 32 banks for HT-sized accesses
  HT* S = reinterpret_cast<HT*>(smem + (size_t)(blockDim.x / LPP) * A.RS) + (size_t)team * A.HS;
  HT* Y = S + m * Pp;
  HT* rho = Y + m * Pp;  // 1 / (s.y) per pair, stored like the pairs themselves
  double* alp = base + 4 * Pp;
  double2* tg = reinterpret_cast<double2*>(alp + (m + (m & 1)));

  // zero the slice once: the padding entries of every vector must stay zero
  for (int j = sub; j < A.RS; j += LPP) base[j] = 0.0;
  for (int j = sub; j < A.HS; j += LPP) S[j] = H::pack(0.0);
  __syncwarp();

  const int64_t total = A.Nt * (int64_t)A.restarts;
  // team-uniform scalars
  int state = ST_IDLE, cur = 0, iter = 0, ls = 0, hcount = 0, hpos = 0;
  int64_t pid = -1, tgt = 0;
  double f = 0.0, alpha = 1.0, gd = 0.0, gamma = 1.0, f_chk = 0.0;
  bool slow = false;
  unsigned long long evals = 0;
  bool exhausted = false;
  cd vcol[CPL][4];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) vcol[c][a] = mkc(0.0, 0.0);

  // The whole tick is WARP-CONVERGENT: every lane executes every section and per-team decisions only predicate
  // stores / scalar updates.  Team reductions can therefore use full-mask xor shuffles (which never leave a team);
  // team-masked shuffles inside divergent branches compile to a WARPSYNC.COLLECTIVE + BSSY/BSYNC sequence of ~10
  // instructions each and made up a fifth of the kernel before this restructuring.
  constexpr unsigned FULL = 0xffffffffu;
  while (true) {
    // ---------------- fetch work for idle teams (convergent loop) --------------------------------
    while (true) {
      const bool need = (state == ST_IDLE && !exhausted);
      if (!__any_sync(FULL, need)) break;
      unsigned long long w = 0;
      if (need && sub == 0) w = atomicAdd(A.next, 1ULL);
      w = __shfl_sync(FULL, w, lane & ~(LPP - 1));
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
          if (!skip && A.solved_in && *((volatile const int32_t*)(A.solved_in + t)) != 0) {
            skip = true;  // solved at a smaller template size: propagate, so that the flags stay cumulative
            if (sub == 0) A.solved[t] = 1;
          }
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
            for (int c = sub; c < P; c += LPP) {
              const int j = EXACT ? kt.p1q[c / 6][c % 6] : c;  // API index of internal entry c
              double v = A.x0 ? A.x0[pid * A.ldx0 + j] : philox_param(A.seed, (uint64_t)pid, j, A.x0_lo, A.x0_span);
              if (EXTRAS && A.lower) v = fmin(fmax(v, A.lower[j]), A.upper[j]);  // start inside the box
              x1[c] = v;
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c)
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const double2 v = *reinterpret_cast<const double2*>(A.V + t * 32 + (a * 4 + sub * CPL + c) * 2);
                vcol[c][a] = mkc(v.x, v.y);
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
    const double ft = loss_grad_team<LPP, GM, true, EXACT>(kt, xt, tg, gt, vcol, A.cost_kind, sub, nullptr);
    const bool live = (state != ST_IDLE);
    if (live) ++evals;

    const double* x = base + 2 * cur * Pp;
    const double* g = x + Pp;
    const bool first = live && (state == ST_INIT);
    const bool accepted = live && (first || (ft <= f + kArmijo * alpha * gd));  // Armijo; NaN compares false
    const bool rejected = live && !accepted;
    bool done = false;

    // ---------------- accept path (all lanes execute; effects predicated on `accepted`) -----------
    double q[NPL];  // this lane's slice of the working vector (entries j = sub + LPP i)
    {
      HT* s_new = S + hpos * Pp;
      HT* y_new = Y + hpos * Pp;
      const bool pair = accepted && !first;
      double sy = 0.0, yy = 0.0, gmax = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        q[i] = 0.0;
        if (EXACT || i < npl) {
          const int j = sub + LPP * i;
          const double gj = gt[j];
          double gp = gj;  // projected gradient: components pushing against an active bound are dropped
          if (EXTRAS && A.lower && j < P) {
            const double xj = xt[j];
            if ((xj <= A.lower[j] && gj > 0.0) || (xj >= A.upper[j] && gj < 0.0)) gp = 0.0;
          }
          q[i] = gp;
          gmax = fmax(gmax, fabs(gp));
          const HT sf = H::pack(xt[j] - x[j]), yf = H::pack(gj - g[j]);
          if (pair) {  // history pair s = xt - x, y = gt - g (rounded to HT; the curvature uses the rounded values)
            s_new[j] = sf;
            y_new[j] = yf;
          }
          sy = fma(H::unpack(sf), H::unpack(yf), sy);
          yy = fma(H::unpack(yf), H::unpack(yf), yy);
        }
      }
      gmax = team_max<LPP>(gmax);
      sy = team_sum<LPP>(sy);
      yy = team_sum<LPP>(yy);
      if (accepted) {
        if (pair) {
          if (sy > 1e-14 * yy && yy > 0.0) {  // cautious update: keep only positive-curvature pairs
            if (sub == 0) rho[hpos] = H::pack(1.0 / sy);
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
          if (EXACT || i < npl) a0 = fma(H::unpack(s[LPP * i]), q[i], a0);
          if (i + 1 < NPL && (EXACT || i + 1 < npl)) a1 = fma(H::unpack(s[LPP * (i + 1)]), q[i + 1], a1);
        }
        const double asum = team_sum<LPP>(a0 + a1);
        const double a = on ? asum * H::unpack(rho[slot]) : 0.0;
        if (on && sub == 0) alp[slot] = a;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (EXACT || i < npl) q[i] = fma(-a, H::unpack(y[LPP * i]), q[i]);
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
          if (EXACT || i < npl) b0 = fma(H::unpack(y[LPP * i]), q[i], b0);
          if (i + 1 < NPL && (EXACT || i + 1 < npl)) b1 = fma(H::unpack(y[LPP * (i + 1)]), q[i + 1], b1);
        }
        const double bsum = team_sum<LPP>(b0 + b1);
        const double c = on ? alp[slot] - bsum * H::unpack(rho[slot]) : 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (EXACT || i < npl) q[i] = fma(c, H::unpack(s[LPP * i]), q[i]);
      }
      // d = -q ; gd = g.d ; gg = g.g  (g = gt: the accepted gradient)
      double gdn = 0.0, gg = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i)
        if (EXACT || i < npl) {
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
            if (EXACT || i < npl) q[i] = -gt[sub + LPP * i];
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
            if (EXACT || i < npl) xn[sub + LPP * i] = fma(alpha, q[i], xt[sub + LPP * i]);
        }
      }
      if (EXTRAS && A.lower) {
        // box constraints: project the trial point; the line search then runs along the projected segment, whose
        // directional derivative is g.(x_new - x)/alpha
        double gde = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (EXACT || i < npl) {
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
          if (EXACT || i < npl) {
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

    // ---------------- backtrack path (warp-uniform branch; effects predicated on `rejected`) ------------
    // Most ticks no team of the warp has a rejected trial (L-BFGS steps are usually accepted at alpha = 1), so the whole
    // section sits behind a warp vote; inside it every lane executes every reduction (full-mask shuffles stay legal).
    if (__any_sync(FULL, rejected)) {
      // cubic through (0, f, gd) and (alpha, ft, gdt), safeguarded to [0.1, 0.5] alpha; the direction is recovered
      // from the failed trial point: d = (xt - x) / alpha
      double dx[NPL];
      double gdt = 0.0;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        dx[i] = 0.0;
        if (EXACT || i < npl) {
          const int j = sub + LPP * i;
          dx[i] = xt[j] - x[j];
          gdt = fma(gt[j], dx[i], gdt);
        }
      }
      gdt = team_sum<LPP>(gdt);
      double ratio = 0.0;
      bool sd = false;  // curvature model is bad: restart from steepest descent
      if (rejected) {
        const double ia = 1.0 / alpha;
        gdt *= ia;
        double an = 0.5 * alpha;
        if (ft == ft && gdt == gdt) {
          const double d1 = gd + gdt - 3.0 * (ft - f) * ia;
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
        ratio = an * ia;
        alpha = an;
        ++ls;
        if (ls > 30) {
          if (hcount > 0) sd = true;
          else done = true;  // no progress possible at working precision
        }
      }
      if (__any_sync(FULL, sd)) {  // rare
        double gg = 0.0;
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (EXACT || i < npl) gg = fma(g[sub + LPP * i], g[sub + LPP * i], gg);
        gg = team_sum<LPP>(gg);
        if (sd) {
          hcount = 0;
#pragma unroll
          for (int i = 0; i < NPL; ++i)
            if (EXACT || i < npl) dx[i] = -g[sub + LPP * i];
          gd = -gg;
          alpha = fmin(1.0, rsqrt(fmax(gg, 1e-300)));
          ratio = alpha;
          ls = 0;
        }
      }
      if (rejected && !done) {
#pragma unroll
        for (int i = 0; i < NPL; ++i)
          if (EXACT || i < npl) xt[sub + LPP * i] = fma(ratio, dx[i], x[sub + LPP * i]);
      }
    }
    if (done) {
      const double* xf = base + 2 * cur * Pp;
      if (sub == 0) {
        A.out_loss[pid] = f;
        A.out_iters[pid] = iter;
        if (A.early_exit && f < A.success_threshold) atomicExch(A.solved + tgt, 1);
        if (A.best_key && f == f && f >= 0.0 && f < DBL_MAX)
          atomicMin(A.best_key + tgt, best_key_pack(f, f < A.success_threshold, kt.k, (int)(pid - tgt * A.restarts)));
      }
      for (int c = sub; c < P; c += LPP) A.out_x[pid * P + (EXACT ? kt.p1q[c / 6][c % 6] : c)] = xf[c];
      state = ST_IDLE;
    }
    __syncwarp();
  }
  if (A.out_evals && sub == 0 && evals) atomicAdd(A.out_evals, evals);
}

template <int LPP, int MAXT, int GM, int NPL, bool EXACT, class H, bool EXTRAS>
static int launch_lbfgs(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  auto kern = lbfgs_kernel<LPP, MAXT, GM, NPL, EXACT, H, EXTRAS>;
  if (c.threads > MAXT) return SLAM_ERR_INVALID;
  SLAM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  kern<<<c.grid, c.threads, c.smem, st>>>(A, kt);
  SLAM_CUDA_CHECK(cudaGetLastError());
  return SLAM_OK;
}

// CTA sizes the two team widths are compiled for: 4 lanes -> 384 threads (register cap 168), 2 lanes -> 256 threads
// (register cap 255; each lane carries two matrix columns and the kernels use 200-211; ptxas applies the 384-thread cap of 168
// registers to any CTA size above 256, which spills)
constexpr int kMaxT4 = 384;
constexpr int kMaxT4x = 512;  // register cap 128: only the exact GM_SYM kernels fit without spills
constexpr int kMaxT2 = 256;

// generic (guarded) instantiations; `EX` adds the exact-length ones used by the headline templates P = 6(k+1)
template <int GM, class H, bool EXTRAS>
static int dispatch_generic(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
  if (c.lpp == 4) {
    if (c.npl <= 8) return launch_lbfgs<4, kMaxT4, GM, 8, false, H, EXTRAS>(kt, A, c, st);
    if (c.npl <= 14) return launch_lbfgs<4, kMaxT4, GM, 14, false, H, EXTRAS>(kt, A, c, st);
    if (c.npl <= 24) return launch_lbfgs<4, kMaxT4, GM, 24, false, H, EXTRAS>(kt, A, c, st);
  } else if (c.lpp == 2) {
    if (c.npl <= 16) return launch_lbfgs<2, kMaxT2, GM, 16, false, H, EXTRAS>(kt, A, c, st);
    if (c.npl <= 28) return launch_lbfgs<2, kMaxT2, GM, 28, false, H, EXTRAS>(kt, A, c, st);
  }
  return SLAM_ERR_UNSUPPORTED;
}

template <int GM, class H>
static int dispatch_exact(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st) {
#define SLAM_EXACT(L, MT, N) \
  if (c.lpp == L && c.maxt == MT && c.npl == N) return launch_lbfgs<L, MT, GM, N, true, H, false>(kt, A, c, st);
  SLAM_EXACT(4, kMaxT4x, 3)  // 16 warps/SM where 128 teams fit in shared memory (k <= 3)
  SLAM_EXACT(4, kMaxT4x, 5)
  SLAM_EXACT(4, kMaxT4x, 6)
  SLAM_EXACT(4, kMaxT4, 3)   // P = 12
  SLAM_EXACT(4, kMaxT4, 5)   // P = 18 (padded to 20)
  SLAM_EXACT(4, kMaxT4, 6)   // P = 24
  SLAM_EXACT(4, kMaxT4, 8)   // P = 30 (padded to 32)
  SLAM_EXACT(4, kMaxT4, 9)   // P = 36
  SLAM_EXACT(4, kMaxT4, 11)  // P = 42 (padded to 44)
  SLAM_EXACT(2, kMaxT2, 6)
  SLAM_EXACT(2, kMaxT2, 9)
  SLAM_EXACT(2, kMaxT2, 12)
  SLAM_EXACT(2, kMaxT2, 15)
  SLAM_EXACT(2, kMaxT2, 18)
  SLAM_EXACT(2, kMaxT2, 21)
#undef SLAM_EXACT
  return SLAM_ERR_UNSUPPORTED;
}

// per-gate-mode entry points (one translation unit each)
int lbfgs_launch_sym(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st);
int lbfgs_launch_block(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st);
int lbfgs_launch_dense(const KTemplate& kt, const LbfgsArgs& A, const LbfgsCfg& c, cudaStream_t st);
bool lbfgs_has_exact(int lpp, int npl);

}  // namespace slam
