// slam_objective.cuh -- generic objective of the thread-per-problem optimisers (K5b Nelder-Mead, K5c FD-gradient L-BFGS):
// any template the forward evaluator handles (closed-form, fixed, smush gates; U3 or RZ layers) against any in-scope
// cost functional (trace-based, Weyl/Makhlin-coordinate based, canonical-reduced)  (src/slam/cost_function.py:133-221,
// src/slam/optimizer.py:191-214).
#pragma once
#include "slam_fwd1.cuh"
#include "slam_weyl.cuh"

namespace slam {

struct StridedParams {
  const double* p;
  int64_t stride;
  __device__ __forceinline__ double get(int j) const { return p[(int64_t)j * stride]; }
};

// cost of U against a target described by (V, its Weyl coords, its Makhlin invariants)
struct TargetInfo {
  const double* V;  // 32 doubles, row-major
  double c[3], g[3];
  // circuit-cost constraint of this problem (augmented Lagrangian; pen_mu = 0: none)
  double pen_mu, pen_max, pen_lambda;
};

// CircuitTemplateV2.circuit_cost (basisv2.py:98-127): RiSwap alpha (custom_gates.py:568-572); ConversionGain and its smush
// form (|gc| + |gg|) t / (pi/2) (custom_gates.py:208-212, 252-257); every other gate scores 0
template <class PS>
__device__ __forceinline__ double circuit_cost(const KTemplate& kt, const PS& ps) {
  double c = 0.0;
  for (int g = 0; g < kt.k; ++g) {
    if (kt.gate_kind == SLAM_GATE_RISWAP) {
      c += slot_val(kt, ps, g, 0);
    } else if (kt.gate_kind == SLAM_GATE_CG || kt.gate_kind == SLAM_GATE_SMUSH) {
      const double tt = slot_val(kt, ps, g, kt.n_slots - 1);
      c += (fabs(slot_val(kt, ps, g, 2)) + fabs(slot_val(kt, ps, g, 3))) * tt * (1.0 / 1.5707963267948966);
    }
  }
  return c;
}

__device__ __forceinline__ double generic_cost(const cd R[4][4] /*[col][row]*/, const TargetInfo& t, int kind) {
  if (kind <= SLAM_COST_BASIC_INVERSE) {
    cd T = mkc(0.0, 0.0);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const cd v = mkc(t.V[(r * 4 + c) * 2], t.V[(r * 4 + c) * 2 + 1]);
        const cd u = R[c][r];
        T.re = fma(v.re, u.re, fma(v.im, u.im, T.re));
        T.im = fma(v.re, u.im, fma(-v.im, u.re, T.im));
      }
    double loss, dl;
    cost_from_abs(kind, sqrt(fma(T.re, T.re, T.im * T.im)), loss, dl);
    return loss;
  }
  cd M[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) M[r][c] = R[c][r];
  double cc[3], gg[3];
  const bool need_g = (kind == SLAM_COST_MAKHLIN_FUNCTIONAL || kind == SLAM_COST_MAKHLIN_EUCLIDEAN);
  weyl_makhlin(M, SLAM_WEYL_ROUND8, need_g ? nullptr : cc, need_g ? gg : nullptr);  // 8-dp rounded, as the reference
  if (kind == SLAM_COST_MAKHLIN_FUNCTIONAL || kind == SLAM_COST_MAKHLIN_EUCLIDEAN) {
    const double d0 = t.g[0] - gg[0], d1 = t.g[1] - gg[1], d2 = t.g[2] - gg[2];
    const double s = d0 * d0 + d1 * d1 + d2 * d2;
    return kind == SLAM_COST_MAKHLIN_FUNCTIONAL ? s : sqrt(s);
  }
  const double d0 = cc[0] - t.c[0], d1 = cc[1] - t.c[1], d2 = cc[2] - t.c[2];
  if (kind == SLAM_COST_WEYL_EUCLIDEAN) return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  // canonical-reduced costs: Tr(Can(c_t)^dagger Can(c_u)) from the coordinate differences (cost_function.py:176-189)
  const double h = 1.5707963267948966;
  const double a0 = h * (d0 - d1 + d2), a1 = h * (d0 + d1 - d2), a2 = -h * (d0 + d1 + d2), a3 = h * (-d0 + d1 + d2);
  double s, c, tr = 0.0, ti = 0.0;
  sincos(a0, &s, &c); tr += c; ti += s;
  sincos(a1, &s, &c); tr += c; ti += s;
  sincos(a2, &s, &c); tr += c; ti += s;
  sincos(a3, &s, &c); tr += c; ti += s;
  double loss, dl;
  cost_from_abs(kind == SLAM_COST_BASIC_REDUCED ? SLAM_COST_BASIC : SLAM_COST_SQUARE, sqrt(tr * tr + ti * ti), loss, dl);
  return loss;
}

// Weyl / Makhlin data of the target, needed by the coordinate-based functionals only
__device__ __forceinline__ void target_info_init(TargetInfo& ti, const double* V, int cost_kind) {
  ti.V = V;
  ti.pen_mu = ti.pen_max = ti.pen_lambda = 0.0;
  if (cost_kind > SLAM_COST_BASIC_INVERSE) {
    cd M[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) M[r][c] = mkc(V[(r * 4 + c) * 2], V[(r * 4 + c) * 2 + 1]);
    weyl_makhlin(M, SLAM_WEYL_ROUND8, ti.c, ti.g);
  }
}

// objective value at the parameters `ps`:  cost(U(x), V), or 1 - fidelity * circuit_fidelity(x) for BasicCostInverse
// (optimizer.py:200-201; circuit_fidelity = product over riswap gates of RiSwapGate(alpha).cost() = alpha, basisv2.py:129-141)
template <class PS>
__device__ __forceinline__ double objective_value(const KTemplate& kt, const PS& ps, const TargetInfo& ti, int cost_kind) {
  cd R[4][4];
  fwd1_chain(kt, ps, R);
  double c = generic_cost(R, ti, cost_kind);
  if (cost_kind == SLAM_COST_BASIC_INVERSE) {
    double F = 1.0;
    if (kt.gate_kind == SLAM_GATE_RISWAP)
      for (int g = 0; g < kt.k; ++g) F *= slot_val(kt, ps, g, 0);
    c = 1.0 - c * F;
  }
  if (ti.pen_mu > 0.0) {
    const double v = fmax(0.0, circuit_cost(kt, ps) - ti.pen_max + ti.pen_lambda / ti.pen_mu);
    c = fma(0.5 * ti.pen_mu * v, v, c);
  }
  return c;
}

// parameters of a workspace vector, optionally with one entry shifted: x + h e_j (j < 0: no shift)
struct ShiftedParams {
  const double* p;
  int64_t stride;
  int j;
  double h;
  __device__ __forceinline__ double get(int i) const {
    const double v = p[(int64_t)i * stride];
    return i == j ? v + h : v;
  }
};

// ONE out-of-line copy of the objective per kernel.  The thread-per-problem optimisers evaluate it from several places
// (line search, every finite difference, every simplex move); inlined at each of them the kernels were 320-720 KB of
// code against a 32 KB instruction cache.  `kt` points at the kernel's __grid_constant__ parameter.
static __device__ __noinline__ double objective_value_nl(const KTemplate* kt, const double* p, int64_t stride, int j,
                                                         double h, const TargetInfo* ti, int cost_kind) {
  ShiftedParams ps{p, stride, j, h};
  return objective_value(*kt, ps, *ti, cost_kind);
}

}  // namespace slam
