// slam_core.cuh -- device core shared by the eval / loss+grad / L-BFGS / coverage kernels.
//
// Work decomposition ("team" design):
//   a template evaluation is a chain U = L_k G_{k-1} L_{k-1} ... G_0 L_0 of 4x4 complex128 matrices.
//   One problem is handled by a TEAM of LPP lanes (LPP = 1, 2 or 4) of one warp.  Each lane owns
//   CPL = 4/LPP columns of the running right-product R_i and the matching rows of the running
//   left-product W_i, so every matrix product in both sweeps is a lane-local 4-vector update and
//   only scalar reductions (the trace, the per-layer partial derivatives) cross lanes by shuffle.
//   The backward sweep is a reversible adjoint: because every factor is unitary, R_i is recovered
//   from R_{i+1} by multiplying with the adjoint factors, so nothing is stored per layer.
//
// Reference behaviour restated here (paths relative to the reference root):
//   layer/gate order ........ src/slam/basis.py:152-169, src/slam/basisv2.py:262-299
//   U3 / RZ matrices ........ qiskit UGate / RZGate (reached from basis.py:157,168; basisv2.py:267-298)
//   RiSwap .................. src/slam/utils/gates/custom_gates.py:582-595
//   ConversionGain .......... custom_gates.py:163-184 -> src/slam/hamiltonian.py:84-111 (closed form)
//   cost functionals ........ src/slam/cost_function.py:133-173
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/slam_b200.h"

namespace slam {

// ------------------------------------------------------------------------------------------------
// kernel-side template descriptor (built on the host from SlamTemplateDesc, passed by value)
// ------------------------------------------------------------------------------------------------
enum GateMode : int {
  GM_SYM = 0,    // constant gate [[co,0,0,i so],[0,ci,i si,0],[0,i si,ci,0],[i so,0,0,co]] (RiSwap, CG with zero phases)
  GM_BLOCK = 1,  // two 2x2 blocks with complex off-diagonals; constant or bound to parameters
  GM_DENSE = 2,  // constant dense 4x4 per repetition (fixed gates, constant smush gates)
  GM_SMUSH = 3   // parameter-bound smush gate: per-slice exp(-i dt H) (forward evaluation only)
};

// Field order matters: kernel parameters beyond the first 4 KB of the parameter space are served by a slower path
// (measured: +33 % sweep time when the hot tables crossed it), so the fields the GM_SYM hot path reads come first and
// the big, rarely used tables (slot bindings, dense gates) last.  Kernels also take this struct as their LAST parameter.
struct KTemplate {
  int k;         // 2Q gate applications
  int P;         // parameters
  int gmode;     // GateMode
  int gate_kind; // SlamGateKind
  int T;         // smush slices
  int n_slots;
  int vz_only;
  int n_trig;    // entries of the per-problem trig cache: 6(k+1) [+ 4k for parameter-bound block gates]
  short p1q[SLAM_MAX_K + 1][6];
  int gate_bound[SLAM_MAX_K];        // GM_BLOCK: 1 if any slot of gate g is a parameter
  double gsym[SLAM_MAX_K][4];        // GM_SYM: co, so, ci, si
  double gblk[SLAM_MAX_K][8];        // GM_BLOCK constant gates: (cos,sin) of phi_c, phi_g, a_c, a_g
  short slot_param[SLAM_MAX_K][SLAM_MAX_SLOTS];
  double slot_const[SLAM_MAX_K][SLAM_MAX_SLOTS];
  double dense[SLAM_MAX_K][32];      // GM_DENSE
};

// ------------------------------------------------------------------------------------------------
// complex helpers (plain structs; everything lives in registers)
// ------------------------------------------------------------------------------------------------
struct cd {
  double re, im;
};
__device__ __forceinline__ cd mkc(double r, double i) { return cd{r, i}; }
__device__ __forceinline__ cd cmul(cd a, cd b) { return cd{fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re)}; }
__device__ __forceinline__ cd cmulc(cd a, cd b) {  // a * conj(b)
  return cd{fma(a.re, b.re, a.im * b.im), fma(a.im, b.re, -(a.re * b.im))};
}
__device__ __forceinline__ void cacc(cd& acc, cd a, cd b) {  // acc += a*b
  acc.re = fma(a.re, b.re, fma(-a.im, b.im, acc.re));
  acc.im = fma(a.re, b.im, fma(a.im, b.re, acc.im));
}
// m*a + n*b, optionally with m and n conjugated
template <bool CONJ>
__device__ __forceinline__ cd lin2(cd m, cd a, cd n, cd b) {
  cd o;
  if (!CONJ) {
    o.re = fma(m.re, a.re, fma(-m.im, a.im, fma(n.re, b.re, -(n.im * b.im))));
    o.im = fma(m.re, a.im, fma(m.im, a.re, fma(n.re, b.im, n.im * b.re)));
  } else {
    o.re = fma(m.re, a.re, fma(m.im, a.im, fma(n.re, b.re, n.im * b.im)));
    o.im = fma(m.re, a.im, fma(-m.im, a.re, fma(n.re, b.im, -(n.im * b.re))));
  }
  return o;
}

// ------------------------------------------------------------------------------------------------
// sincos with the coefficients in the constant bank.  The (cos, sin) pairs are 12 % of the instructions of an
// evaluation; the CUDA library routine re-materialises its 64-bit polynomial coefficients through the uniform
// datapath on every call (28 UMOVs).  Here they are constant-bank operands of the DFMAs.
// Cody-Waite reduction by pi/2 in three FMAs (valid for |x| <= 105615, the library's own fast-path limit; larger or
// non-finite arguments go to the library), fdlibm kernel polynomials on |r| <= pi/4.  Max error measured against
// libm over 4e6 arguments: 2.3e-16 absolute, 1e-15 relative (tests/test_library_abi.py::test_fast_sincos_matches_libm).
// ------------------------------------------------------------------------------------------------
#define SLAM_TRIG_TABLE                                                                                         \
  {0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,                    \
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,                      \
   2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,                       \
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,                       \
   -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11}
static __constant__ double kTrigDev[16] = SLAM_TRIG_TABLE;
static const double kTrigHost[16] = SLAM_TRIG_TABLE;
#ifdef __CUDA_ARCH__
#define SLAM_TRIGC(i) kTrigDev[i]
#else
#define SLAM_TRIGC(i) kTrigHost[i]
#endif

__host__ __device__ __forceinline__ void fast_sincos(double x, double* sp, double* cp) {
  if (!(fabs(x) <= 105615.0)) {
    sincos(x, sp, cp);
    return;
  }
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: the low mantissa word of t is rint(x * 2/pi)
  const double t = fma(x, SLAM_TRIGC(0), magic);
#ifdef __CUDA_ARCH__
  const int q = __double2loint(t);
#else
  long long tb;
  memcpy(&tb, &t, sizeof(tb));
  const int q = (int)(unsigned)(tb & 0xffffffffLL);
#endif
  const double qd = t - magic;
  double r = fma(qd, -SLAM_TRIGC(1), x);
  r = fma(qd, -SLAM_TRIGC(2), r);
  r = fma(qd, -SLAM_TRIGC(3), r);
  const double z = r * r;
  double ps = fma(z, SLAM_TRIGC(9), SLAM_TRIGC(8));
  double pc = fma(z, SLAM_TRIGC(15), SLAM_TRIGC(14));
  ps = fma(z, ps, SLAM_TRIGC(7));
  pc = fma(z, pc, SLAM_TRIGC(13));
  ps = fma(z, ps, SLAM_TRIGC(6));
  pc = fma(z, pc, SLAM_TRIGC(12));
  ps = fma(z, ps, SLAM_TRIGC(5));
  pc = fma(z, pc, SLAM_TRIGC(11));
  ps = fma(z, ps, SLAM_TRIGC(4));
  pc = fma(z, pc, SLAM_TRIGC(10));
  const double sr = fma(r * z, ps, r);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  // x = q pi/2 + r:  sin x = {sr, cr, -sr, -cr}[q mod 4],  cos x = {cr, -sr, -cr, sr}[q mod 4]
  double sv = (q & 1) ? cr : sr;
  double cv = (q & 1) ? sr : cr;
  if (q & 2) sv = -sv;
  if ((q + 1) & 2) cv = -cv;
  *sp = sv;
  *cp = cv;
}

enum { OP_N = 0, OP_T = 1, OP_H = 2 };  // apply M, M^T, M^dagger

// v <- (op M acting on qubit Q) v  for a 4-vector indexed (q1 q0); m = {m00, m01, m10, m11}
template <int Q, int OP>
__device__ __forceinline__ void apply1q(cd v[4], const cd m[4]) {
  const cd e01 = (OP == OP_N) ? m[1] : m[2];
  const cd e10 = (OP == OP_N) ? m[2] : m[1];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int lo = (Q == 0) ? 2 * p : p;
    const int hi = (Q == 0) ? 2 * p + 1 : p + 2;
    const cd a = v[lo], b = v[hi];
    v[lo] = lin2<OP == OP_H>(m[0], a, e01, b);
    v[hi] = lin2<OP == OP_H>(e10, a, m[3], b);
  }
}

// 1Q gate matrices from cached (cos, sin) pairs
__device__ __forceinline__ void build_u3(double2 t, double2 p, double2 l, cd m[4]) {
  // t = (cos, sin)(theta/2), p = (cos, sin)(phi), l = (cos, sin)(lam)
  m[0] = mkc(t.x, 0.0);
  m[1] = mkc(-(l.x * t.y), -(l.y * t.y));
  m[2] = mkc(p.x * t.y, p.y * t.y);
  const double er = fma(p.x, l.x, -(p.y * l.y)), ei = fma(p.y, l.x, p.x * l.y);
  m[3] = mkc(er * t.x, ei * t.x);
}
__device__ __forceinline__ void build_rz(double2 h, cd m[4]) {  // h = (cos, sin)(lam/2)
  m[0] = mkc(h.x, -h.y);
  m[1] = mkc(0.0, 0.0);
  m[2] = mkc(0.0, 0.0);
  m[3] = mkc(h.x, h.y);
}

// Factored U3: U(theta, phi, lam) = D(phi) Ry(theta) D(lam) with D(a) = diag(1, e^{ia}), Ry = [[c,-s],[s,c]], (c,s) =
// (cos,sin)(theta/2).  Applying the three factors costs the same 16 FP64 instructions per amplitude pair as a general
// complex 2x2 but needs no matrix build, and in the backward sweep the cuts between the factors give every parameter
// derivative as a short diagonal / antisymmetric contraction of the row vector w and the column vector r at the SAME cut:
//   T = w_x . r_x at every cut x;   dT/dphi = i sum_hi w r (cut after D(phi));   dT/dlam = i sum_hi w r and
//   dT/dtheta = 1/2 sum_pairs (w_hi r_lo - w_lo r_hi) (cut between D(lam) and Ry, since dRy/dtheta = Ry J / 2).
// "hi" = the two amplitudes whose bit Q is set; pairs (lo, hi) differ in bit Q.
template <int Q, bool CONJ>
__device__ __forceinline__ void phase1q(cd v[4], double2 e) {  // hi <- e^{+-ia} hi
  const double s = CONJ ? -e.y : e.y;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int hi = (Q == 0) ? 2 * p + 1 : p + 2;
    const cd b = v[hi];
    v[hi] = mkc(fma(e.x, b.re, -(s * b.im)), fma(e.x, b.im, s * b.re));
  }
}
template <int Q, bool TR>
__device__ __forceinline__ void rot1q(cd v[4], double2 t) {  // (lo,hi) <- Ry (lo,hi), or Ry^T when TR
  const double s = TR ? -t.y : t.y;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int lo = (Q == 0) ? 2 * p : p;
    const int hi = (Q == 0) ? 2 * p + 1 : p + 2;
    const cd a = v[lo], b = v[hi];
    v[lo] = mkc(fma(t.x, a.re, -(s * b.re)), fma(t.x, a.im, -(s * b.im)));
    v[hi] = mkc(fma(t.x, b.re, s * a.re), fma(t.x, b.im, s * a.im));
  }
}
// column vector: v <- U3 v (qubit Q); t[0..2] = (cos,sin) of theta/2, phi, lam
template <int Q>
__device__ __forceinline__ void u3_fwd(cd v[4], const double2* t) {
  phase1q<Q, false>(v, t[2]);
  rot1q<Q, false>(v, t[0]);
  phase1q<Q, false>(v, t[1]);
}
// sum over the hi amplitudes of Im(w r), and the antisymmetric pair sum Re(w_hi r_lo - w_lo r_hi)
template <int Q>
__device__ __forceinline__ double diag_im(const cd w[4], const cd r[4], double acc) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int hi = (Q == 0) ? 2 * p + 1 : p + 2;
    acc = fma(w[hi].re, r[hi].im, fma(w[hi].im, r[hi].re, acc));
  }
  return acc;
}
template <int Q>
__device__ __forceinline__ double anti_re(const cd w[4], const cd r[4], double acc) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int lo = (Q == 0) ? 2 * p : p;
    const int hi = (Q == 0) ? 2 * p + 1 : p + 2;
    acc = fma(w[hi].re, r[lo].re, fma(-w[hi].im, r[lo].im, acc));
    acc = fma(-w[lo].re, r[hi].re, fma(w[lo].im, r[hi].im, acc));
  }
  return acc;
}
// backward through the U3 of qubit Q: r <- U3^dagger r, w <- w U3, accumulating this lane's partial derivatives
// (d[0] theta, d[1] phi, d[2] lam; the 1/2 of theta and the signs are applied by the caller)
template <int Q>
__device__ __forceinline__ void u3_bwd(cd r[4], cd w[4], const double2* t, double d[3]) {
  d[1] = diag_im<Q>(w, r, d[1]);
  phase1q<Q, true>(r, t[1]);
  phase1q<Q, false>(w, t[1]);
  rot1q<Q, true>(r, t[0]);
  rot1q<Q, true>(w, t[0]);
  d[2] = diag_im<Q>(w, r, d[2]);
  d[0] = anti_re<Q>(w, r, d[0]);
  phase1q<Q, true>(r, t[2]);
  phase1q<Q, false>(w, t[2]);
}

// ------------------------------------------------------------------------------------------------
// 2Q gates
// ------------------------------------------------------------------------------------------------
struct BlockGate {  // G[0,0]=G[3,3]=co, G[0,3]=bo, G[3,0]=go, G[1,1]=G[2,2]=ci, G[1,2]=bi, G[2,1]=gi
  double co, ci;
  cd bo, go, bi, gi;
};

// from (cos,sin) of phi_c, phi_g, a_c = gc*t, a_g = gg*t  (closed form of exp(-i t (gc H_c + gg H_g)))
__device__ __forceinline__ BlockGate block_from_trig(double2 pc, double2 pg, double2 ac, double2 ag) {
  BlockGate g;
  g.co = ag.x;
  g.ci = ac.x;
  g.bi = mkc(-(pc.y * ac.y), -(pc.x * ac.y));  // -i e^{-i phi_c} sin a_c
  g.gi = mkc(pc.y * ac.y, -(pc.x * ac.y));     // -i e^{+i phi_c} sin a_c
  g.bo = mkc(-(pg.y * ag.y), -(pg.x * ag.y));
  g.go = mkc(pg.y * ag.y, -(pg.x * ag.y));
  return g;
}

template <int OP>
__device__ __forceinline__ void block_apply(cd v[4], const BlockGate& g) {
  const cd b_o = (OP == OP_N) ? g.bo : g.go;  // effective [0,3]
  const cd g_o = (OP == OP_N) ? g.go : g.bo;  // effective [3,0]
  const cd b_i = (OP == OP_N) ? g.bi : g.gi;
  const cd g_i = (OP == OP_N) ? g.gi : g.bi;
  const cd v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
  constexpr bool C = (OP == OP_H);
  v[0] = lin2<C>(mkc(g.co, 0.0), v0, b_o, v3);
  v[3] = lin2<C>(g_o, v0, mkc(g.co, 0.0), v3);
  v[1] = lin2<C>(mkc(g.ci, 0.0), v1, b_i, v2);
  v[2] = lin2<C>(g_i, v1, mkc(g.ci, 0.0), v2);
}

// symmetric gate with purely imaginary off-diagonals: G = G^T, G^dagger = conj(G)
template <int OP>
__device__ __forceinline__ void sym_apply(cd v[4], double co, double so, double ci, double si) {
  if (OP == OP_H) {
    so = -so;
    si = -si;
  }
  const cd v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
  v[0] = mkc(fma(co, v0.re, -(so * v3.im)), fma(co, v0.im, so * v3.re));
  v[3] = mkc(fma(co, v3.re, -(so * v0.im)), fma(co, v3.im, so * v0.re));
  v[1] = mkc(fma(ci, v1.re, -(si * v2.im)), fma(ci, v1.im, si * v2.re));
  v[2] = mkc(fma(ci, v2.re, -(si * v1.im)), fma(ci, v2.im, si * v1.re));
}

template <int OP>
__device__ __forceinline__ void dense_apply(cd v[4], const double* __restrict__ g) {  // g: 32 doubles row-major
  cd o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    cd acc = mkc(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = (OP == OP_N) ? (i * 4 + j) : (j * 4 + i);
      const cd m = mkc(g[2 * e], (OP == OP_H) ? -g[2 * e + 1] : g[2 * e + 1]);
      cacc(acc, m, v[j]);
    }
    o[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = o[i];
}

// ------------------------------------------------------------------------------------------------
// team reductions
// ------------------------------------------------------------------------------------------------
template <int LPP>
__device__ __forceinline__ double team_sum(double v) {
  if (LPP >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
  if (LPP >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
template <int LPP>
__device__ __forceinline__ double team_max(double v) {
  if (LPP >= 2) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  if (LPP >= 4) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// Sum six per-lane partials over the team and leave each total on exactly one lane (reduce-scatter):
// 5 double shuffles instead of the 12 of a full butterfly, and the stores are spread over the lanes.
//   LPP = 4: lane 0 owns {0,1}, lane 1 {2}, lane 2 {3,4}, lane 3 {5};  LPP = 2: lane 0 {0,1,2}, lane 1 {3,4,5}.
template <int LPP, class Store>
__device__ __forceinline__ void reduce_scatter6(const double d[6], int sub, Store store) {
  if (LPP == 1) {
#pragma unroll
    for (int s = 0; s < 6; ++s) store(s, d[s]);
  } else if (LPP == 2) {
    const bool hi = sub & 1;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const double keep = hi ? d[3 + t] : d[t], give = hi ? d[t] : d[3 + t];
      const double v = keep + __shfl_xor_sync(0xffffffffu, give, 1);
      store(hi ? 3 + t : t, v);
    }
  } else {
    const bool up = sub & 2, hi = sub & 1;
    double m[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const double keep = up ? d[3 + t] : d[t], give = up ? d[t] : d[3 + t];
      m[t] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    }
    const int b = up ? 3 : 0;
    const double v = __shfl_xor_sync(0xffffffffu, hi ? m[0] : m[2], 1);
    const double w = m[1] + __shfl_xor_sync(0xffffffffu, m[1], 1);
    if (hi) {
      store(b + 2, m[2] + v);
    } else {
      store(b + 0, m[0] + v);
      store(b + 1, w);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// trig cache: entry e < 6(k+1): layer e/6, slot e%6 (q0: theta/2, phi, lam; q1: theta/2, phi, lam)
//             entries 6(k+1)+4g+{0,1,2,3}: gate g: phi_c, phi_g, a_c, a_g   (parameter-bound block gates)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double slot_value(const KTemplate& kt, const double* xs, int g, int s) {
  const int p = kt.slot_param[g][s];
  return p >= 0 ? xs[p] : kt.slot_const[g][s];
}

// GM_SYM kernels are the hot instantiations: RZ layers (vz_only) and gate trig entries are compiled out of them
// (compile_template never selects GM_SYM for vz_only templates), which keeps the tick body of K5 smaller.
// CANON = canonical parameter layout: the caller keeps x (and the gradient) in circuit creation order, entry 6*layer+slot,
// every layer present, no parameter-bound gate -- no index tables are read (K5 permutes to/from the API order at the
// problem boundaries).
template <int LPP, int GM, bool CANON = false>
__device__ __forceinline__ void fill_trig(const KTemplate& kt, const double* xs, double2* tg, int sub) {
  if (CANON) {
    const int n1 = 6 * (kt.k + 1);
    for (int e = sub; e < n1; e += LPP) {
      double2 cs;
      const double a = (e % 3 == 0) ? 0.5 * xs[e] : xs[e];
      fast_sincos(a, &cs.y, &cs.x);
      tg[e] = cs;
    }
    return;
  }
  const bool vz = (GM != GM_SYM) && kt.vz_only;
  const int n1 = 6 * (kt.k + 1);
  for (int e = sub; e < kt.n_trig; e += LPP) {
    double2 cs = make_double2(1.0, 0.0);
    if (e < n1) {
      const int layer = e / 6, s = e - 6 * layer;
      const int p = kt.p1q[layer][s];
      if (p >= 0) {
        const bool half = vz ? true : (s == 0 || s == 3);
        const double a = half ? 0.5 * xs[p] : xs[p];
        fast_sincos(a, &cs.y, &cs.x);
      }
    } else if (GM == GM_BLOCK) {
      const int q = e - n1;
      const int g = q >> 2, w = q & 3;
      if (kt.gate_kind == SLAM_GATE_RISWAP) {
        // RiSwap(alpha) in block form: phi_c = pi (so that -i e^{-i phi_c} = +i), a_c = pi*alpha/2, a_g = 0
        if (w == 0) cs = make_double2(-1.0, 0.0);
        if (w == 2) sincos(1.5707963267948966 * slot_value(kt, xs, g, 0), &cs.y, &cs.x);
      } else {  // SLAM_GATE_CG slots: phi_c, phi_g, gc, gg, t
        double a;
        if (w == 0) a = slot_value(kt, xs, g, 0);
        else if (w == 1) a = slot_value(kt, xs, g, 1);
        else if (w == 2) a = slot_value(kt, xs, g, 2) * slot_value(kt, xs, g, 4);
        else a = slot_value(kt, xs, g, 3) * slot_value(kt, xs, g, 4);
        sincos(a, &cs.y, &cs.x);
      }
    }
    tg[e] = cs;
  }
}

// block gate g of this problem: from the trig cache if parameter-bound, else from the constant table
__device__ __forceinline__ BlockGate load_block(const KTemplate& kt, const double2* tg, int g) {
  if (kt.gate_bound[g]) {
    const double2* q = tg + 6 * (kt.k + 1) + 4 * g;
    return block_from_trig(q[0], q[1], q[2], q[3]);
  }
  const double* c = kt.gblk[g];
  return block_from_trig(make_double2(c[0], c[1]), make_double2(c[2], c[3]), make_double2(c[4], c[5]),
                         make_double2(c[6], c[7]));
}

template <int GM, int OP>
__device__ __forceinline__ void gate_apply(const KTemplate& kt, const BlockGate& bg, int g, cd v[4]) {
  // (CANON callers pass g = 0: one constant gate for every repetition, read as immediate constant-bank operands)
  if (GM == GM_SYM) sym_apply<OP>(v, kt.gsym[g][0], kt.gsym[g][1], kt.gsym[g][2], kt.gsym[g][3]);
  else if (GM == GM_BLOCK) block_apply<OP>(v, bg);
  else dense_apply<OP>(v, kt.dense[g]);
}

template <int GM>
__device__ __forceinline__ bool layer_present(const KTemplate& kt, int i) {  // false: dropped by no_exterior_1q
  return !(kt.p1q[i][0] < 0 && kt.p1q[i][3] < 0);
}
template <int GM>
__device__ __forceinline__ bool layer_is_rz(const KTemplate& kt) { return (GM != GM_SYM) && kt.vz_only; }

// ------------------------------------------------------------------------------------------------
// forward sweep: r[c] = column (sub*CPL + c) of U(x)
// ------------------------------------------------------------------------------------------------
template <int LPP, int GM, bool CANON = false>
__device__ __forceinline__ void forward_chain(const KTemplate& kt, const double2* tg, int sub, cd r[4 / LPP][4]) {
  constexpr int CPL = 4 / LPP;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) r[c][a] = mkc((a == sub * CPL + c) ? 1.0 : 0.0, 0.0);
  for (int i = 0; i <= kt.k; ++i) {
    if (CANON || layer_present<GM>(kt, i)) {
      const double2* t = tg + 6 * i;
      if (!CANON && layer_is_rz<GM>(kt)) {
        cd A[4], B[4];
        build_rz(t[0], B);
        build_rz(t[3], A);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          apply1q<0, OP_N>(r[c], B);
          apply1q<1, OP_N>(r[c], A);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          u3_fwd<0>(r[c], t);
          u3_fwd<1>(r[c], t + 3);
        }
      }
    }
    if (i < kt.k) {
      BlockGate bg;
      if (GM == GM_BLOCK) bg = load_block(kt, tg, i);
#pragma unroll
      for (int c = 0; c < CPL; ++c) gate_apply<GM, OP_N>(kt, bg, CANON ? 0 : i, r[c]);
    }
  }
}

// cost functional from |T|: value and d(loss)/d|T|   (src/slam/cost_function.py:133-173)
__device__ __forceinline__ void cost_from_abs(int cost_kind, double a, double& loss, double& dl_da) {
  if (cost_kind == SLAM_COST_SQUARE) {
    loss = 1.0 - (a * a + 4.0) / 20.0;
    dl_da = -a / 10.0;
  } else if (cost_kind == SLAM_COST_BASIC_INVERSE) {
    loss = a * 0.25;
    dl_da = 0.25;
  } else {
    loss = 1.0 - a * 0.25;
    dl_da = -0.25;
  }
}

// ------------------------------------------------------------------------------------------------
// loss + analytic gradient for one problem, executed by a team.
//   xs : this problem's parameters (API order)            [shared or global]
//   tg : this problem's trig cache, n_trig entries        [shared]  (filled here)
//   gs : this problem's gradient (API order), written if WANT_GRAD.  Every parameter is bound to exactly one
//        slot (checked on the host), so each entry is stored once; unused entries keep the zero fill.
//   vcol[c][a] = V[a][sub*CPL + c]  (this lane's target columns)
// returns loss (identical on every lane of the team); *T_out = Tr(V^dag U)
// ------------------------------------------------------------------------------------------------
template <int LPP, int GM, bool WANT_GRAD, bool CANON = false>
__device__ __forceinline__ double loss_grad_team(const KTemplate& kt, const double* xs, double2* tg, double* gs,
                                                 const cd vcol[4 / LPP][4], int cost_kind, int sub, cd* T_out) {
  constexpr int CPL = 4 / LPP;
  fill_trig<LPP, GM, CANON>(kt, xs, tg, sub);
  if (WANT_GRAD && !CANON)  // (canonical layout: every entry is stored below)
    for (int j = sub; j < kt.P; j += LPP) gs[j] = 0.0;
  __syncwarp();

  cd r[CPL][4];
  forward_chain<LPP, GM, CANON>(kt, tg, sub, r);

  // T = Tr(V^dag U) = sum_col sum_a conj(V[a][col]) U[a][col]
  cd Tp = mkc(0.0, 0.0);
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      Tp.re = fma(vcol[c][a].re, r[c][a].re, fma(vcol[c][a].im, r[c][a].im, Tp.re));
      Tp.im = fma(vcol[c][a].re, r[c][a].im, fma(-vcol[c][a].im, r[c][a].re, Tp.im));
    }
  cd T = mkc(team_sum<LPP>(Tp.re), team_sum<LPP>(Tp.im));
  const double absT = sqrt(fma(T.re, T.re, T.im * T.im));
  double loss, dl_da;
  cost_from_abs(cost_kind, absT, loss, dl_da);
  if (T_out) *T_out = T;
  if (!WANT_GRAD) return loss;

  // w[c] = row (sub*CPL+c) of (dl/d|T|) * conj(T)/|T| * V^dagger ; d loss = Re( sum w dM r )
  const double inv = absT > 0.0 ? dl_da / absT : 0.0;
  const cd ph = mkc(T.re * inv, -T.im * inv);
  cd w[CPL][4];
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) w[c][a] = cmulc(ph, vcol[c][a]);

  for (int i = kt.k; i >= 0; --i) {
    if (CANON || layer_present<GM>(kt, i)) {
      const double2* t = tg + 6 * i;
      if (!CANON && layer_is_rz<GM>(kt)) {
        // RZ = diag(e^{-i l/2}, e^{+i l/2}); d/dl = (i/2) diag(-e^{-i l/2}, e^{+i l/2}).  With r, w at the cut after
        // the layer: dT/dl = (i/2) (sum_hi w r - sum_lo w r)  ->  d loss = -(1/2) (Im sum_hi - Im sum_lo)
        cd A[4], B[4];
        build_rz(t[0], B);
        build_rz(t[3], A);
        double d0 = 0.0, d3 = 0.0;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          double all = 0.0;
#pragma unroll
          for (int a = 0; a < 4; ++a) all = fma(w[c][a].re, r[c][a].im, fma(w[c][a].im, r[c][a].re, all));
          d0 += all - 2.0 * diag_im<0>(w[c], r[c], 0.0);  // Im sum_lo - Im sum_hi (qubit 0)
          d3 += all - 2.0 * diag_im<1>(w[c], r[c], 0.0);
          apply1q<1, OP_H>(r[c], A);
          apply1q<0, OP_H>(r[c], B);
          apply1q<1, OP_T>(w[c], A);
          apply1q<0, OP_T>(w[c], B);
        }
        d0 = team_sum<LPP>(0.5 * d0);
        d3 = team_sum<LPP>(0.5 * d3);
        if (sub == 0) {
          if (kt.p1q[i][0] >= 0) gs[kt.p1q[i][0]] = d0;
          if (kt.p1q[i][3] >= 0) gs[kt.p1q[i][3]] = d3;
        }
      } else {
        double dq[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          u3_bwd<1>(r[c], w[c], t + 3, dq[1]);
          u3_bwd<0>(r[c], w[c], t, dq[0]);
        }
        double d[6];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          d[3 * q + 0] = 0.5 * dq[q][0];
          d[3 * q + 1] = -dq[q][1];
          d[3 * q + 2] = -dq[q][2];
        }
        reduce_scatter6<LPP>(d, sub, [&](int s, double v) {
          if (CANON) {
            gs[6 * i + s] = v;
          } else {
            const int p = kt.p1q[i][s];
            if (p >= 0) gs[p] = v;
          }
        });
      }
    }
    if (i > 0) {
      const int g = i - 1;
      BlockGate bg;
      if (GM == GM_BLOCK) bg = load_block(kt, tg, g);
#pragma unroll
      for (int c = 0; c < CPL; ++c) gate_apply<GM, OP_H>(kt, bg, CANON ? 0 : g, r[c]);  // r <- G^dagger r
      if (GM == GM_BLOCK && kt.gate_bound[g]) {
        // environment of the gate: E[i][j] = sum_c w[c][i] r[c][j] on the block pattern
        cd Ed_o = mkc(0, 0), Ed_i = mkc(0, 0), E03 = mkc(0, 0), E30 = mkc(0, 0), E12 = mkc(0, 0), E21 = mkc(0, 0);
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          cacc(Ed_o, w[c][0], r[c][0]);
          cacc(Ed_o, w[c][3], r[c][3]);
          cacc(Ed_i, w[c][1], r[c][1]);
          cacc(Ed_i, w[c][2], r[c][2]);
          cacc(E03, w[c][0], r[c][3]);
          cacc(E30, w[c][3], r[c][0]);
          cacc(E12, w[c][1], r[c][2]);
          cacc(E21, w[c][2], r[c][1]);
        }
        const double2* q = tg + 6 * (kt.k + 1) + 4 * g;
        const double2 pc = q[0], pg = q[1], ac = q[2], ag = q[3];
        // d/da_c: dci = -sin a_c ; dbi = -i e^{-i pc} cos a_c ; dgi = -i e^{+i pc} cos a_c
        const cd dbi_a = mkc(-(pc.y * ac.x), -(pc.x * ac.x)), dgi_a = mkc(pc.y * ac.x, -(pc.x * ac.x));
        const cd dbo_a = mkc(-(pg.y * ag.x), -(pg.x * ag.x)), dgo_a = mkc(pg.y * ag.x, -(pg.x * ag.x));
        double d_ac = -ac.y * Ed_i.re + cmul(dbi_a, E12).re + cmul(dgi_a, E21).re;
        double d_ag = -ag.y * Ed_o.re + cmul(dbo_a, E03).re + cmul(dgo_a, E30).re;
        // d/dphi_c: dbi = -i*bi, dgi = +i*gi  -> Re(-i bi E12 + i gi E21) = Im(bi E12) - Im(gi E21)
        double d_pc = cmul(bg.bi, E12).im - cmul(bg.gi, E21).im;
        double d_pg = cmul(bg.bo, E03).im - cmul(bg.go, E30).im;
        d_ac = team_sum<LPP>(d_ac);
        d_ag = team_sum<LPP>(d_ag);
        d_pc = team_sum<LPP>(d_pc);
        d_pg = team_sum<LPP>(d_pg);
        if (sub == 0) {
          if (kt.gate_kind == SLAM_GATE_RISWAP) {
            const int p = kt.slot_param[g][0];
            if (p >= 0) gs[p] = 1.5707963267948966 * d_ac;
          } else {
            const double gc = slot_value(kt, xs, g, 2), gg = slot_value(kt, xs, g, 3), tt = slot_value(kt, xs, g, 4);
            int p;
            if ((p = kt.slot_param[g][0]) >= 0) gs[p] = d_pc;
            if ((p = kt.slot_param[g][1]) >= 0) gs[p] = d_pg;
            if ((p = kt.slot_param[g][2]) >= 0) gs[p] = tt * d_ac;
            if ((p = kt.slot_param[g][3]) >= 0) gs[p] = tt * d_ag;
            if ((p = kt.slot_param[g][4]) >= 0) gs[p] = gc * d_ac + gg * d_ag;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) gate_apply<GM, OP_T>(kt, bg, CANON ? 0 : g, w[c]);  // w <- w G
    }
  }
  __syncwarp();
  return loss;
}

}  // namespace slam
