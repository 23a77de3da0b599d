// slam_fwd1.cuh -- thread-per-problem forward evaluation of a template (all gate modes, incl. K4 smush).
//
// Used by the forward-only kernels: smush template evaluation (K1 for GM_SMUSH), the fused coverage
// Monte-Carlo (K6) and the parallel-drive trajectory (K4b).  One thread owns the whole 4x4 running product.
//
// K4: a smush gate is the ordered product over T time slices of exp(-i dt H_s) with
//   H_s = gx_s (e^{i pa} A + h.c.) + gy_s (e^{i pb} B + h.c.) + gc (e^{i pc} A B^dag + h.c.) + gg (e^{i pg} A B + h.c.)
//         + gz1 A^dag A + gz2 B^dag B                                   (src/slam/hamiltonian.py:114-182)
// with qutip's create(2) = |1><0|, A = a (x) I, B = I (x) a.  H_s is a general 4x4 Hermitian matrix (no closed
// form once the phases are non-zero), so each slice is exponentiated by scaling-and-squaring of the degree-16/17
// cos/sin series of the Hermitian generator (||dt H / 2^s|| <= 0.7 -> truncation < 1e-17), evaluated with 9 triangular
// products of commuting Hermitian matrices; this replaces qutip's Qobj.expm -> scipy Pade.
#pragma once
#include "slam_core.cuh"

namespace slam {

// ---- parameter sources -----------------------------------------------------------------------------
struct GlobalParams {  // row of a [B, ldx] array
  const double* row;
  __device__ __forceinline__ double get(int j) const { return row[j]; }
};

// ---- Hermitian slice generator -----------------------------------------------------------------------
struct Herm4 {  // h13 = h02, h23 = h01 ; lower triangle = conjugates
  double d0, d1, d2, d3;
  cd h01, h02, h03, h12;
};

__device__ __forceinline__ void herm_mul(const Herm4& H, const cd v[4], cd o[4]) {
  o[0] = mkc(H.d0 * v[0].re, H.d0 * v[0].im);
  cacc(o[0], H.h01, v[1]);
  cacc(o[0], H.h02, v[2]);
  cacc(o[0], H.h03, v[3]);
  o[1] = cmulc(v[0], H.h01);  // conj(h01) v0
  o[1].re = fma(H.d1, v[1].re, o[1].re);
  o[1].im = fma(H.d1, v[1].im, o[1].im);
  cacc(o[1], H.h12, v[2]);
  cacc(o[1], H.h02, v[3]);
  o[2] = cmulc(v[0], H.h02);
  const cd t = cmulc(v[1], H.h12);
  o[2].re += t.re;
  o[2].im += t.im;
  o[2].re = fma(H.d2, v[2].re, o[2].re);
  o[2].im = fma(H.d2, v[2].im, o[2].im);
  cacc(o[2], H.h01, v[3]);
  o[3] = cmulc(v[0], H.h03);
  const cd t1 = cmulc(v[1], H.h02), t2 = cmulc(v[2], H.h01);
  o[3].re += t1.re + t2.re;
  o[3].im += t1.im + t2.im;
  o[3].re = fma(H.d3, v[3].re, o[3].re);
  o[3].im = fma(H.d3, v[3].im, o[3].im);
}

// ---- general Hermitian 4x4 (upper triangle) and products of COMMUTING Hermitian matrices -----------------------
struct HermG {
  double d[4];
  cd u[6];  // (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
};
__device__ __forceinline__ constexpr int hg_idx(int i, int j) { return i == 0 ? j - 1 : (i == 1 ? j + 1 : 5); }  // i < j
__device__ __forceinline__ cd hg_get(const HermG& A, int i, int j) {  // i != j; resolved at compile time after unrolling
  if (i < j) return A.u[hg_idx(i, j)];
  const cd v = A.u[hg_idx(j, i)];
  return mkc(v.re, -v.im);
}
// C = A B for commuting Hermitian A, B (then C is Hermitian): only the upper triangle is formed, ~100 FP64
// instructions instead of the 256 of a general complex 4x4 product
__device__ __forceinline__ void hg_mul(const HermG& A, const HermG& B, HermG& C) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc = A.d[i] * B.d[i];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k != i) {
        const cd a = hg_get(A, i, k), b = hg_get(B, i, k);  // A_ik conj(B_ik): real part
        acc = fma(a.re, b.re, fma(a.im, b.im, acc));
      }
    C.d[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      const cd aij = A.u[hg_idx(i, j)], bij = B.u[hg_idx(i, j)];
      cd acc = mkc(A.d[i] * bij.re, A.d[i] * bij.im);
      acc.re = fma(aij.re, B.d[j], acc.re);
      acc.im = fma(aij.im, B.d[j], acc.im);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k != i && k != j) cacc(acc, hg_get(A, i, k), hg_get(B, k, j));
      C.u[hg_idx(i, j)] = acc;
    }
}
// Taylor coefficients of cos (a_j = (-1)^j / (2j)!) and sin x / x (b_j = (-1)^j / (2j+1)!), highest power first; the
// leading sin entry is a zero so that both series run through the same loop
static __constant__ double kCosSinCoef[9][2] = {
    {1.0 / 20922789888000.0, 0.0},
    {-1.0 / 87178291200.0, -1.0 / 1307674368000.0},
    {1.0 / 479001600.0, 1.0 / 6227020800.0},
    {-1.0 / 3628800.0, -1.0 / 39916800.0},
    {1.0 / 40320.0, 1.0 / 362880.0},
    {-1.0 / 720.0, -1.0 / 5040.0},
    {1.0 / 24.0, 1.0 / 120.0},
    {-0.5, -1.0 / 6.0},
    {1.0, 1.0}};

// Y[col][row] = exp(-i dt H).  rho = upper bound of ||dt H||_2.
//
// exp(-i A) = cos A - i sin A with A = theta H Hermitian, theta = dt / 2^s chosen so that ||A|| <= 0.7:
//   cos A = sum_{j<=8} a_j K^j,   sin A = A sum_{j<=7} b_j K^j,   K = A^2     (truncation < 1e-17 at ||A|| = 0.7)
// Both series are Horner recurrences in K^2 with linear blocks, X <- (c_2i I + c_2i+1 K) + K^2 X (Paterson-Stockmeyer with
// block size 2), in ONE runtime loop of three steps; every product is between commuting Hermitian matrices, so only upper
// triangles are formed (~100 FP64 instructions per product instead of 256 dense): 9 triangular products per slice (A A,
// K K, 3 x 2 in the loop, A P) where the plain Horner form in K needed 18.  The loop body stays ~4 KB on purpose: the
// forward-only kernels are instruction-fetch sensitive (a straight-line Paterson-Stockmeyer schedule with block size 3 was
// measured slower than the compact Horner loop last round), and this form keeps the compact loop.
__device__ __forceinline__ void herm_expm(const Herm4& H, double dt, double rho, cd Y[4][4]) {
  int s = 0;
  if (rho > 0.7) s = min(ilogb(rho * (1.0 / 0.7)) + 1, 40);
  const double theta = ldexp(dt, -s);
  HermG A;
  A.d[0] = theta * H.d0;
  A.d[1] = theta * H.d1;
  A.d[2] = theta * H.d2;
  A.d[3] = theta * H.d3;
  A.u[0] = mkc(theta * H.h01.re, theta * H.h01.im);
  A.u[1] = mkc(theta * H.h02.re, theta * H.h02.im);
  A.u[2] = mkc(theta * H.h03.re, theta * H.h03.im);
  A.u[3] = mkc(theta * H.h12.re, theta * H.h12.im);
  A.u[4] = A.u[1];  // h13 = h02, h23 = h01 (structure of the smush generator)
  A.u[5] = A.u[0];
  // Horner in K2 = K^2 with linear blocks (Paterson-Stockmeyer, block size 2): a_j = kCosSinCoef[8 - j][0], b_j = ...[1]
  //   cos:  C <- (a_2i I + a_2i+1 K) + K2 C   from  C = a_6 I + a_7 K + a_8 K2,   i = 2, 1, 0
  //   sin:  P <- (b_2i I + b_2i+1 K) + K2 P   from  P = b_6 I + b_7 K
  HermG K, K2, C, P;
  hg_mul(A, A, K);
  hg_mul(K, K, K2);
  {
    const double a6 = kCosSinCoef[2][0], a7 = kCosSinCoef[1][0], a8 = kCosSinCoef[0][0];
    const double b6 = kCosSinCoef[2][1], b7 = kCosSinCoef[1][1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      C.d[i] = fma(a8, K2.d[i], fma(a7, K.d[i], a6));
      P.d[i] = fma(b7, K.d[i], b6);
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      C.u[e] = mkc(fma(a8, K2.u[e].re, a7 * K.u[e].re), fma(a8, K2.u[e].im, a7 * K.u[e].im));
      P.u[e] = mkc(b7 * K.u[e].re, b7 * K.u[e].im);
    }
  }
#pragma unroll 1
  for (int i = 2; i >= 0; --i) {
    HermG Cn, Pn;
    hg_mul(K2, C, Cn);
    hg_mul(K2, P, Pn);
    const double a0 = kCosSinCoef[8 - 2 * i][0], a1 = kCosSinCoef[7 - 2 * i][0];
    const double b0 = kCosSinCoef[8 - 2 * i][1], b1 = kCosSinCoef[7 - 2 * i][1];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      C.d[r] = Cn.d[r] + fma(a1, K.d[r], a0);
      P.d[r] = Pn.d[r] + fma(b1, K.d[r], b0);
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      C.u[e] = mkc(fma(a1, K.u[e].re, Cn.u[e].re), fma(a1, K.u[e].im, Cn.u[e].im));
      P.u[e] = mkc(fma(b1, K.u[e].re, Pn.u[e].re), fma(b1, K.u[e].im, Pn.u[e].im));
    }
  }
  HermG S;
  hg_mul(A, P, S);
  // Y = C - i S, Y[col][row]
#pragma unroll
  for (int r = 0; r < 4; ++r) Y[r][r] = mkc(C.d[r], -S.d[r]);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i + 1; j < 4; ++j) {
      const cd c = C.u[hg_idx(i, j)], sv = S.u[hg_idx(i, j)];
      Y[j][i] = mkc(c.re + sv.im, c.im - sv.re);    // element (i, j):  c - i s
      Y[i][j] = mkc(c.re - sv.im, -c.im - sv.re);   // element (j, i):  conj(c) - i conj(s)
    }
#pragma unroll 1
  for (int q = 0; q < s; ++q) {
    cd Z[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        cd acc = cmul(Y[0][r], Y[c][0]);
        cacc(acc, Y[1][r], Y[c][1]);
        cacc(acc, Y[2][r], Y[c][2]);
        cacc(acc, Y[3][r], Y[c][3]);
        Z[c][r] = acc;
      }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < 4; ++r) Y[c][r] = Z[c][r];
  }
}

// phases of one smush gate: e^{-i phi} for (a, b, c, g) and the static couplings
struct SmushGate {
  cd ea, eb, ec, eg;  // e^{-i phi_a}, e^{-i phi_b}, e^{-i phi_c}, e^{-i phi_g}
  double gc, gg, gz1, gz2;
};

__device__ __forceinline__ SmushGate smush_gate(double pa, double pb, double pc, double pg, double gc, double gg,
                                                double gz1, double gz2) {
  SmushGate G;
  double s, c;
  fast_sincos(pa, &s, &c); G.ea = mkc(c, -s);
  fast_sincos(pb, &s, &c); G.eb = mkc(c, -s);
  fast_sincos(pc, &s, &c); G.ec = mkc(c, -s);
  fast_sincos(pg, &s, &c); G.eg = mkc(c, -s);
  G.gc = gc; G.gg = gg; G.gz1 = gz1; G.gz2 = gz2;
  return G;
}

__device__ __forceinline__ void smush_slice(const SmushGate& G, double gx, double gy, double dt, cd Y[4][4]) {
  Herm4 H;
  H.d0 = G.gz1 + G.gz2;  // A^dag A = |0><0| (x) I, B^dag B = I (x) |0><0|  (a = create(2) = |1><0|)
  H.d1 = G.gz1;
  H.d2 = G.gz2;
  H.d3 = 0.0;
  H.h01 = mkc(gy * G.eb.re, gy * G.eb.im);
  H.h02 = mkc(gx * G.ea.re, gx * G.ea.im);
  H.h03 = mkc(G.gg * G.eg.re, G.gg * G.eg.im);
  H.h12 = mkc(G.gc * G.ec.re, G.gc * G.ec.im);
  const double rho = fabs(dt) * (fabs(gx) + fabs(gy) + fmax(fabs(G.gc), fabs(G.gg)) + fabs(G.gz1) + fabs(G.gz2));
  herm_expm(H, dt, rho, Y);
}

// R <- Y R  (both [col][row])
__device__ __forceinline__ void left_mul(const cd Y[4][4], cd R[4][4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cd n[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      cd acc = cmul(Y[0][r], R[c][0]);
      cacc(acc, Y[1][r], R[c][1]);
      cacc(acc, Y[2][r], R[c][2]);
      cacc(acc, Y[3][r], R[c][3]);
      n[r] = acc;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) R[c][r] = n[r];
  }
}

// ---- full forward chain, one thread, R[col][row] ------------------------------------------------------
template <class PS>
__device__ __forceinline__ double slot_val(const KTemplate& kt, const PS& ps, int g, int s) {
  const int p = kt.slot_param[g][s];
  return p >= 0 ? ps.get(p) : kt.slot_const[g][s];
}

template <class PS>
__device__ __forceinline__ void fwd1_layer(const KTemplate& kt, const PS& ps, int i, cd R[4][4]) {
  if (kt.p1q[i][0] < 0 && kt.p1q[i][3] < 0) return;
  cd A[4], B[4];
  double s, c;
  if (kt.vz_only) {
    sincos(0.5 * ps.get(kt.p1q[i][0]), &s, &c);
    build_rz(make_double2(c, s), B);
    sincos(0.5 * ps.get(kt.p1q[i][3]), &s, &c);
    build_rz(make_double2(c, s), A);
  } else {
    double2 t[6];
#pragma unroll 1  // one copy of the (cos, sin) routine: the forward-only kernels are instruction-fetch bound
    for (int q = 0; q < 6; ++q) {
      const double v = ps.get(kt.p1q[i][q]);
      fast_sincos((q == 0 || q == 3) ? 0.5 * v : v, &s, &c);
      t[q] = make_double2(c, s);
    }
    build_u3(t[0], t[1], t[2], B);
    build_u3(t[3], t[4], t[5], A);
  }
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    apply1q<0, OP_N>(R[col], B);
    apply1q<1, OP_N>(R[col], A);
  }
}

// GMT >= 0 fixes the gate mode at compile time (the other modes' code is not generated); -1 = runtime dispatch on kt.gmode
// SYNC: CTA-wide barriers at the phase boundaries (every layer / gate / slice).  All threads of a forward-only kernel run
// the same control flow, so the barriers cost nothing but keep the CTA's warps inside the SAME few KB of code at any time;
// without them the warps spread over the whole kernel (50-80 KB) and thrash the 32 KB L1.5 instruction cache.
template <class PS, int GMT = -1, bool SYNC = false>
__device__ __forceinline__ void fwd1_gate(const KTemplate& kt, const PS& ps, int g, cd R[4][4]) {
  const int gmode = (GMT >= 0) ? GMT : kt.gmode;
  if (gmode == GM_SYM) {
#pragma unroll
    for (int col = 0; col < 4; ++col) sym_apply<OP_N>(R[col], kt.gsym[g][0], kt.gsym[g][1], kt.gsym[g][2], kt.gsym[g][3]);
  } else if (gmode == GM_BLOCK) {
    BlockGate bg;
    if (kt.gate_bound[g]) {
      double2 q[4] = {make_double2(1, 0), make_double2(1, 0), make_double2(1, 0), make_double2(1, 0)};
      if (kt.gate_kind == SLAM_GATE_RISWAP) {
        q[0] = make_double2(-1.0, 0.0);
        sincos(1.5707963267948966 * slot_val(kt, ps, g, 0), &q[2].y, &q[2].x);
      } else {
        const double tt = slot_val(kt, ps, g, 4);
        sincos(slot_val(kt, ps, g, 0), &q[0].y, &q[0].x);
        sincos(slot_val(kt, ps, g, 1), &q[1].y, &q[1].x);
        sincos(slot_val(kt, ps, g, 2) * tt, &q[2].y, &q[2].x);
        sincos(slot_val(kt, ps, g, 3) * tt, &q[3].y, &q[3].x);
      }
      bg = block_from_trig(q[0], q[1], q[2], q[3]);
    } else {
      const double* c = kt.gblk[g];
      bg = block_from_trig(make_double2(c[0], c[1]), make_double2(c[2], c[3]), make_double2(c[4], c[5]),
                           make_double2(c[6], c[7]));
    }
#pragma unroll
    for (int col = 0; col < 4; ++col) block_apply<OP_N>(R[col], bg);
  } else if (gmode == GM_DENSE) {
#pragma unroll
    for (int col = 0; col < 4; ++col) dense_apply<OP_N>(R[col], kt.dense[g]);
  } else {  // GM_SMUSH
    const int T = kt.T;
    const bool ph1q = kt.gate_kind == SLAM_GATE_SMUSH_1QPHASE;
    const int o = ph1q ? 8 : 4;  // first gx slot
    SmushGate G;
    if (ph1q)
      G = smush_gate(slot_val(kt, ps, g, 0), slot_val(kt, ps, g, 1), slot_val(kt, ps, g, 2), slot_val(kt, ps, g, 3),
                     slot_val(kt, ps, g, 4), slot_val(kt, ps, g, 5), slot_val(kt, ps, g, 6), slot_val(kt, ps, g, 7));
    else
      G = smush_gate(0.0, 0.0, slot_val(kt, ps, g, 0), slot_val(kt, ps, g, 1), slot_val(kt, ps, g, 2),
                     slot_val(kt, ps, g, 3), 0.0, 0.0);
    const double dt = slot_val(kt, ps, g, o + 2 * T) / (double)T;  // timestep = t / N (hamiltonian.py:136-137)
    for (int it = 0; it < T; ++it) {
      if (SYNC) __syncthreads();
      cd Y[4][4];
      smush_slice(G, slot_val(kt, ps, g, o + it), slot_val(kt, ps, g, o + T + it), dt, Y);
      left_mul(Y, R);  // later slices multiply on the left (hamiltonian.py:143)
    }
  }
}

template <class PS, int GMT = -1, bool SYNC = false>
__device__ __forceinline__ void fwd1_chain(const KTemplate& kt, const PS& ps, cd R[4][4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) R[c][r] = mkc(c == r ? 1.0 : 0.0, 0.0);
  for (int i = 0; i <= kt.k; ++i) {
    if (SYNC) __syncthreads();
    fwd1_layer(kt, ps, i, R);
    if (SYNC) __syncthreads();
    if (i < kt.k) fwd1_gate<PS, GMT, SYNC>(kt, ps, i, R);
  }
}

}  // namespace slam
