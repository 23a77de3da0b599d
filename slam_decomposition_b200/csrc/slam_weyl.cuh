// slam_weyl.cuh -- K3 device function: Weyl-chamber coordinates and Makhlin invariants of one 4x4 unitary.
//
// Algorithm (restating weylchamber.c1c2c3 / g1g2g3, which the reference calls at basis_abc.py:80-84,
// optimizer.py:85,103,224, cost_function.py:199-221, parallel_drive_volume.py:225, pd_playground.py:199):
//   U_B = Q^dagger U Q (magic basis),  m = U_B^T U_B  (complex symmetric, unitary)
//   eigenvalues of m  ==  eigenvalues of U (sy sy) U^T (sy sy), which weylchamber feeds to LAPACK zgeev.
//   Here: Re(m), Im(m) are commuting real symmetric matrices, so a cyclic Jacobi with JOINT rotations
//   diagonalises both; it is backward stable at the degenerate spectra of CNOT / iSWAP / SWAP / identity,
//   where closed-form quartic roots would lose half the digits.
//   phases / sqrt(det U) -> sort -> integer shift -> (c1,c2,c3), mirror when c3 < 0  (Childs et al. PRA 68 052311)
//   g1 + i g2 = tr(m)^2 / (16 det U),  g3 = Re((tr(m)^2 - tr(m^2)) / (4 det U))     (Makhlin)
#pragma once
#include "slam_core.cuh"

namespace slam {

__host__ __device__ constexpr int sym_idx(int i, int j) {  // packed upper-triangular index of a 4x4 symmetric matrix
  return (i <= j) ? (i * 4 - i * (i - 1) / 2 + (j - i)) : (j * 4 - j * (j - 1) / 2 + (i - j));
}

template <int P, int Q>
__device__ __forceinline__ void joint_rotation(double a[10], double b[10]) {
  constexpr int pp = sym_idx(P, P), qq = sym_idx(Q, Q), pq = sym_idx(P, Q);
  const double x1 = a[pp] - a[qq], y1 = 2.0 * a[pq];
  const double x2 = b[pp] - b[qq], y2 = 2.0 * b[pq];
  const double ton = fma(x1, x1, x2 * x2) - fma(y1, y1, y2 * y2);
  const double toff = 2.0 * fma(x1, y1, x2 * y2);
  const double r2 = fma(ton, ton, toff * toff);
  if (!(r2 > 1e-300)) return;  // the pair is already jointly diagonal (or numerically null)
  // (ton, toff) has polar angle phi = 2*angle(x, y); the joint rotation angle is theta = phi/4
  // (Cardoso-Souloumiac closed form), evaluated with half-angle formulas so that the exactly
  // degenerate case ton < 0, toff = 0 (equal diagonals, non-zero off-diagonal) rotates by pi/4.
  const double inv_r = rsqrt(r2);
  const double big2 = 0.5 * fma(fabs(ton), inv_r, 1.0);  // in [1/2, 1]
  const double inv_big = rsqrt(big2);
  const double big = big2 * inv_big;
  const double small = 0.5 * fabs(toff) * inv_r * inv_big;
  const double C = ton >= 0.0 ? big : small;              // cos(phi/2) >= 0
  const double S = copysign(ton >= 0.0 ? small : big, toff);  // sin(phi/2)
  const double c2 = 0.5 * (1.0 + C);
  const double inv_c = rsqrt(c2);
  const double c = c2 * inv_c, s = 0.5 * S * inv_c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k == P || k == Q) continue;
    const int kp = sym_idx(k, P), kq = sym_idx(k, Q);
    const double ap = a[kp], aq = a[kq], bp = b[kp], bq = b[kq];
    a[kp] = fma(c, ap, s * aq);
    a[kq] = fma(c, aq, -(s * ap));
    b[kp] = fma(c, bp, s * bq);
    b[kq] = fma(c, bq, -(s * bp));
  }
  const double cc = c * c, ss = s * s, cs = c * s;
  {
    const double app = a[pp], aqq = a[qq], apq = a[pq];
    a[pp] = fma(cc, app, fma(2.0 * cs, apq, ss * aqq));
    a[qq] = fma(ss, app, fma(-2.0 * cs, apq, cc * aqq));
    a[pq] = fma(cc - ss, apq, cs * (aqq - app));
  }
  {
    const double bpp = b[pp], bqq = b[qq], bpq = b[pq];
    b[pp] = fma(cc, bpp, fma(2.0 * cs, bpq, ss * bqq));
    b[qq] = fma(ss, bpp, fma(-2.0 * cs, bpq, cc * bqq));
    b[pq] = fma(cc - ss, bpq, cs * (bqq - bpp));
  }
}

__device__ __forceinline__ cd det4(const cd M[4][4]) {  // M[row][col]; Laplace expansion over rows (0,1)x(2,3)
  auto minor = [&](int r0, int r1, int i, int j) {
    cd t = cmul(M[r0][i], M[r1][j]);
    const cd u = cmul(M[r0][j], M[r1][i]);
    return mkc(t.re - u.re, t.im - u.im);
  };
  cd d = mkc(0.0, 0.0);
  const cd t0 = cmul(minor(0, 1, 0, 1), minor(2, 3, 2, 3));
  const cd t1 = cmul(minor(0, 1, 0, 2), minor(2, 3, 1, 3));
  const cd t2 = cmul(minor(0, 1, 0, 3), minor(2, 3, 1, 2));
  const cd t3 = cmul(minor(0, 1, 1, 2), minor(2, 3, 0, 3));
  const cd t4 = cmul(minor(0, 1, 1, 3), minor(2, 3, 0, 2));
  const cd t5 = cmul(minor(0, 1, 2, 3), minor(2, 3, 0, 1));
  d.re = t0.re - t1.re + t2.re + t3.re - t4.re + t5.re;
  d.im = t0.im - t1.im + t2.im + t3.im - t4.im + t5.im;
  return d;
}

__device__ __forceinline__ double round8(double x) {  // round(x, 8) as CPython does, up to near-ties
  return rint(x * 1e8) / 1e8 + 0.0;
}

// U[row][col].  c (3) and/or g (3) may be null.
__device__ __forceinline__ void weyl_makhlin(const cd U[4][4], int flags, double* c_out, double* g_out) {
  // X = U Q, then UB = Q^dagger X, with the two 1/sqrt(2) folded into a single 1/2
  cd X[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    X[i][0] = mkc(U[i][0].re + U[i][3].re, U[i][0].im + U[i][3].im);
    X[i][1] = mkc(-(U[i][1].im + U[i][2].im), U[i][1].re + U[i][2].re);  // i (U1 + U2)
    X[i][2] = mkc(U[i][1].re - U[i][2].re, U[i][1].im - U[i][2].im);
    X[i][3] = mkc(-(U[i][0].im - U[i][3].im), U[i][0].re - U[i][3].re);  // i (U0 - U3)
  }
  cd UB[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    UB[0][j] = mkc(0.5 * (X[0][j].re + X[3][j].re), 0.5 * (X[0][j].im + X[3][j].im));
    UB[1][j] = mkc(0.5 * (X[1][j].im + X[2][j].im), -0.5 * (X[1][j].re + X[2][j].re));  // -i (X1 + X2)
    UB[2][j] = mkc(0.5 * (X[1][j].re - X[2][j].re), 0.5 * (X[1][j].im - X[2][j].im));
    UB[3][j] = mkc(0.5 * (X[0][j].im - X[3][j].im), -0.5 * (X[0][j].re - X[3][j].re));  // -i (X0 - X3)
  }
  // m = UB^T UB (upper triangle)
  double a[10], b[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      cd acc = mkc(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < 4; ++k) cacc(acc, UB[k][i], UB[k][j]);
      a[sym_idx(i, j)] = acc.re;
      b[sym_idx(i, j)] = acc.im;
    }
  const cd det = det4(U);

  if (g_out) {
    // tr(m), tr(m^2) = sum_ij m_ij^2 (m symmetric)
    cd tr = mkc(a[0] + a[4] + a[7] + a[9], b[0] + b[4] + b[7] + b[9]);
    cd tr2 = mkc(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const cd mij = mkc(a[sym_idx(i, j)], b[sym_idx(i, j)]);
        cacc(tr2, mij, mij);
      }
    const cd trsq = cmul(tr, tr);
    const double dn = 1.0 / fma(det.re, det.re, det.im * det.im);
    const cd dinv = mkc(det.re * dn, -det.im * dn);
    const cd g12 = cmul(trsq, dinv);
    const cd g3c = cmul(mkc(trsq.re - tr2.re, trsq.im - tr2.im), dinv);
    double g1 = g12.re / 16.0 + 0.0, g2 = g12.im / 16.0 + 0.0, g3 = g3c.re / 4.0 + 0.0;
    if (flags & SLAM_WEYL_ROUND8) {
      g1 = round8(g1);
      g2 = round8(g2);
      g3 = round8(g3);
    }
    g_out[0] = g1;
    g_out[1] = g2;
    g_out[2] = g3;
  }
  if (!c_out) return;

  // joint Jacobi sweeps on (Re m, Im m).  Not unrolled: the kernels that inline this function are instruction-fetch bound
  // (ncu: stall_no_instruction 3-5 per issue in the coverage kernels with the 12 sweeps unrolled, 46 KB of code for K3 alone)
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) off += fma(a[sym_idx(i, j)], a[sym_idx(i, j)], b[sym_idx(i, j)] * b[sym_idx(i, j)]);
    // Round-off keeps the off-diagonal mass of the rotated pair at ~1e-31 (six entries of ~1e-16, ||m||_F = 2), so the
    // exit test must sit above that floor: a threshold below it never fires and every matrix runs all 12 sweeps (ncu: 12
    // sweeps per warp with the former 1e-33; the quadratic convergence reaches the floor after 4-5).  off < 1e-26 bounds the
    // eigen-phase error by sqrt(off) = 1e-13 even for exactly degenerate pairs (typical: off / gap ~ 1e-26).
    if (off < 1e-26) break;
    joint_rotation<0, 1>(a, b);
    joint_rotation<0, 2>(a, b);
    joint_rotation<0, 3>(a, b);
    joint_rotation<1, 2>(a, b);
    joint_rotation<1, 3>(a, b);
    joint_rotation<2, 3>(a, b);
  }
  // two_S = angle(ev / sqrt(det)) / pi, principal square root.  The five atan2 run through ONE copy of the routine (a rolled
  // loop over a small local array): inlined five times they were 660 of the function's 2950 instructions, and the kernels
  // that use it are instruction-fetch bound (hot code above the 32 KB L1.5 instruction cache).
  double ang_y[5] = {det.im, b[sym_idx(0, 0)], b[sym_idx(1, 1)], b[sym_idx(2, 2)], b[sym_idx(3, 3)]};
  double ang_x[5] = {det.re, a[sym_idx(0, 0)], a[sym_idx(1, 1)], a[sym_idx(2, 2)], a[sym_idx(3, 3)]};
  double ang[5];
#pragma unroll 1
  for (int j = 0; j < 5; ++j) ang[j] = atan2(ang_y[j], ang_x[j]);
  const double half_det_phase = 0.5 * ang[0];
  const double inv_pi = 0.31830988618379067154, pi = 3.14159265358979323846;
  double S[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double th = ang[j + 1] - half_det_phase;
    if (th > pi) th -= 2.0 * pi;
    if (th <= -pi) th += 2.0 * pi;
    double two_s = th * inv_pi;
    // weylchamber: two_S[two_S <= -0.5] += 2.  SWAP-class inputs put ALL FOUR phases exactly on this branch point, where the
    // sign of the 1e-16 round-off decides the representative (four misses give (1.5, -0.5, 0.5); the reference needs its own
    // work-around there, speed_limit_pass.py:369-377).  A 1e-13 window resolves the tie towards the shifted branch.
    if (two_s <= -0.5 + 1e-13) two_s += 2.0;
    S[j] = 0.5 * two_s;
  }
  // sort descending (5-comparator network)
  auto cswap = [&](int i, int j) {
    const double hi = fmax(S[i], S[j]), lo = fmin(S[i], S[j]);
    S[i] = hi;
    S[j] = lo;
  };
  cswap(0, 1);
  cswap(2, 3);
  cswap(0, 2);
  cswap(1, 3);
  cswap(1, 2);
  const int n = (int)rint(S[0] + S[1] + S[2] + S[3]);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < n) S[j] -= 1.0;
  // roll(S, -n): out[i] = S[(i + n) % 4]
  double R[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int src = (i + n) & 3;
    R[i] = src == 0 ? S[0] : src == 1 ? S[1] : src == 2 ? S[2] : S[3];
  }
  double c1 = R[0] + R[1], c2 = R[0] + R[2], c3 = R[1] + R[2];
  if (c3 < 0.0) {
    c1 = 1.0 - c1;
    c3 = -c3;
  }
  if (flags & SLAM_WEYL_ROUND8) {
    c1 = round8(c1);
    c2 = round8(c2);
    c3 = round8(c3);
  }
  if ((flags & SLAM_WEYL_FOLD) && c1 > 0.5) c1 = 1.0 - c1;
  c_out[0] = c1 + 0.0;
  c_out[1] = c2 + 0.0;
  c_out[2] = c3 + 0.0;
}

}  // namespace slam
