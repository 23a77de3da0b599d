// slam_adj1.cuh -- thread-per-problem loss + ANALYTIC gradient for templates whose 2Q gate is a parameter-bound smush
// gate (GM_SMUSH): the adjoint sweep of slam_core.cuh extended through the time-sliced exponentials of
// src/slam/hamiltonian.py:114-182.  Replaces the (P+1)-evaluation finite-difference gradient scipy builds for these
// templates (opt.minimize(jac=None), src/slam/optimizer.py:270-278) by one backward pass.
//
// Per slice E = exp(-i dt H), H a general 4x4 Hermitian matrix of the gate parameters:
//   * H = Q diag(lam) Q^dagger by cyclic complex Jacobi (backward stable, 3-5 sweeps for a 4x4);
//   * the running right product is un-applied in the eigenbasis (r_prev = Q e^{+i dt lam} Q^dagger r), so nothing is
//     stored per slice (reversible adjoint, as for the 1Q layers);
//   * the derivative of the exponential is the Daleckii-Krein form dE = Q [Phi o (Q^dagger dA Q)] Q^dagger with
//     dA = -i dt dH and the divided differences Phi_ab = e^{-i dt (lam_a+lam_b)/2} sinc(dt (lam_a-lam_b)/2)
//     (Phi_aa = e^{-i dt lam_a}); contracting it with the environment sum_c r_prev[c] (x) w[c] gives one 4x4 matrix
//     G = conj(Q) [Phi o M] Q^T whose entries are d(loss)/d(H entries): every gate parameter (amplitudes, phases,
//     couplings, Z terms) is a short linear combination of them, and d/d(dt) = sum_a lam_a Im(Phi_aa M_aa).
// The formulas are validated against central differences of scipy.linalg.expm in tests/test_gpu_smush_adjoint.py.
#pragma once
#include "slam_fwd1.cuh"

namespace slam {

__device__ __forceinline__ void hg_set(HermG& A, int i, int j, cd v) {  // i != j
  if (i < j) A.u[hg_idx(i, j)] = v;
  else A.u[hg_idx(j, i)] = mkc(v.re, -v.im);
}

// a*x - b*y and a*x + b*y for real a and complex b (Jacobi column updates)
__device__ __forceinline__ cd rot_minus(double a, cd x, cd b, cd y) {
  return mkc(fma(a, x.re, fma(-b.re, y.re, b.im * y.im)), fma(a, x.im, fma(-b.re, y.im, -(b.im * y.re))));
}
__device__ __forceinline__ cd rot_plus(double a, cd x, cd b, cd y) {
  return mkc(fma(a, x.re, fma(b.re, y.re, -(b.im * y.im))), fma(a, x.im, fma(b.re, y.im, b.im * y.re)));
}

// Eigen-decomposition of a Hermitian 4x4 by cyclic Jacobi with complex rotations.  On return A.d = eigenvalues and
// Q[a][i] = component i of eigenvector a (A.u is annihilated to ~1e-17 ||A||).
__device__ __forceinline__ void herm_eig4(HermG& A, cd Q[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i) Q[a][i] = mkc(a == i ? 1.0 : 0.0, 0.0);
#pragma unroll 1
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0, tot = 0.0;
#pragma unroll
    for (int e = 0; e < 6; ++e) off = fma(A.u[e].re, A.u[e].re, fma(A.u[e].im, A.u[e].im, off));
#pragma unroll
    for (int i = 0; i < 4; ++i) tot = fma(A.d[i], A.d[i], tot);
    if (off <= 1e-30 * (tot + off)) break;  // above the round-off floor (~1e-32 ||A||_F^2), so the test always fires
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const cd h = A.u[hg_idx(p, q)];
        const double a2 = fma(h.re, h.re, h.im * h.im);
        if (a2 > 1e-280) {
          const double inv = rsqrt(a2), absh = a2 * inv;
          const double er = h.re * inv, ei = -h.im * inv;  // e^{-i arg h}
          // rotation angle phi of the real 2x2 problem [[dp, |h|], [|h|, dq]]: cos 2phi = |dq - dp| / r, sin 2phi = 2|h| / r,
          // r = hypot(dq - dp, 2|h|); half-angle formulas with rsqrt only (the divide + sqrt form of the textbook tangent was
          // the hottest line of the adjoint kernels: 12 % of the stall samples)
          const double da = A.d[q] - A.d[p];
          const double inv_r = rsqrt(fma(da, da, 4.0 * a2));
          const double u = 0.5 * fma(fabs(da), inv_r, 1.0);  // cos^2 phi, in [1/2, 1]
          const double ic = rsqrt(u);
          const double c = u * ic;
          const double s = copysign(absh * inv_r * ic, da);
          const double t = s * ic;  // tan phi
          A.d[p] = fma(-t, absh, A.d[p]);
          A.d[q] = fma(t, absh, A.d[q]);
          A.u[hg_idx(p, q)] = mkc(0.0, 0.0);
          const cd se = mkc(s * er, s * ei), ce = mkc(c * er, c * ei);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k != p && k != q) {
              const cd hkp = hg_get(A, k, p), hkq = hg_get(A, k, q);
              hg_set(A, k, p, rot_minus(c, hkp, se, hkq));
              hg_set(A, k, q, rot_plus(s, hkp, ce, hkq));
            }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const cd qp = Q[p][i], qq = Q[q][i];
            Q[p][i] = rot_minus(c, qp, se, qq);
            Q[q][i] = rot_plus(s, qp, ce, qq);
          }
        }
      }
  }
}

// gradient accumulators of one smush gate (summed over its slices)
struct SmushGrad {
  double pa, pb, pc, pg, gc, gg, gz1, gz2, dt;
};

// ---- eigen-decomposition hand-over from the forward to the backward sweep ---------------------------------------------
// The backward pass needs H = Q diag(lam) Q^dagger of every slice.  When the forward sweep of the SAME evaluation also
// exponentiates through it (E = Q e^{-i dt lam} Q^dagger, instead of the cos/sin series of slam_fwd1.cuh: about the same
// cost, one Jacobi + two 4x4 products per slice) and keeps (Q, lam), the backward sweep skips its Jacobi -- half of a
// slice's backward work, a quarter of the whole loss + gradient evaluation.  The 36 doubles per slice live in a
// thread-local array (local memory: interleaved across the warp by the hardware, L1-cached), bounded to kEigSlices
// slices; templates with more slices use the recomputing form.
constexpr int kEigSlices = 8;
struct NoEigStore {
  static constexpr bool on = false;
  __device__ __forceinline__ void put(int, const cd (*)[4], const double*) {}
  __device__ __forceinline__ void get(int, cd (*)[4], double*) const {}
};
struct LocalEigStore {
  static constexpr bool on = true;
  double buf[kEigSlices * 36];
  __device__ __forceinline__ void put(int slice, const cd (*Q)[4], const double* lam) {
    double* p = buf + slice * 36;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p[(a * 4 + i) * 2] = Q[a][i].re;
        p[(a * 4 + i) * 2 + 1] = Q[a][i].im;
      }
      p[32 + a] = lam[a];
    }
  }
  __device__ __forceinline__ void get(int slice, cd (*Q)[4], double* lam) const {
    const double* p = buf + slice * 36;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Q[a][i] = mkc(p[(a * 4 + i) * 2], p[(a * 4 + i) * 2 + 1]);
      lam[a] = p[32 + a];
    }
  }
};

// the slice generator of hamiltonian.py:114-182 as a general Hermitian 4x4 (same entries as smush_slice in slam_fwd1.cuh)
__device__ __forceinline__ void smush_generator(const SmushGate& G, double gx, double gy, HermG& A) {
  A.d[0] = G.gz1 + G.gz2;
  A.d[1] = G.gz1;
  A.d[2] = G.gz2;
  A.d[3] = 0.0;
  A.u[0] = mkc(gy * G.eb.re, gy * G.eb.im);  // (0,1)
  A.u[1] = mkc(gx * G.ea.re, gx * G.ea.im);  // (0,2)
  A.u[2] = mkc(G.gg * G.eg.re, G.gg * G.eg.im);  // (0,3)
  A.u[3] = mkc(G.gc * G.ec.re, G.gc * G.ec.im);  // (1,2)
  A.u[4] = A.u[1];                               // (1,3)
  A.u[5] = A.u[0];                               // (2,3)
}

// Forward through one slice in eigen-form: R[c] <- Q e^{-i dt lam} Q^dagger R[c]; (Q, lam) handed to the backward sweep.
template <bool SYNC, class ES>
__device__ __forceinline__ void smush_slice_fwd_eig(const SmushGate& G, double gx, double gy, double dt, cd R[4][4], ES& es,
                                                    int slice) {
  HermG A;
  smush_generator(G, gx, gy, A);
  cd Q[4][4];
  if (SYNC) __syncthreads();
  herm_eig4(A, Q);
  if (SYNC) __syncthreads();
  es.put(slice, Q, A.d);
  cd ph[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double s, c;
    fast_sincos(-dt * A.d[a], &s, &c);
    ph[a] = mkc(c, s);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cd t[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      cd x = cmulc(R[c][0], Q[a][0]);  // (Q^dagger r)_a = sum_i conj(Q[a][i]) r_i
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const cd z = cmulc(R[c][i], Q[a][i]);
        x.re += z.re;
        x.im += z.im;
      }
      t[a] = cmul(x, ph[a]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cd x = cmul(Q[0][i], t[0]);
#pragma unroll
      for (int a = 1; a < 4; ++a) cacc(x, Q[a][i], t[a]);
      R[c][i] = x;
    }
  }
}

// Backward through one slice.  In: R[c] = column c of the right product INCLUDING the slice, W[c] = row c of the left
// environment EXCLUDING it.  Out: R[c] without the slice, W[c] including it; *d_gx, *d_gy = d loss / d amplitudes of this
// slice; gate-level derivatives accumulated into `acc`.
// SYNC: CTA-wide barriers between the sub-phases (eigen-decomposition | into the eigenbasis | divided differences and G |
// parameter derivatives and back-transform), each 5-15 KB of code: see slam_fwd1.cuh (fwd1_gate) for why.
template <bool SYNC = false, class ES = NoEigStore>
__device__ __forceinline__ void smush_slice_bwd(const SmushGate& G, double gx, double gy, double dt, cd R[4][4], cd W[4][4],
                                                double* d_gx, double* d_gy, SmushGrad& acc, const ES& es = ES(),
                                                int slice = 0) {
  HermG A;
  cd Q[4][4];
  if (ES::on) {
    if (SYNC) __syncthreads();
    es.get(slice, Q, A.d);  // (Q, lam) of this slice from the forward sweep
  } else {
    smush_generator(G, gx, gy, A);
    if (SYNC) __syncthreads();
    herm_eig4(A, Q);
    if (SYNC) __syncthreads();
  }
  cd hp[4], ph[4];  // e^{-i dt lam / 2}, e^{-i dt lam}
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double s, c;
    fast_sincos(-0.5 * dt * A.d[a], &s, &c);
    hp[a] = mkc(c, s);
    ph[a] = mkc(fma(c, c, -(s * s)), 2.0 * c * s);
  }
  // into the eigenbasis: R[c][a] <- conj(ph_a) (Q^dagger r_c)_a  (the slice un-applied),  W[c][a] <- (w_c Q)_a
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cd tr[4], tw[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      cd x = cmulc(R[c][0], Q[a][0]);
      cd y = cmul(W[c][0], Q[a][0]);
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        const cd z = cmulc(R[c][i], Q[a][i]);
        x.re += z.re;
        x.im += z.im;
        cacc(y, W[c][i], Q[a][i]);
      }
      tr[a] = cmulc(x, ph[a]);
      tw[a] = y;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      R[c][a] = tr[a];
      W[c][a] = tw[a];
    }
  }
  if (SYNC) __syncthreads();
  // N = Phi o M,  M_ab = sum_c W[c][a] R[c][b]
  cd N[4][4];
  double ddt = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      cd m = cmul(W[0][a], R[0][b]);
#pragma unroll
      for (int c = 1; c < 4; ++c) cacc(m, W[c][a], R[c][b]);
      cd phi;
      if (a == b) {
        phi = ph[a];
      } else {
        const double delta = 0.5 * dt * (A.d[a] - A.d[b]);
        const double sd = -cmulc(hp[a], hp[b]).im;  // sin(delta)
        const double d2 = delta * delta;  // series below |delta| = 0.05 (truncation < 1e-20), where sd / delta loses digits
        const double ser = fma(d2, fma(d2, fma(d2, fma(d2, 1.0 / 362880.0, -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
        const double sinc = fabs(delta) > 0.05 ? sd / delta : ser;
        const cd pp = cmul(hp[a], hp[b]);
        phi = mkc(pp.re * sinc, pp.im * sinc);
      }
      N[a][b] = cmul(phi, m);
      if (a == b) ddt = fma(A.d[a], N[a][a].im, ddt);
    }
  acc.dt += ddt;
  // G = conj(Q) N Q^T :  Z[a][j] = sum_b N[a][b] Q[b][j],  G[i][j] = sum_a conj(Q[a][i]) Z[a][j]
  cd Gm[4][4];
  {
    cd Z[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cd z = cmul(N[a][0], Q[0][j]);
#pragma unroll
        for (int b = 1; b < 4; ++b) cacc(z, N[a][b], Q[b][j]);
        Z[a][j] = z;
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cd g = cmulc(Z[0][j], Q[0][i]);
#pragma unroll
        for (int a = 1; a < 4; ++a) {
          const cd z = cmulc(Z[a][j], Q[a][i]);
          g.re += z.re;
          g.im += z.im;
        }
        Gm[i][j] = g;
      }
  }
  if (SYNC) __syncthreads();
  // d loss / d(Re H_ij) = dt Im(G_ij + G_ji),  d loss / d(Im H_ij) = dt Re(G_ij - G_ji),  d loss / d(H_ii) = dt Im G_ii
  // entries: (0,1),(2,3) = gy e^{-i pb};  (0,2),(1,3) = gx e^{-i pa};  (0,3) = gg e^{-i pg};  (1,2) = gc e^{-i pc}
  {
    const double sx = dt * (Gm[0][1].im + Gm[1][0].im + Gm[2][3].im + Gm[3][2].im);
    const double sy = dt * (Gm[0][1].re - Gm[1][0].re + Gm[2][3].re - Gm[3][2].re);
    *d_gy = fma(sx, G.eb.re, sy * G.eb.im);
    acc.pb = fma(gy, fma(sx, G.eb.im, -(sy * G.eb.re)), acc.pb);
  }
  {
    const double sx = dt * (Gm[0][2].im + Gm[2][0].im + Gm[1][3].im + Gm[3][1].im);
    const double sy = dt * (Gm[0][2].re - Gm[2][0].re + Gm[1][3].re - Gm[3][1].re);
    *d_gx = fma(sx, G.ea.re, sy * G.ea.im);
    acc.pa = fma(gx, fma(sx, G.ea.im, -(sy * G.ea.re)), acc.pa);
  }
  {
    const double sx = dt * (Gm[0][3].im + Gm[3][0].im), sy = dt * (Gm[0][3].re - Gm[3][0].re);
    acc.gg += fma(sx, G.eg.re, sy * G.eg.im);
    acc.pg = fma(G.gg, fma(sx, G.eg.im, -(sy * G.eg.re)), acc.pg);
  }
  {
    const double sx = dt * (Gm[1][2].im + Gm[2][1].im), sy = dt * (Gm[1][2].re - Gm[2][1].re);
    acc.gc += fma(sx, G.ec.re, sy * G.ec.im);
    acc.pc = fma(G.gc, fma(sx, G.ec.im, -(sy * G.ec.re)), acc.pc);
  }
  acc.gz1 = fma(dt, Gm[0][0].im + Gm[1][1].im, acc.gz1);
  acc.gz2 = fma(dt, Gm[0][0].im + Gm[2][2].im, acc.gz2);
  // back to the computational basis: r_prev = Q r~,  w_new = (w~ ph) Q^dagger
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    cd wt[4], nr[4], nw[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) wt[a] = cmul(W[c][a], ph[a]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cd x = cmul(Q[0][i], R[c][0]);
      cd y = cmulc(wt[0], Q[0][i]);
#pragma unroll
      for (int a = 1; a < 4; ++a) {
        cacc(x, Q[a][i], R[c][a]);
        const cd z = cmulc(wt[a], Q[a][i]);
        y.re += z.re;
        y.im += z.im;
      }
      nr[i] = x;
      nw[i] = y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      R[c][i] = nr[i];
      W[c][i] = nw[i];
    }
  }
}

// loss only (forward evaluation + trace functional), one thread
template <class PS, bool SYNC = false>
__device__ __forceinline__ double fwd1_loss(const KTemplate& kt, const PS& ps, const double* __restrict__ V, int cost_kind,
                                            cd* T_out) {
  cd R[4][4];
  fwd1_chain<PS, GM_SMUSH, SYNC>(kt, ps, R);  // (only launched for parameter-bound smush templates)
  cd T = mkc(0.0, 0.0);
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double vr = V[(a * 4 + c) * 2], vi = V[(a * 4 + c) * 2 + 1];
      T.re = fma(vr, R[c][a].re, fma(vi, R[c][a].im, T.re));
      T.im = fma(vr, R[c][a].im, fma(-vi, R[c][a].re, T.im));
    }
  double loss, dl_da;
  cost_from_abs(cost_kind, sqrt(fma(T.re, T.re, T.im * T.im)), loss, dl_da);
  if (T_out) *T_out = T;
  return loss;
}

// gradient sinks
struct RowGrad {  // row of a [B, ldg] array (row == nullptr: discard -- padding lanes of a phase-locked CTA)
  double* row;
  __device__ __forceinline__ void set(int j, double v) const {
    if (j >= 0 && row) row[j] = v;
  }
};
struct StridedGrad {  // workspace vector interleaved across threads
  double* p;
  int64_t stride;
  __device__ __forceinline__ void set(int j, double v) const {
    if (j >= 0) p[(int64_t)j * stride] = v;
  }
};

// ------------------------------------------------------------------------------------------------
// loss + gradient of cost(template(x), V) for a GM_SMUSH template, one thread.
//   ES = LocalEigStore: eigen-form forward sweep that hands (Q, lam) of every slice to the backward sweep (k T <= kEigSlices)
//   ps.get(j)      : parameter j (API order)
//   gs.set(j, v)   : d loss / d parameter j  (every parameter is bound to exactly one slot -- checked on the host --
//                    so each entry is set once)
// returns the loss; *T_out = Tr(V^dagger U) if non-null
// ------------------------------------------------------------------------------------------------
// forward chain of a GM_SMUSH template with the slices in eigen-form (see LocalEigStore); slice index = gate * T + slice
template <class PS, bool SYNC, class ES>
__device__ __forceinline__ void fwd1_chain_eig(const KTemplate& kt, const PS& ps, cd R[4][4], ES& es) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) R[c][r] = mkc(c == r ? 1.0 : 0.0, 0.0);
  const int Tn = kt.T;
  const bool ph1q = kt.gate_kind == SLAM_GATE_SMUSH_1QPHASE;
  const int o = ph1q ? 8 : 4;  // first gx slot
  for (int i = 0; i <= kt.k; ++i) {
    if (SYNC) __syncthreads();
    fwd1_layer(kt, ps, i, R);
    if (i < kt.k) {
      const int g = i;
      SmushGate G;
      if (ph1q)
        G = smush_gate(slot_val(kt, ps, g, 0), slot_val(kt, ps, g, 1), slot_val(kt, ps, g, 2), slot_val(kt, ps, g, 3),
                       slot_val(kt, ps, g, 4), slot_val(kt, ps, g, 5), slot_val(kt, ps, g, 6), slot_val(kt, ps, g, 7));
      else
        G = smush_gate(0.0, 0.0, slot_val(kt, ps, g, 0), slot_val(kt, ps, g, 1), slot_val(kt, ps, g, 2),
                       slot_val(kt, ps, g, 3), 0.0, 0.0);
      const double dt = slot_val(kt, ps, g, o + 2 * Tn) / (double)Tn;
      for (int it = 0; it < Tn; ++it)
        smush_slice_fwd_eig<SYNC>(G, slot_val(kt, ps, g, o + it), slot_val(kt, ps, g, o + Tn + it), dt, R, es, g * Tn + it);
    }
  }
}

template <class PS, class GS, bool SYNC = false, class ES = NoEigStore>
__device__ __forceinline__ double adj1_loss_grad(const KTemplate& kt, const PS& ps, const double* __restrict__ V,
                                                 int cost_kind, GS& gs, cd* T_out) {
  cd R[4][4];  // [col][row]
  ES es;
  if (ES::on) fwd1_chain_eig<PS, SYNC>(kt, ps, R, es);
  else fwd1_chain<PS, GM_SMUSH, SYNC>(kt, ps, R);
  cd T = mkc(0.0, 0.0);
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const double vr = V[(a * 4 + c) * 2], vi = V[(a * 4 + c) * 2 + 1];
      T.re = fma(vr, R[c][a].re, fma(vi, R[c][a].im, T.re));
      T.im = fma(vr, R[c][a].im, fma(-vi, R[c][a].re, T.im));
    }
  const double absT = sqrt(fma(T.re, T.re, T.im * T.im));
  double loss, dl_da;
  cost_from_abs(cost_kind, absT, loss, dl_da);
  if (T_out) *T_out = T;
  // W[c][a] = (dl/d|T|) conj(T)/|T| conj(V[a][c]):  d loss = Re sum_c W[c] . dM . R[c]
  const double inv = absT > 0.0 ? dl_da / absT : 0.0;
  const cd phs = mkc(T.re * inv, -T.im * inv);
  cd W[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) W[c][a] = cmulc(phs, mkc(V[(a * 4 + c) * 2], V[(a * 4 + c) * 2 + 1]));

  const int Tn = kt.T;
  const bool ph1q = kt.gate_kind == SLAM_GATE_SMUSH_1QPHASE;
  const int o = ph1q ? 8 : 4;  // first gx slot
  for (int i = kt.k; i >= 0; --i) {
    if (SYNC) __syncthreads();
    if (!(kt.p1q[i][0] < 0 && kt.p1q[i][3] < 0)) {
      if (kt.vz_only) {
        // RZ = diag(e^{-i l/2}, e^{+i l/2}) on each qubit: d loss/dl = (1/2) (Im sum_lo - Im sum_hi) w r at the cut after the layer
        double s, c;
        cd A[4], B[4];
        sincos(0.5 * ps.get(kt.p1q[i][0]), &s, &c);
        build_rz(make_double2(c, s), B);
        sincos(0.5 * ps.get(kt.p1q[i][3]), &s, &c);
        build_rz(make_double2(c, s), A);
        double d0 = 0.0, d3 = 0.0;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          double all = 0.0;
#pragma unroll
          for (int a = 0; a < 4; ++a) all = fma(W[c4][a].re, R[c4][a].im, fma(W[c4][a].im, R[c4][a].re, all));
          d0 += all - 2.0 * diag_im<0>(W[c4], R[c4], 0.0);
          d3 += all - 2.0 * diag_im<1>(W[c4], R[c4], 0.0);
          apply1q<1, OP_H>(R[c4], A);
          apply1q<0, OP_H>(R[c4], B);
          apply1q<1, OP_T>(W[c4], A);
          apply1q<0, OP_T>(W[c4], B);
        }
        gs.set(kt.p1q[i][0], 0.5 * d0);
        gs.set(kt.p1q[i][3], 0.5 * d3);
      } else {
        double2 t[6];
#pragma unroll 1
        for (int q = 0; q < 6; ++q) {
          const double v = ps.get(kt.p1q[i][q]);
          double s, c;
          fast_sincos((q == 0 || q == 3) ? 0.5 * v : v, &s, &c);
          t[q] = make_double2(c, s);
        }
        double dq[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          u3_bwd<1>(R[c], W[c], t + 3, dq[1]);
          u3_bwd<0>(R[c], W[c], t, dq[0]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          gs.set(kt.p1q[i][3 * q + 0], 0.5 * dq[q][0]);
          gs.set(kt.p1q[i][3 * q + 1], -dq[q][1]);
          gs.set(kt.p1q[i][3 * q + 2], -dq[q][2]);
        }
      }
    }
    if (i > 0) {
      const int g = i - 1;
      double pa = 0.0, pb = 0.0, pc, pg, gc, gg, gz1 = 0.0, gz2 = 0.0;
      if (ph1q) {
        pa = slot_val(kt, ps, g, 0);
        pb = slot_val(kt, ps, g, 1);
        pc = slot_val(kt, ps, g, 2);
        pg = slot_val(kt, ps, g, 3);
        gc = slot_val(kt, ps, g, 4);
        gg = slot_val(kt, ps, g, 5);
        gz1 = slot_val(kt, ps, g, 6);
        gz2 = slot_val(kt, ps, g, 7);
      } else {
        pc = slot_val(kt, ps, g, 0);
        pg = slot_val(kt, ps, g, 1);
        gc = slot_val(kt, ps, g, 2);
        gg = slot_val(kt, ps, g, 3);
      }
      const SmushGate G = smush_gate(pa, pb, pc, pg, gc, gg, gz1, gz2);
      const double dt = slot_val(kt, ps, g, o + 2 * Tn) / (double)Tn;
      SmushGrad acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int it = Tn - 1; it >= 0; --it) {
        double dgx, dgy;
        smush_slice_bwd<SYNC, ES>(G, slot_val(kt, ps, g, o + it), slot_val(kt, ps, g, o + Tn + it), dt, R, W, &dgx, &dgy, acc,
                                  es, g * Tn + it);
        int p;
        if ((p = kt.slot_param[g][o + it]) >= 0) gs.set(p, dgx);
        if ((p = kt.slot_param[g][o + Tn + it]) >= 0) gs.set(p, dgy);
      }
      int p;
      if (ph1q) {
        if ((p = kt.slot_param[g][0]) >= 0) gs.set(p, acc.pa);
        if ((p = kt.slot_param[g][1]) >= 0) gs.set(p, acc.pb);
        if ((p = kt.slot_param[g][6]) >= 0) gs.set(p, acc.gz1);
        if ((p = kt.slot_param[g][7]) >= 0) gs.set(p, acc.gz2);
      }
      const int b = ph1q ? 2 : 0;
      if ((p = kt.slot_param[g][b + 0]) >= 0) gs.set(p, acc.pc);
      if ((p = kt.slot_param[g][b + 1]) >= 0) gs.set(p, acc.pg);
      if ((p = kt.slot_param[g][b + 2]) >= 0) gs.set(p, acc.gc);
      if ((p = kt.slot_param[g][b + 3]) >= 0) gs.set(p, acc.gg);
      if ((p = kt.slot_param[g][o + 2 * Tn]) >= 0) gs.set(p, acc.dt / (double)Tn);
    }
  }
  return loss;
}

}  // namespace slam
