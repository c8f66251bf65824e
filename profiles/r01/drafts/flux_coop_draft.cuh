// DRAFT -- not compiled into the library, not tested, not measured.  Kept so that the round-2 work on a co-operative
// (several lanes per solve) flux kernel for nstr >= 12 can start from it; see DESIGN.md section 7 for the
// instruction-efficiency estimate that argues for finishing it only if > 45 % FP64-pipe utilisation looks reachable.
// flux_coop.cuh -- K1c: the flux solve with T = nn/2 co-operating lanes per solve (sub-warp teams).
//
// Same mathematics as the thread-per-solve kernel (disort_solver_body.inc: symmetrised reduced eigenproblem,
// particular solutions from the eigen-decomposition, R/T from two SPD inverses, adding sweep), reorganised for
// large stream counts where one thread cannot hold a solve: at nstr = 16 the per-solve state is ~400 doubles, the
// thread-per-solve kernel spills 3.5 KB per thread and streams 350 KB of unrolled code through a 32 KB
// instruction cache (profiles/r01).  Here a TEAM of T adjacent lanes owns one solve:
//   * the nn x nn matrices of a layer live in per-team SHARED-MEMORY tiles; a lane computes a slice (two rows, or
//     two columns where the work is column-separable) in registers with compile-time indices, publishes it, and
//     the team meets at a __syncwarp;
//   * triangular recurrences are made column-separable: the small factorisations (Cholesky of Qt, Pt, S1, S2;
//     unpivoted LU of the M-matrix I - R Ra) are repeated by every lane in registers, the O(n^3) solves behind them
//     (X = L^-T V, Z = L V, S^-1, the nn+1 right-hand sides of the adding step) are split by columns;
//   * the eigen-solve is a ONE-SIDED Jacobi (Hestenes) on B = Cp^T L  (Pt = Cp Cp^T, Qt = L L^T, so B^T B =
//     L^T Pt L = G): rotating two columns needs three dot products (partial sums per lane, summed through shared
//     memory) and then touches only the lane's own rows; the Brent-Luk tournament permutes columns = registers.
// Identities used to avoid reductions: X Z^T = I, Z^T X = I, Qt^-1 = X X^T.
//
// Written as phases over a CoopLane (registers) and a CoopSM (shared tiles) so that tests/hostsim can run the T
// lanes of a team one after another between the same synchronisation points.
#pragma once

#include "disort_device.cuh"

namespace b200 {
namespace coop {

template <int NN>
struct CoopSM {
  static constexpr int T = NN / 2;
  double m[5][NN][NN];   // matrix tiles; m[4] = Ra (persistent across layers)
  double xc[NN];         // extra column of the right-hand-side tile m[3] (source column); rk^2 before that
  double v[6][NN];       // vector slots; v[0] = da (persistent)
  double red[T][2 * NN]; // partial sums published by each lane; red[k][12..15] = rotation parameters of pair k
  double pad[2];         // 16 bytes: consecutive teams start in different 16-byte bank groups
};
static_assert(sizeof(CoopSM<8>) == 3536, "CoopSM<8> layout");

template <int NN>
struct CoopLane {
  double mu[2], rmu[2], t[2], eh[2];  // constants of the lane's rows
  double Ra[2][NN];                   // own rows of the stack reflection matrix (also mirrored in sm.m[4])
  double da[2];
  double up[NN];                      // back sweep: upward radiation (replicated)
};

// replicated per-layer scalars
template <int NN>
struct CoopRep {
  double gl[2 * NN];
  double dtp, oprim, rmu0, umu0, cb, etop, ebot, pk_top, pk_bot;
  bool beam, planck;
};

#define COOP_UNROLL _Pragma("unroll")

// Cholesky of an SPD matrix whose lower triangle is read from a tile into packed registers; returns rinv.
template <int NN>
B200_HD bool chol_from_tile(const double (&tl)[NN][NN], double (&Lp)[NN * (NN + 1) / 2], double (&rinv)[NN]) {
  bool ok = true;
COOP_UNROLL
  for (int i = 0; i < NN; ++i)
COOP_UNROLL
    for (int k = 0; k <= i; ++k) Lp[sym(i, k)] = tl[i][k];
COOP_UNROLL
  for (int c = 0; c < NN; ++c) {
    double d = Lp[sym(c, c)];
COOP_UNROLL
    for (int k = 0; k < c; ++k) d = fma(-Lp[sym(c, k)], Lp[sym(c, k)], d);
    ok = ok && (d > 0.0);
    const double r = B200_RSQRT(d);
    rinv[c] = r;
    Lp[sym(c, c)] = d * r;
COOP_UNROLL
    for (int i = c + 1; i < NN; ++i) {
      double x = Lp[sym(i, c)];
COOP_UNROLL
      for (int k = 0; k < c; ++k) x = fma(-Lp[sym(i, k)], Lp[sym(c, k)], x);
      Lp[sym(i, c)] = x * r;
    }
  }
  return ok;
}

// x <- (L L^T)^-1 x for one vector held in registers
template <int NN>
B200_HD void chol_solve(const double (&Lp)[NN * (NN + 1) / 2], const double (&rinv)[NN], double (&x)[NN]) {
COOP_UNROLL
  for (int i = 0; i < NN; ++i) {
    double s = x[i];
COOP_UNROLL
    for (int k = 0; k < i; ++k) s = fma(-Lp[sym(i, k)], x[k], s);
    x[i] = s * rinv[i];
  }
COOP_UNROLL
  for (int i = NN - 1; i >= 0; --i) {
    double s = x[i];
COOP_UNROLL
    for (int k = i + 1; k < NN; ++k) s = fma(-Lp[sym(k, i)], x[k], s);
    x[i] = s * rinv[i];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Phase 1 (rows): Pt, Qt rows -> m[0], m[1]; beam projections as -> v[1]; own ad kept.
// ---------------------------------------------------------------------------------------------------------
template <int NN, typename QT>
B200_HD void ph1_matrices(const CoopLane<NN>& ln, CoopSM<NN>& sm, const CoopRep<NN>& rp, const QT& qt, const double* y0,
                          int h, double (&ad)[2], double (&as_own)[2]) {
  constexpr int N = 2 * NN;
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
    double P[NN], Q[NN], asv = 0.0, adv = 0.0;
COOP_UNROLL
    for (int k = 0; k < NN; ++k) {
      P[k] = (k == i) ? ln.rmu[a] : 0.0;
      Q[k] = P[k];
    }
COOP_UNROLL
    for (int l = 0; l < N; ++l) {
      const double yi = qt.yhat[l][i];
      const double gi = rp.gl[l] * yi;
      if (l & 1) {
        adv = fma(-rp.gl[l] * y0[l], yi, adv);
COOP_UNROLL
        for (int k = 0; k < NN; ++k) P[k] = fma(-gi, qt.yhat[l][k], P[k]);
      } else {
        asv = fma(rp.gl[l] * y0[l], yi, asv);
COOP_UNROLL
        for (int k = 0; k < NN; ++k) Q[k] = fma(-gi, qt.yhat[l][k], Q[k]);
      }
    }
COOP_UNROLL
    for (int k = 0; k < NN; ++k) {
      sm.m[0][i][k] = P[k];
      sm.m[1][i][k] = Q[k];
    }
    sm.v[1][i] = asv;
    as_own[a] = asv;
    ad[a] = adv;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Phase 2 (replicated factorisations): L = chol(Qt), Cp = chol(Pt) in registers; own rows published
// (L -> m[2], Cp -> m[3], zeros above the diagonal); own rhs = Pt as - ad/mu0 -> v[2].  Keeps L, rinvL.
// ---------------------------------------------------------------------------------------------------------
template <int NN>
B200_HD bool ph2_factor(CoopSM<NN>& sm, const CoopRep<NN>& rp, int h, const double (&ad)[2],
                        double (&Lp)[NN * (NN + 1) / 2], double (&rinvL)[NN]) {
  constexpr int T = NN / 2;
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
    double r = -ad[a] * rp.rmu0;
COOP_UNROLL
    for (int k = 0; k < NN; ++k) r = fma(sm.m[0][i][k], sm.v[1][k], r);
    sm.v[2][i] = r;
  }
  double Cp[NN * (NN + 1) / 2], rinvP[NN];
  bool ok = chol_from_tile<NN>(sm.m[1], Lp, rinvL);
  ok = chol_from_tile<NN>(sm.m[0], Cp, rinvP) && ok;
COOP_UNROLL
  for (int hh = 0; hh < T; ++hh)
    if (h == hh) {
COOP_UNROLL
      for (int a = 0; a < 2; ++a)
COOP_UNROLL
        for (int k = 0; k < NN; ++k) {
          sm.m[2][2 * hh + a][k] = (k <= 2 * hh + a) ? Lp[sym(2 * hh + a, k)] : 0.0;
          sm.m[3][2 * hh + a][k] = (k <= 2 * hh + a) ? Cp[sym(2 * hh + a, k)] : 0.0;
        }
    }
  return ok;
}

// ---------------------------------------------------------------------------------------------------------
// Phase 3 (rows): B = Cp^T L (own rows) and V = I (own rows) into registers.
// ---------------------------------------------------------------------------------------------------------
template <int NN>
B200_HD void ph3_factor_product(const CoopSM<NN>& sm, int h, double (&B)[2][NN], double (&V)[2][NN]) {
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      double s = 0.0;
COOP_UNROLL
      for (int k = 0; k < NN; ++k) s = fma(sm.m[3][k][i], sm.m[2][k][c], s);  // Cp[k][i] L[k][c]
      B[a][c] = s;
      V[a][c] = (c == i) ? 1.0 : 0.0;
    }
  }
}

// ---- one-sided Jacobi round, part a: partial dot products of the NN/2 column pairs -> red[h] ------------
template <int NN>
B200_HD void jac_dots(CoopSM<NN>& sm, int h, const double (&B)[2][NN]) {
COOP_UNROLL
  for (int k = 0; k < NN / 2; ++k) {
    const int p = 2 * k, q = 2 * k + 1;
    sm.red[h][3 * k + 0] = fma(B[0][p], B[0][p], B[1][p] * B[1][p]);
    sm.red[h][3 * k + 1] = fma(B[0][q], B[0][q], B[1][q] * B[1][q]);
    sm.red[h][3 * k + 2] = fma(B[0][p], B[0][q], B[1][p] * B[1][q]);
  }
}

// ---- part b: lane k turns the summed dot products of pair k into a rotation -> par[k]; returns whether the
//      pair still needed a rotation (relative threshold, as jacobi_spd) ------------------------------------
template <int NN>
B200_HD bool jac_params(CoopSM<NN>& sm, int h) {
  constexpr int T = NN / 2;
  const double eps2 = B200_JACOBI_EPS2;
  bool rot_any = false;
COOP_UNROLL
  for (int k = 0; k < T; ++k)
    if (h == k) {
      double al = 0.0, be = 0.0, ga = 0.0;
COOP_UNROLL
      for (int j = 0; j < T; ++j) {
        al += sm.red[j][3 * k + 0];
        be += sm.red[j][3 * k + 1];
        ga += sm.red[j][3 * k + 2];
      }
      const bool rot = ga * ga > eps2 * fabs(al * be);
      const double d = be - al, b2 = 2.0 * ga;
      const double rr = sqrt(fma(d, d, b2 * b2));
      const double den = d + copysign(rr, d);
      const double t = rot ? b2 * B200_RCP(den) : 0.0;
      const double c = B200_RSQRT(fma(t, t, 1.0));
      const double s = t * c;
      sm.red[k][NN + NN / 2 + 0] = t;
      sm.red[k][NN + NN / 2 + 1] = s;
      sm.red[k][NN + NN / 2 + 2] = s * B200_RCP(1.0 + c);
      sm.red[k][NN + NN / 2 + 3] = rot ? 1.0 : 0.0;
      rot_any = rot;
    }
  return rot_any;
}

// ---- part c: apply the NN/2 rotations to the own rows of B and V, then the Brent-Luk column permutation ---
template <int NN>
B200_HD bool jac_apply(const CoopSM<NN>& sm, double (&B)[2][NN], double (&V)[2][NN]) {
  bool any = false;
COOP_UNROLL
  for (int k = 0; k < NN / 2; ++k) {
    const int p = 2 * k, q = 2 * k + 1;
    const double s = sm.red[k][NN + NN / 2 + 1], tau = sm.red[k][NN + NN / 2 + 2];
    any = any || (sm.red[k][NN + NN / 2 + 3] != 0.0);
COOP_UNROLL
    for (int a = 0; a < 2; ++a) {
      double x = B[a][p], y = B[a][q];
      B[a][p] = fma(-s, fma(tau, x, y), x);
      B[a][q] = fma(s, fma(-tau, y, x), y);
      x = V[a][p];
      y = V[a][q];
      V[a][p] = fma(-s, fma(tau, x, y), x);
      V[a][q] = fma(s, fma(-tau, y, x), y);
    }
  }
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    double nb[NN], nv[NN];
COOP_UNROLL
    for (int j = 0; j < NN; ++j) {
      nb[j] = B[a][bl_src(j, NN)];
      nv[j] = V[a][bl_src(j, NN)];
    }
COOP_UNROLL
    for (int j = 0; j < NN; ++j) {
      B[a][j] = nb[j];
      V[a][j] = nv[j];
    }
  }
  return any;
}

// ---- after the sweeps: publish V rows -> m[0] and partial column norms -> red (NN values per lane) -------
template <int NN>
B200_HD void ph3_publish(CoopSM<NN>& sm, int h, const double (&B)[2][NN], const double (&V)[2][NN]) {
COOP_UNROLL
  for (int a = 0; a < 2; ++a)
COOP_UNROLL
    for (int c = 0; c < NN; ++c) sm.m[0][2 * h + a][c] = V[a][c];
COOP_UNROLL
  for (int c = 0; c < NN; ++c) sm.red[h][c] = fma(B[0][c], B[0][c], B[1][c] * B[1][c]);
}

// ---------------------------------------------------------------------------------------------------------
// Phase 4 (columns c = 2h, 2h+1): eigenvalue k_c, X[:,c] = L^-T V[:,c] -> m[1], Z[:,c] = L V[:,c] -> m[3],
// beam coefficient coef_c -> v[1], thermal xe_c = X[:,c].ehat -> v[3], d1_c -> v[4], d2_c -> v[5],
// rk_c^2 -> v[2] is NOT possible (v[2] = rhs is read here), so rk2 goes to red[0] after the team sync (phase 5
// reads it from v[2] once rhs is dead): it is returned to the caller instead and published in phase 4b.
// ---------------------------------------------------------------------------------------------------------
template <int NN, typename QT>
B200_HD int ph4_columns(CoopSM<NN>& sm, const CoopRep<NN>& rp, const QT& qt, int h, const double (&Lp)[NN * (NN + 1) / 2],
                        const double (&rinvL)[NN]) {
  constexpr int T = NN / 2;
  int st = ST_OK;
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int c = 2 * h + a;
    double lam = 0.0;
COOP_UNROLL
    for (int j = 0; j < T; ++j) lam += sm.red[j][c];
    if (!(lam > 0.0)) { st = ST_EIGEN_FAIL; lam = 1.0; }
    const double rk = B200_RSQRT(lam), kk = lam * rk;
    sm.xc[c] = rk * rk;
    double vcol[NN], x[NN], z[NN];
COOP_UNROLL
    for (int i = 0; i < NN; ++i) vcol[i] = sm.m[0][i][c];
COOP_UNROLL
    for (int i = 0; i < NN; ++i) {  // Z = L V
      double s = 0.0;
COOP_UNROLL
      for (int k = 0; k <= i; ++k) s = fma(Lp[sym(i, k)], vcol[k], s);
      z[i] = s;
    }
COOP_UNROLL
    for (int i = NN - 1; i >= 0; --i) {  // X = L^-T V
      double s = vcol[i];
COOP_UNROLL
      for (int k = i + 1; k < NN; ++k) s = fma(-Lp[sym(k, i)], x[k], s);
      x[i] = s * rinvL[i];
    }
    double cz = 0.0, xe = 0.0;
COOP_UNROLL
    for (int i = 0; i < NN; ++i) {
      sm.m[1][i][c] = x[i];
      sm.m[3][i][c] = z[i];
      cz = fma(z[i], sm.v[2][i], cz);
      xe = fma(x[i], qt.ehat[i], xe);
    }
    const double den = fma(kk, kk, -rp.rmu0 * rp.rmu0);
    if (rp.beam && den == 0.0) st = ST_SINGULAR;
    sm.v[1][c] = rp.beam ? cz * B200_RCP(den) : 0.0;
    sm.v[3][c] = xe;
    double em = 1.0 - exp(-kk * rp.dtp);
    em = fmax(em, 1.0e-300);
    const double q2 = 2.0 - em, kr = kk * B200_RCP(em * q2);
    sm.v[4][c] = kr * (q2 * q2);
    sm.v[5][c] = kr * (em * em);
  }
  return st;
}

// ---------------------------------------------------------------------------------------------------------
// Phase 5 (rows): particular solutions (own entries), S1 / S2 rows -> m[0], m[2]; publishes rk^2 -> v[2],
// p-(top) -> red[0][0..NN), p+(bottom) -> red[1][0..NN) for the sources.  Keeps ptu, pbd, own entries.
// ---------------------------------------------------------------------------------------------------------
template <int NN>
B200_HD void ph5_rows(CoopSM<NN>& sm, const CoopRep<NN>& rp, int h, const double (&as_own)[2], double (&ptu)[2],
                      double (&pbd)[2]) {
  double ptd[2], pbu[2];
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
    double xr[NN], zr[NN];
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      xr[c] = sm.m[1][i][c];
      zr[c] = sm.m[3][i][c];
    }
    ptu[a] = ptd[a] = pbu[a] = pbd[a] = 0.0;
    if (rp.beam) {
      double sg = 0.0, le = 0.0;
COOP_UNROLL
      for (int c = 0; c < NN; ++c) {
        sg = fma(xr[c], sm.v[1][c], sg);
        le = fma(zr[c], sm.v[1][c], le);
      }
      const double dl = rp.umu0 * (as_own[a] - le);
      const double zu = 0.5 * rp.cb * (sg + dl), zd = 0.5 * rp.cb * (sg - dl);
      ptu[a] = zu * rp.etop; ptd[a] = zd * rp.etop;
      pbu[a] = zu * rp.ebot; pbd[a] = zd * rp.ebot;
    }
    if (rp.planck) {
      // qv = Qt^-1 ehat = X (X^T ehat);  rv = Pt^-1 qv = Z Lam^-1 (Z^T qv) with Z^T qv = Z^T X xe = xe
      double qv = 0.0, rv = 0.0;
COOP_UNROLL
      for (int c = 0; c < NN; ++c) {
        qv = fma(xr[c], sm.v[3][c], qv);
        rv = fma(zr[c] * sm.xc[c], sm.v[3][c], rv);  // xc[c] = rk_c^2 (phase 4)
      }
      const double b0 = rp.pk_top;
      const double b1 = (rp.dtp > 0.0) ? (rp.pk_bot - b0) / rp.dtp : 0.0;
      const double om = 1.0 - rp.oprim, bb = b0 + b1 * rp.dtp, r1 = om * b1 * rv;
      ptu[a] += om * b0 * qv + r1; ptd[a] += om * b0 * qv - r1;
      pbu[a] += om * bb * qv + r1; pbd[a] += om * bb * qv - r1;
    }
    // S1 = I + X D1 X^T, S2 = I + X D2 X^T (own rows)
    double x1[NN], x2[NN];
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      x1[c] = xr[c] * sm.v[4][c];
      x2[c] = xr[c] * sm.v[5][c];
    }
COOP_UNROLL
    for (int k = 0; k < NN; ++k) {
      double s1 = (k == i) ? 1.0 : 0.0, s2 = s1;
COOP_UNROLL
      for (int c = 0; c < NN; ++c) {
        const double xk = sm.m[1][k][c];
        s1 = fma(x1[c], xk, s1);
        s2 = fma(x2[c], xk, s2);
      }
      sm.m[0][i][k] = s1;
      sm.m[2][i][k] = s2;
    }
    sm.red[0][i] = ptd[a];
    sm.red[1][i] = pbu[a];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Phase 6 (replicated Cholesky of S1, S2; own columns of the inverses = own rows, the matrices are symmetric):
// R, T rows in registers; sources; A = I - R Ra rows -> m[1]; right-hand sides [T | R da + s+] rows -> m[3].
// ---------------------------------------------------------------------------------------------------------
template <int NN>
B200_HD bool ph6_rt(const CoopLane<NN>& ln, CoopSM<NN>& sm, int h, const double (&ptu)[2], const double (&pbd)[2],
                    double (&R)[2][NN], double (&Tm)[2][NN], double (&smn)[2]) {
  double C1[NN * (NN + 1) / 2], C2[NN * (NN + 1) / 2], ri1[NN], ri2[NN];
  bool ok = chol_from_tile<NN>(sm.m[0], C1, ri1);
  ok = chol_from_tile<NN>(sm.m[2], C2, ri2) && ok;
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
    double e1[NN], e2[NN];
COOP_UNROLL
    for (int k = 0; k < NN; ++k) { e1[k] = (k == i) ? 1.0 : 0.0; e2[k] = e1[k]; }
    chol_solve<NN>(C1, ri1, e1);
    chol_solve<NN>(C2, ri2, e2);
COOP_UNROLL
    for (int k = 0; k < NN; ++k) {
      R[a][k] = ((k == i) ? 1.0 : 0.0) - e1[k] - e2[k];
      Tm[a][k] = e2[k] - e1[k];
    }
    // layer sources: s+ = p+(top) - R p-(top) - T p+(bot),  s- = p-(bot) - T p-(top) - R p+(bot)
    double sp = ptu[a], sn = pbd[a], rd = 0.0;
COOP_UNROLL
    for (int k = 0; k < NN; ++k) {
      sp = fma(-R[a][k], sm.red[0][k], sp);
      sp = fma(-Tm[a][k], sm.red[1][k], sp);
      sn = fma(-Tm[a][k], sm.red[0][k], sn);
      sn = fma(-R[a][k], sm.red[1][k], sn);
      rd = fma(R[a][k], sm.v[0][k], rd);
    }
    smn[a] = sn;
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      double s = (c == i) ? 1.0 : 0.0;
COOP_UNROLL
      for (int k = 0; k < NN; ++k) s = fma(-R[a][k], sm.m[4][k][c], s);
      sm.m[1][i][c] = s;
      sm.m[3][i][c] = Tm[a][c];
    }
    sm.xc[i] = sp + rd;
  }
  (void)ln;
  return ok;
}

// ---------------------------------------------------------------------------------------------------------
// Phase 7 (replicated unpivoted LU of the M-matrix A; own right-hand-side columns 2h, 2h+1 and, in every lane,
// the source column): M1 columns -> m[0]; v1 returned replicated.
// ---------------------------------------------------------------------------------------------------------
template <int NN>
B200_HD bool ph7_solve(CoopSM<NN>& sm, int h, double (&v1)[NN]) {
  double A[NN][NN], rd[NN];
  bool ok = true;
COOP_UNROLL
  for (int i = 0; i < NN; ++i)
COOP_UNROLL
    for (int c = 0; c < NN; ++c) A[i][c] = sm.m[1][i][c];
COOP_UNROLL
  for (int k = 0; k < NN; ++k) {
    ok = ok && (A[k][k] > 0.0);
    const double rp = B200_RCP(A[k][k]);
    rd[k] = rp;
COOP_UNROLL
    for (int r = k + 1; r < NN; ++r) {
      const double f = A[r][k] * rp;
      A[r][k] = f;
COOP_UNROLL
      for (int c = k + 1; c < NN; ++c) A[r][c] = fma(-f, A[k][c], A[r][c]);
    }
  }
COOP_UNROLL
  for (int a = 0; a < 3; ++a) {
    const int c = 2 * h + a;
    double x[NN];
COOP_UNROLL
    for (int i = 0; i < NN; ++i) x[i] = (a < 2) ? sm.m[3][i][c] : sm.xc[i];
COOP_UNROLL
    for (int i = 1; i < NN; ++i)
COOP_UNROLL
      for (int k = 0; k < i; ++k) x[i] = fma(-A[i][k], x[k], x[i]);
COOP_UNROLL
    for (int i = NN - 1; i >= 0; --i) {
      double s = x[i];
COOP_UNROLL
      for (int k = i + 1; k < NN; ++k) s = fma(-A[i][k], x[k], s);
      x[i] = s * rd[i];
    }
    if (a < 2) {
COOP_UNROLL
      for (int i = 0; i < NN; ++i) sm.m[0][i][c] = x[i];
    } else {
COOP_UNROLL
      for (int i = 0; i < NN; ++i) v1[i] = x[i];
    }
  }
  return ok;
}

// ---------------------------------------------------------------------------------------------------------
// Phase 8 (rows): save the level record (own rows of M1, v1, c^T Ra; lane 0 adds c^T da + direct beam);
// RM = Ra M1 rows -> m[2]; dv = da + Ra v1 -> v[1].
// Record layout per lane: [0..NN) M1 row 0, [NN..2NN) M1 row 1, v1 own (2), cRa own (2), cd (lane 0 only) = 2NN+5.
// ---------------------------------------------------------------------------------------------------------
template <int NN, typename QT>
B200_HD void ph8_save(const CoopLane<NN>& ln, CoopSM<NN>& sm, const QT& qt, int h, const double (&v1)[NN], double direct,
                      double* rec, long rstride) {
  double cd = 0.0;
COOP_UNROLL
  for (int k = 0; k < NN; ++k) cd = fma(qt.t[k], sm.v[0][k], cd);  // v[0] = da
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
    double m1[NN], cra = 0.0, dvv = ln.da[a];
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      m1[c] = sm.m[0][i][c];
      rec[(long)(a * NN + c) * rstride] = m1[c];
      cra = fma(ln.Ra[a][c], qt.t[c], cra);  // Ra symmetric: (t^T Ra)_i = (Ra t)_i
      dvv = fma(ln.Ra[a][c], v1[c], dvv);
    }
    double vi = 0.0;
COOP_UNROLL
    for (int hh = 0; hh < NN / 2; ++hh)
      if (h == hh) vi = v1[2 * hh + a];
    rec[(long)(2 * NN + a) * rstride] = vi;
    rec[(long)(2 * NN + 2 + a) * rstride] = 2.0 * kPi * cra;
    sm.v[1][i] = dvv;
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      double s = 0.0;
COOP_UNROLL
      for (int k = 0; k < NN; ++k) s = fma(ln.Ra[a][k], sm.m[0][k][c], s);
      sm.m[2][i][c] = s;
    }
  }
  rec[(long)(2 * NN + 4) * rstride] = 2.0 * kPi * cd + direct;
}

// ---------------------------------------------------------------------------------------------------------
// Phase 9 (rows): new stack Ra' = R + T (Ra M1), da' = s- + T (da + Ra v1); published to m[4], v[0].
// ---------------------------------------------------------------------------------------------------------
template <int NN>
B200_HD void ph9_stack(CoopLane<NN>& ln, CoopSM<NN>& sm, int h, const double (&R)[2][NN], const double (&Tm)[2][NN],
                       const double (&smn)[2], bool frozen) {
COOP_UNROLL
  for (int a = 0; a < 2; ++a) {
    const int i = 2 * h + a;
    double d = smn[a];
COOP_UNROLL
    for (int k = 0; k < NN; ++k) d = fma(Tm[a][k], sm.v[1][k], d);
    double nr[NN];
COOP_UNROLL
    for (int c = 0; c < NN; ++c) {
      double s = R[a][c];
COOP_UNROLL
      for (int k = 0; k < NN; ++k) s = fma(Tm[a][k], sm.m[2][k][c], s);
      nr[c] = s;
    }
    if (!frozen) {
      ln.da[a] = d;
      sm.v[0][i] = d;
COOP_UNROLL
      for (int c = 0; c < NN; ++c) {
        ln.Ra[a][c] = nr[c];
        sm.m[4][i][c] = nr[c];
      }
    }
  }
}

#undef COOP_UNROLL

}  // namespace coop
}  // namespace b200
