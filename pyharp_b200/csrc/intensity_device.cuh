// intensity_device.cuh -- K3: the general DISORT solve with user levels, user angles, the
// azimuth Fourier series and the Nakajima-Tanaka intensity corrections
// (c_disort with !onlyfl / usrtau: c_sol_eigen, c_upbeam, c_upisot, c_terpev, c_terpso,
// c_set_matrix, c_solve0, c_fluxes, c_user_intensities, c_intensity_components,
// c_intensity_correction -- SURVEY.md 8 rows a4, a6-a16; reached from pyharp through
// disort->forward + disort->get_rad, tests/test_disort.cpp:49-52).
//
// Parallel structure (one CTA owns one solve, SURVEY.md north star): the layers of a column are
// independent until the boundary-value solve, so the CTA's threads take layers in the eigen /
// particular-solution phase, rows in the banded elimination, and (level, angle) outputs in the
// source-function integration.  Phases are separated by CTA barriers; all per-solve state
// lives in a per-CTA global workspace (sizes are run-time: nstr <= 64, any nlyr).
// The reduced eigenproblem uses the same symmetrisation as the flux kernel (Cholesky + Jacobi
// on the nn x nn matrix scaled by sqrt(mu w)); its eigenvectors are mapped back to DISORT's
// (G+ + G-), (G+ - G-) so that everything downstream follows the published formulas.
//
// Also compiled by g++ for tests/hostsim (single "thread", barriers vanish); the product only
// instantiates it in __global__ kernels.
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>

#include "disort_device.cuh"

#if defined(__CUDACC__)
#define B200_FN_ATTR __host__ __device__
#else
#define B200_FN_ATTR
#endif
#if defined(__CUDA_ARCH__)
#define B200_SYNC() __syncthreads()
#else
#define B200_SYNC() ((void)0)
#endif
// strided loop over the CTA's threads
#define B200_PAR_FOR(i, n) for (int i = tid; i < (n); i += nthr)

namespace b200 {

struct IntensityParams {
  int nwave, ncol, nlyr, nprop, nstr, nmom;
  int ntau, numu, nphi;  // effective sizes: ntau = nlyr+1 unless USRTAU; numu = nstr unless USRANG
  unsigned flags;
  double accur;
  const double* prop;
  const double *umu0, *fbeam, *albedo, *btemp, *ttemp, *temis, *fisot, *phi0;
  const double* temf;  // (ncol, nlyr+1), level 0 = surface
  const double *wave_lower, *wave_upper;
  const double *utau, *umu, *phi;  // user grids (device); utau/umu unused without the flags
  const double *cmu, *cwt;         // nn quadrature nodes / weights
  double* flx;                     // (nsolve, ntau, 2), level order reversed (pyharp bottom-up)
  double* uu;                      // (nsolve, nphi, ntau, numu) or null when ONLYFL
  int32_t* status;
  double* ws;       // per-CTA workspaces
  long ws_per_cta;  // doubles
  long s0, nlocal;
};

// columns of at most this many layers run the per-layer phase CTA-cooperatively (one layer at a time) instead of one
// thread per layer
#ifndef B200_K3_COOP_MAX_LAYERS
#define B200_K3_COOP_MAX_LAYERS 16
#endif
constexpr int kCoopMaxLayers = B200_K3_COOP_MAX_LAYERS;

// ---- workspace map (offsets in doubles) -------------------------------------------------
struct IntensityWs {
  long hdr, tauc, taucpr, dtaucp, oprim, flyr, ssa, gl, expbea, utau, utaupr, layru, pkag, xr0, xr1;
  long ylmc, ylm0, ylmu, umu, gc, kk, ll, zz, zp0, zp1, gu, zbeam, z0u, z1u, cband, b, uum, red, flux5;
  long thread_scratch, per_thread, total;
  long shared_scr;  // (coop) the layer matrices of the cooperative per-layer phase
  int coop;
};

B200_FN_ATTR inline IntensityWs intensity_ws_map(int nlyr, int nstr, int ntau, int numu, int nthr) {
  IntensityWs m;
  const long L = nlyr, N = nstr, nn = nstr / 2, T = ntau, U = numu;
  long o = 0;
  auto take = [&](long n) { long r = o; o += (n + 1) & ~1L; return r; };
  m.hdr = take(16);
  m.tauc = take(L + 1); m.taucpr = take(L + 1); m.dtaucp = take(L); m.oprim = take(L);
  m.flyr = take(L); m.ssa = take(L); m.gl = take(L * N); m.expbea = take(L + 1);
  m.utau = take(T); m.utaupr = take(T); m.layru = take(T); m.pkag = take(L + 1);
  m.xr0 = take(L); m.xr1 = take(L);
  m.ylmc = take(N * N); m.ylm0 = take(N); m.ylmu = take(N * U); m.umu = take(U);
  m.gc = take(L * N * N); m.kk = take(L * N); m.ll = take(L * N); m.zz = take(L * N);
  m.zp0 = take(L * N); m.zp1 = take(L * N); m.gu = take(L * U * N); m.zbeam = take(L * U);
  m.z0u = take(L * U); m.z1u = take(L * U);
  m.cband = take(N * L * (9 * nn - 2)); m.b = take(N * L);
  m.uum = take(T * (U > N ? U : N)); m.red = take(nthr); m.flux5 = take(5 * T);
  const long layer_scr = 6 * nn * nn + 2 * N * N + 8 * N + 8;
  const long corr_scr = 3 * L + 8;
  // columns of at most kCoopMaxLayers layers run the per-layer phase CTA-cooperatively: ONE set of layer matrices shared
  // by the CTA plus 2N+2 doubles per thread, instead of a full set per thread (64 x 8 KB at nstr = 16).  (The whole
  // workspace of a single-layer solve is then 44 KB; keeping it in shared memory instead of global memory was measured
  // and changes nothing -- 0.715 against 0.693 ms per call: the L1 serves the re-reads, the barriers are the cost)
  m.coop = (L <= kCoopMaxLayers) ? 1 : 0;
  if (m.coop) {
    m.shared_scr = take(layer_scr);
    const long small = 2 * N + 2;
    m.per_thread = ((small > corr_scr ? small : corr_scr) + 1) & ~1L;
  } else {
    m.shared_scr = 0;
    m.per_thread = ((layer_scr > corr_scr ? layer_scr : corr_scr) + 1) & ~1L;
  }
  m.thread_scratch = take(m.per_thread * nthr);
  m.total = o;
  return m;
}

// header slots
enum { H_STATUS = 0, H_NCUT, H_LYRCUT, H_NAZ, H_KCONV, H_DONE, H_TPLANK, H_BPLANK, H_SING };

// normalised associated Legendre functions Y_l^m(mu), l = m..lmax (c_legendre_poly); ylm[l*stride + col]
B200_FN_ATTR inline void lepoly_one(int m, int lmax, double mu, double* ylm, int stride, int col) {
  if (m == 0) {
    ylm[col] = 1.0;
    if (lmax >= 1) ylm[stride + col] = mu;
    for (int l = 2; l <= lmax; ++l)
      ylm[l * stride + col] = ((2 * l - 1) * mu * ylm[(l - 1) * stride + col] - (l - 1) * ylm[(l - 2) * stride + col]) / l;
  } else {
    const double tmp1 = 1.0 - mu * mu;
    double ymm = 1.0;
    for (int mp = 1; mp <= m; ++mp) ymm = -sqrt((double)(2 * mp - 1)) / sqrt((double)(2 * mp)) * sqrt(tmp1) * ymm;
    if (m <= lmax) ylm[m * stride + col] = ymm;
    if (m + 1 <= lmax) ylm[(m + 1) * stride + col] = sqrt((double)(2 * m + 1)) * mu * ymm;
    for (int l = m + 2; l <= lmax; ++l) {
      const double t1 = sqrt((double)(l - m)) * sqrt((double)(l + m));
      const double t2 = sqrt((double)(l - m - 1)) * sqrt((double)(l + m - 1));
      ylm[l * stride + col] = ((2 * l - 1) * mu * ylm[(l - 1) * stride + col] - t2 * ylm[(l - 2) * stride + col]) / t1;
    }
  }
}

// dense LU with partial pivoting (n x n, row-major) and one solve; returns false when singular
B200_FN_ATTR inline bool lu_factor(double* a, int n, int* ipvt) {
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double amax = fabs(a[k * n + k]);
    for (int r = k + 1; r < n; ++r)
      if (fabs(a[r * n + k]) > amax) { amax = fabs(a[r * n + k]); p = r; }
    ipvt[k] = p;
    if (amax == 0.0) { ok = false; continue; }
    if (p != k)
      for (int c = k; c < n; ++c) {  // LINPACK style: earlier multipliers stay put, lu_solve interleaves the swaps
        double t = a[k * n + c]; a[k * n + c] = a[p * n + c]; a[p * n + c] = t;
      }
    const double rp = 1.0 / a[k * n + k];
    for (int r = k + 1; r < n; ++r) {
      const double f = a[r * n + k] * rp;
      a[r * n + k] = f;
      for (int c = k + 1; c < n; ++c) a[r * n + c] -= f * a[k * n + c];
    }
  }
  return ok;
}

// the same factorisation with the row updates below the pivot spread over the CTA (COOP; every thread must call it).
// Thread 0 searches the pivot and swaps the rows, a barrier publishes them, the rows r > k are updated by different
// threads, a second barrier closes the step.  !COOP: the statements of lu_factor on one thread.
template <bool COOP>
B200_FN_ATTR inline bool lu_factor_team(double* a, int n, int* ipvt, int ltid, int lnthr) {
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    if (ltid == 0) {
      int p = k;
      double amax = fabs(a[k * n + k]);
      for (int r = k + 1; r < n; ++r)
        if (fabs(a[r * n + k]) > amax) { amax = fabs(a[r * n + k]); p = r; }
      ipvt[k] = p;
      if (amax != 0.0 && p != k)
        for (int c = k; c < n; ++c) {
          double t = a[k * n + c]; a[k * n + c] = a[p * n + c]; a[p * n + c] = t;
        }
    }
    if (COOP) { B200_SYNC(); }
    const double piv = a[k * n + k];
    if (piv == 0.0) { ok = false; if (COOP) { B200_SYNC(); } continue; }  // (the whole column below the diagonal is zero)
    const double rp = 1.0 / piv;
    for (int r = k + 1 + ltid; r < n; r += lnthr) {
      const double f = a[r * n + k] * rp;
      a[r * n + k] = f;
      for (int c = k + 1; c < n; ++c) a[r * n + c] -= f * a[k * n + c];
    }
    if (COOP) { B200_SYNC(); }
  }
  return ok;
}

B200_FN_ATTR inline void lu_solve(const double* a, int n, const int* ipvt, double* b) {
  for (int k = 0; k < n; ++k) {
    const int p = ipvt[k];
    if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; }
    for (int r = k + 1; r < n; ++r) b[r] -= a[r * n + k] * b[k];
  }
  for (int k = n - 1; k >= 0; --k) {
    double x = b[k];
    for (int c = k + 1; c < n; ++c) x -= a[k * n + c] * b[c];
    b[k] = x / a[k * n + k];
  }
}

// cyclic Jacobi on a symmetric positive-definite n x n matrix (full storage); eigenvectors in the
// columns of y; relative off-diagonal criterion as in the flux kernel.  Returns false on no convergence.
B200_FN_ATTR inline bool jacobi_dyn(double* g, double* y, int n) {
  const double eps2 = 1.0e-31;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) y[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep <= 30; ++sweep) {
    bool conv = true;
    for (int p = 0; p < n - 1 && conv; ++p)
      for (int q = p + 1; q < n; ++q)
        if (g[q * n + p] * g[q * n + p] > eps2 * fabs(g[p * n + p] * g[q * n + q])) { conv = false; break; }
    if (conv) return true;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double gpq = g[q * n + p], gpp = g[p * n + p], gqq = g[q * n + q];
        if (!(gpq * gpq > eps2 * fabs(gpp * gqq))) continue;
        const double d = gqq - gpp, b2 = 2.0 * gpq;
        const double r = sqrt(d * d + b2 * b2);
        const double t = b2 / (d + copysign(r, d));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c, tau = s / (1.0 + c);
        g[p * n + p] = gpp - t * gpq;
        g[q * n + q] = gqq + t * gpq;
        g[q * n + p] = g[p * n + q] = 0.0;
        for (int r2 = 0; r2 < n; ++r2) {
          if (r2 != p && r2 != q) {
            const double a = g[r2 * n + p], b = g[r2 * n + q];
            const double na = a - s * (b + tau * a), nb = b + s * (a - tau * b);
            g[r2 * n + p] = g[p * n + r2] = na;
            g[r2 * n + q] = g[q * n + r2] = nb;
          }
          const double a = y[r2 * n + p], b = y[r2 * n + q];
          y[r2 * n + p] = a - s * (b + tau * a);
          y[r2 * n + q] = b + s * (a - tau * b);
        }
      }
  }
  return false;
}

// the same iteration with each rotation's row / column updates spread over the CTA (COOP; every thread must call it): all
// threads read the 2 x 2 block and form the rotation redundantly, a barrier, the updates of the n rows by different
// threads (thread 0 the 2 x 2 block itself), a barrier.  Same statements per element as jacobi_dyn.
template <bool COOP>
B200_FN_ATTR inline bool jacobi_team(double* g, double* y, int n, int ltid, int lnthr) {
  const double eps2 = 1.0e-31;
  for (int idx = ltid; idx < n * n; idx += lnthr) y[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
  if (COOP) { B200_SYNC(); }
  for (int sweep = 0; sweep <= 30; ++sweep) {
    bool conv = true;
    for (int p = 0; p < n - 1 && conv; ++p)
      for (int q = p + 1; q < n; ++q)
        if (g[q * n + p] * g[q * n + p] > eps2 * fabs(g[p * n + p] * g[q * n + q])) { conv = false; break; }
    if (conv) return true;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double gpq = g[q * n + p], gpp = g[p * n + p], gqq = g[q * n + q];
        if (!(gpq * gpq > eps2 * fabs(gpp * gqq))) continue;  // (uniform: every thread reads the same values)
        const double d = gqq - gpp, b2 = 2.0 * gpq;
        const double r = sqrt(d * d + b2 * b2);
        const double t = b2 / (d + copysign(r, d));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c, tau = s / (1.0 + c);
        if (COOP) { B200_SYNC(); }  // everybody has read the block
        if (ltid == 0) {
          g[p * n + p] = gpp - t * gpq;
          g[q * n + q] = gqq + t * gpq;
          g[q * n + p] = g[p * n + q] = 0.0;
        }
        for (int r2 = ltid; r2 < n; r2 += lnthr) {
          if (r2 != p && r2 != q) {
            const double a = g[r2 * n + p], b = g[r2 * n + q];
            const double na = a - s * (b + tau * a), nb = b + s * (a - tau * b);
            g[r2 * n + p] = g[p * n + r2] = na;
            g[r2 * n + q] = g[q * n + r2] = nb;
          }
          const double a = y[r2 * n + p], b = y[r2 * n + q];
          y[r2 * n + p] = a - s * (b + tau * a);
          y[r2 * n + q] = b + s * (a - tau * b);
        }
        if (COOP) { B200_SYNC(); }
      }
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// One layer, one azimuth mode: eigen-solution (c_sol_eigen), beam (c_upbeam) and thermal
// (c_upisot) particular solutions, interpolation to user angles (c_terpev / c_terpso).
// Returns a status code.
// ------------------------------------------------------------------------------------------
//
// Two execution shapes of the same statements (template COOP):
//   COOP = false  one thread owns the layer (the CTA's threads take different layers): ltid = 0, lnthr = 1, no barriers;
//   COOP = true   the whole CTA works on ONE layer -- for columns with fewer layers than threads (tests/test_disort.cpp
//                 has a single layer: its per-layer phase ran on one thread while 63 waited, 2.1 ms per 10-solve call).
//                 Loops over independent outputs are spread over the CTA (LFOR) -- including the row updates of the
//                 pivoted LU of c_upbeam / c_upisot and of every Jacobi rotation --, the 8 x 8 Cholesky and the
//                 triangular solves stay on one thread, CTA barriers separate the phases.  `scr` holds the layer's matrices (shared by the CTA when COOP), `prv` this thread's own
//                 vectors.  Every thread must call it with the same arguments.
#define LFOR(i, n) for (int i = ltid; i < (n); i += lnthr)
#define LSYNC() do { if (COOP) { B200_SYNC(); } } while (0)
template <bool COOP>
B200_FN_ATTR inline int intensity_layer(const IntensityParams& P, const IntensityWs& M, double* w, double* scr, double* prv, int lc,
                                        int mazim, bool planck, bool user_angles, double fbeam, double umu0, int ltid, int lnthr) {
  const int N = P.nstr, nn = N / 2, U = P.numu;
  const double* gl = w + M.gl + (long)lc * N;
  const double* ylmc = w + M.ylmc;
  const double* ylm0 = w + M.ylm0;
  const double* ylmu = w + M.ylmu;
  double* gc = w + M.gc + (long)lc * N * N;
  double* kk = w + M.kk + (long)lc * N;
  double* Pt = scr; double* Lq = Pt + nn * nn; double* G = Lq + nn * nn; double* Y = G + nn * nn;
  double* X = Y + nn * nn; double* Z = X + nn * nn;
  double* cc = Z + nn * nn; double* arr = cc + N * N;
  double* zj = arr + N * N; double* z0 = zj + N; double* z1 = z0 + N; double* wks = z1 + N;  // wks: 2N+2, shared
  int* ipvt = reinterpret_cast<int*>(wks + 2 * N + 2);                                       // N ints <= N doubles
  double* wk = COOP ? prv : wks;  // this thread's own 2N+2 doubles (COOP: its private scratch; else the same storage)
  const double delm0 = (mazim == 0) ? 1.0 : 0.0;
  int st = ST_OK;

  // symmetric reduced matrices: Pt = diag(1/mu) - sum_{l-m odd} gl Yh Yh^T, Qt likewise over even l-m
  LFOR(idx, nn * nn) {
    const int i = idx / nn, k = idx - i * nn;
    if (k > i) continue;
    double so = 0.0, se = 0.0;
    for (int l = mazim; l < N; ++l) {
      const double v = gl[l] * ylmc[l * N + i] * ylmc[l * N + k];
      if ((l - mazim) & 1) so += v; else se += v;
    }
    const double e = sqrt(P.cwt[i] / P.cmu[i]) * sqrt(P.cwt[k] / P.cmu[k]);
    const double dg = (i == k) ? 1.0 / P.cmu[i] : 0.0;
    Pt[i * nn + k] = Pt[k * nn + i] = dg - e * so;
    Lq[i * nn + k] = Lq[k * nn + i] = dg - e * se;
  }
  LSYNC();
  // Cholesky Qt = L L^T (lower, in place)
  if (ltid == 0) {
    for (int c = 0; c < nn; ++c) {
      double d = Lq[c * nn + c];
      for (int k = 0; k < c; ++k) d -= Lq[c * nn + k] * Lq[c * nn + k];
      if (!(d > 0.0)) { st = ST_SINGULAR; d = 1.0; }
      const double r = sqrt(d);
      Lq[c * nn + c] = r;
      for (int i = c + 1; i < nn; ++i) {
        double v = Lq[i * nn + c];
        for (int k = 0; k < c; ++k) v -= Lq[i * nn + k] * Lq[c * nn + k];
        Lq[i * nn + c] = v / r;
      }
      for (int i = 0; i < c; ++i) Lq[i * nn + c] = 0.0;
    }
  }
  LSYNC();
  // G = L^T Pt L  (X used as temporary Pt L)
  LFOR(idx, nn * nn) {
    const int i = idx / nn, c = idx - i * nn;
    double v = 0.0;
    for (int k = c; k < nn; ++k) v += Pt[i * nn + k] * Lq[k * nn + c];
    X[i * nn + c] = v;
  }
  LSYNC();
  LFOR(idx, nn * nn) {
    const int i = idx / nn, c = idx - i * nn;
    if (c > i) continue;
    double v = 0.0;
    for (int k = i; k < nn; ++k) v += Lq[k * nn + i] * X[k * nn + c];
    G[i * nn + c] = G[c * nn + i] = v;
  }
  LSYNC();
  if (!jacobi_team<COOP>(G, Y, nn, ltid, lnthr) && st == ST_OK) st = ST_EIGEN_FAIL;
  LSYNC();
  // X = L^-T Y, Z = L Y (columns = eigenmodes); s = T^-1 X, d = -T^-1 Z / k with T = sqrt(mu w)
  LFOR(c, nn) {
    for (int i = nn - 1; i >= 0; --i) {
      double v = Y[i * nn + c];
      for (int k = i + 1; k < nn; ++k) v -= Lq[k * nn + i] * X[k * nn + c];
      X[i * nn + c] = v / Lq[i * nn + i];
    }
    for (int i = 0; i < nn; ++i) {
      double v = 0.0;
      for (int k = 0; k <= i; ++k) v += Lq[i * nn + k] * Y[k * nn + c];
      Z[i * nn + c] = v;
    }
  }
  LSYNC();
  LFOR(idx, nn * nn) {
    const int jq = idx / nn, iq = idx - jq * nn;
    double lam = G[jq * nn + jq];
    if (!(lam > 0.0)) { if (st == ST_OK) st = ST_EIGEN_FAIL; lam = 1.0; }
    const double k = sqrt(lam);
    if (iq == 0) {
      kk[jq + nn] = k;
      kk[nn - 1 - jq] = -k;
    }
    const double rt = 1.0 / sqrt(P.cmu[iq] * P.cwt[iq]);
    const double s = rt * X[iq * nn + jq], d = -rt * Z[iq * nn + jq] / k;
    const double up = 0.5 * (s + d), dn = 0.5 * (s - d);
    gc[(iq + nn) * N + jq + nn] = up;
    gc[(nn - 1 - iq) * N + jq + nn] = dn;
    gc[(iq + nn) * N + nn - 1 - jq] = dn;
    gc[(nn - 1 - iq) * N + nn - 1 - jq] = up;
  }
  // full scattering kernel in CMU order (first nn = +mu, then -mu): C(mu_i, mu_j)
  LFOR(idx, nn * nn) {
    const int iq = idx / nn, jq = idx - iq * nn;
    double sp = 0.0, sm = 0.0, sg = 1.0;
    for (int l = mazim; l < N; ++l) {
      const double v = gl[l] * ylmc[l * N + iq] * ylmc[l * N + jq];
      sp += v;
      sm += sg * v;
      sg = -sg;
    }
    const double cp = 0.5 * sp * P.cwt[jq], cm = 0.5 * sm * P.cwt[jq];
    cc[iq * N + jq] = cp; cc[iq * N + jq + nn] = cm;
    cc[(iq + nn) * N + jq] = cm; cc[(iq + nn) * N + jq + nn] = cp;
  }
  LSYNC();
  auto cmu_of = [&](int iq) { return iq < nn ? P.cmu[iq] : -P.cmu[iq - nn]; };
  auto ylmc_of = [&](int l, int iq) {
    return iq < nn ? ylmc[l * N + iq] : (((l - mazim) & 1) ? -ylmc[l * N + iq - nn] : ylmc[l * N + iq - nn]);
  };
  if (fbeam > 0.0) {  // c_upbeam
    double* zz = w + M.zz + (long)lc * N;
    LFOR(iq, N) {
      for (int jq = 0; jq < N; ++jq) arr[iq * N + jq] = -cc[iq * N + jq];
      arr[iq * N + iq] += 1.0 + cmu_of(iq) / umu0;
      double sum = 0.0;
      for (int l = mazim; l < N; ++l) sum += gl[l] * ylmc_of(l, iq) * ylm0[l];
      zj[iq] = (2.0 - delm0) * fbeam * sum / (4.0 * kPi);
    }
    LSYNC();
    if (!lu_factor_team<COOP>(arr, N, ipvt, ltid, lnthr) && st == ST_OK) st = ST_SINGULAR;
    if (ltid == 0) {
      lu_solve(arr, N, ipvt, zj);
      for (int iq = 0; iq < nn; ++iq) { zz[iq + nn] = zj[iq]; zz[nn - 1 - iq] = zj[iq + nn]; }
    }
    LSYNC();
  }
  if (planck && mazim == 0) {  // c_upisot
    double* zp0 = w + M.zp0 + (long)lc * N;
    double* zp1 = w + M.zp1 + (long)lc * N;
    const double oprim = w[M.oprim + lc], xr0 = w[M.xr0 + lc], xr1 = w[M.xr1 + lc];
    LFOR(iq, N) {
      for (int jq = 0; jq < N; ++jq) arr[iq * N + jq] = -cc[iq * N + jq];
      arr[iq * N + iq] += 1.0;
      z1[iq] = (1.0 - oprim) * xr1;
    }
    LSYNC();
    if (!lu_factor_team<COOP>(arr, N, ipvt, ltid, lnthr) && st == ST_OK) st = ST_SINGULAR;
    if (ltid == 0) {
      lu_solve(arr, N, ipvt, z1);
      for (int iq = 0; iq < N; ++iq) z0[iq] = (1.0 - oprim) * xr0 + cmu_of(iq) * z1[iq];
      lu_solve(arr, N, ipvt, z0);
      for (int iq = 0; iq < nn; ++iq) {
        zp0[iq + nn] = z0[iq]; zp1[iq + nn] = z1[iq];
        zp0[nn - 1 - iq] = z0[iq + nn]; zp1[nn - 1 - iq] = z1[iq + nn];
      }
    }
    LSYNC();
  }
  if (user_angles) {  // c_terpev + c_terpso
    double* gu = w + M.gu + (long)lc * U * N;
    // row r of gc: r < nn -> -cmu[nn-1-r], else +cmu[r-nn]; weight and Y_l^m at that angle
    auto wrow = [&](int r) { return r < nn ? P.cwt[nn - 1 - r] : P.cwt[r - nn]; };
    auto yrow = [&](int l, int r) {
      return r < nn ? (((l - mazim) & 1) ? -ylmc[l * N + nn - 1 - r] : ylmc[l * N + nn - 1 - r]) : ylmc[l * N + r - nn];
    };
    LFOR(jq, N) {  // (eigenvectors are independent: wk is this thread's own)
      for (int l = mazim; l < N; ++l) {
        double sum = 0.0;
        for (int r = 0; r < N; ++r) sum += wrow(r) * yrow(l, r) * gc[r * N + jq];
        wk[l] = 0.5 * gl[l] * sum;
      }
      for (int iu = 0; iu < U; ++iu) {
        double sum = 0.0;
        for (int l = mazim; l < N; ++l) sum += wk[l] * ylmu[l * U + iu];
        gu[iu * N + jq] = sum;
      }
    }
    if (fbeam > 0.0) {
      const double* zz = w + M.zz + (long)lc * N;
      const double fact = (2.0 - delm0) * fbeam / (4.0 * kPi);
      LFOR(l, N) {
        if (l < mazim) continue;
        double ps = 0.0;
        for (int r = 0; r < N; ++r) ps += wrow(r) * yrow(l, r) * zz[r];
        wks[l] = 0.5 * gl[l] * ps;
      }
      LSYNC();
      LFOR(iu, U) {
        double sum = 0.0;
        for (int l = mazim; l < N; ++l) sum += ylmu[l * U + iu] * (wks[l] + fact * gl[l] * ylm0[l]);
        w[M.zbeam + (long)lc * U + iu] = sum;
      }
      LSYNC();
    }
    if (planck && mazim == 0) {
      const double* zp0 = w + M.zp0 + (long)lc * N;
      const double* zp1 = w + M.zp1 + (long)lc * N;
      double* psi0 = wks; double* psi1 = wks + N;
      LFOR(l, N) {
        double p0 = 0.0, p1 = 0.0;
        for (int r = 0; r < N; ++r) {
          const double c = wrow(r) * yrow(l, r);
          p0 += c * zp0[r];
          p1 += c * zp1[r];
        }
        psi0[l] = 0.5 * gl[l] * p0;
        psi1[l] = 0.5 * gl[l] * p1;
      }
      LSYNC();
      const double om = 1.0 - w[M.oprim + lc];
      LFOR(iu, U) {
        double s0 = 0.0, s1 = 0.0;
        for (int l = 0; l < N; ++l) { s0 += ylmu[l * U + iu] * psi0[l]; s1 += ylmu[l * U + iu] * psi1[l]; }
        w[M.z0u + (long)lc * U + iu] = s0 + om * w[M.xr0 + lc];
        w[M.z1u + (long)lc * U + iu] = s1 + om * w[M.xr1 + lc];
      }
    }
  }
  LSYNC();  // (COOP: the next layer re-uses the shared matrices)
  return st;
}
#undef LFOR
#undef LSYNC

// c_xi_func / c_single_scat / c_secondary_scat (Nakajima-Tanaka TMS/IMS)
B200_FN_ATTR inline double nt_xifunc(double u1, double u2, double u3, double tau) {
  const double x1 = 1.0 / u1 - 1.0 / u2, x2 = 1.0 / u1 - 1.0 / u3, e1 = exp(-tau / u1);
  if (u2 == u3 && u1 == u2) return tau * tau * e1 / (2.0 * u1 * u2);
  if (u2 == u3 && u1 != u2) return ((tau - 1.0 / x1) * exp(-tau / u2) + e1 / x1) / (x1 * u1 * u2);
  if (u2 != u3 && u1 == u2) return ((exp(-tau / u3) - e1) / x2 - tau * e1) / (x2 * u1 * u2);
  if (u2 != u3 && u1 == u3) return ((exp(-tau / u2) - e1) / x1 - tau * e1) / (x1 * u1 * u2);
  return ((exp(-tau / u3) - e1) / x2 - (exp(-tau / u2) - e1) / x1) / (x2 - x1) / (u1 * u2);
}

B200_FN_ATTR inline double nt_sinsca(double dither, int layru, int nlyr, const double* phase, const double* omega,
                                     const double* tau, double umu, double umu0, double utau, double fbeam) {
  double s = 0.0, exp0 = exp(-utau / umu0);
  if (fabs(umu + umu0) <= dither) {
    for (int l = 0; l < layru; ++l) s += omega[l] * phase[l] * (tau[l + 1] - tau[l]);
    return fbeam / (4.0 * kPi * umu0) * exp0 * (s + omega[layru] * phase[layru] * (utau - tau[layru]));
  }
  if (umu > 0.0) {
    for (int l = layru; l < nlyr; ++l) {
      const double e1 = exp(-((tau[l + 1] - utau) / umu + tau[l + 1] / umu0));
      s += omega[l] * phase[l] * (exp0 - e1);
      exp0 = e1;
    }
  } else {
    for (int l = layru; l >= 0; --l) {
      const double e1 = exp(-((tau[l] - utau) / umu + tau[l] / umu0));
      s += omega[l] * phase[l] * (exp0 - e1);
      exp0 = e1;
    }
  }
  return fbeam / (4.0 * kPi * (1.0 + umu / umu0)) * s;
}

// pmom access: moment k of DISORT layer lc straight from pyharp's prop tensor
B200_FN_ATTR inline double prop_pmom(const double* pr, int nlyr, int nprop, int lc, int k) {
  if (k == 0) return 1.0;
  return (k + 1 < nprop) ? pr[(long)(nlyr - 1 - lc) * nprop + k + 1] : 0.0;
}

// ------------------------------------------------------------------------------------------
// One solve by the calling thread group (tid in [0, nthr)); `w` is the group's workspace.
// ------------------------------------------------------------------------------------------
B200_FN_ATTR inline void solve_intensity_one(const IntensityParams& P, long s, double* w, int tid, int nthr) {
  const int L = P.nlyr, N = P.nstr, nn = N / 2, T = P.ntau, U = P.numu, nprop = P.nprop;
  const IntensityWs M = intensity_ws_map(L, N, T, U, nthr);
  const bool usrtau = P.flags & F_USRTAU, planck = P.flags & F_PLANCK, onlyfl = P.flags & F_ONLYFL;
  const bool user_angles = (P.flags & F_USRANG) && !onlyfl;
  const long sg = s + P.s0;
  const int n = (int)(sg / P.ncol), j = (int)(sg - (long)n * P.ncol);
  const double umu0 = P.umu0 ? P.umu0[s] : 1.0, fbeam = P.fbeam ? P.fbeam[s] : 0.0;
  const double albedo = P.albedo ? P.albedo[s] : 0.0, fisot = P.fisot ? P.fisot[s] : 0.0;
  const double phi0 = P.phi0 ? P.phi0[s] : 0.0;
  const double* pr = P.prop + (long)s * L * nprop;
  double* hdr = w + M.hdr;
  double* scr = w + M.thread_scratch + (long)tid * M.per_thread;
  double* flx = P.flx + (long)s * T * 2;
  double* uu = P.uu ? P.uu + (long)s * P.nphi * T * U : nullptr;
  const int nphi = onlyfl ? 0 : P.nphi;

  // ---- c_check_inputs + c_disort_set (serial: O(L N)) ----
  if (tid == 0) {
    int st = ST_OK;
    if (!(fbeam >= 0.0) || !(fisot >= 0.0) || !(albedo >= 0.0 && albedo <= 1.0)) st = ST_BAD_INPUT;
    if (fbeam > 0.0 && !(umu0 > 0.0 && umu0 <= 1.0)) st = ST_BAD_INPUT;
    if (fbeam > 0.0 && st == ST_OK)
      for (int iq = 0; iq < nn; ++iq)
        if (fabs((umu0 - P.cmu[iq]) / umu0) < 1.0e-4) st = ST_BEAM_ANGLE;
    double tplank = 0.0, bplank = 0.0;
    if (planck) {
      const double wlo = P.wave_lower ? P.wave_lower[n] : 0.0, whi = P.wave_upper ? P.wave_upper[n] : 0.0;
      const double btemp = P.btemp ? P.btemp[s] : 0.0, ttemp = P.ttemp ? P.ttemp[s] : 0.0;
      const double temis = P.temis ? P.temis[s] : 0.0;
      if (!(btemp >= 0.0) || !(ttemp >= 0.0) || !(temis >= 0.0 && temis <= 1.0) || !(wlo >= 0.0) || !(whi >= wlo))
        st = ST_BAD_INPUT;
      else {
        tplank = temis * plkavg(wlo, whi, ttemp);
        bplank = plkavg(wlo, whi, btemp);
        for (int lev = 0; lev <= L; ++lev) {
          const double t = P.temf[(long)j * (L + 1) + (L - lev)];
          if (!(t >= 0.0)) st = ST_BAD_INPUT;
          w[M.pkag + lev] = (t >= 0.0) ? plkavg(wlo, whi, t) : 0.0;
        }
      }
    }
    hdr[H_TPLANK] = tplank;
    hdr[H_BPLANK] = bplank;
    double abstau = 0.0;
    int ncut = L;
    w[M.tauc] = 0.0;
    w[M.taucpr] = 0.0;
    w[M.expbea] = (fbeam > 0.0) ? 1.0 : 0.0;
    for (int lc = 0; lc < L; ++lc) {
      const double* q = pr + (long)(L - 1 - lc) * nprop;
      double ssa = (nprop > 1) ? q[1] : 0.0;
      const double dtau = q[0];
      if (!(dtau >= 0.0) || !(ssa >= 0.0 && ssa <= 1.0)) st = ST_BAD_INPUT;
      if (ssa == 1.0) ssa = 1.0 - 1.0e-10;  // DISORT's dither of exactly conservative layers
      for (int k = 1; k <= P.nmom; ++k) {
        const double pm = prop_pmom(pr, L, nprop, lc, k);
        if (!(pm >= -1.0 && pm <= 1.0)) st = ST_BAD_INPUT;
      }
      if (abstau < 10.0) ncut = lc + 1;
      abstau += (1.0 - ssa) * dtau;
      const double f = prop_pmom(pr, L, nprop, lc, N);
      w[M.tauc + lc + 1] = w[M.tauc + lc] + dtau;
      w[M.ssa + lc] = ssa;
      w[M.flyr + lc] = f;
      const double oprim = ssa * (1.0 - f) / (1.0 - f * ssa);
      w[M.oprim + lc] = oprim;
      w[M.dtaucp + lc] = (1.0 - f * ssa) * dtau;
      w[M.taucpr + lc + 1] = w[M.taucpr + lc] + w[M.dtaucp + lc];
      for (int k = 0; k < N; ++k)
        w[M.gl + (long)lc * N + k] = (2 * k + 1) * oprim * (prop_pmom(pr, L, nprop, lc, k) - f) / (1.0 - f);
      w[M.expbea + lc + 1] = (fbeam > 0.0) ? exp(-w[M.taucpr + lc + 1] / umu0) : 0.0;
    }
    const bool lyrcut = (abstau >= 10.0 && !planck && L > 1);
    if (!lyrcut) ncut = L;
    for (int lu = 0; lu < T; ++lu) {
      const double ut = usrtau ? P.utau[lu] : w[M.tauc + lu];
      if (usrtau && (ut < 0.0 || ut > w[M.tauc + L] * (1.0 + 1e-12) + 1e-300)) st = ST_BAD_INPUT;
      int lc;
      for (lc = 0; lc < L - 1; ++lc)
        if (ut >= w[M.tauc + lc] && ut <= w[M.tauc + lc + 1]) break;
      w[M.utau + lu] = ut;
      w[M.layru + lu] = (double)lc;
      w[M.utaupr + lu] = w[M.taucpr + lc] + (1.0 - w[M.ssa + lc] * w[M.flyr + lc]) * (ut - w[M.tauc + lc]);
    }
    if (user_angles) {
      for (int iu = 0; iu < U; ++iu) {
        w[M.umu + iu] = P.umu[iu];
        if (!(P.umu[iu] >= -1.0 && P.umu[iu] <= 1.0) || P.umu[iu] == 0.0) st = ST_BAD_INPUT;
      }
    } else {
      for (int iu = 0; iu < nn; ++iu) { w[M.umu + iu] = -P.cmu[nn - 1 - iu]; w[M.umu + iu + nn] = P.cmu[iu]; }
    }
    if (planck)
      for (int lc = 0; lc < L; ++lc) {
        const double dt = w[M.dtaucp + lc];
        const double xr1 = (dt > 0.0) ? (w[M.pkag + lc + 1] - w[M.pkag + lc]) / dt : 0.0;
        w[M.xr1 + lc] = xr1;
        w[M.xr0 + lc] = w[M.pkag + lc] - xr1 * w[M.taucpr + lc];
      }
    int naz = N - 1;
    const double* um = w + M.umu;
    if (fbeam == 0.0 || fabs(1.0 - umu0) < 1.0e-5 || onlyfl || (U == 1 && fabs(1.0 - um[0]) < 1.0e-5) ||
        (U == 1 && fabs(1.0 + um[0]) < 1.0e-5) || (U == 2 && fabs(1.0 + um[0]) < 1.0e-5 && fabs(1.0 - um[1]) < 1.0e-5))
      naz = 0;
    // Azimuth modes above the highest non-zero (delta-M scaled) moment of the column contribute exactly nothing: with
    // gl[l] = 0 for all l >= m the scattering kernel of mode m vanishes, so do its beam source and -- m > 0, Lambertian
    // surface -- its boundary terms, and the homogeneous system has the zero solution.  cdisort walks such modes until its
    // convergence test has passed twice (isotropic scattering, tests/test_disort.cpp: modes 0, 1, 2); the sum is the same.
    if (naz > 0) {
      int lmax = 0;
      for (int lc = 0; lc < L; ++lc)
        for (int k = 1; k < N; ++k)
          if (w[M.gl + (long)lc * N + k] != 0.0 && k > lmax) lmax = k;
      if (lmax < naz) naz = lmax;
    }
    if ((P.flags & F_INTCOR) && !(P.flags & F_OLDINTCOR) && !onlyfl && fbeam > 0.0)
      for (int lc = 0; lc < L; ++lc)
        if (w[M.flyr + lc] != 0.0 && st == ST_OK) st = ST_UNSUPPORTED;  // Buras-Emde correction not on pyharp's path
    hdr[H_STATUS] = (double)st;
    hdr[H_NCUT] = (double)ncut;
    hdr[H_LYRCUT] = lyrcut ? 1.0 : 0.0;
    hdr[H_NAZ] = (double)naz;
    hdr[H_KCONV] = 0.0;
    hdr[H_DONE] = 0.0;
  }
  B200_SYNC();
  int st = (int)hdr[H_STATUS];
  const int ncut = (int)hdr[H_NCUT];
  const bool lyrcut = hdr[H_LYRCUT] != 0.0;
  const int naz = (int)hdr[H_NAZ];
  const double tplank = hdr[H_TPLANK], bplank = hdr[H_BPLANK];
  const double* cmu = P.cmu;
  const double* cwt = P.cwt;
  const int kl = 3 * nn - 1, ku = kl, bw = 2 * kl + ku + 1, ncol = N * ncut;
  double* ab = w + M.cband;
  double* b = w + M.b;
  double* ll = w + M.ll;
  const double* gcA = w + M.gc;
  const double* kkA = w + M.kk;
  const double* zzA = w + M.zz;
  const double* zp0A = w + M.zp0;
  const double* zp1A = w + M.zp1;
  const double* taucpr = w + M.taucpr;
  const double* dtaucp = w + M.dtaucp;
  const double* expbea = w + M.expbea;
  const double* utaupr = w + M.utaupr;
  const bool beam = fbeam > 0.0;
#define GCL(lc, i, jq) gcA[((long)(lc) * N + (i)) * N + (jq)]
#define KKL(lc, jq) kkA[(long)(lc) * N + (jq)]
#define ZZL(lc, i) zzA[(long)(lc) * N + (i)]
#define ZP0(lc, i) zp0A[(long)(lc) * N + (i)]
#define ZP1(lc, i) zp1A[(long)(lc) * N + (i)]
#define AB(i, jc) ab[(long)(i) * bw + ((jc) - (i) + kl)]

  for (int mazim = 0; mazim <= naz && st == ST_OK; ++mazim) {
    const double delm0 = (mazim == 0) ? 1.0 : 0.0;
    // ---- Legendre tables for this mode ----
    if (tid == 0 && beam) lepoly_one(mazim, N - 1, -umu0, w + M.ylm0, 1, 0);
    if (user_angles) { B200_PAR_FOR(iu, U) lepoly_one(mazim, N - 1, w[M.umu + iu], w + M.ylmu, U, iu); }
    { B200_PAR_FOR(iq, nn) lepoly_one(mazim, N - 1, cmu[iq], w + M.ylmc, N, iq); }
    B200_SYNC();
    // ---- per-layer eigen / particular solutions: layers are independent ----
    {
      int lst = ST_OK;
      if (M.coop) {  // few layers: the CTA works on one layer at a time (see intensity_layer)
        for (int lc = 0; lc < ncut; ++lc) {
          const int r = intensity_layer<true>(P, M, w, w + M.shared_scr, scr, lc, mazim, planck, user_angles, fbeam, umu0, tid, nthr);
          if (r != ST_OK && lst == ST_OK) lst = r;
        }
      } else {
        B200_PAR_FOR(lc, ncut) {
          const int r = intensity_layer<false>(P, M, w, scr, scr, lc, mazim, planck, user_angles, fbeam, umu0, 0, 1);
          if (r != ST_OK && lst == ST_OK) lst = r;
        }
      }
      w[M.red + tid] = (double)lst;
    }
    { B200_PAR_FOR(i, ncol * bw) ab[i] = 0.0; }
    B200_SYNC();
    for (int t = 0; t < nthr; ++t)
      if (w[M.red + t] != 0.0 && st == ST_OK) st = (int)w[M.red + t];
    if (st != ST_OK) break;
    // ---- c_set_matrix: every layer fills its own block of columns ----
    const bool noreflect = lyrcut || (mazim > 0);
    B200_PAR_FOR(lc, ncut) {
      for (int jq = 0; jq < N; ++jq) {
        const int col = lc * N + jq;
        const double ex = exp(-fabs(KKL(lc, jq)) * dtaucp[lc]);
        const double etop = (jq < nn) ? ex : 1.0, ebot = (jq < nn) ? 1.0 : ex;
        if (lc == 0) {
          for (int r = 0; r < nn; ++r) AB(r, col) = GCL(0, nn - 1 - r, jq) * etop;
        } else {
          const int row0 = nn + (lc - 1) * N;
          for (int iq = 0; iq < N; ++iq) AB(row0 + iq, col) = -GCL(lc, iq, jq) * etop;
        }
        if (lc < ncut - 1) {
          const int row0 = nn + lc * N;
          for (int iq = 0; iq < N; ++iq) AB(row0 + iq, col) = GCL(lc, iq, jq) * ebot;
        } else {
          const int row0 = nn + (ncut - 1) * N;
          for (int r = 0; r < nn; ++r) {
            double v = GCL(lc, nn + r, jq);
            if (!noreflect) {
              double sum = 0.0;
              for (int k = 0; k < nn; ++k) sum += cwt[k] * cmu[k] * albedo * GCL(lc, nn - 1 - k, jq);
              v -= (1.0 + delm0) * sum;
            }
            AB(row0 + r, col) = v * ebot;
          }
        }
      }
    }
    // ---- c_solve0 right-hand side ----
    if (mazim > 0) {
      B200_PAR_FOR(r, nn) {
        b[r] = -ZZL(0, nn - 1 - r);
        b[ncol - nn + r] = -ZZL(ncut - 1, nn + r) * expbea[ncut];
      }
      B200_PAR_FOR(i, (ncut - 1) * N) {
        const int lc = i / N, iq = i - lc * N;
        b[nn + i] = (ZZL(lc + 1, iq) - ZZL(lc, iq)) * expbea[lc + 1];
      }
    } else {
      const double taub = taucpr[ncut];
      B200_PAR_FOR(r, nn) {
        double v = fisot + tplank;
        if (beam) v -= ZZL(0, nn - 1 - r);
        if (planck) v -= ZP0(0, nn - 1 - r);
        b[r] = v;
        const int iq = nn + r;
        v = 0.0;
        if (beam) v -= ZZL(ncut - 1, iq) * expbea[ncut];
        if (planck) v -= ZP0(ncut - 1, iq) + ZP1(ncut - 1, iq) * taub;
        if (!lyrcut) {
          double sum = 0.0;
          for (int k = 0; k < nn; ++k) {
            double t = 0.0;
            if (beam) t += ZZL(ncut - 1, nn - 1 - k) * expbea[ncut];
            if (planck) t += ZP0(ncut - 1, nn - 1 - k) + ZP1(ncut - 1, nn - 1 - k) * taub;
            sum += cwt[k] * cmu[k] * albedo * t;
          }
          v += 2.0 * sum + (1.0 - albedo) * bplank;
          if (beam) v += albedo * umu0 * fbeam / kPi * expbea[ncut];
        }
        b[ncol - nn + r] = v;
      }
      B200_PAR_FOR(i, (ncut - 1) * N) {
        const int lc = i / N, iq = i - lc * N;
        double v = 0.0;
        if (beam) v += (ZZL(lc + 1, iq) - ZZL(lc, iq)) * expbea[lc + 1];
        if (planck) v += ZP0(lc + 1, iq) - ZP0(lc, iq) + (ZP1(lc + 1, iq) - ZP1(lc, iq)) * taucpr[lc + 1];
        b[nn + i] = v;
      }
    }
    B200_SYNC();
    // ---- banded elimination with partial pivoting (c_sgbco / c_sgbsl): rows below the pivot
    //      are updated by different threads ----
    {
      bool sing = false;
      for (int k = 0; k < ncol; ++k) {
        const int rmax = (k + kl < ncol - 1) ? k + kl : ncol - 1;
        if (tid == 0) {
          int p = k;
          double amax = fabs(AB(k, k));
          for (int r = k + 1; r <= rmax; ++r)
            if (fabs(AB(r, k)) > amax) { amax = fabs(AB(r, k)); p = r; }
          if (p != k) {
            const int cmax = (k + kl + ku < ncol - 1) ? k + kl + ku : ncol - 1;
            for (int c = k; c <= cmax; ++c) {
              // AB(p, c) is inside row p's stored band because c <= k + kl + ku <= p + kl + ku
              const double t = AB(k, c);
              AB(k, c) = AB(p, c);
              AB(p, c) = t;
            }
            const double t = b[k]; b[k] = b[p]; b[p] = t;
          }
          hdr[H_SING] = (amax == 0.0) ? 1.0 : 0.0;
        }
        B200_SYNC();
        if (hdr[H_SING] != 0.0) { sing = true; }
        else {
          const int cmax = (k + kl + ku < ncol - 1) ? k + kl + ku : ncol - 1;
          const double rp = 1.0 / AB(k, k);
          for (int r = k + 1 + tid; r <= rmax; r += nthr) {
            const double f = AB(r, k) * rp;
            if (f != 0.0) {
              for (int c = k + 1; c <= cmax; ++c) AB(r, c) -= f * AB(k, c);
              b[r] -= f * b[k];
            }
          }
        }
        B200_SYNC();
      }
      if (sing) { st = ST_SINGULAR; break; }
      if (tid == 0) {
        for (int k = ncol - 1; k >= 0; --k) {
          const int cmax = (k + kl + ku < ncol - 1) ? k + kl + ku : ncol - 1;
          double x = b[k];
          for (int c = k + 1; c <= cmax; ++c) x -= AB(k, c) * b[c];
          b[k] = x / AB(k, k);
        }
        for (int i = 0; i < ncol; ++i) ll[i] = b[i];
      }
      B200_SYNC();
    }
    // ---- c_fluxes (mode 0) ----
    if (mazim == 0) {
      double* f5 = w + M.flux5;
      B200_PAR_FOR(lu, T) {
        double rfldir = 0.0, rfldn = 0.0, flup = 0.0;
        const int lyu = (int)w[M.layru + lu];
        if (!(lyrcut && lyu >= ncut)) {
          double fact = 0.0, fldir = 0.0, rdir = 0.0;
          if (beam) {
            fact = exp(-utaupr[lu] / umu0);
            fldir = umu0 * (fbeam * fact);
            rdir = umu0 * fbeam * exp(-w[M.utau + lu] / umu0);
          }
          double sdn = 0.0, sup = 0.0;
          for (int iq = 0; iq < N; ++iq) {
            double zint = 0.0;
            for (int jq = 0; jq < nn; ++jq)
              zint += GCL(lyu, iq, jq) * ll[(long)lyu * N + jq] * exp(-KKL(lyu, jq) * (utaupr[lu] - taucpr[lyu + 1]));
            for (int jq = nn; jq < N; ++jq)
              zint += GCL(lyu, iq, jq) * ll[(long)lyu * N + jq] * exp(-KKL(lyu, jq) * (utaupr[lu] - taucpr[lyu]));
            double u = zint;
            if (beam) u += ZZL(lyu, iq) * fact;
            if (planck) u += ZP0(lyu, iq) + ZP1(lyu, iq) * utaupr[lu];
            if (iq < nn) sdn += cwt[nn - 1 - iq] * cmu[nn - 1 - iq] * u;
            else sup += cwt[iq - nn] * cmu[iq - nn] * u;
          }
          flup = 2.0 * kPi * sup;
          rfldir = rdir;
          rfldn = 2.0 * kPi * sdn + fldir - rdir;
        }
        f5[lu * 2] = flup;
        f5[lu * 2 + 1] = rfldir + rfldn;
      }
    }
    if (onlyfl) break;
    // ---- azimuthal intensity components ----
    double* uum = w + M.uum;
    if (user_angles) {  // c_user_intensities
      const double* guA = w + M.gu;
      const double* zbeamA = w + M.zbeam;
      const double* z0uA = w + M.z0u;
      const double* z1uA = w + M.z1u;
#define GUL(lc, iu, jq) (guA[((long)(lc) * U + (iu)) * N + (jq)] * ll[(long)(lc) * N + (jq)])
      B200_PAR_FOR(idx, T * U) {
        const int lu = idx / U, iu = idx - lu * U;
        const int lyu = (int)w[M.layru + lu];
        double res = 0.0;
        if (!(lyrcut && lyu >= ncut)) {
          const double exp0 = beam ? exp(-utaupr[lu] / umu0) : 0.0;
          const double mu = w[M.umu + iu];
          const bool neg = mu < 0.0;
          const int lyrstr = neg ? 0 : lyu + 1, lyrend = neg ? lyu - 1 : ncut - 1;
          const double sgn = neg ? -1.0 : 1.0;
          double palint = 0.0, plkint = 0.0, exp1 = 0.0, exp2 = 0.0;
          for (int lc = lyrstr; lc <= lyrend; ++lc) {
            const double dtau = dtaucp[lc];
            exp1 = exp((utaupr[lu] - taucpr[lc]) / mu);
            exp2 = exp((utaupr[lu] - taucpr[lc + 1]) / mu);
            if (planck && mazim == 0) {
              const double f0n = sgn * (exp1 - exp2);
              const double f1n = sgn * ((taucpr[lc] + mu) * exp1 - (taucpr[lc + 1] + mu) * exp2);
              plkint += z0uA[(long)lc * U + iu] * f0n + z1uA[(long)lc * U + iu] * f1n;
            }
            if (beam) {
              const double denom = 1.0 + mu / umu0;
              const double expn = (fabs(denom) < 0.0001) ? (dtau / umu0) * exp0
                                                         : (exp1 * expbea[lc] - exp2 * expbea[lc + 1]) * sgn / denom;
              palint += zbeamA[(long)lc * U + iu] * expn;
            }
            for (int iq = 0; iq < nn; ++iq) {
              const double wkq = exp(KKL(lc, iq) * dtau), denom = 1.0 + mu * KKL(lc, iq);
              const double expn = (fabs(denom) < 0.0001) ? dtau / mu * exp2 : sgn * (exp1 * wkq - exp2) / denom;
              palint += GUL(lc, iu, iq) * expn;
            }
            for (int iq = nn; iq < N; ++iq) {
              const double wkq = exp(KKL(lc, N - 1 - iq) * dtau), denom = 1.0 + mu * KKL(lc, iq);
              const double expn = (fabs(denom) < 0.0001) ? -dtau / mu * exp1 : sgn * (exp1 - exp2 * wkq) / denom;
              palint += GUL(lc, iu, iq) * expn;
            }
          }
          const double dtau1 = utaupr[lu] - taucpr[lyu], dtau2 = utaupr[lu] - taucpr[lyu + 1];
          const bool skip = (fabs(dtau1) < 1.0e-6 && neg) || (fabs(dtau2) < 1.0e-6 && !neg);
          if (!skip) {
            if (neg) exp1 = exp(dtau1 / mu);
            else exp2 = exp(dtau2 / mu);
            if (beam) {
              const double denom = 1.0 + mu / umu0;
              double expn;
              if (fabs(denom) < 0.0001) expn = (dtau1 / umu0) * exp0;
              else if (neg) expn = (exp0 - expbea[lyu] * exp1) / denom;
              else expn = (exp0 - expbea[lyu + 1] * exp2) / denom;
              palint += zbeamA[(long)lyu * U + iu] * expn;
            }
            const double dtau = dtaucp[lyu];
            for (int iq = 0; iq < nn; ++iq) {
              const double k = KKL(lyu, iq), denom = 1.0 + mu * k;
              double expn;
              if (fabs(denom) < 0.0001) expn = -dtau2 / mu * exp2;
              else if (neg) expn = (exp(-k * dtau2) - exp(k * dtau) * exp1) / denom;
              else expn = (exp(-k * dtau2) - exp2) / denom;
              palint += GUL(lyu, iu, iq) * expn;
            }
            for (int iq = nn; iq < N; ++iq) {
              const double k = KKL(lyu, iq), denom = 1.0 + mu * k;
              double expn;
              if (fabs(denom) < 0.0001) expn = -dtau1 / mu * exp1;
              else if (neg) expn = (exp(-k * dtau1) - exp1) / denom;
              else expn = (exp(-k * dtau1) - exp(-k * dtau) * exp2) / denom;
              palint += GUL(lyu, iu, iq) * expn;
            }
            if (planck && mazim == 0) {
              const double expn = neg ? exp1 : exp2;
              const double fact = (neg ? taucpr[lyu] : taucpr[lyu + 1]) + mu;
              plkint += z0uA[(long)lyu * U + iu] * (1.0 - expn) + z1uA[(long)lyu * U + iu] * (utaupr[lu] + mu - fact * expn);
            }
          }
          double bndint = 0.0;
          if (neg && mazim == 0) {
            bndint = (fisot + tplank) * exp(utaupr[lu] / mu);
          } else if (!neg && !(lyrcut || mazim > 0)) {
            const int lb = L - 1;
            double bnddfu = 0.0;
            for (int iq = nn - 1; iq >= 0; --iq) {
              double df = 0.0;
              for (int jq = 0; jq < nn; ++jq) df += GCL(lb, iq, jq) * ll[(long)lb * N + jq];
              for (int jq = nn; jq < N; ++jq) df += GCL(lb, iq, jq) * ll[(long)lb * N + jq] * exp(-KKL(lb, jq) * dtaucp[lb]);
              if (beam) df += ZZL(lb, iq) * expbea[L];
              if (planck) df += delm0 * (ZP0(lb, iq) + ZP1(lb, iq) * taucpr[L]);
              const int c = nn - 1 - iq;
              bnddfu += (1.0 + delm0) * albedo * cmu[c] * cwt[c] * df;
            }
            const double bnddir = beam ? umu0 * fbeam / kPi * albedo * expbea[L] : 0.0;
            bndint = (bnddfu + bnddir + delm0 * (1.0 - albedo) * bplank) * exp((utaupr[lu] - taucpr[L]) / mu);
          }
          res = palint + plkint + bndint;
        }
        uum[idx] = res;
      }
#undef GUL
    } else {  // c_intensity_components: intensities at the quadrature angles
      B200_PAR_FOR(idx, T * N) {
        const int lu = idx / N, iq = idx - lu * N;
        const int lyu = (int)w[M.layru + lu];
        double u = 0.0;
        if (!(lyrcut && lyu >= ncut)) {
          for (int jq = 0; jq < nn; ++jq)
            u += GCL(lyu, iq, jq) * ll[(long)lyu * N + jq] * exp(-KKL(lyu, jq) * (utaupr[lu] - taucpr[lyu + 1]));
          for (int jq = nn; jq < N; ++jq)
            u += GCL(lyu, iq, jq) * ll[(long)lyu * N + jq] * exp(-KKL(lyu, jq) * (utaupr[lu] - taucpr[lyu]));
          if (beam) u += ZZL(lyu, iq) * exp(-utaupr[lu] / umu0);
          if (planck && mazim == 0) u += ZP0(lyu, iq) + ZP1(lyu, iq) * utaupr[lu];
        }
        uum[idx] = u;
      }
    }
    B200_SYNC();
    // ---- Fourier sum over azimuth with DISORT's convergence test ----
    {
      double azerr = 0.0;
      B200_PAR_FOR(idx, nphi * T * U) {
        const int jp = idx / (T * U), rem = idx - jp * (T * U);
        if (mazim == 0) {
          uu[idx] = uum[rem];
        } else {
          const double phirad = (kPi / 180.0) * (P.phi[jp] - phi0);
          const double azterm = uum[rem] * cos(mazim * phirad);
          const double nv = uu[idx] + azterm;
          uu[idx] = nv;
          const double a = fabs(azterm), bb = fabs(nv);
          double r = (bb == 0.0) ? ((a == 0.0) ? 1.0 : DBL_MAX) : a / bb;
          if (a == 0.0) r = 0.0;
          if (r > azerr) azerr = r;
        }
      }
      w[M.red + tid] = azerr;
      B200_SYNC();
      if (tid == 0 && mazim > 0) {
        double e = 0.0;
        for (int t = 0; t < nthr; ++t) e = fmax(e, w[M.red + t]);
        if (e <= P.accur) hdr[H_KCONV] += 1.0;
        if (hdr[H_KCONV] >= 2.0) hdr[H_DONE] = 1.0;
      }
      B200_SYNC();
      if (hdr[H_DONE] != 0.0) break;
    }
  }

  // ---- c_intensity_correction (Nakajima-Tanaka TMS / IMS) ----
  if (st == ST_OK && !onlyfl && (P.flags & F_INTCOR) && beam) {
    const double dither = 10.0 * DBL_EPSILON;
    double* phasa = scr; double* phasm = scr + L; double* phast = scr + 2 * L;
    const double* flyr = w + M.flyr;
    const double* ssr = w + M.ssa;
    const double* tauc = w + M.tauc;
    B200_PAR_FOR(idx, U * nphi) {
      const int iu = idx / nphi, jp = idx - iu * nphi;
      const double mu = w[M.umu + iu];
      const double phirad = (kPi / 180.0) * (P.phi[jp] - phi0);
      const double ctheta = -umu0 * mu + sqrt((1.0 - umu0 * umu0) * (1.0 - mu * mu)) * cos(phirad);
      for (int lc = 0; lc < ncut; ++lc) { phasa[lc] = 1.0; phasm[lc] = 1.0; }
      double plm1 = 1.0, plm2 = 0.0;
      for (int k = 1; k <= P.nmom; ++k) {
        const double pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
        plm2 = plm1;
        plm1 = pl;
        for (int lc = 0; lc < ncut; ++lc) {
          const double pm = prop_pmom(pr, L, nprop, lc, k);
          phasa[lc] += (2 * k + 1) * pl * pm;
          if (k <= N - 1) phasm[lc] += (2 * k + 1) * pl * (pm - flyr[lc]) / (1.0 - flyr[lc]);
        }
      }
      for (int lc = 0; lc < ncut; ++lc) phast[lc] = phasa[lc] / (1.0 - flyr[lc] * ssr[lc]);
      for (int lu = 0; lu < T; ++lu) {
        const int lyu = (int)w[M.layru + lu];
        if (lyrcut && lyu >= ncut - 1) continue;
        const double ussndm = nt_sinsca(dither, lyu, ncut, phasm, w + M.oprim, taucpr, mu, umu0, utaupr[lu], fbeam);
        const double ussp = nt_sinsca(dither, lyu, ncut, phast, ssr, tauc, mu, umu0, w[M.utau + lu], fbeam);
        double duims = 0.0;
        if (mu < 0.0) {  // c_secondary_scat
          const double utau = w[M.utau + lu];
          const double dtau = utau - tauc[lyu];
          double wbar = ssr[lyu] * dtau, fbar = flyr[lyu] * wbar, stau = dtau;
          for (int l = 0; l < lyu; ++l) {
            const double dt = tauc[l + 1] - tauc[l];
            wbar += dt * ssr[l];
            fbar += ssr[l] * dt * flyr[l];
            stau += dt;
          }
          if (!(wbar <= 1.0e-4 || fbar <= 1.0e-4 || stau <= 1.0e-4 || fbeam <= 1.0e-4)) {
            fbar /= wbar;
            wbar /= stau;
            double pspike = 1.0, q1 = 1.0, q2 = 0.0;
            for (int k = 1; k <= N - 1; ++k) {
              const double pl = ((2 * k - 1) * ctheta * q1 - (k - 1) * q2) / k;
              q2 = q1; q1 = pl;
              pspike += (2 * k + 1) * pl;
            }
            for (int k = N; k <= P.nmom; ++k) {
              const double pl = ((2 * k - 1) * ctheta * q1 - (k - 1) * q2) / k;
              q2 = q1; q1 = pl;
              double gbar = prop_pmom(pr, L, nprop, lyu, k) * ssr[lyu] * dtau;
              for (int l = 0; l < lyu; ++l) gbar += prop_pmom(pr, L, nprop, l, k) * ssr[l] * (tauc[l + 1] - tauc[l]);
              gbar = (fbar * wbar * stau <= 1.0e-4) ? 0.0 : gbar / (fbar * wbar * stau);
              pspike += (2.0 * gbar - gbar * gbar) * (2 * k + 1) * pl;
            }
            const double umu0p = umu0 / (1.0 - fbar * wbar);
            duims = fbeam / (4.0 * kPi) * (fbar * wbar) * (fbar * wbar) / (1.0 - fbar * wbar) * pspike *
                    nt_xifunc(-mu, umu0p, umu0p, utau);
          }
        }
        uu[((long)jp * T + lu) * U + iu] += ussp - ussndm - duims;
      }
    }
  }
  B200_SYNC();
  // ---- outputs: pyharp's bottom-up level order for the fluxes ----
  {
    const double nanv = nan("");
    const double* f5 = w + M.flux5;
    B200_PAR_FOR(i, T) {
      const int lu = T - 1 - i;
      flx[i * 2] = (st != ST_OK) ? nanv : f5[lu * 2];
      flx[i * 2 + 1] = (st != ST_OK) ? nanv : f5[lu * 2 + 1];
    }
    if (st != ST_OK && uu) { B200_PAR_FOR(i, nphi * T * U) uu[i] = nanv; }
    if (tid == 0) P.status[s] = st;
  }
  B200_SYNC();
#undef GCL
#undef KKL
#undef ZZL
#undef ZP0
#undef ZP1
#undef AB
}

}  // namespace b200
