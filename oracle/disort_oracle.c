/*
 * disort_oracle.c -- TEST INFRASTRUCTURE ONLY (see disort_oracle.h).
 *
 * CPU restatement, routine by routine, of the DISORT v2 / cdisort 2.1.3 solve that
 * pyharp calls per (wave, column): reference call sites tests/test_disort.cpp:49,
 * examples/amars_sw.cpp:280, examples/amars_lw.cpp:80,
 * src/rtsolver/rt_solver_disort.cpp_:147,266.  The algorithm source is the
 * un-vendored dependency pydisort@afee3ec897f -> cdisort213/cdisort.c
 * (cmake/pydisort.cmake:9-11); each function below names the cdisort routine
 * (and the DISORT Fortran routine it descends from) that it restates.
 * PARITY AGAINST cdisort ITSELF IS UNPINNED (source unobtainable here); pinned
 * against tests/golden/ (DISORT test-problem values) and an independent solver.
 *
 * Build: see oracle/Makefile.  -DORC_REAL="long double" gives an extended-precision
 * build used only to estimate the FP64 oracle's own rounding error.
 */
#include "disort_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORC_REAL
#define ORC_REAL double
#endif
typedef ORC_REAL REAL;

#if defined(ORC_LONG_DOUBLE)
#define R_EXP expl
#define R_SQRT sqrtl
#define R_FABS fabsl
#define R_COS cosl
#define R_TAN tanl
#define R_LOG logl
#define R_EPS LDBL_EPSILON
#else
#define R_EXP exp
#define R_SQRT sqrt
#define R_FABS fabs
#define R_COS cos
#define R_TAN tan
#define R_LOG log
#define R_EPS DBL_EPSILON
#endif

#define PI_R ((REAL)3.14159265358979323846264338327950288L)

/* status codes (mirrored in include/b200disort.h) */
#define ST_OK 0
#define ST_BAD_INPUT 1
#define ST_EIGEN_FAIL 2
#define ST_SINGULAR 3
#define ST_UNSUPPORTED 4
#define ST_BEAM_ANGLE 5

int orc_real_bytes(void) { return (int)sizeof(REAL); }

/* ------------------------------------------------------------------------- */
/* c_gaussian_quadrature (QGAUSN): Gauss-Legendre nodes/weights on (0,1).     */
/* ------------------------------------------------------------------------- */
static void qgausn(int m, REAL* gmu, REAL* gwt) {
  const int maxit = 1000;
  const REAL tol = 10 * R_EPS;
  if (m == 1) {
    gmu[0] = 0.5;
    gwt[0] = 1.0;
    return;
  }
  REAL en = m;
  int np1 = m + 1;
  REAL nnp1 = (REAL)m * np1;
  REAL cona = (REAL)(m - 1) / (8.0 * (REAL)m * m * m);
  int lim = m / 2;
  for (int k = 1; k <= lim; ++k) {
    REAL t = (4 * k - 1) * PI_R / (4 * m + 2);
    REAL x = R_COS(t + cona / R_TAN(t));
    REAL p = 0, pm1 = 0, pm2 = 0, tmp = 0, ppr, p2pri, xi;
    int niter = 0;
    for (;;) {
      ++niter;
      pm2 = 1.0;
      pm1 = x;
      for (int nn = 2; nn <= m; ++nn) {
        p = ((2 * nn - 1) * x * pm1 - (nn - 1) * pm2) / nn;
        pm2 = pm1;
        pm1 = p;
      }
      tmp = 1.0 / (1.0 - x * x);
      ppr = en * (pm2 - x * p) * tmp;
      p2pri = (2.0 * x * ppr - nnp1 * p) * tmp;
      xi = x - (p / ppr) * (1.0 + (p / ppr) * p2pri / (2.0 * ppr));
      if (R_FABS(xi - x) > tol && niter <= maxit) {
        x = xi;
        continue;
      }
      break;
    }
    gmu[k - 1] = -x;
    gwt[k - 1] = 2.0 / (tmp * (en * pm2) * (en * pm2));
    gmu[np1 - k - 1] = -gmu[k - 1];
    gwt[np1 - k - 1] = gwt[k - 1];
  }
  if (m % 2 != 0) {
    gmu[lim] = 0.0;
    REAL prod = 1.0;
    for (int k = 3; k <= m; k += 2) prod = prod * k / (k - 1);
    gwt[lim] = 2.0 / (prod * prod);
  }
  for (int k = 0; k < m; ++k) {
    gmu[k] = 0.5 * gmu[k] + 0.5;
    gwt[k] = 0.5 * gwt[k];
  }
}

void orc_qgausn(int m, double* gmu, double* gwt) {
  REAL a[256], b[256];
  qgausn(m, a, b);
  for (int i = 0; i < m; ++i) {
    gmu[i] = (double)a[i];
    gwt[i] = (double)b[i];
  }
}

/* ------------------------------------------------------------------------- */
/* c_legendre_poly (LEPOLY): normalised associated Legendre polynomials       */
/* Y_l^m(mu) = sqrt((l-m)!/(l+m)!) P_l^m(mu), l = m..twonm1.                  */
/* ylm is [(twonm1+1)][nmu].  Unlike the Fortran (which reuses Y_{m-1}^{m-1}  */
/* left in the array by the previous azimuth pass) Y_m^m is rebuilt by the    */
/* same recurrence from m' = 1..m, so calls are order-independent.            */
/* ------------------------------------------------------------------------- */
static void lepoly(int nmu, int m, int twonm1, const REAL* mu, REAL* ylm) {
#define YLM(l, i) ylm[(l) * nmu + (i)]
  if (m == 0) {
    for (int i = 0; i < nmu; ++i) {
      YLM(0, i) = 1.0;
      if (twonm1 >= 1) YLM(1, i) = mu[i];
    }
    for (int l = 2; l <= twonm1; ++l)
      for (int i = 0; i < nmu; ++i)
        YLM(l, i) = ((2 * l - 1) * mu[i] * YLM(l - 1, i) - (l - 1) * YLM(l - 2, i)) / l;
  } else {
    for (int i = 0; i < nmu; ++i) {
      REAL tmp1 = 1.0 - mu[i] * mu[i];
      REAL ymm = 1.0; /* Y_0^0 */
      for (int mp = 1; mp <= m; ++mp)
        ymm = -R_SQRT((REAL)(2 * mp - 1)) / R_SQRT((REAL)(2 * mp)) * R_SQRT(tmp1) * ymm;
      if (m <= twonm1) YLM(m, i) = ymm;
      if (m + 1 <= twonm1) YLM(m + 1, i) = R_SQRT((REAL)(2 * m + 1)) * mu[i] * ymm;
    }
    for (int l = m + 2; l <= twonm1; ++l) {
      REAL tmp1 = R_SQRT((REAL)(l - m)) * R_SQRT((REAL)(l + m));
      REAL tmp2 = R_SQRT((REAL)(l - m - 1)) * R_SQRT((REAL)(l + m - 1));
      for (int i = 0; i < nmu; ++i)
        YLM(l, i) = ((2 * l - 1) * mu[i] * YLM(l - 1, i) - tmp2 * YLM(l - 2, i)) / tmp1;
    }
  }
#undef YLM
}

void orc_lepoly(int nmu, int m, int twonm1, const double* mu, double* ylm) {
  REAL* a = (REAL*)malloc(sizeof(REAL) * nmu);
  REAL* y = (REAL*)calloc((size_t)(twonm1 + 1) * nmu, sizeof(REAL));
  for (int i = 0; i < nmu; ++i) a[i] = mu[i];
  lepoly(nmu, m, twonm1, a, y);
  for (int i = 0; i < (twonm1 + 1) * nmu; ++i) ylm[i] = (double)y[i];
  free(a);
  free(y);
}

/* ------------------------------------------------------------------------- */
/* c_planck_func1 (PLKAVG): Planck function integrated between two            */
/* wavenumbers, W/m^2.  Same series, cut-offs and legacy constants.           */
/* ------------------------------------------------------------------------- */
static REAL plkf(REAL x) { return x * x * x / (R_EXP(x) - 1.0); }

static REAL plkavg(REAL wnumlo, REAL wnumhi, REAL t) {
  const REAL a1 = 1.0 / 3.0, a2 = -1.0 / 8.0, a3 = 1.0 / 60.0, a4 = -1.0 / 5040.0,
             a5 = 1.0 / 272160.0, a6 = -1.0 / 13305600.0;
  const REAL c2 = 1.438786, sigma = 5.67032e-8, vcut = 1.5;
  const REAL vcp[7] = {10.25, 5.7, 3.9, 2.9, 2.3, 1.9, 0.0};
  const REAL vmax = R_LOG((REAL)DBL_MAX);
  const REAL epsil = R_EPS;
  const REAL sigdpi = sigma / PI_R;
  const REAL conc = 15.0 / (PI_R * PI_R * PI_R * PI_R);
  REAL v[2], p[2] = {0, 0}, d[2] = {0, 0};

  if (t < 0.0 || wnumhi < wnumlo || wnumlo < 0.0) return -1.0; /* caller flags bad input */
  if (t < 1.0e-4) return 0.0;
  v[0] = c2 * wnumlo / t;
  v[1] = c2 * wnumhi / t;
  if (v[0] > epsil && v[1] < vmax && (wnumhi - wnumlo) / wnumhi < 1.0e-2) {
    /* wavenumbers very close: Simpson rule iterated to convergence */
    REAL hh = v[1] - v[0], oldval = 0.0, val = 0.0;
    REAL val0 = plkf(v[0]) + plkf(v[1]);
    for (int n = 1; n <= 10; ++n) {
      REAL del = hh / (2 * n);
      val = val0;
      for (int k = 1; k <= 2 * n - 1; ++k) val += 2 * (1 + k % 2) * plkf(v[0] + k * del);
      val = del / 3.0 * val;
      if (R_FABS((val - oldval) / val) <= 1.0e-6) break;
      oldval = val;
    }
    return sigdpi * t * t * t * t * conc * val;
  }
  int smallv = 0;
  for (int i = 0; i < 2; ++i) {
    if (v[i] < vcut) {
      ++smallv;
      REAL vsq = v[i] * v[i];
      p[i] = conc * vsq * v[i] *
             (a1 + v[i] * (a2 + v[i] * (a3 + vsq * (a4 + vsq * (a5 + vsq * a6)))));
    } else {
      int mmax = 0;
      do {
        ++mmax;
      } while (v[i] < vcp[mmax - 1]);
      REAL ex = R_EXP(-v[i]), exm = 1.0;
      d[i] = 0.0;
      for (int m = 1; m <= mmax; ++m) {
        REAL mv = m * v[i];
        exm = ex * exm;
        d[i] += exm * (6.0 + mv * (6.0 + mv * (3.0 + mv))) / ((REAL)m * m * m * m);
      }
      d[i] *= conc;
    }
  }
  REAL ans;
  if (smallv == 2)
    ans = p[1] - p[0];
  else if (smallv == 1)
    ans = 1.0 - p[0] - d[1];
  else
    ans = d[0] - d[1];
  return sigdpi * t * t * t * t * ans;
}

double orc_plkavg(double wnumlo, double wnumhi, double t) {
  return (double)plkavg(wnumlo, wnumhi, t);
}

/* ------------------------------------------------------------------------- */
/* LINPACK sgefa/sgesl restated (dense LU, partial pivoting), row-major a.    */
/* c_sgeco's condition estimate is only used for a warning in cdisort and is  */
/* not reproduced.                                                            */
/* ------------------------------------------------------------------------- */
static int sgefa(REAL* a, int n, int* ipvt) {
  int info = 0;
  for (int k = 0; k < n - 1; ++k) {
    int l = k;
    REAL amax = R_FABS(a[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (R_FABS(a[i * n + k]) > amax) {
        amax = R_FABS(a[i * n + k]);
        l = i;
      }
    ipvt[k] = l;
    if (a[l * n + k] == 0.0) {
      info = k + 1;
      continue;
    }
    if (l != k) /* LINPACK swaps only the active columns; multipliers stay put */
      for (int j = k; j < n; ++j) {
        REAL t = a[l * n + j];
        a[l * n + j] = a[k * n + j];
        a[k * n + j] = t;
      }
    REAL piv = -1.0 / a[k * n + k];
    for (int i = k + 1; i < n; ++i) a[i * n + k] *= piv;
    for (int i = k + 1; i < n; ++i) {
      REAL t = a[i * n + k];
      for (int j = k + 1; j < n; ++j) a[i * n + j] += t * a[k * n + j];
    }
  }
  ipvt[n - 1] = n - 1;
  if (a[(n - 1) * n + n - 1] == 0.0) info = n;
  return info;
}

static void sgesl(const REAL* a, int n, const int* ipvt, REAL* b) {
  /* forward elimination interleaves the row interchanges, as LINPACK sgesl does */
  for (int k = 0; k < n - 1; ++k) {
    int l = ipvt[k];
    REAL t = b[l];
    if (l != k) {
      b[l] = b[k];
      b[k] = t;
    }
    for (int i = k + 1; i < n; ++i) b[i] += a[i * n + k] * t;
  }
  for (int k = n - 1; k >= 0; --k) {
    b[k] /= a[k * n + k];
    REAL t = -b[k];
    for (int i = 0; i < k; ++i) b[i] += t * a[i * n + k];
  }
}

/* ------------------------------------------------------------------------- */
/* LINPACK sgbfa/sgbsl restated: banded LU with partial pivoting.             */
/* Storage: ab[i][c], c = j - i + kl, row width w = 2*kl+ku+1 (row i holds    */
/* columns i-kl .. i+kl+ku; the extra kl columns take pivoting fill-in).      */
/* Pivot choice (largest magnitude in the column among rows k..k+kl, first    */
/* on ties) is LINPACK's.                                                     */
/* ------------------------------------------------------------------------- */
static int sgbfa_solve(REAL* ab, int n, int kl, int ku, REAL* b) {
  const int w = 2 * kl + ku + 1;
#define AB(i, j) ab[(size_t)(i) * w + ((j) - (i) + kl)]
  int info = 0;
  for (int k = 0; k < n; ++k) {
    int lm = (kl < n - 1 - k) ? kl : n - 1 - k;
    int l = k;
    REAL amax = R_FABS(AB(k, k));
    for (int i = k + 1; i <= k + lm; ++i)
      if (R_FABS(AB(i, k)) > amax) {
        amax = R_FABS(AB(i, k));
        l = i;
      }
    if (AB(l, k) == 0.0) {
      info = k + 1;
      continue;
    }
    int ju = k + kl + ku;
    if (ju > n - 1) ju = n - 1;
    if (l != k) {
      for (int j = k; j <= ju; ++j) {
        REAL t = AB(l, j);
        AB(l, j) = AB(k, j);
        AB(k, j) = t;
      }
      REAL t = b[l];
      b[l] = b[k];
      b[k] = t;
    }
    REAL piv = 1.0 / AB(k, k);
    for (int i = k + 1; i <= k + lm; ++i) {
      REAL t = -AB(i, k) * piv;
      if (t != 0.0) {
        for (int j = k + 1; j <= ju; ++j) AB(i, j) += t * AB(k, j);
        b[i] += t * b[k];
      }
      AB(i, k) = 0.0;
    }
  }
  if (info) return info;
  for (int k = n - 1; k >= 0; --k) {
    int ju = k + kl + ku;
    if (ju > n - 1) ju = n - 1;
    REAL s = b[k];
    for (int j = k + 1; j <= ju; ++j) s -= AB(k, j) * b[j];
    b[k] = s / AB(k, k);
  }
#undef AB
  return 0;
}

/* ------------------------------------------------------------------------- */
/* c_asymmetric_matrix (ASYMTX): eigenvalues/vectors of a real general matrix */
/* known to have real eigenvalues.  ASYMTX is EISPACK BALANC + ELMHES/ELTRAN  */
/* + HQR2 (real-eigenvalue branch only) + BALBAK merged; this is the same     */
/* sequence (same radix-16 balancing, 0.95 test, exceptional shifts 0.75 /    */
/* -0.4375 at iterations 10 and 20, 30-iteration limit).                      */
/* aa: n x n row-major (not modified).  evec: n x n row-major, column j is    */
/* the eigenvector of eval[j].  Returns 0, or k>0 if eigenvalue k failed.     */
/* ------------------------------------------------------------------------- */
static REAL r_sign(REAL a, REAL b) { return b >= 0 ? R_FABS(a) : -R_FABS(a); }

static int asymtx(const REAL* aa, int n, REAL* evec, REAL* eval) {
  const REAL c1 = 0.4375, c2 = 0.5, c3 = 0.75, c4 = 0.95, c5 = 16.0, c6 = 256.0;
  const REAL tol = R_EPS;
  (void)c2;
  if (n == 1) {
    eval[0] = aa[0];
    evec[0] = 1.0;
    return 0;
  }
  if (n == 2) {
    REAL discri = (aa[0] - aa[3]) * (aa[0] - aa[3]) + 4.0 * aa[1] * aa[2];
    if (discri < 0.0) return 1;
    REAL sgn = (aa[0] < aa[3]) ? -1.0 : 1.0;
    eval[0] = 0.5 * (aa[0] + aa[3] + sgn * R_SQRT(discri));
    eval[1] = 0.5 * (aa[0] + aa[3] - sgn * R_SQRT(discri));
    evec[0] = 1.0;
    evec[3] = 1.0;
    if (aa[0] == aa[3] && (aa[2] == 0.0 || aa[1] == 0.0)) {
      REAL rnorm = R_FABS(aa[0]) + R_FABS(aa[1]) + R_FABS(aa[2]) + R_FABS(aa[3]);
      REAL w = tol * rnorm;
      evec[2] = aa[2] / w;  /* evec(2,1) */
      evec[1] = -aa[1] / w; /* evec(1,2) */
    } else {
      evec[2] = aa[2] / (eval[0] - aa[3]);
      evec[1] = aa[1] / (eval[1] - aa[0]);
    }
    return 0;
  }
  REAL* a = (REAL*)malloc(sizeof(REAL) * n * n);
  REAL* z = evec;
  REAL* scale = (REAL*)malloc(sizeof(REAL) * n);
  int* intch = (int*)malloc(sizeof(int) * n);
  memcpy(a, aa, sizeof(REAL) * n * n);
#define A(i, j) a[(i) * n + (j)]
#define Z(i, j) z[(i) * n + (j)]
  /* ---- BALANC ---- */
  int low = 0, high = n - 1;
  {
    int done;
    /* search for rows isolating an eigenvalue and push them down */
    do {
      done = 1;
      for (int j = high; j >= 0; --j) {
        int iso = 1;
        for (int i = 0; i <= high; ++i)
          if (i != j && A(j, i) != 0.0) {
            iso = 0;
            break;
          }
        if (iso) {
          scale[high] = j;
          if (j != high) {
            for (int i = 0; i <= high; ++i) {
              REAL t = A(i, j);
              A(i, j) = A(i, high);
              A(i, high) = t;
            }
            for (int i = low; i < n; ++i) {
              REAL t = A(j, i);
              A(j, i) = A(high, i);
              A(high, i) = t;
            }
          }
          if (high == 0) {
            /* all isolated: triangular permutation; fall through with high=low */
            done = 1;
            goto balanced_trivial;
          }
          --high;
          done = 0;
          break;
        }
      }
    } while (!done);
    /* search for columns isolating an eigenvalue and push them left */
    do {
      done = 1;
      for (int j = low; j <= high; ++j) {
        int iso = 1;
        for (int i = low; i <= high; ++i)
          if (i != j && A(i, j) != 0.0) {
            iso = 0;
            break;
          }
        if (iso) {
          scale[low] = j;
          if (j != low) {
            for (int i = 0; i <= high; ++i) {
              REAL t = A(i, j);
              A(i, j) = A(i, low);
              A(i, low) = t;
            }
            for (int i = low; i < n; ++i) {
              REAL t = A(j, i);
              A(j, i) = A(low, i);
              A(low, i) = t;
            }
          }
          ++low;
          done = 0;
          break;
        }
      }
    } while (!done);
  balanced_trivial:
    for (int i = low; i <= high; ++i) scale[i] = 1.0;
    /* iterative loop for norm reduction */
    int noconv;
    do {
      noconv = 0;
      for (int i = low; i <= high; ++i) {
        REAL col = 0.0, row = 0.0;
        for (int j = low; j <= high; ++j)
          if (j != i) {
            col += R_FABS(A(j, i));
            row += R_FABS(A(i, j));
          }
        if (col == 0.0 || row == 0.0) continue;
        REAL f = 1.0, g = row / c5, s = col + row;
        while (col < g) {
          f *= c5;
          col *= c6;
        }
        g = row * c5;
        while (col >= g) {
          f /= c5;
          col /= c6;
        }
        if ((col + row) / f < c4 * s) {
          scale[i] *= f;
          noconv = 1;
          for (int j = low; j < n; ++j) A(i, j) /= f;
          for (int j = 0; j <= high; ++j) A(j, i) *= f;
        }
      }
    } while (noconv);
  }
  /* ---- ELMHES: reduce to upper Hessenberg by stabilised elementary transforms ---- */
  for (int m = low + 1; m < high; ++m) {
    REAL x = 0.0;
    int i = m;
    for (int j = m; j <= high; ++j)
      if (R_FABS(A(j, m - 1)) > R_FABS(x)) {
        x = A(j, m - 1);
        i = j;
      }
    intch[m] = i;
    if (i != m) {
      for (int j = m - 1; j < n; ++j) {
        REAL t = A(i, j);
        A(i, j) = A(m, j);
        A(m, j) = t;
      }
      for (int j = 0; j <= high; ++j) {
        REAL t = A(j, i);
        A(j, i) = A(j, m);
        A(j, m) = t;
      }
    }
    if (x != 0.0) {
      for (int ii = m + 1; ii <= high; ++ii) {
        REAL y = A(ii, m - 1);
        if (y != 0.0) {
          y /= x;
          A(ii, m - 1) = y;
          for (int j = m; j < n; ++j) A(ii, j) -= y * A(m, j);
          for (int j = 0; j <= high; ++j) A(j, m) += y * A(j, ii);
        }
      }
    }
  }
  /* ---- ELTRAN: accumulate the transformations ---- */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Z(i, j) = (i == j) ? 1.0 : 0.0;
  for (int mp = high - 1; mp >= low + 1; --mp) {
    for (int i = mp + 1; i <= high; ++i) Z(i, mp) = A(i, mp - 1);
    int i = intch[mp];
    if (i != mp) {
      for (int j = mp; j <= high; ++j) {
        Z(mp, j) = Z(i, j);
        Z(i, j) = 0.0;
      }
      Z(i, mp) = 1.0;
    }
  }
  /* zero below the subdiagonal (ELMHES leaves multipliers there) */
  for (int i = 2; i < n; ++i)
    for (int j = 0; j < i - 1; ++j) A(i, j) = 0.0;

  /* ---- HQR2, real-eigenvalue branch ---- */
  int ier = 0;
  REAL rnorm = 0.0;
  {
    int k = 0;
    for (int i = 0; i < n; ++i) {
      for (int j = k; j < n; ++j) rnorm += R_FABS(A(i, j));
      k = i;
      if (i < low || i > high) eval[i] = A(i, i);
    }
  }
  int en = high;
  REAL t = 0.0;
  while (en >= low) {
    int its = 0;
    for (;;) {
      int l;
      REAL x, y, w, p = 0, q = 0, r = 0, s, zz;
      for (l = en; l > low; --l) {
        s = R_FABS(A(l - 1, l - 1)) + R_FABS(A(l, l));
        if (s == 0.0) s = rnorm;
        if (R_FABS(A(l, l - 1)) <= tol * s) break;
      }
      x = A(en, en);
      if (l == en) { /* one root found */
        A(en, en) = x + t;
        eval[en] = A(en, en);
        --en;
        break;
      }
      y = A(en - 1, en - 1);
      w = A(en, en - 1) * A(en - 1, en);
      if (l == en - 1) { /* two roots found */
        p = (y - x) * 0.5;
        q = p * p + w;
        zz = R_SQRT(R_FABS(q));
        A(en, en) = x + t;
        x = A(en, en);
        A(en - 1, en - 1) = y + t;
        if (q < 0.0) { /* complex pair: not allowed for the DISORT matrix */
          ier = en + 1;
          goto done;
        }
        zz = p + r_sign(zz, p);
        eval[en - 1] = x + zz;
        eval[en] = eval[en - 1];
        if (zz != 0.0) eval[en] = x - w / zz;
        x = A(en, en - 1);
        s = R_FABS(x) + R_FABS(zz);
        p = x / s;
        q = zz / s;
        r = R_SQRT(p * p + q * q);
        p /= r;
        q /= r;
        for (int j = en - 1; j < n; ++j) { /* row modification */
          zz = A(en - 1, j);
          A(en - 1, j) = q * zz + p * A(en, j);
          A(en, j) = q * A(en, j) - p * zz;
        }
        for (int i = 0; i <= en; ++i) { /* column modification */
          zz = A(i, en - 1);
          A(i, en - 1) = q * zz + p * A(i, en);
          A(i, en) = q * A(i, en) - p * zz;
        }
        for (int i = low; i <= high; ++i) { /* accumulate */
          zz = Z(i, en - 1);
          Z(i, en - 1) = q * zz + p * Z(i, en);
          Z(i, en) = q * Z(i, en) - p * zz;
        }
        en -= 2;
        break;
      }
      if (its == 30) {
        ier = en + 1;
        goto done;
      }
      if (its == 10 || its == 20) { /* exceptional shift */
        t += x;
        for (int i = low; i <= en; ++i) A(i, i) -= x;
        s = R_FABS(A(en, en - 1)) + R_FABS(A(en - 1, en - 2));
        x = c3 * s;
        y = x;
        w = -c1 * s * s;
      }
      ++its;
      /* look for two consecutive small sub-diagonal elements */
      int m;
      for (m = en - 2; m >= l; --m) {
        zz = A(m, m);
        r = x - zz;
        s = y - zz;
        p = (r * s - w) / A(m + 1, m) + A(m, m + 1);
        q = A(m + 1, m + 1) - zz - r - s;
        r = A(m + 2, m + 1);
        s = R_FABS(p) + R_FABS(q) + R_FABS(r);
        p /= s;
        q /= s;
        r /= s;
        if (m == l) break;
        REAL uu = R_FABS(A(m, m - 1)) * (R_FABS(q) + R_FABS(r));
        REAL vv = R_FABS(p) * (R_FABS(A(m - 1, m - 1)) + R_FABS(zz) + R_FABS(A(m + 1, m + 1)));
        if (uu <= tol * vv) break;
      }
      for (int i = m + 2; i <= en; ++i) {
        A(i, i - 2) = 0.0;
        if (i != m + 2) A(i, i - 3) = 0.0;
      }
      /* double QR step involving rows l..en and columns m..en */
      for (int k = m; k <= en - 1; ++k) {
        int notlas = (k != en - 1);
        if (k != m) {
          p = A(k, k - 1);
          q = A(k + 1, k - 1);
          r = notlas ? A(k + 2, k - 1) : 0.0;
          x = R_FABS(p) + R_FABS(q) + R_FABS(r);
          if (x == 0.0) continue;
          p /= x;
          q /= x;
          r /= x;
        }
        s = r_sign(R_SQRT(p * p + q * q + r * r), p);
        if (k == m) {
          if (l != m) A(k, k - 1) = -A(k, k - 1);
        } else {
          A(k, k - 1) = -s * x;
        }
        p += s;
        x = p / s;
        y = q / s;
        zz = r / s;
        q /= p;
        r /= p;
        for (int j = k; j < n; ++j) { /* row modification */
          p = A(k, j) + q * A(k + 1, j);
          if (notlas) {
            p += r * A(k + 2, j);
            A(k + 2, j) -= p * zz;
          }
          A(k + 1, j) -= p * y;
          A(k, j) -= p * x;
        }
        int mmin = (en < k + 3) ? en : k + 3;
        for (int i = 0; i <= mmin; ++i) { /* column modification */
          p = x * A(i, k) + y * A(i, k + 1);
          if (notlas) {
            p += zz * A(i, k + 2);
            A(i, k + 2) -= p * r;
          }
          A(i, k + 1) -= p * q;
          A(i, k) -= p;
        }
        for (int i = low; i <= high; ++i) { /* accumulate transformations */
          p = x * Z(i, k) + y * Z(i, k + 1);
          if (notlas) {
            p += zz * Z(i, k + 2);
            Z(i, k + 2) -= p * r;
          }
          Z(i, k + 1) -= p * q;
          Z(i, k) -= p;
        }
      }
    }
  }
  /* all roots found: back-substitute for vectors of the upper-triangular form */
  if (rnorm != 0.0) {
    for (int e = n - 1; e >= 0; --e) {
      REAL p = eval[e];
      int m = e;
      A(e, e) = 1.0;
      for (int i = e - 1; i >= 0; --i) {
        REAL w = A(i, i) - p;
        REAL r = 0.0;
        for (int j = m; j <= e; ++j) r += A(i, j) * A(j, e);
        m = i;
        REAL tt = w;
        if (tt == 0.0) tt = tol * rnorm;
        A(i, e) = -r / tt;
        /* overflow control */
        tt = R_FABS(A(i, e));
        if (tt != 0.0) {
          REAL tst1 = tt, tst2 = tst1 + 1.0 / tst1;
          if (tst2 <= tst1)
            for (int j = i; j <= e; ++j) A(j, e) /= tt;
        }
      }
    }
    /* vectors of isolated roots */
    for (int i = 0; i < n; ++i)
      if (i < low || i > high)
        for (int j = i; j < n; ++j) Z(i, j) = A(i, j);
    /* multiply by transformation matrix to give vectors of the balanced matrix */
    for (int j = n - 1; j >= low; --j) {
      int m = (j < high) ? j : high;
      for (int i = low; i <= high; ++i) {
        REAL zz = 0.0;
        for (int k = low; k <= m; ++k) zz += Z(i, k) * A(k, j);
        Z(i, j) = zz;
      }
    }
  }
  /* ---- BALBAK ---- */
  for (int i = low; i <= high; ++i)
    for (int j = 0; j < n; ++j) Z(i, j) *= scale[i];
  for (int i = low - 1; i >= 0; --i) {
    int k = (int)scale[i];
    if (k != i)
      for (int j = 0; j < n; ++j) {
        REAL tt = Z(i, j);
        Z(i, j) = Z(k, j);
        Z(k, j) = tt;
      }
  }
  for (int i = high + 1; i < n; ++i) {
    int k = (int)scale[i];
    if (k != i)
      for (int j = 0; j < n; ++j) {
        REAL tt = Z(i, j);
        Z(i, j) = Z(k, j);
        Z(k, j) = tt;
      }
  }
done:
#undef A
#undef Z
  free(a);
  free(scale);
  free(intch);
  return ier;
}

int orc_asymtx(const double* aa, int n, double* evec, double* eval) {
  REAL* a = (REAL*)malloc(sizeof(REAL) * n * n);
  REAL* z = (REAL*)malloc(sizeof(REAL) * n * n);
  REAL* e = (REAL*)malloc(sizeof(REAL) * n);
  for (int i = 0; i < n * n; ++i) a[i] = aa[i];
  int ier = asymtx(a, n, z, e);
  for (int i = 0; i < n * n; ++i) evec[i] = (double)z[i];
  for (int i = 0; i < n; ++i) eval[i] = (double)e[i];
  free(a);
  free(z);
  free(e);
  return ier;
}

/* ------------------------------------------------------------------------- */
/* Workspace for one solve                                                    */
/* ------------------------------------------------------------------------- */
typedef struct {
  int nlyr, nstr, nn, ntau, numu, nphi, ncut, lyrcut;
  REAL *cmu, *cwt;                  /* [nstr] */
  REAL *tauc, *taucpr;              /* [nlyr+1] */
  REAL *dtaucp, *oprim, *flyr, *ssa; /* [nlyr] */
  REAL* gl;                         /* [nlyr][nstr] */
  REAL* expbea;                     /* [nlyr+1] */
  REAL *utau, *utaupr;              /* [ntau] */
  int* layru;                       /* [ntau] */
  REAL* umu;                        /* [numu] */
  REAL *pkag, *xr0, *xr1;           /* [nlyr+1], [nlyr], [nlyr] */
  REAL *ylmc, *ylm0, *ylmu;         /* [nstr][nstr], [nstr], [nstr][numu] */
  REAL *gc, *kk, *ll;               /* [nlyr][nstr][nstr], [nlyr][nstr] x2 */
  REAL *zz, *zplk0, *zplk1;         /* [nlyr][nstr] */
  REAL *gu, *zbeam, *z0u, *z1u;     /* [nlyr][numu][nstr], [nlyr][numu] x3 */
  REAL *zj, *z0, *z1;               /* [nstr] CMU ordering */
  REAL *cc, *amb, *apb, *arr, *evecc, *eval; /* scratch */
  REAL *cband, *b;
  REAL *u0c;                        /* [ntau][nstr] */
  REAL *uum;                        /* [ntau][numu] */
  REAL *wk;
} work;

static REAL* ralloc(size_t n) { return (REAL*)calloc(n ? n : 1, sizeof(REAL)); }

/* ------------------------------------------------------------------------- */
/* c_sol_eigen (SOLEIG): eigenproblem of STWJ(8b) for one layer and azimuth m */
/* ------------------------------------------------------------------------- */
static int soleig(work* w, int lc, int mazim) {
  const int nstr = w->nstr, nn = w->nn;
  const REAL* gl = w->gl + (size_t)lc * nstr;
  REAL *cc = w->cc, *amb = w->amb, *apb = w->apb, *arr = w->arr, *evecc = w->evecc,
       *eval = w->eval;
  REAL* gc = w->gc + (size_t)lc * nstr * nstr;
  REAL* kk = w->kk + (size_t)lc * nstr;
#define YLMC(l, i) w->ylmc[(l) * nstr + (i)]
#define CC(i, j) cc[(i) * nstr + (j)]
#define GC(i, j) gc[(i) * nstr + (j)]
  /* kernel of the scattering integral at the quadrature angles, STWJ(8b) */
  for (int iq = 0; iq < nn; ++iq) {
    for (int jq = 0; jq < nstr; ++jq) {
      REAL sum = 0.0;
      for (int l = mazim; l <= nstr - 1; ++l) sum += gl[l] * YLMC(l, iq) * YLMC(l, jq);
      CC(iq, jq) = 0.5 * sum * w->cwt[jq];
    }
    for (int jq = 0; jq < nn; ++jq) {
      /* fill remainder by symmetry C(-mui,muj)=C(mui,-muj), C(-mui,-muj)=C(mui,muj) */
      CC(iq + nn, jq) = CC(iq, jq + nn);
      CC(iq + nn, jq + nn) = CC(iq, jq);
      REAL alpha = CC(iq, jq) / w->cmu[iq];
      REAL beta = CC(iq, jq + nn) / w->cmu[iq];
      amb[iq * nn + jq] = alpha - beta;
      apb[iq * nn + jq] = alpha + beta;
    }
    amb[iq * nn + iq] -= 1.0 / w->cmu[iq];
    apb[iq * nn + iq] -= 1.0 / w->cmu[iq];
  }
  /* (alfa-beta)*(alfa+beta), STWJ(12) */
  for (int iq = 0; iq < nn; ++iq)
    for (int jq = 0; jq < nn; ++jq) {
      REAL sum = 0.0;
      for (int kq = 0; kq < nn; ++kq) sum += amb[iq * nn + kq] * apb[kq * nn + jq];
      arr[iq * nn + jq] = sum;
    }
  int ier = asymtx(arr, nn, evecc, eval);
  if (ier > 0) return ST_EIGEN_FAIL;
  for (int iq = 0; iq < nn; ++iq) {
    eval[iq] = R_SQRT(R_FABS(eval[iq]));
    kk[iq + nn] = eval[iq];
    kk[nn - 1 - iq] = -eval[iq]; /* negative eigenvalue precedes its positive twin */
  }
  /* eigenvectors (G+)+(G-) = evecc; (G+)-(G-) from STWJ(11); stack, SS(10) */
  for (int jq = 0; jq < nn; ++jq) {
    for (int iq = 0; iq < nn; ++iq) {
      REAL sum = 0.0;
      for (int kq = 0; kq < nn; ++kq) sum += apb[iq * nn + kq] * evecc[kq * nn + jq];
      REAL gpmigm = sum / eval[jq];
      REAL gpplgm = evecc[iq * nn + jq];
      REAL up = 0.5 * (gpplgm + gpmigm); /* component at +cmu[iq] for +k */
      REAL dn = 0.5 * (gpplgm - gpmigm); /* component at -cmu[iq] for +k */
      GC(iq + nn, jq + nn) = up;
      GC(nn - 1 - iq, jq + nn) = dn;
      GC(iq + nn, nn - 1 - jq) = dn;
      GC(nn - 1 - iq, nn - 1 - jq) = up;
    }
  }
#undef GC
  return ST_OK;
}

/* ------------------------------------------------------------------------- */
/* c_upbeam (UPBEAM): particular solution for the pseudo-beam source          */
/* ------------------------------------------------------------------------- */
static int upbeam(work* w, int lc, int mazim, REAL fbeam, REAL umu0) {
  const int nstr = w->nstr, nn = w->nn;
  const REAL* gl = w->gl + (size_t)lc * nstr;
  REAL* cc = w->cc;
  REAL* arr = w->arr;
  REAL* zj = w->zj;
  REAL* zz = w->zz + (size_t)lc * nstr;
  int ipvt[512];
  REAL delm0 = (mazim == 0) ? 1.0 : 0.0;
  for (int iq = 0; iq < nstr; ++iq) {
    for (int jq = 0; jq < nstr; ++jq) arr[iq * nstr + jq] = -CC(iq, jq);
    arr[iq * nstr + iq] += 1.0 + w->cmu[iq] / umu0;
    REAL sum = 0.0;
    for (int k = mazim; k <= nstr - 1; ++k) sum += gl[k] * YLMC(k, iq) * w->ylm0[k];
    zj[iq] = (2.0 - delm0) * fbeam * sum / (4.0 * PI_R);
  }
  if (sgefa(arr, nstr, ipvt) != 0) return ST_SINGULAR;
  sgesl(arr, nstr, ipvt, zj);
  for (int iq = 0; iq < nn; ++iq) {
    zz[iq + nn] = zj[iq];
    zz[nn - 1 - iq] = zj[iq + nn];
  }
  return ST_OK;
}

/* ------------------------------------------------------------------------- */
/* c_upisot (UPISOT): particular solution for thermal emission, linear in tau */
/* ------------------------------------------------------------------------- */
static int upisot(work* w, int lc) {
  const int nstr = w->nstr, nn = w->nn;
  REAL* cc = w->cc;
  REAL* arr = w->arr;
  REAL *z0 = w->z0, *z1 = w->z1;
  REAL* zplk0 = w->zplk0 + (size_t)lc * nstr;
  REAL* zplk1 = w->zplk1 + (size_t)lc * nstr;
  REAL oprim = w->oprim[lc], xr0 = w->xr0[lc], xr1 = w->xr1[lc];
  int ipvt[512];
  for (int iq = 0; iq < nstr; ++iq) {
    for (int jq = 0; jq < nstr; ++jq) arr[iq * nstr + jq] = -CC(iq, jq);
    arr[iq * nstr + iq] += 1.0;
    z1[iq] = (1.0 - oprim) * xr1;
  }
  if (sgefa(arr, nstr, ipvt) != 0) return ST_SINGULAR;
  sgesl(arr, nstr, ipvt, z1);
  for (int iq = 0; iq < nstr; ++iq) z0[iq] = (1.0 - oprim) * xr0 + w->cmu[iq] * z1[iq];
  sgesl(arr, nstr, ipvt, z0);
  for (int iq = 0; iq < nn; ++iq) {
    zplk0[iq + nn] = z0[iq];
    zplk1[iq + nn] = z1[iq];
    zplk0[nn - 1 - iq] = z0[iq + nn];
    zplk1[nn - 1 - iq] = z1[iq + nn];
  }
  return ST_OK;
}
#undef CC

/* ------------------------------------------------------------------------- */
/* c_terpev (TERPEV) + c_terpso (TERPSO): eigenvectors and source terms       */
/* interpolated to the user angles, SD(8,9), STWL(31c).  Written in the       */
/* stacked GC/ZZ ordering (rows ascending in mu) rather than EVECC/ZJ order;  */
/* the sums are identical.                                                    */
/* ------------------------------------------------------------------------- */
static void terp(work* w, int lc, int mazim, int planck, REAL fbeam) {
  const int nstr = w->nstr, nn = w->nn, numu = w->numu;
  const REAL* gl = w->gl + (size_t)lc * nstr;
  const REAL* gc = w->gc + (size_t)lc * nstr * nstr;
  REAL* gu = w->gu + (size_t)lc * numu * nstr;
  REAL delm0 = (mazim == 0) ? 1.0 : 0.0;
  REAL* wk = w->wk;
  /* row r of GC: r<nn -> -cmu[nn-1-r] (index nn + (nn-1-r) in CMU order), else +cmu[r-nn] */
#define ROW2CMU(r) ((r) < nn ? nn + (nn - 1 - (r)) : (r) - nn)
  for (int jq = 0; jq < nstr; ++jq) {
    for (int l = mazim; l <= nstr - 1; ++l) {
      REAL sum = 0.0;
      for (int r = 0; r < nstr; ++r) {
        int c = ROW2CMU(r);
        sum += w->cwt[c] * YLMC(l, c) * gc[r * nstr + jq];
      }
      wk[l] = 0.5 * gl[l] * sum;
    }
    for (int iu = 0; iu < numu; ++iu) {
      REAL sum = 0.0;
      for (int l = mazim; l <= nstr - 1; ++l) sum += wk[l] * w->ylmu[l * numu + iu];
      gu[iu * nstr + jq] = sum;
    }
  }
  if (fbeam > 0.0) {
    const REAL* zz = w->zz + (size_t)lc * nstr;
    REAL fact = (2.0 - delm0) * fbeam / (4.0 * PI_R);
    for (int l = mazim; l <= nstr - 1; ++l) {
      REAL psum = 0.0;
      for (int r = 0; r < nstr; ++r) {
        int c = ROW2CMU(r);
        psum += w->cwt[c] * YLMC(l, c) * zz[r];
      }
      wk[l] = 0.5 * gl[l] * psum;
    }
    for (int iu = 0; iu < numu; ++iu) {
      REAL sum = 0.0;
      for (int l = mazim; l <= nstr - 1; ++l)
        sum += w->ylmu[l * numu + iu] * (wk[l] + fact * gl[l] * w->ylm0[l]);
      w->zbeam[(size_t)lc * numu + iu] = sum;
    }
  }
  if (planck && mazim == 0) {
    const REAL* zplk0 = w->zplk0 + (size_t)lc * nstr;
    const REAL* zplk1 = w->zplk1 + (size_t)lc * nstr;
    REAL* psi0 = wk;
    REAL* psi1 = wk + nstr;
    for (int l = 0; l <= nstr - 1; ++l) {
      REAL p0 = 0.0, p1 = 0.0;
      for (int r = 0; r < nstr; ++r) {
        int c = ROW2CMU(r);
        p0 += w->cwt[c] * YLMC(l, c) * zplk0[r];
        p1 += w->cwt[c] * YLMC(l, c) * zplk1[r];
      }
      psi0[l] = 0.5 * gl[l] * p0;
      psi1[l] = 0.5 * gl[l] * p1;
    }
    for (int iu = 0; iu < numu; ++iu) {
      REAL s0 = 0.0, s1 = 0.0;
      for (int l = 0; l <= nstr - 1; ++l) {
        s0 += w->ylmu[l * numu + iu] * psi0[l];
        s1 += w->ylmu[l * numu + iu] * psi1[l];
      }
      w->z0u[(size_t)lc * numu + iu] = s0 + (1.0 - w->oprim[lc]) * w->xr0[lc];
      w->z1u[(size_t)lc * numu + iu] = s1 + (1.0 - w->oprim[lc]) * w->xr1[lc];
    }
  }
#undef ROW2CMU
}

/* ------------------------------------------------------------------------- */
/* c_set_matrix (SETMTX) + c_solve0 (SOLVE0): assemble the banded system of   */
/* STWJ(20) with the scaling transformation STWJ(22), factor and solve        */
/* (LINPACK sgbco/sgbsl) for the constants of integration LL.                 */
/* Lambertian surface only (BDR = albedo, BEM = 1-albedo): pyharp always      */
/* passes "lamber" (tests/test_disort.cpp:17, examples/amars_*.cpp).          */
/* ------------------------------------------------------------------------- */
static int setmtx_solve0(work* w, int mazim, int planck, REAL fbeam, REAL umu0, REAL fisot,
                         REAL albedo, REAL tplank, REAL bplank) {
  const int nstr = w->nstr, nn = w->nn, ncut = w->ncut, lyrcut = w->lyrcut;
  const int ncol = nstr * ncut;
  const int kl = 3 * nn - 1, ku = 3 * nn - 1, bw = 2 * kl + ku + 1;
  REAL delm0 = (mazim == 0) ? 1.0 : 0.0;
  REAL* ab = w->cband;
  REAL* b = w->b;
  memset(ab, 0, sizeof(REAL) * (size_t)ncol * bw);
  memset(b, 0, sizeof(REAL) * (size_t)ncol);
#define AB(i, j) ab[(size_t)(i) * bw + ((j) - (i) + kl)]
#define GCL(lc, i, j) w->gc[((size_t)(lc) * nstr + (i)) * nstr + (j)]
#define KKL(lc, j) w->kk[(size_t)(lc) * nstr + (j)]
#define ZZL(lc, i) w->zz[(size_t)(lc) * nstr + (i)]
#define ZP0(lc, i) w->zplk0[(size_t)(lc) * nstr + (i)]
#define ZP1(lc, i) w->zplk1[(size_t)(lc) * nstr + (i)]
  const int noreflect = lyrcut || (mazim > 0); /* lamber: no azimuthal dependence */
  for (int lc = 0; lc < ncut; ++lc) {
    REAL* wk = w->wk;
    for (int iq = 0; iq < nn; ++iq) wk[iq] = R_EXP(KKL(lc, iq) * w->dtaucp[lc]); /* kk<0 */
    for (int jq = 0; jq < nstr; ++jq) {
      int col = lc * nstr + jq;
      /* scaled exponentials: negative-k terms referenced to the layer bottom,
         positive-k terms to the layer top (STWJ 22) */
      REAL etop = (jq < nn) ? wk[jq] : 1.0;
      REAL ebot = (jq < nn) ? 1.0 : wk[nstr - 1 - jq];
      if (lc == 0) { /* top boundary: downward streams */
        for (int r = 0; r < nn; ++r) AB(r, col) = GCL(0, nn - 1 - r, jq) * etop;
      } else { /* continuity with the layer above */
        int row0 = nn + (lc - 1) * nstr;
        for (int iq = 0; iq < nstr; ++iq) AB(row0 + iq, col) = -GCL(lc, iq, jq) * etop;
      }
      if (lc < ncut - 1) { /* continuity with the layer below */
        int row0 = nn + lc * nstr;
        for (int iq = 0; iq < nstr; ++iq) AB(row0 + iq, col) = GCL(lc, iq, jq) * ebot;
      } else { /* bottom boundary: upward streams */
        int row0 = nn + (ncut - 1) * nstr;
        for (int r = 0; r < nn; ++r) {
          REAL v = GCL(lc, nn + r, jq);
          if (!noreflect) {
            REAL sum = 0.0;
            for (int k = 0; k < nn; ++k)
              sum += w->cwt[k] * w->cmu[k] * albedo * GCL(lc, nn - 1 - k, jq);
            v -= (1.0 + delm0) * sum;
          }
          AB(row0 + r, col) = v * ebot;
        }
      }
    }
  }
  /* right-hand side, SOLVE0 */
  const int beam = (fbeam > 0.0);
  if (mazim > 0) {
    for (int r = 0; r < nn; ++r) {
      b[r] = -ZZL(0, nn - 1 - r);
      b[ncol - nn + r] = -ZZL(ncut - 1, nn + r) * w->expbea[ncut];
    }
    for (int lc = 0; lc < ncut - 1; ++lc)
      for (int iq = 0; iq < nstr; ++iq)
        b[nn + lc * nstr + iq] = (ZZL(lc + 1, iq) - ZZL(lc, iq)) * w->expbea[lc + 1];
  } else {
    REAL taub = w->taucpr[ncut];
    for (int r = 0; r < nn; ++r) {
      REAL v = fisot + tplank;
      if (beam) v -= ZZL(0, nn - 1 - r);
      if (planck) v -= ZP0(0, nn - 1 - r);
      b[r] = v;
    }
    for (int r = 0; r < nn; ++r) {
      int iq = nn + r;
      REAL v = 0.0;
      if (beam) v -= ZZL(ncut - 1, iq) * w->expbea[ncut];
      if (planck) v -= ZP0(ncut - 1, iq) + ZP1(ncut - 1, iq) * taub;
      if (!lyrcut) {
        REAL sum = 0.0;
        for (int k = 0; k < nn; ++k) {
          REAL t = 0.0;
          if (beam) t += ZZL(ncut - 1, nn - 1 - k) * w->expbea[ncut];
          if (planck) t += ZP0(ncut - 1, nn - 1 - k) + ZP1(ncut - 1, nn - 1 - k) * taub;
          sum += w->cwt[k] * w->cmu[k] * albedo * t;
        }
        v += 2.0 * sum + (1.0 - albedo) * bplank;
        if (beam) v += albedo * umu0 * fbeam / PI_R * w->expbea[ncut];
      }
      b[ncol - nn + r] = v;
    }
    for (int lc = 0; lc < ncut - 1; ++lc)
      for (int iq = 0; iq < nstr; ++iq) {
        REAL v = 0.0;
        if (beam) v += (ZZL(lc + 1, iq) - ZZL(lc, iq)) * w->expbea[lc + 1];
        if (planck)
          v += ZP0(lc + 1, iq) - ZP0(lc, iq) + (ZP1(lc + 1, iq) - ZP1(lc, iq)) * w->taucpr[lc + 1];
        b[nn + lc * nstr + iq] = v;
      }
  }
  if (sgbfa_solve(ab, ncol, kl, ku, b) != 0) return ST_SINGULAR;
  for (int lc = 0; lc < ncut; ++lc)
    for (int jq = 0; jq < nstr; ++jq) w->ll[(size_t)lc * nstr + jq] = b[lc * nstr + jq];
#undef AB
  return ST_OK;
}

/* ------------------------------------------------------------------------- */
/* c_fluxes (FLUXES): upward/downward fluxes, mean intensity, flux divergence */
/* ------------------------------------------------------------------------- */
static void fluxes(work* w, int planck, REAL fbeam, REAL umu0, double* rfldir, double* rfldn,
                   double* flup, double* dfdt, double* uavg) {
  const int nstr = w->nstr, nn = w->nn, ntau = w->ntau;
  for (int lu = 0; lu < ntau; ++lu) {
    rfldir[lu] = rfldn[lu] = flup[lu] = dfdt[lu] = uavg[lu] = 0.0;
    for (int iq = 0; iq < nstr; ++iq) w->u0c[lu * nstr + iq] = 0.0;
    int lyu = w->layru[lu];
    if (w->lyrcut && lyu >= w->ncut) continue; /* no radiation reaches this level */
    REAL fact = 0.0, dirint = 0.0, fldir = 0.0, rdir = 0.0;
    if (fbeam > 0.0) {
      fact = R_EXP(-w->utaupr[lu] / umu0);
      dirint = fbeam * fact;
      fldir = umu0 * (fbeam * fact);
      rdir = umu0 * fbeam * R_EXP(-w->utau[lu] / umu0);
    }
    REAL sfldn = 0.0, sflup = 0.0, savg = 0.0;
    for (int iq = 0; iq < nstr; ++iq) {
      REAL zint = 0.0;
      for (int jq = 0; jq < nn; ++jq)
        zint += GCL(lyu, iq, jq) * w->ll[(size_t)lyu * nstr + jq] *
                R_EXP(-KKL(lyu, jq) * (w->utaupr[lu] - w->taucpr[lyu + 1]));
      for (int jq = nn; jq < nstr; ++jq)
        zint += GCL(lyu, iq, jq) * w->ll[(size_t)lyu * nstr + jq] *
                R_EXP(-KKL(lyu, jq) * (w->utaupr[lu] - w->taucpr[lyu]));
      REAL u = zint;
      if (fbeam > 0.0) u += ZZL(lyu, iq) * fact;
      if (planck) u += ZP0(lyu, iq) + ZP1(lyu, iq) * w->utaupr[lu];
      w->u0c[lu * nstr + iq] = u;
      if (iq < nn) {
        int c = nn - 1 - iq;
        savg += w->cwt[c] * u;
        sfldn += w->cwt[c] * w->cmu[c] * u;
      } else {
        int c = iq - nn;
        savg += w->cwt[c] * u;
        sflup += w->cwt[c] * w->cmu[c] * u;
      }
    }
    REAL fup = 2.0 * PI_R * sflup, fdn = 2.0 * PI_R * sfldn;
    REAL fdntot = fdn + fldir;
    REAL av = (2.0 * PI_R * savg + dirint) / (4.0 * PI_R);
    REAL plsorc = planck ? w->xr0[lyu] + w->xr1[lyu] * w->utaupr[lu] : 0.0;
    flup[lu] = (double)fup;
    rfldir[lu] = (double)rdir;
    rfldn[lu] = (double)(fdntot - rdir);
    uavg[lu] = (double)av;
    dfdt[lu] = (double)((1.0 - w->ssa[lyu]) * 4.0 * PI_R * (av - plsorc));
  }
}

/* ------------------------------------------------------------------------- */
/* c_user_intensities (USRINT): azimuthal intensity components at user angles */
/* by analytic integration of the source function, STWL(35-37), incl. the     */
/* L'Hospital limits at 1+umu/umu0 ~ 0 and 1+umu*k ~ 0.                       */
/* ------------------------------------------------------------------------- */
static void usrint(work* w, int mazim, int planck, REAL fbeam, REAL umu0, REAL fisot, REAL albedo,
                   REAL tplank, REAL bplank) {
  const int nstr = w->nstr, nn = w->nn, ntau = w->ntau, numu = w->numu, ncut = w->ncut;
  const int nlyr = w->nlyr;
  REAL delm0 = (mazim == 0) ? 1.0 : 0.0;
#define GUL(lc, iu, jq) (w->gu[((size_t)(lc) * numu + (iu)) * nstr + (jq)] * w->ll[(size_t)(lc) * nstr + (jq)])
  for (int lu = 0; lu < ntau; ++lu) {
    REAL exp0 = (fbeam > 0.0) ? R_EXP(-w->utaupr[lu] / umu0) : 0.0;
    int lyu = w->layru[lu];
    for (int iu = 0; iu < numu; ++iu) {
      w->uum[lu * numu + iu] = 0.0;
      if (w->lyrcut && lyu >= ncut) continue;
      REAL mu = w->umu[iu];
      int negumu = (mu < 0.0);
      int lyrstr, lyrend;
      REAL sgn;
      if (negumu) {
        lyrstr = 0;
        lyrend = lyu - 1;
        sgn = -1.0;
      } else {
        lyrstr = lyu + 1;
        lyrend = ncut - 1;
        sgn = 1.0;
      }
      REAL palint = 0.0, plkint = 0.0;
      REAL exp1 = 0.0, exp2 = 0.0;
      for (int lc = lyrstr; lc <= lyrend; ++lc) {
        REAL dtau = w->dtaucp[lc];
        exp1 = R_EXP((w->utaupr[lu] - w->taucpr[lc]) / mu);
        exp2 = R_EXP((w->utaupr[lu] - w->taucpr[lc + 1]) / mu);
        if (planck && mazim == 0) {
          REAL f0n = sgn * (exp1 - exp2);
          REAL f1n = sgn * ((w->taucpr[lc] + mu) * exp1 - (w->taucpr[lc + 1] + mu) * exp2);
          plkint += w->z0u[(size_t)lc * numu + iu] * f0n + w->z1u[(size_t)lc * numu + iu] * f1n;
        }
        if (fbeam > 0.0) {
          REAL denom = 1.0 + mu / umu0, expn;
          if (R_FABS(denom) < 0.0001)
            expn = (dtau / umu0) * exp0;
          else
            expn = (exp1 * w->expbea[lc] - exp2 * w->expbea[lc + 1]) * sgn / denom;
          palint += w->zbeam[(size_t)lc * numu + iu] * expn;
        }
        for (int iq = 0; iq < nn; ++iq) { /* kk negative */
          REAL wkq = R_EXP(KKL(lc, iq) * dtau);
          REAL denom = 1.0 + mu * KKL(lc, iq), expn;
          if (R_FABS(denom) < 0.0001)
            expn = dtau / mu * exp2;
          else
            expn = sgn * (exp1 * wkq - exp2) / denom;
          palint += GUL(lc, iu, iq) * expn;
        }
        for (int iq = nn; iq < nstr; ++iq) { /* kk positive */
          REAL wkq = R_EXP(KKL(lc, nstr - 1 - iq) * dtau);
          REAL denom = 1.0 + mu * KKL(lc, iq), expn;
          if (R_FABS(denom) < 0.0001)
            expn = -dtau / mu * exp1;
          else
            expn = sgn * (exp1 - exp2 * wkq) / denom;
          palint += GUL(lc, iu, iq) * expn;
        }
      }
      /* contribution from the user level to the next computational level */
      REAL dtau1 = w->utaupr[lu] - w->taucpr[lyu];
      REAL dtau2 = w->utaupr[lu] - w->taucpr[lyu + 1];
      int skip = (R_FABS(dtau1) < 1.0e-6 && negumu) || (R_FABS(dtau2) < 1.0e-6 && !negumu);
      if (!skip) {
        if (negumu) exp1 = R_EXP(dtau1 / mu);
        if (!negumu) exp2 = R_EXP(dtau2 / mu);
        if (fbeam > 0.0) {
          REAL denom = 1.0 + mu / umu0, expn;
          if (R_FABS(denom) < 0.0001)
            expn = (dtau1 / umu0) * exp0;
          else if (negumu)
            expn = (exp0 - w->expbea[lyu] * exp1) / denom;
          else
            expn = (exp0 - w->expbea[lyu + 1] * exp2) / denom;
          palint += w->zbeam[(size_t)lyu * numu + iu] * expn;
        }
        REAL dtau = w->dtaucp[lyu];
        for (int iq = 0; iq < nn; ++iq) { /* kk negative */
          REAL k = KKL(lyu, iq);
          REAL denom = 1.0 + mu * k, expn;
          if (R_FABS(denom) < 0.0001)
            expn = -dtau2 / mu * exp2;
          else if (negumu)
            expn = (R_EXP(-k * dtau2) - R_EXP(k * dtau) * exp1) / denom;
          else
            expn = (R_EXP(-k * dtau2) - exp2) / denom;
          palint += GUL(lyu, iu, iq) * expn;
        }
        for (int iq = nn; iq < nstr; ++iq) { /* kk positive */
          REAL k = KKL(lyu, iq);
          REAL denom = 1.0 + mu * k, expn;
          if (R_FABS(denom) < 0.0001)
            expn = -dtau1 / mu * exp1;
          else if (negumu)
            expn = (R_EXP(-k * dtau1) - exp1) / denom;
          else
            expn = (R_EXP(-k * dtau1) - R_EXP(-k * dtau) * exp2) / denom;
          palint += GUL(lyu, iu, iq) * expn;
        }
        if (planck && mazim == 0) {
          REAL expn, fact;
          if (negumu) {
            expn = exp1;
            fact = w->taucpr[lyu] + mu;
          } else {
            expn = exp2;
            fact = w->taucpr[lyu + 1] + mu;
          }
          REAL f0n = 1.0 - expn;
          REAL f1n = w->utaupr[lu] + mu - fact * expn;
          plkint += w->z0u[(size_t)lyu * numu + iu] * f0n + w->z1u[(size_t)lyu * numu + iu] * f1n;
        }
      }
      /* intensity components attenuated at both boundaries */
      REAL bndint = 0.0;
      if (negumu && mazim == 0) {
        bndint = (fisot + tplank) * R_EXP(w->utaupr[lu] / mu);
      } else if (!negumu) {
        if (!(w->lyrcut || mazim > 0)) { /* lamber: azimuth-independent only */
          const int lb = nlyr - 1;
          REAL bnddfu = 0.0;
          for (int iq = nn - 1; iq >= 0; --iq) { /* downward streams at the bottom */
            REAL dfuint = 0.0;
            for (int jq = 0; jq < nn; ++jq) dfuint += GCL(lb, iq, jq) * w->ll[(size_t)lb * nstr + jq];
            for (int jq = nn; jq < nstr; ++jq)
              dfuint += GCL(lb, iq, jq) * w->ll[(size_t)lb * nstr + jq] *
                        R_EXP(-KKL(lb, jq) * w->dtaucp[lb]);
            if (fbeam > 0.0) dfuint += ZZL(lb, iq) * w->expbea[nlyr];
            dfuint += delm0 * (ZP0(lb, iq) + ZP1(lb, iq) * w->taucpr[nlyr]);
            int c = nn - 1 - iq;
            bnddfu += (1.0 + delm0) * albedo * w->cmu[c] * w->cwt[c] * dfuint;
          }
          REAL bnddir = 0.0;
          if (fbeam > 0.0) bnddir = umu0 * fbeam / PI_R * albedo * w->expbea[nlyr];
          bndint = (bnddfu + bnddir + delm0 * (1.0 - albedo) * bplank) *
                   R_EXP((w->utaupr[lu] - w->taucpr[nlyr]) / mu);
        }
      }
      w->uum[lu * numu + iu] = palint + plkint + bndint;
    }
  }
#undef GUL
}

/* ------------------------------------------------------------------------- */
/* c_intensity_components (CMPINT): intensities at the quadrature angles      */
/* (no user angles), STWJ(5), STWL(6)                                         */
/* ------------------------------------------------------------------------- */
static void cmpint(work* w, int mazim, int planck, REAL fbeam, REAL umu0) {
  const int nstr = w->nstr, nn = w->nn, ntau = w->ntau;
  for (int lu = 0; lu < ntau; ++lu) {
    int lyu = w->layru[lu];
    for (int iq = 0; iq < nstr; ++iq) {
      w->uum[lu * nstr + iq] = 0.0;
      if (w->lyrcut && lyu >= w->ncut) continue;
      REAL zint = 0.0;
      for (int jq = 0; jq < nn; ++jq)
        zint += GCL(lyu, iq, jq) * w->ll[(size_t)lyu * nstr + jq] *
                R_EXP(-KKL(lyu, jq) * (w->utaupr[lu] - w->taucpr[lyu + 1]));
      for (int jq = nn; jq < nstr; ++jq)
        zint += GCL(lyu, iq, jq) * w->ll[(size_t)lyu * nstr + jq] *
                R_EXP(-KKL(lyu, jq) * (w->utaupr[lu] - w->taucpr[lyu]));
      REAL u = zint;
      if (fbeam > 0.0) u += ZZL(lyu, iq) * R_EXP(-w->utaupr[lu] / umu0);
      if (planck && mazim == 0) u += ZP0(lyu, iq) + ZP1(lyu, iq) * w->utaupr[lu];
      w->uum[lu * nstr + iq] = u;
    }
  }
}
#undef GCL
#undef KKL
#undef ZZL
#undef ZP0
#undef ZP1
#undef YLMC

/* ------------------------------------------------------------------------- */
/* c_single_scat (SINSCA), c_xi_func (XIFUNC), c_secondary_scat (SECSCA) and  */
/* c_intensity_correction (INTCOR): Nakajima-Tanaka TMS/IMS corrections,      */
/* NT(1988) Eqs. 13-14, STWL(65-68, A.7-A.15).                                */
/* ------------------------------------------------------------------------- */
static REAL sinsca(REAL dither, int layru, int nlyr, const REAL* phase, const REAL* omega,
                   const REAL* tau, REAL umu, REAL umu0, REAL utau, REAL fbeam) {
  REAL s = 0.0;
  REAL exp0 = R_EXP(-utau / umu0);
  if (R_FABS(umu + umu0) <= dither) {
    for (int lyr = 0; lyr < layru; ++lyr) s += omega[lyr] * phase[lyr] * (tau[lyr + 1] - tau[lyr]);
    return fbeam / (4.0 * PI_R * umu0) * exp0 *
           (s + omega[layru] * phase[layru] * (utau - tau[layru]));
  }
  if (umu > 0.0) {
    for (int lyr = layru; lyr < nlyr; ++lyr) {
      REAL exp1 = R_EXP(-((tau[lyr + 1] - utau) / umu + tau[lyr + 1] / umu0));
      s += omega[lyr] * phase[lyr] * (exp0 - exp1);
      exp0 = exp1;
    }
  } else {
    for (int lyr = layru; lyr >= 0; --lyr) {
      REAL exp1 = R_EXP(-((tau[lyr] - utau) / umu + tau[lyr] / umu0));
      s += omega[lyr] * phase[lyr] * (exp0 - exp1);
      exp0 = exp1;
    }
  }
  return fbeam / (4.0 * PI_R * (1.0 + umu / umu0)) * s;
}

static REAL xifunc(REAL umu1, REAL umu2, REAL umu3, REAL tau) {
  REAL x1 = 1.0 / umu1 - 1.0 / umu2;
  REAL x2 = 1.0 / umu1 - 1.0 / umu3;
  REAL exp1 = R_EXP(-tau / umu1);
  if (umu2 == umu3 && umu1 == umu2) return tau * tau * exp1 / (2.0 * umu1 * umu2);
  if (umu2 == umu3 && umu1 != umu2)
    return ((tau - 1.0 / x1) * R_EXP(-tau / umu2) + exp1 / x1) / (x1 * umu1 * umu2);
  if (umu2 != umu3 && umu1 == umu2)
    return ((R_EXP(-tau / umu3) - exp1) / x2 - tau * exp1) / (x2 * umu1 * umu2);
  if (umu2 != umu3 && umu1 == umu3)
    return ((R_EXP(-tau / umu2) - exp1) / x1 - tau * exp1) / (x1 * umu1 * umu2);
  return ((R_EXP(-tau / umu3) - exp1) / x2 - (R_EXP(-tau / umu2) - exp1) / x1) / (x2 - x1) /
         (umu1 * umu2);
}

static REAL secsca(REAL ctheta, const REAL* flyr, int layru, int nmom, int nstr, const REAL* pmom,
                   const REAL* ssalb, const REAL* dtauc, const REAL* tauc, REAL umu, REAL umu0,
                   REAL utau, REAL fbeam) {
  REAL dtau = utau - tauc[layru];
  REAL wbar = ssalb[layru] * dtau;
  REAL fbar = flyr[layru] * wbar;
  REAL stau = dtau;
  for (int lyr = 0; lyr < layru; ++lyr) {
    wbar += dtauc[lyr] * ssalb[lyr];
    fbar += ssalb[lyr] * dtauc[lyr] * flyr[lyr];
    stau += dtauc[lyr];
  }
  if (wbar <= 1.0e-4 || fbar <= 1.0e-4 || stau <= 1.0e-4 || fbeam <= 1.0e-4) return 0.0;
  fbar /= wbar;
  wbar /= stau;
  REAL pspike = 1.0, plm1 = 1.0, plm2 = 0.0, pl;
  for (int k = 1; k <= nstr - 1; ++k) {
    pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
    plm2 = plm1;
    plm1 = pl;
    pspike += (2 * k + 1) * pl;
  }
  for (int k = nstr; k <= nmom; ++k) {
    pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
    plm2 = plm1;
    plm1 = pl;
    REAL gbar = pmom[(size_t)layru * (nmom + 1) + k] * ssalb[layru] * dtau;
    for (int lyr = 0; lyr < layru; ++lyr)
      gbar += pmom[(size_t)lyr * (nmom + 1) + k] * ssalb[lyr] * dtauc[lyr];
    if (fbar * wbar * stau <= 1.0e-4)
      gbar = 0.0;
    else
      gbar /= (fbar * wbar * stau);
    pspike += (2.0 * gbar - gbar * gbar) * (2 * k + 1) * pl;
  }
  REAL umu0p = umu0 / (1.0 - fbar * wbar);
  return fbeam / (4.0 * PI_R) * (fbar * wbar) * (fbar * wbar) / (1.0 - fbar * wbar) * pspike *
         xifunc(-umu, umu0p, umu0p, utau);
}

/* ------------------------------------------------------------------------- */
/* c_disort driver (DISORT main): c_disort_set (SETDIS), Planck terms,        */
/* azimuth loop, layer loop, boundary solve, fluxes, intensities, NT fix.     */
/* ------------------------------------------------------------------------- */
int orc_disort(const orc_problem* pb, const double* dtauc_in, const double* ssalb_in,
               const double* pmom_in, const double* temper_in, const double* utau_in,
               const double* umu_in, const double* phi_in, double* rfldir, double* rfldn,
               double* flup, double* dfdt, double* uavg, double* uu) {
  const int nlyr = pb->nlyr, nstr = pb->nstr, nmom = pb->nmom, nn = nstr / 2;
  const int usrtau = !!(pb->flags & ORC_USRTAU), usrang = !!(pb->flags & ORC_USRANG);
  const int planck = !!(pb->flags & ORC_PLANCK), onlyfl = !!(pb->flags & ORC_ONLYFL);
  const int lamber = !!(pb->flags & ORC_LAMBER);
  const REAL fbeam = pb->fbeam, umu0 = pb->umu0, fisot = pb->fisot, albedo = pb->albedo;
  int status = ST_OK;

  /* ---- c_check_inputs (CHEKIN): reject, never clamp ---- */
  if (nlyr < 1 || nstr < 2 || (nstr & 1) || nmom < nstr || nstr > 256) return ST_BAD_INPUT;
  if (!lamber) return ST_UNSUPPORTED; /* BRDF surfaces are not on pyharp's path */
  for (int lc = 0; lc < nlyr; ++lc) {
    if (!(dtauc_in[lc] >= 0.0)) return ST_BAD_INPUT;
    if (!(ssalb_in[lc] >= 0.0 && ssalb_in[lc] <= 1.0)) return ST_BAD_INPUT;
    for (int k = 0; k <= nmom; ++k)
      if (!(pmom_in[(size_t)lc * (nmom + 1) + k] >= -1.0 && pmom_in[(size_t)lc * (nmom + 1) + k] <= 1.0))
        return ST_BAD_INPUT;
    if (planck && !(temper_in[lc] >= 0.0)) return ST_BAD_INPUT;
  }
  if (planck && !(temper_in[nlyr] >= 0.0)) return ST_BAD_INPUT;
  if (!(fbeam >= 0.0) || !(fisot >= 0.0)) return ST_BAD_INPUT;
  if (fbeam > 0.0 && !(umu0 > 0.0 && umu0 <= 1.0)) return ST_BAD_INPUT;
  if (!(albedo >= 0.0 && albedo <= 1.0)) return ST_BAD_INPUT;
  if (planck && (!(pb->wvnmlo >= 0.0) || !(pb->wvnmhi >= pb->wvnmlo) || !(pb->btemp >= 0.0) ||
                 !(pb->ttemp >= 0.0) || !(pb->temis >= 0.0 && pb->temis <= 1.0)))
    return ST_BAD_INPUT;

  /* DISORT main program: "IF( SSALB(LC).EQ.1.0 ) SSALB(LC) = 1.0 - DITHER" with DITHER = max(10 eps, 1e-10):
     an exactly conservative layer would give a zero eigenvalue */
  double* ssalb_d = (double*)malloc(sizeof(double) * nlyr);
  for (int lc = 0; lc < nlyr; ++lc) ssalb_d[lc] = (ssalb_in[lc] == 1.0) ? 1.0 - 1.0e-10 : ssalb_in[lc];
  const double* ssalb_orig = ssalb_in;
  (void)ssalb_orig;
  ssalb_in = ssalb_d;

  work W;
  work* w = &W;
  memset(w, 0, sizeof(W));
  w->nlyr = nlyr;
  w->nstr = nstr;
  w->nn = nn;
  w->ntau = usrtau ? pb->ntau : nlyr + 1;
  w->nphi = (onlyfl) ? 0 : pb->nphi;
  const int user_angles = usrang && !onlyfl;
  w->numu = user_angles ? pb->numu : nstr;
  const int ntau = w->ntau, numu = w->numu, nphi = w->nphi;
  if (usrtau && ntau < 1) return ST_BAD_INPUT;
  if (user_angles && numu < 1) return ST_BAD_INPUT;
  if (!onlyfl && nphi < 1) return ST_BAD_INPUT;

  w->cmu = ralloc(nstr);
  w->cwt = ralloc(nstr);
  w->tauc = ralloc(nlyr + 1);
  w->taucpr = ralloc(nlyr + 1);
  w->dtaucp = ralloc(nlyr);
  w->oprim = ralloc(nlyr);
  w->flyr = ralloc(nlyr);
  w->ssa = ralloc(nlyr);
  w->gl = ralloc((size_t)nlyr * nstr);
  w->expbea = ralloc(nlyr + 1);
  w->utau = ralloc(ntau);
  w->utaupr = ralloc(ntau);
  w->layru = (int*)calloc(ntau, sizeof(int));
  w->umu = ralloc(numu);
  w->pkag = ralloc(nlyr + 1);
  w->xr0 = ralloc(nlyr);
  w->xr1 = ralloc(nlyr);
  w->ylmc = ralloc((size_t)nstr * nstr);
  w->ylm0 = ralloc(nstr);
  w->ylmu = ralloc((size_t)nstr * numu);
  w->gc = ralloc((size_t)nlyr * nstr * nstr);
  w->kk = ralloc((size_t)nlyr * nstr);
  w->ll = ralloc((size_t)nlyr * nstr);
  w->zz = ralloc((size_t)nlyr * nstr);
  w->zplk0 = ralloc((size_t)nlyr * nstr);
  w->zplk1 = ralloc((size_t)nlyr * nstr);
  w->gu = ralloc((size_t)nlyr * numu * nstr);
  w->zbeam = ralloc((size_t)nlyr * numu);
  w->z0u = ralloc((size_t)nlyr * numu);
  w->z1u = ralloc((size_t)nlyr * numu);
  w->zj = ralloc(nstr);
  w->z0 = ralloc(nstr);
  w->z1 = ralloc(nstr);
  w->cc = ralloc((size_t)nstr * nstr);
  w->amb = ralloc((size_t)nn * nn);
  w->apb = ralloc((size_t)nn * nn);
  w->arr = ralloc((size_t)nstr * nstr);
  w->evecc = ralloc((size_t)nstr * nstr);
  w->eval = ralloc(nstr);
  w->cband = ralloc((size_t)nstr * nlyr * (9 * nn - 2));
  w->b = ralloc((size_t)nstr * nlyr);
  w->u0c = ralloc((size_t)ntau * nstr);
  w->uum = ralloc((size_t)ntau * (numu > nstr ? numu : nstr));
  w->wk = ralloc(4 * nstr + 4);
  REAL* phirad = ralloc(nphi > 0 ? nphi : 1);
  REAL* uur = ralloc((size_t)(nphi > 0 ? nphi : 1) * ntau * numu);

  /* ---- c_disort_set (SETDIS) ---- */
  qgausn(nn, w->cmu, w->cwt);
  for (int iq = 0; iq < nn; ++iq) {
    w->cmu[iq + nn] = -w->cmu[iq];
    w->cwt[iq + nn] = w->cwt[iq];
  }
  if (fbeam > 0.0) {
    /* umu0 must not coincide with a quadrature angle (fatal in DISORT 2) */
    for (int iq = 0; iq < nn; ++iq)
      if (R_FABS((umu0 - w->cmu[iq]) / umu0) < 1.0e-4) {
        status = ST_BEAM_ANGLE;
        goto cleanup;
      }
  }
  w->tauc[0] = 0.0;
  for (int lc = 0; lc < nlyr; ++lc) {
    w->tauc[lc + 1] = w->tauc[lc] + (REAL)dtauc_in[lc];
    w->ssa[lc] = ssalb_in[lc];
  }
  {
    const REAL abscut = 10.0;
    REAL abstau = 0.0;
    w->expbea[0] = (fbeam > 0.0) ? 1.0 : 0.0;
    w->taucpr[0] = 0.0;
    w->ncut = nlyr;
    for (int lc = 0; lc < nlyr; ++lc) {
      const double* pm = pmom_in + (size_t)lc * (nmom + 1);
      if (abstau < abscut) w->ncut = lc + 1;
      abstau += (1.0 - (REAL)ssalb_in[lc]) * (REAL)dtauc_in[lc];
      REAL f = pm[nstr]; /* delta-M: separated fraction = moment nstr */
      REAL ss = ssalb_in[lc];
      w->oprim[lc] = ss * (1.0 - f) / (1.0 - f * ss);
      w->dtaucp[lc] = (1.0 - f * ss) * (REAL)dtauc_in[lc];
      w->taucpr[lc + 1] = w->taucpr[lc] + w->dtaucp[lc];
      w->flyr[lc] = f;
      for (int k = 0; k < nstr; ++k) {
        REAL pk = (k == 0) ? 1.0 : (REAL)pm[k]; /* pmom[0] is forced to 1 */
        w->gl[(size_t)lc * nstr + k] = (2 * k + 1) * w->oprim[lc] * (pk - f) / (1.0 - f);
      }
      w->expbea[lc + 1] = (fbeam > 0.0) ? R_EXP(-w->taucpr[lc + 1] / umu0) : 0.0;
    }
    /* no thermal emission: cut the medium below absorption optical depth abscut */
    w->lyrcut = (abstau >= abscut && !planck && nlyr > 1);
    if (!w->lyrcut) w->ncut = nlyr;
  }
  /* user optical depths -> layer index and delta-M scaled depth */
  for (int lu = 0; lu < ntau; ++lu) {
    REAL ut = usrtau ? (REAL)utau_in[lu] : w->tauc[lu];
    if (usrtau && (ut < 0.0 || ut > w->tauc[nlyr] * (1.0 + 1e-12) + 1e-300)) {
      status = ST_BAD_INPUT;
      goto cleanup;
    }
    w->utau[lu] = ut;
    int lc;
    for (lc = 0; lc < nlyr - 1; ++lc)
      if (ut >= w->tauc[lc] && ut <= w->tauc[lc + 1]) break;
    w->layru[lu] = lc;
    w->utaupr[lu] = w->taucpr[lc] + (1.0 - w->ssa[lc] * w->flyr[lc]) * (ut - w->tauc[lc]);
  }
  /* output polar angles */
  if (user_angles) {
    for (int iu = 0; iu < numu; ++iu) {
      w->umu[iu] = umu_in[iu];
      if (!(w->umu[iu] >= -1.0 && w->umu[iu] <= 1.0) || w->umu[iu] == 0.0) {
        status = ST_BAD_INPUT;
        goto cleanup;
      }
    }
  } else {
    for (int iu = 0; iu < nn; ++iu) {
      w->umu[iu] = -w->cmu[nn - 1 - iu];
      w->umu[iu + nn] = w->cmu[iu];
    }
  }

  /* ---- Planck functions ---- */
  REAL tplank = 0.0, bplank = 0.0;
  if (planck) {
    tplank = pb->temis * plkavg(pb->wvnmlo, pb->wvnmhi, pb->ttemp);
    bplank = plkavg(pb->wvnmlo, pb->wvnmhi, pb->btemp);
    for (int lev = 0; lev <= nlyr; ++lev) w->pkag[lev] = plkavg(pb->wvnmlo, pb->wvnmhi, temper_in[lev]);
  }

  /* ---- azimuth loop ---- */
  int naz = nstr - 1;
  if (fbeam == 0.0 || R_FABS(1.0 - umu0) < 1.0e-5 || onlyfl ||
      (numu == 1 && R_FABS(1.0 - w->umu[0]) < 1.0e-5) ||
      (numu == 1 && R_FABS(1.0 + w->umu[0]) < 1.0e-5) ||
      (numu == 2 && R_FABS(1.0 + w->umu[0]) < 1.0e-5 && R_FABS(1.0 - w->umu[1]) < 1.0e-5))
    naz = 0;
  int kconv = 0;
  for (int mazim = 0; mazim <= naz; ++mazim) {
    if (fbeam > 0.0) {
      REAL ang = -umu0;
      lepoly(1, mazim, nstr - 1, &ang, w->ylm0);
    }
    if (user_angles) lepoly(numu, mazim, nstr - 1, w->umu, w->ylmu);
    {
      /* positive quadrature angles, then Y(-mu) = (-1)^(l-m) Y(mu) */
      REAL* tmp = ralloc((size_t)nstr * nn);
      lepoly(nn, mazim, nstr - 1, w->cmu, tmp);
      REAL sgn = -1.0;
      for (int l = mazim; l <= nstr - 1; ++l) {
        sgn = -sgn;
        for (int iq = 0; iq < nn; ++iq) {
          w->ylmc[l * nstr + iq] = tmp[l * nn + iq];
          w->ylmc[l * nstr + iq + nn] = sgn * tmp[l * nn + iq];
        }
      }
      free(tmp);
    }
    for (int lc = 0; lc < w->ncut; ++lc) {
      if ((status = soleig(w, lc, mazim)) != ST_OK) goto cleanup;
      if (fbeam > 0.0)
        if ((status = upbeam(w, lc, mazim, fbeam, umu0)) != ST_OK) goto cleanup;
      if (planck && mazim == 0) {
        w->xr1[lc] = 0.0;
        if (w->dtaucp[lc] > 0.0) w->xr1[lc] = (w->pkag[lc + 1] - w->pkag[lc]) / w->dtaucp[lc];
        w->xr0[lc] = w->pkag[lc] - w->xr1[lc] * w->taucpr[lc];
        if ((status = upisot(w, lc)) != ST_OK) goto cleanup;
      }
      if (user_angles) terp(w, lc, mazim, planck, fbeam);
    }
    if ((status = setmtx_solve0(w, mazim, planck, fbeam, umu0, fisot, albedo, tplank, bplank)) != ST_OK)
      goto cleanup;
    if (mazim == 0) fluxes(w, planck, fbeam, umu0, rfldir, rfldn, flup, dfdt, uavg);
    if (onlyfl) break;
    if (user_angles)
      usrint(w, mazim, planck, fbeam, umu0, fisot, albedo, tplank, bplank);
    else
      cmpint(w, mazim, planck, fbeam, umu0);
    if (mazim == 0) {
      for (int j = 0; j < nphi; ++j)
        for (int lu = 0; lu < ntau; ++lu)
          for (int iu = 0; iu < numu; ++iu)
            uur[((size_t)j * ntau + lu) * numu + iu] = w->uum[lu * numu + iu];
      for (int j = 0; j < nphi; ++j) phirad[j] = (PI_R / 180.0) * ((REAL)phi_in[j] - (REAL)pb->phi0);
    } else {
      REAL azerr = 0.0;
      for (int j = 0; j < nphi; ++j) {
        REAL cosphi = R_COS(mazim * phirad[j]);
        for (int lu = 0; lu < ntau; ++lu)
          for (int iu = 0; iu < numu; ++iu) {
            REAL azterm = w->uum[lu * numu + iu] * cosphi;
            REAL* u = &uur[((size_t)j * ntau + lu) * numu + iu];
            *u += azterm;
            /* RATIO(|azterm|, |uu|) with DISORT's guarded division */
            REAL a = R_FABS(azterm), bb = R_FABS(*u), r;
            if (bb == 0.0)
              r = (a == 0.0) ? 1.0 : (REAL)DBL_MAX;
            else
              r = a / bb;
            if (a == 0.0) r = 0.0;
            if (r > azerr) azerr = r;
          }
      }
      if (azerr <= pb->accur) ++kconv;
      if (kconv >= 2) break;
    }
  }

  /* ---- c_intensity_correction (INTCOR), Nakajima/Tanaka ---- */
  if (!onlyfl && (pb->flags & ORC_INTCOR) && fbeam > 0.0) {
    if (!(pb->flags & ORC_OLDINTCOR)) {
      /* Buras-Emde "new" correction is not restated; only meaningful when delta-M
         truncates (f>0).  With f == 0 every correction is identically zero. */
      int anyf = 0;
      for (int lc = 0; lc < nlyr; ++lc)
        if (w->flyr[lc] != 0.0) anyf = 1;
      if (anyf) {
        status = ST_UNSUPPORTED;
        goto cleanup;
      }
    }
    const REAL dither = 10.0 * R_EPS;
    REAL* phasa = ralloc(nlyr);
    REAL* phasm = ralloc(nlyr);
    REAL* phast = ralloc(nlyr);
    REAL* ssr = ralloc(nlyr);
    REAL* dtr = ralloc(nlyr);
    REAL* pmr = ralloc((size_t)nlyr * (nmom + 1));
    for (int lc = 0; lc < nlyr; ++lc) {
      ssr[lc] = ssalb_in[lc];
      dtr[lc] = dtauc_in[lc];
      for (int k = 0; k <= nmom; ++k) pmr[(size_t)lc * (nmom + 1) + k] = (k == 0) ? 1.0 : pmom_in[(size_t)lc * (nmom + 1) + k];
    }
    const int ncut = w->ncut;
    for (int iu = 0; iu < numu; ++iu)
      for (int jp = 0; jp < nphi; ++jp) {
        REAL mu = w->umu[iu];
        REAL ctheta = -umu0 * mu + R_SQRT((1.0 - umu0 * umu0) * (1.0 - mu * mu)) * R_COS(phirad[jp]);
        for (int lc = 0; lc < ncut; ++lc) {
          phasa[lc] = 1.0;
          phasm[lc] = 1.0;
        }
        REAL plm1 = 1.0, plm2 = 0.0;
        for (int k = 1; k <= nmom; ++k) {
          REAL pl = ((2 * k - 1) * ctheta * plm1 - (k - 1) * plm2) / k;
          plm2 = plm1;
          plm1 = pl;
          for (int lc = 0; lc < ncut; ++lc) phasa[lc] += (2 * k + 1) * pl * pmr[(size_t)lc * (nmom + 1) + k];
          if (k <= nstr - 1)
            for (int lc = 0; lc < ncut; ++lc)
              phasm[lc] += (2 * k + 1) * pl * (pmr[(size_t)lc * (nmom + 1) + k] - w->flyr[lc]) / (1.0 - w->flyr[lc]);
        }
        for (int lc = 0; lc < ncut; ++lc) phast[lc] = phasa[lc] / (1.0 - w->flyr[lc] * ssr[lc]);
        for (int lu = 0; lu < ntau; ++lu) {
          if (w->lyrcut && w->layru[lu] >= ncut - 1) continue;
          REAL ussndm = sinsca(dither, w->layru[lu], ncut, phasm, w->oprim, w->taucpr, mu, umu0, w->utaupr[lu], fbeam);
          REAL ussp = sinsca(dither, w->layru[lu], ncut, phast, ssr, w->tauc, mu, umu0, w->utau[lu], fbeam);
          REAL duims = 0.0;
          if (mu < 0.0)
            duims = secsca(ctheta, w->flyr, w->layru[lu], nmom, nstr, pmr, ssr, dtr, w->tauc, mu, umu0, w->utau[lu], fbeam);
          uur[((size_t)jp * ntau + lu) * numu + iu] += ussp - ussndm - duims;
        }
      }
    free(phasa);
    free(phasm);
    free(phast);
    free(ssr);
    free(dtr);
    free(pmr);
  }
  if (!onlyfl && uu)
    for (size_t i = 0; i < (size_t)nphi * ntau * numu; ++i) uu[i] = (double)uur[i];

cleanup:
  free(w->cmu); free(w->cwt); free(w->tauc); free(w->taucpr); free(w->dtaucp); free(w->oprim);
  free(w->flyr); free(w->ssa); free(w->gl); free(w->expbea); free(w->utau); free(w->utaupr);
  free(w->layru); free(w->umu); free(w->pkag); free(w->xr0); free(w->xr1); free(w->ylmc);
  free(w->ylm0); free(w->ylmu); free(w->gc); free(w->kk); free(w->ll); free(w->zz);
  free(w->zplk0); free(w->zplk1); free(w->gu); free(w->zbeam); free(w->z0u); free(w->z1u);
  free(w->zj); free(w->z0); free(w->z1); free(w->cc); free(w->amb); free(w->apb); free(w->arr);
  free(w->evecc); free(w->eval); free(w->cband); free(w->b); free(w->u0c); free(w->uum);
  free(w->wk); free(phirad); free(uur);
  free(ssalb_d);
  return status;
}

/* ------------------------------------------------------------------------- */
/* pydisort DisortImpl::forward restated: marshal one        */
/* (wave, column) at a time, flipping pyharp's bottom-up layers/levels into   */
/* DISORT's top-down order (src/radiation/radiation_band_exchanger.cpp_:84-94 */
/* and src/rtsolver/rt_solver_disort.cpp_:172-181 show the same flip).        */
/* ------------------------------------------------------------------------- */
int orc_forward(int nwave, int ncol, int nlyr, int nstr, int nmom, int nprop, int flags,
                double accur, int ntau, int numu, int nphi, const double* utau, const double* umu,
                const double* phi, const double* wave_lower, const double* wave_upper,
                const double* prop, const orc_bc* bc, const double* temf, double* flx, double* uu,
                int* status, int nthreads) {
  const int usrtau = !!(flags & ORC_USRTAU);
  const int onlyfl = !!(flags & ORC_ONLYFL);
  const int planck = !!(flags & ORC_PLANCK);
  const int user_angles = (flags & ORC_USRANG) && !onlyfl;
  const int ntau_eff = usrtau ? ntau : nlyr + 1;
  const int numu_eff = user_angles ? numu : nstr;
  const int nm = nmom > nstr ? nmom : nstr; /* nmom_nstr */
  int nbad = 0;
  long nsolve = (long)nwave * ncol;
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 1 ? nthreads : 1) reduction(+ : nbad)
#endif
  for (long s = 0; s < nsolve; ++s) {
    int n = (int)(s / ncol), j = (int)(s % ncol);
    double* dtauc = (double*)malloc(sizeof(double) * nlyr);
    double* ssalb = (double*)malloc(sizeof(double) * nlyr);
    double* pmom = (double*)calloc((size_t)nlyr * (nm + 1), sizeof(double));
    double* temper = (double*)calloc(nlyr + 1, sizeof(double));
    double* out = (double*)calloc((size_t)5 * ntau_eff, sizeof(double));
    double* uus = (!onlyfl) ? (double*)calloc((size_t)nphi * ntau_eff * numu_eff, sizeof(double)) : NULL;
    for (int i = 0; i < nlyr; ++i) {
      const double* p = prop + (((size_t)n * ncol + j) * nlyr + i) * nprop;
      int lc = nlyr - 1 - i;
      dtauc[lc] = p[0];
      ssalb[lc] = (nprop > 1) ? p[1] : 0.0;
      pmom[(size_t)lc * (nm + 1)] = 1.0;
      for (int k = 0; k + 2 < nprop && k + 1 <= nm; ++k) pmom[(size_t)lc * (nm + 1) + 1 + k] = p[2 + k];
    }
    if (planck && temf)
      for (int i = 0; i <= nlyr; ++i) temper[nlyr - i] = temf[(size_t)j * (nlyr + 1) + i];
    orc_problem pb;
    memset(&pb, 0, sizeof(pb));
    pb.nlyr = nlyr;
    pb.nstr = nstr;
    pb.nmom = nm;
    pb.ntau = ntau;
    pb.numu = numu;
    pb.nphi = nphi;
    pb.flags = flags;
    pb.accur = accur;
    pb.wvnmlo = wave_lower ? wave_lower[n] : 0.0;
    pb.wvnmhi = wave_upper ? wave_upper[n] : 0.0;
    size_t o = (size_t)n * ncol + j;
    pb.umu0 = (bc && bc->umu0) ? bc->umu0[o] : 1.0;
    pb.fbeam = (bc && bc->fbeam) ? bc->fbeam[o] : 0.0;
    pb.albedo = (bc && bc->albedo) ? bc->albedo[o] : 0.0;
    pb.btemp = (bc && bc->btemp) ? bc->btemp[o] : 0.0;
    pb.ttemp = (bc && bc->ttemp) ? bc->ttemp[o] : 0.0;
    pb.temis = (bc && bc->temis) ? bc->temis[o] : 0.0;
    pb.fisot = (bc && bc->fisot) ? bc->fisot[o] : 0.0;
    pb.phi0 = (bc && bc->phi0) ? bc->phi0[o] : 0.0;
    int st = orc_disort(&pb, dtauc, ssalb, pmom, temper, utau, umu, phi, out, out + ntau_eff,
                        out + 2 * ntau_eff, out + 3 * ntau_eff, out + 4 * ntau_eff, uus);
    if (status) status[o] = st;
    if (st != 0) ++nbad;
    if (flx) {
      if (!usrtau) {
        for (int i = 0; i <= nlyr; ++i) {
          int m = nlyr - i;
          double* f = flx + (o * (nlyr + 1) + i) * 2;
          f[0] = st ? NAN : out[2 * ntau_eff + m];
          f[1] = st ? NAN : out[m] + out[ntau_eff + m];
        }
      } else {
        /* user levels: same bottom-up flip applied to rad[0..ntau) (for test_disort's
           utau = layer boundaries this is identical to the branch above) */
        for (int i = 0; i < ntau_eff; ++i) {
          int lu = ntau_eff - 1 - i;
          double* f = flx + (o * ntau_eff + i) * 2;
          f[0] = st ? NAN : out[2 * ntau_eff + lu];
          f[1] = st ? NAN : out[lu] + out[ntau_eff + lu];
        }
      }
    }
    if (uu && uus)
      memcpy(uu + o * (size_t)nphi * ntau_eff * numu_eff, uus, sizeof(double) * (size_t)nphi * ntau_eff * numu_eff);
    free(dtauc);
    free(ssalb);
    free(pmom);
    free(temper);
    free(out);
    free(uus);
  }
  return nbad;
}
