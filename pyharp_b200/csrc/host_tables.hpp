// host_tables.hpp -- host-side construction of the quadrature-dependent constant tables.
// Double-Gauss quadrature of DISORT (c_gaussian_quadrature: nn-point Gauss-Legendre mapped
// to (0,1)) and the scaled Legendre table used by the kernels.
#pragma once
#include <cmath>
#include <vector>

#include "disort_device.cuh"

namespace b200 {

// n-point Gauss-Legendre on (0,1), ascending abscissae.  Newton iteration on P_n with the
// Davis-Rabinowitz starting guess; converged to the last bit, so it agrees with any other
// correct implementation to ~1 ulp.
inline void gauss_legendre_01(int n, double* x, double* w) {
  const double pi = 3.14159265358979323846;
  for (int k = 1; k <= (n + 1) / 2; ++k) {
    double z = std::cos(pi * (k - 0.25) / (n + 0.5));
    double pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    // final derivative at the converged root
    {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
    }
    double wt = 2.0 / ((1.0 - z * z) * pp * pp);
    // roots come out descending in z; map to (0,1)
    x[n - k] = 0.5 * z + 0.5;
    x[k - 1] = 0.5 * (-z) + 0.5;
    w[n - k] = 0.5 * wt;
    w[k - 1] = 0.5 * wt;
  }
}

inline void build_quad_tables(int nstr, QuadTables* qt) {
  const int nn = nstr / 2;
  std::vector<double> x(nn), w(nn);
  gauss_legendre_01(nn, x.data(), w.data());
  *qt = QuadTables{};
  for (int i = 0; i < nn; ++i) {
    qt->mu[i] = x[i];
    qt->rmu[i] = 1.0 / x[i];
    qt->wt[i] = w[i];
    qt->t[i] = std::sqrt(x[i] * w[i]);
    qt->ehat[i] = std::sqrt(w[i] / x[i]);
    // Legendre P_l(mu_i), l = 0..nstr-1 (c_legendre_poly, m = 0)
    double pm2 = 1.0, pm1 = x[i];
    for (int l = 0; l < nstr; ++l) {
      double pl;
      if (l == 0) pl = 1.0;
      else if (l == 1) pl = x[i];
      else {
        pl = ((2 * l - 1) * x[i] * pm1 - (l - 1) * pm2) / l;
        pm2 = pm1;
        pm1 = pl;
      }
      qt->yhat[l][i] = qt->ehat[i] * pl;
    }
  }
}

}  // namespace b200
