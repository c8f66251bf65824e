// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the device math of
// pyharp_b200/csrc/disort_device.cuh for the host so the per-solve algorithm can be
// checked against the oracle on a machine without a GPU.  Never linked into, or
// loaded by, the product library; there is no CPU path in the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../pyharp_b200/csrc/disort_device.cuh"
#include "../../pyharp_b200/csrc/host_tables.hpp"

using namespace b200;

template <int NN>
static void run(SolveParams& p, const QuadTables& qt) {
  const long nsolve = (long)p.nwave * p.ncol;
  constexpr int NSAVE = NN * NN + 2 * NN + 1;
  std::vector<double> ws((size_t)p.nlyr * NSAVE);
  p.ws = ws.data();
  p.ws_threads = 1;
  p.s0 = 0;
  p.nlocal = nsolve;
  for (long s = 0; s < nsolve; ++s) solve_flux_one<NN>(p, qt, s, 0);
}

extern "C" int hostsim_flux(int nwave, int ncol, int nlyr, int nstr, int nprop, unsigned flags,
                            const double* prop, const double* umu0, const double* fbeam,
                            const double* albedo, const double* btemp, const double* ttemp,
                            const double* temis, const double* fisot, const double* temf,
                            const double* wave_lower, const double* wave_upper, double* flx,
                            int* status) {
  QuadTables qt;
  build_quad_tables(nstr, &qt);
  SolveParams p{};
  p.nwave = nwave; p.ncol = ncol; p.nlyr = nlyr; p.nprop = nprop; p.nstr = nstr; p.flags = flags;
  p.prop = prop; p.umu0 = umu0; p.fbeam = fbeam; p.albedo = albedo; p.btemp = btemp;
  p.ttemp = ttemp; p.temis = temis; p.fisot = fisot; p.flx = flx; p.status = status;
  p.wave_lower = wave_lower; p.wave_upper = wave_upper;
  std::vector<int> w2b(nwave);
  std::vector<double> pkag;
  if (flags & F_PLANCK) {
    pkag.resize((size_t)nwave * ncol * (nlyr + 1));
    for (int n = 0; n < nwave; ++n) {
      w2b[n] = n;
      for (int j = 0; j < ncol; ++j)
        for (int lev = 0; lev <= nlyr; ++lev)
          pkag[((size_t)n * ncol + j) * (nlyr + 1) + lev] =
              plkavg(wave_lower[n], wave_upper[n], temf[(size_t)j * (nlyr + 1) + (nlyr - lev)]);
    }
    p.pkag = pkag.data();
    p.wave2band = w2b.data();
  }
  switch (nstr) {
    case 2: run<1>(p, qt); break;
    case 4: run<2>(p, qt); break;
    case 8: run<4>(p, qt); break;
    case 16: run<8>(p, qt); break;
    case 32: run<16>(p, qt); break;
    default: return 1;
  }
  return 0;
}
