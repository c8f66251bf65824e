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

static int g_layerwise = 0;  // hostsim_set_layerwise: run the warp-per-solve order (layer operators first, then the sweep)

template <int NN>
static void run(SolveParams& p, const QuadTables& qt) {
  const long nsolve = (long)p.nwave * p.ncol;
  constexpr int NSAVE = NN * NN + 2 * NN + 1;
  if (g_layerwise) {
    if constexpr (NN <= 4) {  // (the warp-per-solve kernel is instantiated for nstr <= 8)
      std::vector<double> ops((size_t)p.nlyr * kOpsDoubles<NN>), rec((size_t)p.nlyr * NSAVE);
      p.s0 = 0;
      p.nlocal = nsolve;
      const bool band = p.bflx_t != nullptr;
      for (long s = 0; s < nsolve; ++s) {
        if (p.nprop <= 2) {
          if (band) solve_flux_layerwise<NN, true, 1>(p, qt, s, ops.data(), rec.data());
          else solve_flux_layerwise<NN, true, 0>(p, qt, s, ops.data(), rec.data());
        } else {
          if (band) solve_flux_layerwise<NN, false, 1>(p, qt, s, ops.data(), rec.data());
          else solve_flux_layerwise<NN, false, 0>(p, qt, s, ops.data(), rec.data());
        }
      }
      return;
    }
  }
  // same driver structure as disort_flux_kernel (flux_inst.cu): where the kernel carries the back sweep of one
  // solve into the thread's next solve, so does this loop (alternating workspace halves, drain at the end)
  const size_t half = (size_t)p.nlyr * NSAVE;
  std::vector<double> ws(2 * half);
  p.ws = ws.data();
  p.ws_threads = 1;
  p.ws_half = (long)half;
  p.s0 = 0;
  p.nlocal = nsolve;
  unsigned ring_pos = 0;  // (device only: position in the warp's shared-memory ring of back-sweep records)
  BackPend<NN> pend;
  int it = 0;
  const bool band = p.bflx_t != nullptr;  // fused band integration (SolveParams::bflx_t)
  for (long s = 0; s < nsolve; ++s, ++it) {  // same dispatch as flux_inst.cu
    if (p.nprop <= 2 && NN <= 4) {
      if (band) solve_flux_one<NN, true, 1>(p, qt, s, 0, true, ring_pos, pend, it & 1);
      else solve_flux_one<NN, true, 0>(p, qt, s, 0, true, ring_pos, pend, it & 1);
    } else {
      if (band) solve_flux_one<NN, false, 1>(p, qt, s, 0, true, ring_pos, pend, it & 1);
      else solve_flux_one<NN, false, 0>(p, qt, s, 0, true, ring_pos, pend, it & 1);
    }
  }
  if constexpr (kCarryBack<NN>) {
    if (band) drain_back_sweep<NN, 1>(p, pend, qt, 0, ring_pos);
    else drain_back_sweep<NN, 0>(p, pend, qt, 0, ring_pos);
  }
}

extern "C" void hostsim_set_layerwise(int on) { g_layerwise = on; }

extern "C" int hostsim_flux(int nwave, int ncol, int nlyr, int nstr, int nprop, unsigned flags,
                            const double* prop, const double* umu0, const double* fbeam,
                            const double* albedo, const double* btemp, const double* ttemp,
                            const double* temis, const double* fisot, const double* temf,
                            const double* wave_lower, const double* wave_upper, double* flx,
                            int* status, const double* weights, double* bflx_t, int bc_bcast) {
  QuadTables qt;
  build_quad_tables(nstr, &qt);
  SolveParams p{};
  p.weights = weights; p.bflx_t = bflx_t; p.bc_bcast = bc_bcast;  // band mode when bflx_t != NULL (zeroed by the caller)
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
    case 6: run<3>(p, qt); break;
    case 8: run<4>(p, qt); break;
    case 12: run<6>(p, qt); break;
    case 16: run<8>(p, qt); break;
    case 24: run<12>(p, qt); break;
    case 32: run<16>(p, qt); break;
    default: return 1;
  }
  return 0;
}

// ---- general (intensity / user-level) path: pyharp_b200/csrc/intensity_device.cuh on the host ----
#include "../../pyharp_b200/csrc/intensity_device.cuh"

extern "C" int hostsim_intensity(int nwave, int ncol, int nlyr, int nstr, int nmom, int nprop, unsigned flags,
                                 double accur, int ntau, int numu, int nphi, const double* utau,
                                 const double* umu, const double* phi, const double* prop, const double* umu0,
                                 const double* fbeam, const double* albedo, const double* btemp,
                                 const double* ttemp, const double* temis, const double* fisot,
                                 const double* phi0, const double* temf, const double* wave_lower,
                                 const double* wave_upper, double* flx, double* uu, int* status) {
  QuadTables qt;
  build_quad_tables(nstr, &qt);
  IntensityParams P{};
  P.nwave = nwave; P.ncol = ncol; P.nlyr = nlyr; P.nprop = nprop; P.nstr = nstr; P.nmom = nmom;
  const bool onlyfl = flags & F_ONLYFL;
  P.ntau = (flags & F_USRTAU) ? ntau : nlyr + 1;
  P.numu = ((flags & F_USRANG) && !onlyfl) ? numu : nstr;
  P.nphi = nphi;
  P.flags = flags; P.accur = accur;
  P.prop = prop; P.umu0 = umu0; P.fbeam = fbeam; P.albedo = albedo; P.btemp = btemp; P.ttemp = ttemp;
  P.temis = temis; P.fisot = fisot; P.phi0 = phi0; P.temf = temf;
  P.wave_lower = wave_lower; P.wave_upper = wave_upper;
  P.utau = utau; P.umu = umu; P.phi = phi;
  P.cmu = qt.mu; P.cwt = qt.wt;
  P.flx = flx; P.uu = onlyfl ? nullptr : uu; P.status = status;
  const IntensityWs M = intensity_ws_map(nlyr, nstr, P.ntau, P.numu, 1);
  std::vector<double> ws((size_t)M.total);
  P.ws = ws.data(); P.ws_per_cta = M.total; P.s0 = 0; P.nlocal = (long)nwave * ncol;
  for (long s = 0; s < P.nlocal; ++s) solve_intensity_one(P, s, ws.data(), 0, 1);
  return 0;
}

// ---- small building blocks of the flux kernel, exposed for unit tests ----
extern "C" int hostsim_pivot_bad(double d) {
  PivotCheck c;
  c.add(d);
  return c.bad() ? 1 : 0;
}

extern "C" void hostsim_jacobi_rotation(double gpq, double gpp, double gqq, double eps2, double* tsc) {
  jacobi_rotation(gpq, gpp, gqq, eps2, tsc[0], tsc[1], tsc[2]);
}
