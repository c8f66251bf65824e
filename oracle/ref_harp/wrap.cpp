// wrap.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" doorways into the REFERENCE's own code for the
// (f) rows of SURVEY.md section 8, compiled together with the reference sources where they lie under
// /root/reference (oracle/ref_harp/Makefile -> oracle/_ref/libharp_ref.so).  Nothing here restates
// an algorithm: every function forwards to harp:: code (src/math/interpn.h, src/opacity/s8_fuller.cpp,
// h2so4_simple.cpp, rfm.cpp, src/utils/layer2level.cpp).  Used by tests/ and by
// tests/golden/make_golden_opacity.py to generate the committed fixtures.
#include <torch/torch.h>

#include <math/interpn.h>
#include <opacity/h2so4_simple.hpp>
#include <opacity/rfm.hpp>
#include <opacity/s8_fuller.hpp>
#include <utils/layer2level.hpp>
#include <utils/spherical_flux_correction.hpp>

static torch::Tensor wrapd(const double* p, std::vector<int64_t> shape) {
  return torch::from_blob(const_cast<double*>(p), shape, torch::kFloat64).clone();
}

extern "C" int ref_interpn(int ndim, int nval, const int64_t* dims, const double* axis, const double* data,
                           const double* coord, long npoint, double* out) {
  if (nval > 8) return 1;
  for (long p = 0; p < npoint; ++p) harp::interpn<8>(out + p * nval, coord + p * ndim, data, axis, dims, ndim, nval);
  return 0;
}

// type: 0 = S8Fuller, 1 = H2SO4Simple.  wave is wavelength (um) when by_wavenumber == 0, else wavenumber (1/cm).
// Returns the table too (kwave[rows], kdata[rows*2], after the species-weight scaling of reset()).
extern "C" int ref_aerosol_forward(int type, const char* file, const double* species_weights, int nspecies, int species_id,
                                   const double* wave, int nwave, int by_wavenumber, const double* conc, int ncol, int nlyr,
                                   double* out, double* kwave_out, double* kdata_out, int* rows_out) {
  try {
    harp::AttenuatorOptions op;
    op.species_weights(std::vector<double>(species_weights, species_weights + nspecies));
    op.species_names(std::vector<std::string>(nspecies, "X"));
    op.species_ids({species_id}).opacity_files({std::string(file)});
    std::map<std::string, torch::Tensor> kw;
    kw[by_wavenumber ? "wavenumber" : "wavelength"] = wrapd(wave, {nwave});
    auto c = wrapd(conc, {ncol, nlyr, nspecies});
    torch::Tensor res, kwave, kdata;
    if (type == 0) {
      harp::S8Fuller m(op);
      res = m->forward(c, kw); kwave = m->kwave; kdata = m->kdata;
    } else {
      harp::H2SO4Simple m(op);
      res = m->forward(c, kw); kwave = m->kwave; kdata = m->kdata;
    }
    res = res.contiguous();
    std::memcpy(out, res.data_ptr<double>(), sizeof(double) * res.numel());
    const int rows = kwave.size(0);
    if (rows_out) *rows_out = rows;
    if (kwave_out) std::memcpy(kwave_out, kwave.contiguous().data_ptr<double>(), sizeof(double) * rows);
    if (kdata_out) std::memcpy(kdata_out, kdata.contiguous().data_ptr<double>(), sizeof(double) * rows * 2);
    return 0;
  } catch (std::exception const& e) {
    std::fprintf(stderr, "ref_aerosol_forward: %s\n", e.what());
    return 2;
  }
}

// RFMImpl::forward on an in-memory table (the NetCDF reader is compiled out: the library is absent and the
// reference ships no .nc file).  kaxis = [wave | ln p | dT], kdata (nw, np, nt), krefatm (2, np).
extern "C" int ref_rfm_forward(const double* kaxis, const double* kdata, const int64_t* kshape, const double* krefatm,
                               const double* pres, const double* temp, const double* conc, int ncol, int nlyr,
                               int nspecies, int species_id, double* out) {
  try {
    harp::RFMImpl r;
    for (int i = 0; i < 3; ++i) r.kshape[i] = (size_t)kshape[i];
    r.kaxis = wrapd(kaxis, {kshape[0] + kshape[1] + kshape[2]});
    r.kdata = wrapd(kdata, {kshape[0], kshape[1], kshape[2]});
    r.krefatm = wrapd(krefatm, {2, kshape[1]});
    r.options.species_ids({species_id});
    std::map<std::string, torch::Tensor> kw;
    kw["pres"] = wrapd(pres, {ncol, nlyr});
    kw["temp"] = wrapd(temp, {ncol, nlyr});
    auto res = r.forward(wrapd(conc, {ncol, nlyr, nspecies}), kw).contiguous();
    std::memcpy(out, res.data_ptr<double>(), sizeof(double) * res.numel());
    return 0;
  } catch (std::exception const& e) {
    std::fprintf(stderr, "ref_rfm_forward: %s\n", e.what());
    return 2;
  }
}

extern "C" int ref_layer2level(const double* var, long nrow, int nlyr, int order, int blower, int bupper, double* out) {
  try {
    harp::Layer2LevelOptions op;
    op.order(order).blower(blower).bupper(bupper).check_positivity(false);
    auto res = harp::layer2level(wrapd(var, {nrow, nlyr}), op).contiguous();
    std::memcpy(out, res.data_ptr<double>(), sizeof(double) * res.numel());
    return 0;
  } catch (std::exception const& e) {
    std::fprintf(stderr, "ref_layer2level: %s\n", e.what());
    return 2;
  }
}

extern "C" int ref_spherical_flux_correction(double* flx, const double* x1f, const double* area, const double* vol,
                                             long nrow, int nx1) {
  try {
    auto f = wrapd(flx, {nrow, nx1});
    auto res = spherical_flux_correction(f, wrapd(x1f, {nx1}), wrapd(area, {nrow, nx1}), wrapd(vol, {nrow, nx1})).contiguous();
    std::memcpy(flx, res.data_ptr<double>(), sizeof(double) * res.numel());
    return 0;
  } catch (std::exception const& e) {
    std::fprintf(stderr, "ref_spherical_flux_correction: %s\n", e.what());
    return 2;
  }
}
