// band_dropin.cpp -- harp::RadiationBand as declared by the REFERENCE's own header
// (src/radiation/radiation_band.hpp), with its member functions defined by this repository's drop-in translation
// unit pyharp_b200/csrc/host/radiation_band_b200.cpp instead of the reference's radiation_band.cpp: one band of two
// RFM correlated-k attenuators (tables read by the reference's own RFMImpl::reset through the NetCDF stand-in) and
// the DISORT solver, driven through RadiationBandImpl::forward(conc, dz, bc, op).  Inputs come from
// band_inputs.nc, outputs go to raw float64 files the Python test compares with the reference chain.  Needs a GPU.
#include <cstdio>
#include <iostream>

#include <radiation/radiation.hpp>

extern "C" {
#include <netcdf.h>
}

static torch::Tensor read_var(int nc, const char* name, std::vector<int64_t> shape) {
  int id;
  TORCH_CHECK(nc_inq_varid(nc, name, &id) == NC_NOERR, "missing variable ", name);
  auto t = torch::empty(shape, torch::kFloat64);
  TORCH_CHECK(nc_get_var_double(nc, id, t.data_ptr<double>()) == NC_NOERR, "cannot read ", name);
  return t;
}

static void dump(const char* path, torch::Tensor t) {
  t = t.to(torch::kCPU, torch::kDouble).contiguous();
  std::FILE* f = std::fopen(path, "wb");
  std::fwrite(t.data_ptr<double>(), sizeof(double), (size_t)t.numel(), f);
  std::fclose(f);
}

int main(int argc, char** argv) {
  const bool on_device = argc > 1 && std::string(argv[1]) == "cuda";
  int nc;
  TORCH_CHECK(nc_open("band_inputs.nc", NC_NETCDF4, &nc) == NC_NOERR, "band_inputs.nc missing");
  size_t nwave, ncol, nlyr;
  int did;
  nc_inq_dimid(nc, "nwave", &did); nc_inq_dimlen(nc, did, &nwave);
  nc_inq_dimid(nc, "ncol", &did); nc_inq_dimlen(nc, did, &ncol);
  nc_inq_dimid(nc, "nlyr", &did); nc_inq_dimlen(nc, did, &nlyr);
  auto conc = read_var(nc, "conc", {(int64_t)ncol, (int64_t)nlyr, 2});
  auto pres = read_var(nc, "pres", {(int64_t)ncol, (int64_t)nlyr});
  auto temp = read_var(nc, "temp", {(int64_t)ncol, (int64_t)nlyr});
  auto dz = read_var(nc, "dz", {(int64_t)nlyr});
  std::map<std::string, torch::Tensor> bc;
  bc["albedo"] = read_var(nc, "albedo", {(int64_t)nwave, (int64_t)ncol});
  bc["btemp"] = read_var(nc, "btemp", {(int64_t)nwave, (int64_t)ncol});
  nc_close(nc);

  harp::AttenuatorOptions aop;
  aop.type("rfm").species_names({"CO2", "H2O"}).species_weights({44.0e-3, 18.0e-3}).opacity_files({"amarsw-ck-B1.nc"});
  harp::RadiationBandOptions bop;
  bop.name("B1").outdirs("(0,0)").solver_name("disort").nlyr((int)nlyr).ncol((int)ncol);
  aop.species_ids({0});
  bop.attenuators()["CO2"] = aop;
  aop.species_ids({1});
  bop.attenuators()["H2O"] = aop;
  bop.wave_lower(std::vector<double>(nwave, 1.)).wave_upper(std::vector<double>(nwave, 150.));
  bop.disort().flags("lamber,quiet,onlyfl,planck");
  bop.disort().ds().nstr = 8;
  bop.disort().ds().nmom = 8;

  harp::RadiationBand band(bop);
  band->pretty_print(std::cout);
  std::cout << std::endl;

  std::map<std::string, torch::Tensor> op;
  op["pres"] = pres;
  op["temp"] = temp;
  if (on_device) {
    auto dev = torch::Device(torch::kCUDA, 0);
    conc = conc.to(dev); dz = dz.to(dev);
    for (auto& kv : op) kv.second = kv.second.to(dev);
    for (auto& kv : bc) kv.second = kv.second.to(dev);
  }
  auto flx = band->forward(conc, dz, bc, op);
  std::cout << "flux shape = " << flx.sizes() << " on " << flx.device() << std::endl;
  dump("band_flx.bin", flx);
  dump("band_prop.bin", harp::shared.at("radiation/B1/optics"));
  dump("band_bflx.bin", harp::shared.at("radiation/B1/bflx"));
  // a second call reuses the handle
  auto again = band->forward(conc, dz, bc, op);
  const bool ok = flx.isfinite().all().item<bool>() && torch::equal(flx, again) && flx.size(0) == (int64_t)nwave &&
                  flx.size(2) == (int64_t)nlyr + 1;
  std::cout << (ok ? "BAND_DROPIN_OK" : "BAND_DROPIN_FAIL") << std::endl;
  return ok ? 0 : 1;
}
