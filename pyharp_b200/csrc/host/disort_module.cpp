// disort_module.cpp -- C++ host side of the B200 DISORT path: the disort:: surface pyharp links
// against (include/disort/*.hpp), implemented on the C ABI of include/b200disort.h.
// Replaces pydisort's torch wrapper (DisortImpl::forward looping c_disort over (wave, column),
// SURVEY.md section 8 rows a1-a5); here one call solves the whole batch on the GPU.
#include <disort/disort.hpp>
#include <disort/scattering_moments.hpp>

#include <b200disort.h>

#include <c10/cuda/CUDAFunctions.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace disort {

// Debug aid: with B200DISORT_DUMP=<file> every forward() appends its inputs and result as raw records
// "<name> <ndim> <dims...>\n" + float64 payload (tests/golden/make_golden_amars.py reads them).  This is how
// the real amars_sw inputs (tables + profile exactly as examples/amars_sw.cpp builds them) become a fixture.
static void dump_tensor(std::FILE* f, const char* name, torch::Tensor t) {
  t = t.to(torch::kCPU, torch::kDouble).contiguous();
  std::fprintf(f, "%s %d", name, (int)t.dim());
  for (int64_t d : t.sizes()) std::fprintf(f, " %lld", (long long)d);
  std::fprintf(f, "\n");
  std::fwrite(t.data_ptr<double>(), sizeof(double), (size_t)t.numel(), f);
}

DisortOptions::DisortOptions() {
  std::memset(&ds_, 0, sizeof(ds_));
  ds_.nlyr = 1;
  ds_.nstr = 4;
  ds_.nmom = 4;
  ds_.nphi = ds_.ntau = ds_.numu = 1;
}

unsigned DisortOptions::flag_bits() {
  unsigned bits = 0;
  disort_flag& f = ds_.flag;
  std::memset(&f, 0, sizeof(f));
  std::stringstream ss(flags_);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    tok.erase(0, tok.find_first_not_of(" \t"));
    tok.erase(tok.find_last_not_of(" \t") + 1);
    if (tok.empty()) continue;
    if (tok == "usrtau") { f.usrtau = 1; bits |= B200DISORT_USRTAU; }
    else if (tok == "usrang") { f.usrang = 1; bits |= B200DISORT_USRANG; }
    else if (tok == "lamber") { f.lamber = 1; bits |= B200DISORT_LAMBER; }
    else if (tok == "planck") { f.planck = 1; bits |= B200DISORT_PLANCK; }
    else if (tok == "onlyfl") { f.onlyfl = 1; bits |= B200DISORT_ONLYFL; }
    else if (tok == "quiet") { f.quiet = 1; bits |= B200DISORT_QUIET; }
    else if (tok == "intensity_correction") { f.intensity_correction = 1; bits |= B200DISORT_INTENSITY_CORRECTION; }
    else if (tok == "old_intensity_correction") { f.old_intensity_correction = 1; bits |= B200DISORT_OLD_INTENSITY_CORRECTION; }
    else if (tok == "print-input") f.prnt[0] = 1;
    else if (tok == "print-fluxes") f.prnt[1] = 1;
    else if (tok == "print-intensity") f.prnt[2] = 1;
    else if (tok == "print-transmissivity") f.prnt[3] = 1;
    else if (tok == "print-phase-function") f.prnt[4] = 1;
    else if (tok == "ibcnd" || tok == "spher" || tok == "general_source" || tok == "output_uum") {
      TORCH_CHECK(false, "disort flag '", tok, "' is outside the B200 solve path");
    } else {
      TORCH_CHECK(false, "unknown disort flag '", tok, "'");
    }
  }
  return bits;
}

DisortImpl::DisortImpl(DisortOptions const& options_) : options(options_) { reset(); }

// The handle (device workspace, tables) is never shared between copies: a copy starts without one and builds its own
// in reset(); ds().utau / umu / phi are re-pointed at the copy's own vectors there.
DisortImpl::DisortImpl(DisortImpl const& other)
    : torch::nn::Cloneable<DisortImpl>(other), options(other.options), handle_(nullptr), device_(other.device_) {
  if (other.handle_) reset();
}

DisortImpl& DisortImpl::operator=(DisortImpl const& other) {
  if (this != &other) {
    release();
    torch::nn::Cloneable<DisortImpl>::operator=(other);
    options = other.options;
    device_ = other.device_;
    status_ = torch::Tensor();
    have_rad_ = false;
    if (other.handle_) reset();
  }
  return *this;
}

DisortImpl::~DisortImpl() { release(); }

void DisortImpl::release() {
  if (handle_) b200disort_destroy(handle_);
  handle_ = nullptr;
}

void DisortImpl::reset() {
  release();
  flags_ = options.flag_bits();
  disort_state& d = options.ds();
  utau_ = options.user_tau();
  umu_ = options.user_mu();
  phi_ = options.user_phi();
  if (!utau_.empty()) d.ntau = (int)utau_.size();
  if (!umu_.empty()) d.numu = (int)umu_.size();
  if (!phi_.empty()) d.nphi = (int)phi_.size();
  d.utau = utau_.empty() ? nullptr : utau_.data();
  d.umu = umu_.empty() ? nullptr : umu_.data();
  d.phi = phi_.empty() ? nullptr : phi_.data();
  d.nmom_nstr = std::max(d.nmom, d.nstr);
  d.accur = options.accur();
  std::strncpy(d.header, options.header().c_str(), sizeof(d.header) - 1);
  if (flags_ & B200DISORT_PLANCK)
    TORCH_CHECK((int)options.wave_lower().size() == options.nwave() && (int)options.wave_upper().size() == options.nwave(),
                "disort: planck needs wave_lower/wave_upper of length nwave");

  // the handle lives on the caller's current CUDA device until a forward() brings a tensor from another one
  create_handle(device_ >= 0 ? device_ : (int)c10::cuda::current_device());
}

void DisortImpl::create_handle(int device) {
  release();
  disort_state const& d = options.ds();
  b200disort_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.nwave = options.nwave(); cfg.ncol = options.ncol(); cfg.nlyr = d.nlyr;
  cfg.nstr = d.nstr; cfg.nmom = d.nmom_nstr;
  cfg.ntau = d.ntau; cfg.numu = d.numu; cfg.nphi = d.nphi;
  cfg.flags = flags_; cfg.accur = d.accur;
  cfg.user_tau = d.utau; cfg.user_mu = d.umu; cfg.user_phi = d.phi;
  cfg.wave_lower = options.wave_lower().empty() ? nullptr : options.wave_lower().data();
  cfg.wave_upper = options.wave_upper().empty() ? nullptr : options.wave_upper().data();
  cfg.device = device;
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  cfg.stream = c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device).stream();
  int rc = b200disort_create(&cfg, &handle_);
  TORCH_CHECK(rc == B200DISORT_OK, "disort: b200disort_create failed: ", b200disort_strerror(rc));
  device_ = device;
  have_rad_ = false;
}

torch::Tensor DisortImpl::forward(torch::Tensor prop, std::map<std::string, torch::Tensor>* bc,
                                  torch::optional<torch::Tensor> temf) {
  TORCH_CHECK(handle_, "disort: module not initialised");
  disort_state const& d = options.ds();
  const int nwave = options.nwave(), ncol = options.ncol();
  TORCH_CHECK(prop.dim() == 4 && prop.size(0) == nwave && prop.size(1) == ncol && prop.size(2) == d.nlyr,
              "disort: prop must be (nwave=", nwave, ", ncol=", ncol, ", nlyr=", d.nlyr, ", nprop), got ", prop.sizes());
  TORCH_CHECK(prop.scalar_type() == torch::kDouble, "disort: prop must be float64");
  const bool on_host = !prop.is_cuda();
  const auto dev = prop.device();
  prop = prop.contiguous();
  // the solve runs on prop's device (one rank per GPU: a caller on cuda:1 must not be sent to cuda:0) and on the
  // caller's current stream of that device, so it is ordered after the producers of prop and before its consumers;
  // the caller's current device is left as it was
  const int want = on_host ? device_ : (int)dev.index();
  if (want != device_) create_handle(want);
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
  {
    int rc = b200disort_set_stream(handle_, c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device_).stream());
    TORCH_CHECK(rc == B200DISORT_OK, "disort: b200disort_set_stream failed: ", b200disort_strerror(rc));
  }

  b200disort_bc b;
  std::memset(&b, 0, sizeof(b));
  std::vector<torch::Tensor> keep;
  if (bc) {
    for (auto const& [key, val] : *bc) {
      const double** slot = nullptr;
      if (key == "umu0") slot = &b.umu0;
      else if (key == "fbeam") slot = &b.fbeam;
      else if (key == "albedo") slot = &b.albedo;
      else if (key == "btemp") slot = &b.btemp;
      else if (key == "ttemp") slot = &b.ttemp;
      else if (key == "temis") slot = &b.temis;
      else if (key == "fisot") slot = &b.fisot;
      else if (key == "phi0") slot = &b.phi0;
      else if (key == "fluor") continue;  // accepted, unused (as in cdisort without general_source)
      else TORCH_CHECK(false, "disort: unknown boundary condition '", key, "'");
      TORCH_CHECK(val.dim() == 2 && val.size(0) == nwave && val.size(1) == ncol, "disort: bc['", key,
                  "'] must be (nwave, ncol)");
      keep.push_back(val.to(dev, torch::kDouble).contiguous());
      *slot = keep.back().data_ptr<double>();
    }
  }
  const double* tptr = nullptr;
  if (flags_ & B200DISORT_PLANCK) {
    TORCH_CHECK(temf.has_value(), "disort: planck flag set but temf not given");
    TORCH_CHECK(temf->dim() == 2 && temf->size(0) == ncol && temf->size(1) == d.nlyr + 1,
                "disort: temf must be (ncol, nlyr + 1)");
    keep.push_back(temf->to(dev, torch::kDouble).contiguous());
    tptr = keep.back().data_ptr<double>();
  }
  const int nlev = (flags_ & B200DISORT_USRTAU) ? d.ntau : d.nlyr + 1;
  auto topt = torch::TensorOptions().dtype(torch::kDouble).device(dev);
  auto flx = torch::empty({nwave, ncol, nlev, 2}, on_host ? topt.pinned_memory(true) : topt);
  auto st = torch::empty({nwave, ncol}, torch::TensorOptions().dtype(torch::kInt32).device(dev).pinned_memory(on_host));
  int rc;
  if (on_host) {
    rc = b200disort_solve_host(handle_, prop.data_ptr<double>(), (int)prop.size(3), &b, tptr, flx.data_ptr<double>(),
                               nullptr, st.data_ptr<int32_t>());
  } else {
    rc = b200disort_solve(handle_, prop.data_ptr<double>(), (int)prop.size(3), &b, tptr, flx.data_ptr<double>(),
                          nullptr, st.data_ptr<int32_t>());
    if (rc == B200DISORT_OK) {
      int nbad = 0;
      rc = b200disort_check_status(handle_, st.data_ptr<int32_t>(), &nbad);
    }
  }
  status_ = st.cpu();
  have_rad_ = !(flags_ & B200DISORT_ONLYFL);
  if (rc == B200DISORT_ERR_SOLVE) {
    auto bad = status_.flatten().nonzero().flatten();
    const int64_t first = bad.numel() ? bad[0].item<int64_t>() : -1;
    TORCH_CHECK(false, "disort: ", bad.numel(), " of ", (int64_t)nwave * ncol, " solves failed (first: solve ", first,
                ", status ", first >= 0 ? status_.flatten()[first].item<int>() : -1, ")");
  }
  TORCH_CHECK(rc == B200DISORT_OK, "disort: solve failed: ", b200disort_strerror(rc));
  if (const char* path = std::getenv("B200DISORT_DUMP")) {
    if (std::FILE* f = std::fopen(path, "ab")) {
      std::fprintf(f, "forward %s nstr %d\n", options.flags().c_str(), d.nstr);
      dump_tensor(f, "prop", prop);
      if (bc)
        for (auto const& [key, val] : *bc) dump_tensor(f, ("bc_" + key).c_str(), val);
      if (temf.has_value()) dump_tensor(f, "temf", *temf);
      dump_tensor(f, "flx", flx);
      std::fclose(f);
    }
  }
  return flx;
}

torch::Tensor DisortImpl::get_rad(torch::TensorOptions op) const {
  TORCH_CHECK(handle_ && have_rad_, "disort: no intensities available (onlyfl set, or forward() not called)");
  disort_state const& d = options.ds();
  const int nlev = (flags_ & B200DISORT_USRTAU) ? d.ntau : d.nlyr + 1;
  const int numu = (flags_ & B200DISORT_USRANG) ? d.numu : d.nstr;
  auto out = torch::empty({options.nwave(), options.ncol(), d.nphi, nlev, numu},
                          torch::TensorOptions().dtype(torch::kDouble).device(op.device()));
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
  int rc = b200disort_get_rad(handle_, out.data_ptr<double>());
  TORCH_CHECK(rc == B200DISORT_OK, "disort: get_rad failed: ", b200disort_strerror(rc));
  return out.to(op.dtype());
}

torch::Tensor scattering_moments(int nmom, PhaseMomentOptions const& options) {
  auto out = torch::zeros({nmom}, torch::kDouble);
  switch (options.type()) {
    case kIsotropic:
      break;
    case kRayleigh:
      if (nmom >= 2) out[1] = 0.1;
      break;
    case kHenyeyGreenstein: {
      TORCH_CHECK(options.gg() > -1. && options.gg() < 1., "scattering_moments: |gg| must be < 1");
      auto a = out.accessor<double, 1>();
      for (int k = 0; k < nmom; ++k) a[k] = std::pow(options.gg(), k + 1);
      break;
    }
    default:
      TORCH_CHECK(false, "scattering_moments: tabulated Garcia-Siewert phase functions are not on pyharp's path");
  }
  return out;
}

}  // namespace disort
