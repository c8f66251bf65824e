// disort/disort.hpp -- header-compatible replacement of pydisort's <disort/disort.hpp>
// (github.com/zoeyzyhu/pydisort @ afee3ec897f, fetched by cmake/pydisort.cmake:9-11) as pyharp
// uses it: tests/test_disort.cpp:13-55, examples/amars_sw.cpp:43-65,216,280,
// examples/amars_lw.cpp:18-38,60,80, src/radiation/radiation_band.cpp:58-68,124-127,
// src/radiation/radiation_band.hpp:35.  Same names, argument meaning, tensor layout and error
// behaviour (TORCH_CHECK exceptions); the solve itself runs on the B200 through the C ABI in
// <b200disort.h> -- there is no CPU path.
#pragma once

#include <map>
#include <string>
#include <vector>

#include <torch/nn/cloneable.h>
#include <torch/nn/module.h>
#include <torch/nn/modules/common.h>
#include <torch/nn/modules/container/any.h>
#include <torch/nn/pimpl.h>
#include <torch/types.h>

#include <cdisort213/cdisort.h>

struct b200disort_handle;

namespace disort {

namespace index {
constexpr int IEX = 0;  // extinction / layer optical depth
constexpr int ISS = 1;  // single-scattering albedo
constexpr int IPM = 2;  // first phase-function moment
constexpr int IUP = 0;  // upward flux
constexpr int IDN = 1;  // downward flux
}  // namespace index

//! Fluent option bag (pydisort's DisortOptions): op.nwave(10) sets, op.nwave() reads.
struct DisortOptions {
  DisortOptions();

#define B200_DISORT_ARG(T, name)                                                 \
 public:                                                                         \
  DisortOptions& name(const T& v) { name##_ = v; return *this; }                 \
  const T& name() const noexcept { return name##_; }                             \
  T& name() noexcept { return name##_; }                                         \
                                                                                 \
 private:                                                                        \
  T name##_
  B200_DISORT_ARG(std::string, header) = "";
  B200_DISORT_ARG(std::string, flags) = "";
  B200_DISORT_ARG(int, nwave) = 1;
  B200_DISORT_ARG(int, ncol) = 1;
  B200_DISORT_ARG(double, accur) = 0.;
  B200_DISORT_ARG(std::vector<double>, wave_lower) = {};
  B200_DISORT_ARG(std::vector<double>, wave_upper) = {};
  B200_DISORT_ARG(std::vector<double>, user_tau) = {};
  B200_DISORT_ARG(std::vector<double>, user_mu) = {};
  B200_DISORT_ARG(std::vector<double>, user_phi) = {};
#undef B200_DISORT_ARG

 public:
  disort_state& ds() { return ds_; }
  disort_state const& ds() const { return ds_; }

  //! parse flags() onto ds().flag.* and return the B200DISORT_* bitmask; throws on unknown tokens
  unsigned flag_bits();

 private:
  disort_state ds_;
};

class DisortImpl : public torch::nn::Cloneable<DisortImpl> {
 public:
  //! options with which this `DisortImpl` was constructed
  DisortOptions options;

  DisortImpl() = default;
  explicit DisortImpl(DisortOptions const& options_);
  //! A copy owns its own solver handle (torch::nn::Cloneable::clone() copy-constructs and then calls reset()).
  DisortImpl(DisortImpl const& other);
  DisortImpl& operator=(DisortImpl const& other);
  ~DisortImpl() override;
  void reset() override;

  //! shape / flags / user grids of the (n, j) solve; all solves of a module share one shape
  disort_state& ds(int n = 0, int j = 0) { return options.ds(); }
  disort_state const& ds(int n = 0, int j = 0) const { return options.ds(); }

  //! Radiative fluxes.
  /*!
   * \param prop optical properties (nwave, ncol, nlyr, nprop), layer 0 = surface
   * \param bc   boundary conditions, each (nwave, ncol): umu0, fbeam, albedo, btemp, ttemp, temis, fisot, phi0
   * \param temf level temperatures (ncol, nlyr + 1), level 0 = surface (needed with `planck`)
   * \return (nwave, ncol, nlyr + 1 | ntau, 2): [..., IUP] upward, [..., IDN] total downward flux
   */
  torch::Tensor forward(torch::Tensor prop, std::map<std::string, torch::Tensor>* bc,
                        torch::optional<torch::Tensor> temf = torch::nullopt);

  //! intensities of the last forward(): (nwave, ncol, nphi, ntau, numu)
  torch::Tensor get_rad(torch::TensorOptions op = torch::TensorOptions().dtype(torch::kDouble)) const;

  //! per-solve status words of the last forward() (nwave, ncol), int32, CPU
  torch::Tensor status() const { return status_; }

 protected:
  FORWARD_HAS_DEFAULT_ARGS({2, torch::nn::AnyValue(torch::optional<torch::Tensor>())})

 private:
  void release();
  void create_handle(int device);
  b200disort_handle* handle_ = nullptr;
  int device_ = -1;  // CUDA device the handle lives on
  unsigned flags_ = 0;
  std::vector<double> utau_, umu_, phi_;
  torch::Tensor status_;
  bool have_rad_ = false;
};
TORCH_MODULE(Disort);

}  // namespace disort
