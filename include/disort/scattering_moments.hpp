// disort/scattering_moments.hpp -- pydisort's phase-function helper as pyharp uses it
// (tests/test_disort.cpp:37-40).
#pragma once
#include <torch/types.h>

namespace disort {

enum PhaseFunctionType { kIsotropic = 0, kRayleigh = 1, kHenyeyGreenstein = 2, kHazeGarciaSiewert = 3, kCloudGarciaSiewert = 4 };

struct PhaseMomentOptions {
  PhaseMomentOptions& type(int v) { type_ = v; return *this; }
  int type() const { return type_; }
  PhaseMomentOptions& gg(double v) { gg_ = v; return *this; }
  double gg() const { return gg_; }

 private:
  int type_ = kIsotropic;
  double gg_ = 0.;
};

//! Legendre moments 1..nmom of the phase function (the 0th moment is implicit = 1), float64 (nmom,)
torch::Tensor scattering_moments(int nmom, PhaseMomentOptions const& options);

}  // namespace disort
