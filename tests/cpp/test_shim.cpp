// test_shim.cpp -- the C++ disort:: surface (include/disort/*.hpp over the C ABI) with the
// assertions the reference's tests/test_disort.cpp lacks.  Needs a GPU (no CPU path).
#include <gtest/gtest.h>

#include <fmt/format.h>

#include <disort/disort.hpp>
#include <disort/disort_formatter.hpp>
#include <disort/scattering_moments.hpp>

#include <c10/cuda/CUDAFunctions.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

static disort::DisortOptions test1a_options() {
  disort::DisortOptions op;
  op.header("running disort example");
  op.flags("usrtau,usrang,lamber,quiet,intensity_correction,old_intensity_correction,print-input,print-phase-function");
  op.nwave(10);
  op.ds().nlyr = 1;
  op.ds().nstr = 16;
  op.ds().nmom = 16;
  op.user_mu({-1, -0.5, -0.1, 0.1, 0.5, 1});
  op.user_phi({0});
  op.user_tau({0, 0.03125});
  return op;
}

// DISORT test problem 1a, pyharp order: level 0 = bottom
static const double kFlup[2] = {0.0, 7.9945097492e-02};
static const double kFdn[2] = {2.2984382852 + 7.9410795371e-02, 3.1415926536};
static const double kUU[2][6] = {{0, 0, 0, 1.1777066223e-01, 2.6417038057e-02, 1.3404130206e-02},
                                 {1.3382626716e-02, 2.6332351413e-02, 1.1589788558e-01, 0, 0, 0}};

static void run_test1a(torch::Device dev) {
  disort::Disort disort(test1a_options());
  auto prop = torch::zeros({disort->options.nwave(), disort->options.ncol(), disort->ds().nlyr, 2 + disort->ds().nstr},
                           torch::kDouble);
  prop.select(3, disort::index::IEX) = disort->ds().utau[1];
  prop.select(3, disort::index::ISS) = 0.2;
  prop.narrow(3, disort::index::IPM, disort->ds().nstr) =
      disort::scattering_moments(disort->ds().nstr, disort::PhaseMomentOptions().type(disort::kIsotropic));
  std::map<std::string, torch::Tensor> bc;
  bc["umu0"] = 0.1 * torch::ones({disort->options.nwave(), disort->options.ncol()}, torch::kDouble);
  bc["fbeam"] = M_PI / bc["umu0"];
  prop = prop.to(dev);
  for (auto& kv : bc) kv.second = kv.second.to(dev);
  auto result = disort->forward(prop, &bc).cpu();
  auto rad = disort->get_rad(prop.options()).cpu();
  EXPECT_EQ(result.size(0), 10);
  EXPECT_EQ(result.size(2), 2);
  EXPECT_EQ(rad.size(3), 2);
  EXPECT_EQ(rad.size(4), 6);
  for (int n = 0; n < 10; ++n) {
    for (int lev = 0; lev < 2; ++lev) {
      EXPECT_NEAR(result[n][0][lev][0].item<double>(), kFlup[lev], 1e-9 * 0.08);
      EXPECT_NEAR(result[n][0][lev][1].item<double>(), kFdn[lev], 1e-9 * 3.2);
    }
    for (int lu = 0; lu < 2; ++lu)
      for (int iu = 0; iu < 6; ++iu) EXPECT_NEAR(rad[n][0][0][lu][iu].item<double>(), kUU[lu][iu], 1e-9 * 0.12);
  }
  EXPECT_TRUE(fmt::format("{}", disort->options).find("nstr = 16") != std::string::npos);
}

TEST(Shim, test1a_host_tensors) { run_test1a(torch::kCPU); }
TEST(Shim, test1a_device_tensors) { run_test1a(torch::Device(torch::kCUDA, 0)); }

TEST(Shim, longwave_host_equals_device_and_any_module) {
  const int nwave = 8, ncol = 3, nlyr = 40;
  disort::DisortOptions op;
  op.flags("lamber,quiet,onlyfl,planck");
  op.nwave(nwave).ncol(ncol);
  op.wave_lower(std::vector<double>(nwave, 1.)).wave_upper(std::vector<double>(nwave, 150.));
  op.ds().nlyr = nlyr;
  op.ds().nstr = 8;
  op.ds().nmom = 8;
  disort::Disort disort(op);
  torch::manual_seed(3);
  auto prop = torch::pow(10., torch::rand({nwave, ncol, nlyr, 1}, torch::kDouble) * 4. - 3.);
  std::map<std::string, torch::Tensor> bc;
  bc["albedo"] = torch::ones({nwave, ncol}, torch::kDouble);
  bc["btemp"] = torch::ones({nwave, ncol}, torch::kDouble) * 300.;
  auto temf = torch::linspace(300., 180., nlyr + 1, torch::kDouble).unsqueeze(0).repeat({ncol, 1});
  auto a = disort->forward(prop, &bc, temf);
  // the way harp::RadiationBandImpl holds and calls the solver (src/radiation/radiation_band.cpp:68,124-127)
  torch::nn::AnyModule any(disort);
  auto b = any.forward(prop.cuda(), &bc, torch::optional<torch::Tensor>(temf)).cpu();
  EXPECT_TRUE(torch::equal(a, b));
  EXPECT_TRUE(a.isfinite().all().item<bool>());
  EXPECT_GT(a.index({0, 0, 0, 0}).item<double>(), 0.);
  EXPECT_THROW(disort->get_rad(), c10::Error);
}

// Cloneable::clone() copy-constructs and calls reset() on the copy -- the copy must own its own solver
// handle (the original stays usable, nothing is destroyed twice), also through copy assignment
TEST(Shim, clone_and_copy_own_their_handles) {
  const int nwave = 4, ncol = 2, nlyr = 12;
  disort::DisortOptions op;
  op.flags("lamber,quiet,onlyfl");
  op.nwave(nwave).ncol(ncol);
  op.ds().nlyr = nlyr;
  op.ds().nstr = 8;
  op.ds().nmom = 8;
  disort::Disort disort(op);
  torch::manual_seed(5);
  auto prop = torch::rand({nwave, ncol, nlyr, 2}, torch::kDouble) * 0.9;
  std::map<std::string, torch::Tensor> bc;
  bc["fbeam"] = torch::ones({nwave, ncol}, torch::kDouble);
  bc["umu0"] = torch::full({nwave, ncol}, 0.7, torch::kDouble);
  auto a = disort->forward(prop, &bc);
  {
    auto c = std::dynamic_pointer_cast<disort::DisortImpl>(disort->clone());
    ASSERT_TRUE(c != nullptr);
    EXPECT_TRUE(torch::equal(c->forward(prop, &bc), a));
    disort::DisortImpl copy(*disort);  // plain copy construction
    EXPECT_TRUE(torch::equal(copy.forward(prop, &bc), a));
    disort::DisortOptions other = op;
    other.nwave(1);
    disort::DisortImpl assigned(other);
    assigned = *disort;  // copy assignment replaces the handle
    EXPECT_TRUE(torch::equal(assigned.forward(prop, &bc), a));
  }  // the copies and their handles are gone
  EXPECT_TRUE(torch::equal(disort->forward(prop, &bc), a));
  // a module holding the solver (harp::RadiationBand does) clones through register_module / AnyModule too
  torch::nn::AnyModule any(disort);
  auto any2 = any.clone();
  EXPECT_TRUE(torch::equal(any2.forward(prop, &bc, torch::optional<torch::Tensor>()), a));
}

// launches go to the caller's CURRENT stream of the tensor's device (not the legacy stream), and the
// caller's current device is left alone; with two GPUs the handle follows the tensor's device
TEST(Shim, current_stream_and_device_are_respected) {
  const int nwave = 6, ncol = 3, nlyr = 20;
  disort::DisortOptions op;
  op.flags("lamber,quiet,onlyfl");
  op.nwave(nwave).ncol(ncol);
  op.ds().nlyr = nlyr;
  op.ds().nstr = 4;
  op.ds().nmom = 4;
  disort::Disort disort(op);
  torch::manual_seed(7);
  auto prop = torch::rand({nwave, ncol, nlyr, 2}, torch::kDouble) * 0.9;
  std::map<std::string, torch::Tensor> bc;
  bc["fbeam"] = torch::ones({nwave, ncol}, torch::kDouble);
  bc["umu0"] = torch::full({nwave, ncol}, 0.4, torch::kDouble);
  auto want = disort->forward(prop, &bc);
  auto dev0 = torch::Device(torch::kCUDA, 0);
  std::map<std::string, torch::Tensor> bcd;
  for (auto& kv : bc) bcd[kv.first] = kv.second.to(dev0);
  {
    auto stream = c10::cuda::getStreamFromPool(false, 0);
    c10::cuda::CUDAStreamGuard sg(stream);
    // the producer of prop runs on this stream right before the solve: no explicit synchronisation in between
    auto pd = (prop.to(dev0) * 2.0) * 0.5;
    auto got = disort->forward(pd, &bcd);
    EXPECT_TRUE(torch::equal(got.cpu(), want));
  }
  if (torch::cuda::device_count() > 1) {
    auto dev1 = torch::Device(torch::kCUDA, 1);
    std::map<std::string, torch::Tensor> bc1;
    for (auto& kv : bc) bc1[kv.first] = kv.second.to(dev1);
    auto got = disort->forward(prop.to(dev1), &bc1);
    EXPECT_TRUE(got.device() == dev1);
    EXPECT_TRUE(torch::equal(got.cpu(), want));
    EXPECT_EQ((int)c10::cuda::current_device(), 0);
    EXPECT_TRUE(torch::equal(disort->forward(prop.to(dev0), &bcd).cpu(), want));  // and back
  }
}

TEST(Shim, bad_input_throws) {
  disort::DisortOptions op;
  op.flags("lamber,onlyfl");
  op.nwave(2).ncol(2);
  op.ds().nlyr = 3;
  op.ds().nstr = 8;
  op.ds().nmom = 8;
  disort::Disort disort(op);
  auto prop = torch::full({2, 2, 3, 2}, 0.5, torch::kDouble);
  prop[1][0][2][1] = 1.5;  // ssalb > 1
  std::map<std::string, torch::Tensor> bc;
  bc["fbeam"] = torch::ones({2, 2}, torch::kDouble);
  bc["umu0"] = torch::full({2, 2}, 0.5, torch::kDouble);
  EXPECT_THROW(disort->forward(prop, &bc), c10::Error);
  EXPECT_EQ(disort->status()[1][0].item<int>(), 1);
  std::map<std::string, torch::Tensor> bad;
  bad["nonsense"] = torch::ones({2, 2}, torch::kDouble);
  EXPECT_THROW(disort->forward(torch::full({2, 2, 3, 2}, 0.5, torch::kDouble), &bad), c10::Error);
  disort::DisortOptions op2;
  op2.flags("lamber,bogus");
  EXPECT_THROW(disort::Disort d2(op2), c10::Error);
}

// harp::call_interpn_cuda<N> (declared at src/math/interpolation.hpp:12-14, defined in
// pyharp_b200/csrc/host/harp_cuda_dispatch.cpp), driven by a TensorIterator built the way
// S8FullerImpl::forward builds it (src/opacity/s8_fuller.cpp:88-103)
namespace harp {
template <int N>
void call_interpn_cuda(at::TensorIterator& iter, at::Tensor kdata, at::Tensor axis, at::Tensor dims, int nval);
}

TEST(Shim, call_interpn_cuda_like_s8fuller) {
  const int nwave = 7, ncol = 2, nlyr = 3, nprop = 2, rows = 9;
  auto dev = torch::Device(torch::kCUDA, 0);
  auto kwave = torch::linspace(0.2, 5.0, rows, torch::kFloat64);
  auto kdata = torch::rand({rows, nprop}, torch::kFloat64);
  auto coord = torch::tensor({0.1, 0.2, 0.77, 1.4, 3.3, 5.0, 9.0}, torch::kFloat64);
  auto out = torch::zeros({nwave, ncol, nlyr, nprop}, torch::TensorOptions().dtype(torch::kFloat64).device(dev));
  auto dims = torch::tensor({rows}, torch::TensorOptions().dtype(torch::kInt64).device(dev));
  auto iter = at::TensorIteratorConfig()
                  .resize_outputs(false)
                  .check_all_same_dtype(true)
                  .declare_static_shape(out.sizes(), /*squash_dims=*/3)
                  .add_output(out)
                  .add_owned_const_input(coord.to(dev).view({-1, 1, 1, 1}).expand({-1, ncol, nlyr, nprop}))
                  .build();
  harp::call_interpn_cuda<2>(iter, kdata.to(dev), kwave.to(dev), dims, nprop);
  auto got = out.cpu();
  auto kw = kwave.accessor<double, 1>();
  auto kd = kdata.accessor<double, 2>();
  for (int w = 0; w < nwave; ++w) {
    const double x = coord[w].item<double>();
    int i1 = 0;
    while (i1 + 1 < rows - 1 && kw[i1 + 1] <= x) ++i1;
    for (int j = 0; j < nprop; ++j) {
      double want;
      if (x <= kw[0]) want = kd[0][j];
      else if (x >= kw[rows - 1]) want = kd[rows - 1][j];
      else want = ((x - kw[i1]) * kd[i1 + 1][j] + (kw[i1 + 1] - x) * kd[i1][j]) / (kw[i1 + 1] - kw[i1]);
      for (int c = 0; c < ncol; ++c)
        for (int l = 0; l < nlyr; ++l) EXPECT_NEAR(got[w][c][l][j].item<double>(), want, 1e-15);
    }
  }
}

int main(int argc, char** argv) {
  testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
