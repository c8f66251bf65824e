// nccl_band_cabi.cpp -- the multi-GPU path of SURVEY.md 8e from a C++ host, C ABI only (no torch, no Python):
// spectral points of a line-by-line shaped batch are sharded over every GPU of the box, each GPU solves its slice with
// the band integration fused into the solve kernel (b200disort_solve_band), and ONE ncclAllReduce of the
// (ncol, nlyr+1, 2) band flux through b200disort_allreduce_band finishes the integration -- the op of
// examples/amars_lw.cpp:84-88 across GPUs.  One process drives all GPUs (ncclCommInitAll + group calls); the result
// must equal the band flux a single GPU computes for the whole batch.  Needs >= 2 GPUs (exits 77 otherwise).
#include <cuda_runtime.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <b200band.h>
#include <b200disort.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
#define BK(x) do { int rc = (x); if (rc != B200DISORT_OK) { std::printf("b200 error %s at line %d\n", b200disort_strerror(rc), __LINE__); return 1; } } while (0)

struct Rank {
  b200disort_handle* h = nullptr;
  cudaStream_t stream = nullptr;
  double *prop = nullptr, *umu0 = nullptr, *fbeam = nullptr, *albedo = nullptr, *w = nullptr, *bflx = nullptr;
  int32_t* status = nullptr;
  int n0 = 0, n1 = 0;
};

static int make_handle(int dev, cudaStream_t stream, int nwave, int nlyr, b200disort_handle** out) {
  b200disort_config cfg{};
  cfg.nwave = nwave; cfg.ncol = 1; cfg.nlyr = nlyr; cfg.nstr = 4; cfg.nmom = 4;
  cfg.ntau = cfg.numu = cfg.nphi = 1;
  cfg.flags = B200DISORT_LAMBER | B200DISORT_ONLYFL | B200DISORT_QUIET;
  cfg.device = dev;
  cfg.stream = stream;
  return b200disort_create(&cfg, out);
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) {
    std::printf("NCCL_BAND_CABI_SKIP: %d GPU(s)\n", ndev);
    return 77;
  }
  const int nlyr = 40, nwave = 4096 * ndev, nprop = 2, nlev = nlyr + 1;
  // seeded line-by-line shaped inputs (tau log-uniform over six decades, ssa in [0, 0.5]), weights d(nu) = 0.1
  std::vector<double> prop((size_t)nwave * nlyr * nprop), umu0(nwave, 0.6), fbeam(nwave, 1.0), albedo(nwave, 0.3), w(nwave, 0.1);
  unsigned long long sd = 20261018ULL;
  auto rnd = [&]() { sd = sd * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(sd >> 11) / 9007199254740992.0; };
  for (int n = 0; n < nwave; ++n)
    for (int k = 0; k < nlyr; ++k) {
      prop[((size_t)n * nlyr + k) * 2] = std::pow(10.0, -6.0 + 6.0 * rnd()) * std::exp(-4.0 * k / nlyr);
      prop[((size_t)n * nlyr + k) * 2 + 1] = 0.5 * rnd();
    }

  std::vector<int> devs(ndev);
  for (int d = 0; d < ndev; ++d) devs[d] = d;
  std::vector<ncclComm_t> comms(ndev);
  if (ncclCommInitAll(comms.data(), ndev, devs.data()) != ncclSuccess) { std::printf("ncclCommInitAll failed\n"); return 1; }

  // ---- the whole batch on GPU 0: the answer
  std::vector<double> want((size_t)nlev * 2);
  {
    CK(cudaSetDevice(0));
    b200disort_handle* h = nullptr;
    BK(make_handle(0, nullptr, nwave, nlyr, &h));
    double *dp, *du, *df, *da, *dw, *db; int32_t* ds;
    CK(cudaMalloc(&dp, prop.size() * 8)); CK(cudaMalloc(&du, nwave * 8)); CK(cudaMalloc(&df, nwave * 8));
    CK(cudaMalloc(&da, nwave * 8)); CK(cudaMalloc(&dw, nwave * 8)); CK(cudaMalloc(&db, want.size() * 8)); CK(cudaMalloc(&ds, nwave * 4));
    CK(cudaMemcpy(dp, prop.data(), prop.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(du, umu0.data(), nwave * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(df, fbeam.data(), nwave * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, albedo.data(), nwave * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, w.data(), nwave * 8, cudaMemcpyHostToDevice));
    b200disort_bc bc{};
    bc.umu0 = du; bc.fbeam = df; bc.albedo = da;
    BK(b200disort_solve_band(h, dp, nprop, &bc, nullptr, dw, db, ds));
    int nbad = 0;
    BK(b200disort_check_status(h, ds, &nbad));
    CK(cudaMemcpy(want.data(), db, want.size() * 8, cudaMemcpyDeviceToHost));
    b200disort_destroy(h);
    cudaFree(dp); cudaFree(du); cudaFree(df); cudaFree(da); cudaFree(dw); cudaFree(db); cudaFree(ds);
  }

  // ---- sharded: rank r takes spectral points [n0, n1)
  std::vector<Rank> R(ndev);
  for (int r = 0; r < ndev; ++r) {
    Rank& k = R[r];
    k.n0 = (int)((long)nwave * r / ndev); k.n1 = (int)((long)nwave * (r + 1) / ndev);
    const int nl = k.n1 - k.n0;
    CK(cudaSetDevice(r));
    CK(cudaStreamCreate(&k.stream));
    BK(make_handle(r, k.stream, nl, nlyr, &k.h));
    CK(cudaMalloc(&k.prop, (size_t)nl * nlyr * nprop * 8)); CK(cudaMalloc(&k.umu0, nl * 8)); CK(cudaMalloc(&k.fbeam, nl * 8));
    CK(cudaMalloc(&k.albedo, nl * 8)); CK(cudaMalloc(&k.w, nl * 8)); CK(cudaMalloc(&k.bflx, want.size() * 8)); CK(cudaMalloc(&k.status, nl * 4));
    CK(cudaMemcpyAsync(k.prop, prop.data() + (size_t)k.n0 * nlyr * nprop, (size_t)nl * nlyr * nprop * 8, cudaMemcpyHostToDevice, k.stream));
    CK(cudaMemcpyAsync(k.umu0, umu0.data() + k.n0, nl * 8, cudaMemcpyHostToDevice, k.stream));
    CK(cudaMemcpyAsync(k.fbeam, fbeam.data() + k.n0, nl * 8, cudaMemcpyHostToDevice, k.stream));
    CK(cudaMemcpyAsync(k.albedo, albedo.data() + k.n0, nl * 8, cudaMemcpyHostToDevice, k.stream));
    CK(cudaMemcpyAsync(k.w, w.data() + k.n0, nl * 8, cudaMemcpyHostToDevice, k.stream));
  }
  for (int r = 0; r < ndev; ++r) {  // every GPU solves its slice (asynchronous on its stream) ...
    Rank& k = R[r];
    b200disort_bc bc{};
    bc.umu0 = k.umu0; bc.fbeam = k.fbeam; bc.albedo = k.albedo;
    BK(b200disort_solve_band(k.h, k.prop, nprop, &bc, nullptr, k.w, k.bflx, k.status));
  }
  // ... and one all-reduce finishes the band integral; one thread drives all communicators, hence the group
  if (ncclGroupStart() != ncclSuccess) return 1;
  for (int r = 0; r < ndev; ++r) BK(b200disort_allreduce_band(R[r].h, R[r].bflx, comms[r]));
  if (ncclGroupEnd() != ncclSuccess) return 1;
  double worst = 0.0, top = 0.0;
  for (double v : want) top = std::fmax(top, std::fabs(v));
  for (int r = 0; r < ndev; ++r) {
    CK(cudaSetDevice(r));
    int nbad = 0;
    BK(b200disort_check_status(R[r].h, R[r].status, &nbad));
    std::vector<double> got(want.size());
    CK(cudaMemcpy(got.data(), R[r].bflx, got.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < got.size(); ++i) worst = std::fmax(worst, std::fabs(got[i] - want[i]) / top);
  }
  for (int r = 0; r < ndev; ++r) {
    cudaSetDevice(r);
    b200disort_destroy(R[r].h);
    ncclCommDestroy(comms[r]);
  }
  std::printf("GPUs %d, spectral points %d (%d per GPU), band flux at the top: up %.12e down %.12e\n", ndev, nwave, nwave / ndev,
              want[(size_t)nlyr * 2], want[(size_t)nlyr * 2 + 1]);
  std::printf("max |sharded - single GPU| / max|F| = %.3e\n", worst);
  const bool ok = worst < 1e-12;
  std::printf(ok ? "NCCL_BAND_CABI_OK\n" : "NCCL_BAND_CABI_FAIL\n");
  return ok ? 0 : 1;
}
