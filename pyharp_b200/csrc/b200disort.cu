// b200disort.cu -- kernels and C ABI of the B200-native DISORT solve path
// (include/b200disort.h).  sm_100a only; no CPU fallback: every entry point that
// computes needs a CUDA device and fails with B200DISORT_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/b200disort.h"
#include "disort_device.cuh"
#include "flux_inst.cuh"
#include "host_tables.hpp"

namespace b200 {

static const FluxInst* inst_of(int nstr) {
  switch (nstr) {
    case 2: return flux_inst_1();
    case 4: return flux_inst_2();
    case 6: return flux_inst_3();
    case 8: return flux_inst_4();
    case 12: return flux_inst_6();
    case 16: return flux_inst_8();
    case 24: return flux_inst_12();
    case 32: return flux_inst_16();
    default: return nullptr;
  }
}

// ---------------------------------------------------------------------------
// K2: Planck pre-pass.  pkag[band][col][lev] = PLKAVG(wlo_band, whi_band, T(col, lev)),
// DISORT (top-down) level order; temf is pyharp bottom-up.  Hoisted out of the solve so
// g-points sharing a band (amars_lw: all of them) do not recompute it.
// ---------------------------------------------------------------------------
__global__ void planck_prepass_kernel(int nband, int ncol, int nlyr, const double* __restrict__ blo,
                                      const double* __restrict__ bhi, const double* __restrict__ temf,
                                      double* __restrict__ pkag, int* __restrict__ bad) {
  const long total = (long)nband * ncol * (nlyr + 1);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int lev = (int)(i % (nlyr + 1));
    const long bj = i / (nlyr + 1);
    const int j = (int)(bj % ncol);
    const int b = (int)(bj / ncol);
    const double t = temf[(long)j * (nlyr + 1) + (nlyr - lev)];
    if (!(t >= 0.0)) {
      *bad = 1;
      pkag[i] = nan("");
    } else {
      pkag[i] = plkavg(blo[b], bhi[b], t);
    }
  }
}

// ---------------------------------------------------------------------------
// K5: weighted band reduction bflx(ncol,nlev,2) = sum_n w_n flx(n,ncol,nlev,2)
// ---------------------------------------------------------------------------
__global__ void reduce_band_kernel(int nwave, long inner, const double* __restrict__ flx,
                                   const double* __restrict__ w, double* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < inner; i += (long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int n = 0; n < nwave; ++n) acc = fma(w[n], flx[(long)n * inner + i], acc);
    out[i] = acc;
  }
}

__global__ void count_status_kernel(const int32_t* __restrict__ st, long n, int* __restrict__ nbad) {
  int local = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    if (st[i] != 0) ++local;
  if (local) atomicAdd(nbad, local);
}

// DFMA-saturating kernel: 8 independent chains per thread
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

}  // namespace b200

using namespace b200;

struct b200disort_handle {
  b200disort_config cfg;
  const b200::FluxInst* inst = nullptr;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  std::vector<double> user_tau, user_mu, user_phi, wave_lower, wave_upper;
  // device-side constants
  double *d_wlo = nullptr, *d_whi = nullptr, *d_blo = nullptr, *d_bhi = nullptr;
  int* d_w2b = nullptr;
  int nband = 0;
  double* d_pkag = nullptr;
  size_t pkag_elems = 0;
  double* d_ws = nullptr;
  size_t ws_bytes = 0;
  int* d_flag = nullptr;  // [0] planck bad-T flag, [1] bad status counter
  // staging for the host-pointer entry point
  double *h_pin = nullptr, *d_stage = nullptr;
  size_t pin_bytes = 0, stage_bytes = 0;
  int32_t* d_status_own = nullptr;
  // ring of event pairs around the solve kernel: get_stats sums every solve since the last call
  static constexpr int kRing = 256;
  cudaEvent_t ev0 = nullptr, evA[kRing] = {nullptr}, evB[kRing] = {nullptr};
  int ring_used = 0;
  double ring_carry_ms = 0.0;  // durations folded in when the ring wrapped
  int solves_since_stats = 0;
  b200disort_stats stats{};
};

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      std::fprintf(stderr, "b200disort: CUDA error '%s' at %s:%d\n", cudaGetErrorString(e__), \
                   __FILE__, __LINE__);                                                      \
      return B200DISORT_ERR_CUDA;                                                            \
    }                                                                                        \
  } while (0)

static int nlev_out(const b200disort_config& c) {
  return (c.flags & B200DISORT_USRTAU) ? c.ntau : c.nlyr + 1;
}

extern "C" const char* b200disort_version(void) { return "b200disort 0.1 (sm_100a, fp64)"; }

extern "C" const char* b200disort_strerror(int code) {
  switch (code) {
    case B200DISORT_OK: return "ok";
    case B200DISORT_ERR_INVALID: return "invalid argument or configuration";
    case B200DISORT_ERR_CUDA: return "CUDA failure (no device, out of memory, or launch error); there is no CPU fallback";
    case B200DISORT_ERR_UNSUPPORTED: return "configuration not supported by the B200 solve path";
    case B200DISORT_ERR_SOLVE: return "one or more solves reported a non-zero status";
    default: return "unknown error";
  }
}

extern "C" int b200disort_create(const b200disort_config* cfg, b200disort_handle** out) {
  if (!cfg || !out) return B200DISORT_ERR_INVALID;
  *out = nullptr;
  if (cfg->nwave < 1 || cfg->ncol < 1 || cfg->nlyr < 1) return B200DISORT_ERR_INVALID;
  if (cfg->nstr < 2 || (cfg->nstr & 1) || cfg->nmom < cfg->nstr) return B200DISORT_ERR_INVALID;
  if (!(cfg->flags & B200DISORT_LAMBER)) return B200DISORT_ERR_UNSUPPORTED;  // BRDF not on pyharp's path
  const FluxInst* inst = inst_of(cfg->nstr);
  if (!inst) return B200DISORT_ERR_UNSUPPORTED;  /* nstr in {2,4,6,8,12,16,24,32} */
  if ((cfg->flags & B200DISORT_USRTAU) && (cfg->ntau < 1 || !cfg->user_tau)) return B200DISORT_ERR_INVALID;
  const bool onlyfl = cfg->flags & B200DISORT_ONLYFL;
  if (!onlyfl && (cfg->flags & B200DISORT_USRANG) && (cfg->numu < 1 || !cfg->user_mu)) return B200DISORT_ERR_INVALID;
  if (!onlyfl && (cfg->nphi < 1 || !cfg->user_phi)) return B200DISORT_ERR_INVALID;
  if ((cfg->flags & B200DISORT_PLANCK) && (!cfg->wave_lower || !cfg->wave_upper)) return B200DISORT_ERR_INVALID;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    std::fprintf(stderr, "b200disort: no CUDA device; this library has no CPU fallback\n");
    return B200DISORT_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(cfg->device));

  b200disort_handle* h = new (std::nothrow) b200disort_handle();
  if (!h) return B200DISORT_ERR_INVALID;
  h->cfg = *cfg;
  h->inst = inst;
  h->stream = (cudaStream_t)cfg->stream;
  if (cfg->user_tau && cfg->ntau > 0) h->user_tau.assign(cfg->user_tau, cfg->user_tau + cfg->ntau);
  if (cfg->user_mu && cfg->numu > 0) h->user_mu.assign(cfg->user_mu, cfg->user_mu + cfg->numu);
  if (cfg->user_phi && cfg->nphi > 0) h->user_phi.assign(cfg->user_phi, cfg->user_phi + cfg->nphi);
  if (cfg->wave_lower) h->wave_lower.assign(cfg->wave_lower, cfg->wave_lower + cfg->nwave);
  if (cfg->wave_upper) h->wave_upper.assign(cfg->wave_upper, cfg->wave_upper + cfg->nwave);
  h->cfg.user_tau = h->user_tau.data();
  h->cfg.user_mu = h->user_mu.data();
  h->cfg.user_phi = h->user_phi.data();
  h->cfg.wave_lower = h->wave_lower.empty() ? nullptr : h->wave_lower.data();
  h->cfg.wave_upper = h->wave_upper.empty() ? nullptr : h->wave_upper.data();

  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  h->sm_count = prop.multiProcessorCount;

  // constant tables for this stream count
  QuadTables qt;
  build_quad_tables(cfg->nstr, &qt);
  CU_TRY(inst->upload_tables(&qt));

  // spectral bands: unique (wvnmlo, wvnmhi) pairs
  if (!h->wave_lower.empty()) {
    std::map<std::pair<double, double>, int> bands;
    std::vector<int> w2b(cfg->nwave);
    std::vector<double> blo, bhi;
    for (int n = 0; n < cfg->nwave; ++n) {
      auto key = std::make_pair(h->wave_lower[n], h->wave_upper[n]);
      auto it = bands.find(key);
      if (it == bands.end()) {
        it = bands.emplace(key, (int)blo.size()).first;
        blo.push_back(key.first);
        bhi.push_back(key.second);
      }
      w2b[n] = it->second;
      if (!(key.first >= 0.0) || !(key.second >= key.first)) {
        delete h;
        return B200DISORT_ERR_INVALID;
      }
    }
    h->nband = (int)blo.size();
    CU_TRY(cudaMalloc(&h->d_wlo, sizeof(double) * cfg->nwave));
    CU_TRY(cudaMalloc(&h->d_whi, sizeof(double) * cfg->nwave));
    CU_TRY(cudaMalloc(&h->d_blo, sizeof(double) * h->nband));
    CU_TRY(cudaMalloc(&h->d_bhi, sizeof(double) * h->nband));
    CU_TRY(cudaMalloc(&h->d_w2b, sizeof(int) * cfg->nwave));
    CU_TRY(cudaMemcpy(h->d_wlo, h->wave_lower.data(), sizeof(double) * cfg->nwave, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_whi, h->wave_upper.data(), sizeof(double) * cfg->nwave, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_blo, blo.data(), sizeof(double) * h->nband, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_bhi, bhi.data(), sizeof(double) * h->nband, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(h->d_w2b, w2b.data(), sizeof(int) * cfg->nwave, cudaMemcpyHostToDevice));
  }
  CU_TRY(cudaMalloc(&h->d_flag, sizeof(int) * 2));
  CU_TRY(cudaMemset(h->d_flag, 0, sizeof(int) * 2));
  CU_TRY(cudaEventCreate(&h->ev0));
  for (int i = 0; i < b200disort_handle::kRing; ++i) {
    CU_TRY(cudaEventCreate(&h->evA[i]));
    CU_TRY(cudaEventCreate(&h->evB[i]));
  }
  *out = h;
  return B200DISORT_OK;
}

extern "C" void b200disort_destroy(b200disort_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_wlo); cudaFree(h->d_whi); cudaFree(h->d_blo); cudaFree(h->d_bhi);
  cudaFree(h->d_w2b); cudaFree(h->d_pkag); cudaFree(h->d_ws); cudaFree(h->d_flag);
  cudaFree(h->d_stage); cudaFree(h->d_status_own);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->ev0) cudaEventDestroy(h->ev0);
  for (int i = 0; i < b200disort_handle::kRing; ++i) {
    if (h->evA[i]) cudaEventDestroy(h->evA[i]);
    if (h->evB[i]) cudaEventDestroy(h->evB[i]);
  }
  delete h;
}

static int launch_flux(b200disort_handle* h, SolveParams& p) {
  const FluxInst* inst = h->inst;
  const int block = 128;
  int occ = 0;
  CU_TRY(inst->occupancy(block, &occ));
  if (occ < 1) occ = 1;
  const long nsolve = (long)p.nwave * p.ncol;
  long grid = std::min<long>((nsolve + block - 1) / block, (long)h->sm_count * occ);
  const long nthreads = grid * block;
  const size_t need = (size_t)p.nlyr * inst->nsave * nthreads * sizeof(double);
  if (need > h->ws_bytes) {
    if (h->d_ws) CU_TRY(cudaFree(h->d_ws));
    h->d_ws = nullptr;
    h->ws_bytes = 0;
    CU_TRY(cudaMalloc(&h->d_ws, need));
    h->ws_bytes = need;
  }
  p.ws = h->d_ws;
  p.ws_threads = nthreads;
  if (h->ring_used == b200disort_handle::kRing) {  // fold the ring (blocks until those solves finished)
    CU_TRY(cudaEventSynchronize(h->evB[h->ring_used - 1]));
    for (int i = 0; i < h->ring_used; ++i) {
      float ms = 0.f;
      CU_TRY(cudaEventElapsedTime(&ms, h->evA[i], h->evB[i]));
      h->ring_carry_ms += ms;
    }
    h->ring_used = 0;
  }
  CU_TRY(cudaEventRecord(h->evA[h->ring_used], h->stream));
  CU_TRY(inst->launch(p, (unsigned)grid, block, h->stream));
  CU_TRY(cudaEventRecord(h->evB[h->ring_used], h->stream));
  h->ring_used += 1;
  h->solves_since_stats += 1;
  h->stats.kernel_launches += 1;
  h->stats.workspace_bytes = h->ws_bytes;
  return B200DISORT_OK;
}

extern "C" int b200disort_solve(b200disort_handle* h, const double* prop, int nprop,
                                const b200disort_bc* bc, const double* temf, double* flx,
                                double* uu, int32_t* status) {
  if (!h || !prop || !flx || !status || nprop < 1) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  const bool planck = c.flags & B200DISORT_PLANCK;
  const bool onlyfl = c.flags & B200DISORT_ONLYFL;
  if (planck && !temf) return B200DISORT_ERR_INVALID;
  if (!onlyfl || (c.flags & B200DISORT_USRTAU)) {
    (void)uu;
    return B200DISORT_ERR_UNSUPPORTED;  // intensity / user-level path: see b200disort_intensity.cu (next)
  }
  CU_TRY(cudaSetDevice(c.device));
  h->stats.kernel_launches = 0;
  CU_TRY(cudaEventRecord(h->ev0, h->stream));

  SolveParams p{};
  p.nwave = c.nwave; p.ncol = c.ncol; p.nlyr = c.nlyr; p.nprop = nprop; p.nstr = c.nstr;
  p.flags = c.flags;
  p.prop = prop;
  if (bc) {
    p.umu0 = bc->umu0; p.fbeam = bc->fbeam; p.albedo = bc->albedo; p.btemp = bc->btemp;
    p.ttemp = bc->ttemp; p.temis = bc->temis; p.fisot = bc->fisot;
  }
  p.flx = flx;
  p.status = status;
  p.wave_lower = h->d_wlo;
  p.wave_upper = h->d_whi;
  p.wave2band = h->d_w2b;

  if (planck) {
    const size_t elems = (size_t)h->nband * c.ncol * (c.nlyr + 1);
    if (elems > h->pkag_elems) {
      if (h->d_pkag) CU_TRY(cudaFree(h->d_pkag));
      h->d_pkag = nullptr;
      h->pkag_elems = 0;
      CU_TRY(cudaMalloc(&h->d_pkag, elems * sizeof(double)));
      h->pkag_elems = elems;
    }
    const int block = 256;
    const long grid = std::min<long>((long)((elems + block - 1) / block), (long)h->sm_count * 8);
    planck_prepass_kernel<<<(unsigned)grid, block, 0, h->stream>>>(h->nband, c.ncol, c.nlyr, h->d_blo, h->d_bhi,
                                                                   temf, h->d_pkag, h->d_flag);
    CU_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    p.pkag = h->d_pkag;
  }

  return launch_flux(h, p);
}

extern "C" int b200disort_get_stats(b200disort_handle* h, b200disort_stats* out) {
  if (!h || !out) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(h->cfg.device));
  double sum = h->ring_carry_ms;
  float tot = 0.f;
  if (h->ring_used > 0) {
    CU_TRY(cudaEventSynchronize(h->evB[h->ring_used - 1]));
    for (int i = 0; i < h->ring_used; ++i) {
      float ms = 0.f;
      CU_TRY(cudaEventElapsedTime(&ms, h->evA[i], h->evB[i]));
      sum += ms;
    }
    CU_TRY(cudaEventElapsedTime(&tot, h->ev0, h->evB[h->ring_used - 1]));
  }
  h->stats.solve_kernel_ms = (float)sum;
  h->stats.solves_timed = h->solves_since_stats;
  h->stats.total_device_ms = tot;
  *out = h->stats;
  h->ring_used = 0;
  h->ring_carry_ms = 0.0;
  h->solves_since_stats = 0;
  return B200DISORT_OK;
}

extern "C" int b200disort_check_status(b200disort_handle* h, const int32_t* status, int* n_bad) {
  if (!h || !status || !n_bad) return B200DISORT_ERR_INVALID;
  CU_TRY(cudaSetDevice(h->cfg.device));
  const long n = (long)h->cfg.nwave * h->cfg.ncol;
  CU_TRY(cudaMemsetAsync(h->d_flag + 1, 0, sizeof(int), h->stream));
  count_status_kernel<<<(unsigned)std::min<long>((n + 255) / 256, 1024), 256, 0, h->stream>>>(status, n, h->d_flag + 1);
  CU_TRY(cudaGetLastError());
  int flags[2] = {0, 0};
  CU_TRY(cudaMemcpyAsync(flags, h->d_flag, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  *n_bad = flags[1];
  if (flags[0]) {  // a negative / NaN level temperature reached the Planck pre-pass
    CU_TRY(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    return B200DISORT_ERR_SOLVE;
  }
  return flags[1] ? B200DISORT_ERR_SOLVE : B200DISORT_OK;
}

extern "C" int b200disort_solve_host(b200disort_handle* h, const double* prop, int nprop,
                                     const b200disort_bc* bc, const double* temf, double* flx,
                                     double* uu, int32_t* status) {
  if (!h || !prop || !flx || nprop < 1) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  CU_TRY(cudaSetDevice(c.device));
  const size_t nsolve = (size_t)c.nwave * c.ncol;
  const size_t n_prop = nsolve * c.nlyr * nprop;
  const size_t n_temf = temf ? (size_t)c.ncol * (c.nlyr + 1) : 0;
  const double* bcs[8] = {nullptr};
  if (bc) {
    bcs[0] = bc->umu0; bcs[1] = bc->fbeam; bcs[2] = bc->albedo; bcs[3] = bc->btemp;
    bcs[4] = bc->ttemp; bcs[5] = bc->temis; bcs[6] = bc->fisot; bcs[7] = bc->phi0;
  }
  size_t n_bc = 0;
  for (int i = 0; i < 8; ++i) if (bcs[i]) n_bc += nsolve;
  const size_t n_flx = nsolve * nlev_out(c) * 2;
  const size_t n_in = n_prop + n_temf + n_bc;
  const size_t n_status = (nsolve * sizeof(int32_t) + 7) / 8;
  const size_t total = n_in + n_flx + n_status;
  if (total * sizeof(double) > h->stage_bytes) {
    if (h->d_stage) CU_TRY(cudaFree(h->d_stage));
    if (h->h_pin) CU_TRY(cudaFreeHost(h->h_pin));
    h->d_stage = nullptr; h->h_pin = nullptr; h->stage_bytes = 0;
    CU_TRY(cudaMalloc(&h->d_stage, total * sizeof(double)));
    CU_TRY(cudaMallocHost(&h->h_pin, total * sizeof(double)));
    h->stage_bytes = total * sizeof(double);
  }
  // pack inputs into the pinned block, one H2D copy
  double* hp = h->h_pin;
  size_t off = 0;
  std::memcpy(hp + off, prop, n_prop * sizeof(double));
  const double* d_prop = h->d_stage + off;
  off += n_prop;
  const double* d_temf = nullptr;
  if (temf) {
    std::memcpy(hp + off, temf, n_temf * sizeof(double));
    d_temf = h->d_stage + off;
    off += n_temf;
  }
  const double* d_bcs[8] = {nullptr};
  for (int i = 0; i < 8; ++i)
    if (bcs[i]) {
      std::memcpy(hp + off, bcs[i], nsolve * sizeof(double));
      d_bcs[i] = h->d_stage + off;
      off += nsolve;
    }
  CU_TRY(cudaMemcpyAsync(h->d_stage, hp, n_in * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  double* d_flx = h->d_stage + n_in;
  int32_t* d_status = reinterpret_cast<int32_t*>(h->d_stage + n_in + n_flx);
  b200disort_bc dbc{d_bcs[0], d_bcs[1], d_bcs[2], d_bcs[3], d_bcs[4], d_bcs[5], d_bcs[6], d_bcs[7]};
  int rc = b200disort_solve(h, d_prop, nprop, &dbc, d_temf, d_flx, nullptr, d_status);
  if (rc != B200DISORT_OK) return rc;
  (void)uu;
  CU_TRY(cudaMemcpyAsync(hp + n_in, d_flx, (n_flx + n_status) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  int nbad = 0;
  rc = b200disort_check_status(h, d_status, &nbad);  // synchronises the stream
  std::memcpy(flx, hp + n_in, n_flx * sizeof(double));
  if (status) std::memcpy(status, hp + n_in + n_flx, nsolve * sizeof(int32_t));
  return rc;
}

extern "C" int b200disort_get_rad(b200disort_handle* h, double* uu) {
  (void)h; (void)uu;
  return B200DISORT_ERR_UNSUPPORTED;
}

extern "C" int b200disort_reduce_band(b200disort_handle* h, const double* flx, const double* weights,
                                      double* bflx) {
  if (!h || !flx || !weights || !bflx) return B200DISORT_ERR_INVALID;
  const b200disort_config& c = h->cfg;
  CU_TRY(cudaSetDevice(c.device));
  const long inner = (long)c.ncol * nlev_out(c) * 2;
  const int block = 256;
  const long grid = std::min<long>((inner + block - 1) / block, (long)h->sm_count * 8);
  reduce_band_kernel<<<(unsigned)grid, block, 0, h->stream>>>(c.nwave, inner, flx, weights, bflx);
  CU_TRY(cudaGetLastError());
  return B200DISORT_OK;
}

extern "C" int b200disort_fp64_peak(int device, double* tflops) {
  if (!tflops) return B200DISORT_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return B200DISORT_ERR_CUDA;
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CU_TRY(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  const int nb = prop.multiProcessorCount * 4, tpb = 512, iters = 8192;
  dfma_peak_kernel<<<nb, tpb>>>(d, 256, 1.0000001, 1e-9);  // warm-up
  CU_TRY(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CU_TRY(cudaEventRecord(e0));
    dfma_peak_kernel<<<nb, tpb>>>(d, iters, 1.0000001, 1e-9);
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  *tflops = 2.0 * (double)nb * tpb * 64.0 * iters / (best * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return B200DISORT_OK;
}
